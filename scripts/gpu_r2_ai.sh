#!/bin/bash
# round 2, call AI: indirect-projector solves against the fp64 oracle with / without the y recurrence
cd "$GRAFT_REPO_ROOT"
timeout 600 python scripts/dev/indirect_probe.py 2>&1 | tail -24

#!/bin/bash
# round 2, call H: ncu evidence for the one-launch iteration kernel (outputs kept small: raw page as CSV, the
# source page reduced to its hottest lines; the .ncu-rep files are deleted on the box)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out /tmp/prof
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_c2.csv python bench.py --steps 40 --warmup 5 --no-cpu --no-e2e --no-converged > gpurun_out/r2h_ncu_list.log 2>&1
for CFG in c2 c4; do
POGS_B200_NO_GRAPH=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_admm_pass -s 9 -c 2 -o /tmp/prof/admm_$CFG -f python bench.py --config $CFG --steps 8 --warmup 3 --no-cpu --no-e2e --no-converged > gpurun_out/r2h_ncu_full_$CFG.log 2>&1
ncu -i /tmp/prof/admm_$CFG.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_admm_pass_$CFG.csv 2>/dev/null
ncu -i /tmp/prof/admm_$CFG.ncu-rep --page source --csv > /tmp/prof/src_$CFG.csv 2>/dev/null
python - "$CFG" <<'PY'
import csv, sys
cfg = sys.argv[1]
rows = list(csv.reader(open(f"/tmp/prof/src_{cfg}.csv", errors="replace")))
# keep the header block and the 80 lines with the most warp-stall samples
hdr = None
for i, r in enumerate(rows):
    if any("Sampling" in c or "Samples" in c for c in r):
        hdr = i; break
out = rows[: (hdr or 0) + 1]
if hdr is not None:
    col = next((j for j, c in enumerate(rows[hdr]) if "Samples" in c or "Sampling" in c), None)
    def val(r):
        try: return float(r[col].replace(",", ""))
        except Exception: return -1.0
    body = sorted(rows[hdr + 1:], key=val, reverse=True)[:80]
    out += body
csv.writer(open(f"gpurun_out/r02_ncu_source_top_admm_pass_{cfg}.csv", "w")).writerows(out)
PY
done
rm -rf /tmp/prof
ls -la gpurun_out | grep -E "r02_|r2h_"
du -sh gpurun_out

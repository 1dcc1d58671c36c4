#!/bin/bash
# round 2, last call: the final build once more through what the driver runs (full GPU suite, smoke, C2 line at K = 20) and the
# full C5 line (pinned CSR arrays for the e2e call, three products per iteration)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1
tail -3 gpurun_out/r2g_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2g_bench_c2_k20.json 2> gpurun_out/r2g_bench_c2_k20.err
timeout 900 python bench.py --config c5 --steps 100 --warmup 10 > gpurun_out/r2g_bench_c5.json 2> gpurun_out/r2g_bench_c5.err
timeout 600 python bench.py --config c5s --steps 100 --warmup 10 > gpurun_out/r2g_bench_c5s.json 2> gpurun_out/r2g_bench_c5s.err
tail -c 300 gpurun_out/r2g_bench_c2_k20.err gpurun_out/r2g_bench_c5.err
python - <<'PY'
import json
for f in ("r2g_bench_c2_k20","r2g_bench_c5","r2g_bench_c5s"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        c=d.get("converged") or {}
        print(f, round(d["value"],1), "ms/step", round(d["ms_per_step"],4), "e2e", (d.get("e2e") or {}).get("value"), "conv", c.get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "frac", (d.get("roofline") or {}).get("frac"), "parity", ((d.get("sanity") or {}).get("parity") or {}).get("ok"), "failed" if d.get("failed") else "")
    except Exception as e:
        print(f, "ERR", e)
PY

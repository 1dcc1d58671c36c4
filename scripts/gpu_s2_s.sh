set -x
mkdir -p gpurun_out
POGS_B200_SPMV=blocked timeout 600 python -m pytest tests/test_gpu_sparse.py -m gpu -q --timeout=300 -p no:cacheprovider 2>&1 | tail -4
POGS_B200_SPMV=blocked timeout 400 python bench.py --config c5 --steps 50 --warmup 5 > gpurun_out/s2s_bench_c5_blocked.json 2> gpurun_out/s2s_bench_c5_blocked.err; tail -2 gpurun_out/s2s_bench_c5_blocked.err
python - <<'P'
import json
for f in ["c5_blocked"]:
    try:
        d=json.loads([l for l in open("gpurun_out/s2s_bench_%s.json"%f) if l.startswith("{")][0]); print(f, round(d["value"],1), d["ms_per_step"], d["cgls_inner_per_iteration"], round(d["roofline"]["frac"],3), d["setup_ms"], d["converged_run"])
    except Exception as e: print(f, "ERR", e)
P

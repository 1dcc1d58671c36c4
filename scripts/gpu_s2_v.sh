set -x
mkdir -p gpurun_out
export POGS_B200_SYMV=1
timeout 900 python -m pytest tests -m gpu -q --timeout=600 --maxfail=10 -p no:cacheprovider 2>&1 | tail -6
python bench.py --steps 200 --warmup 20 --no-cpu --no-e2e > gpurun_out/s2v_bench_c2.json 2> gpurun_out/s2v_bench_c2.err; tail -2 gpurun_out/s2v_bench_c2.err
python - <<'P'
import json
for f in ["c2"]:
    try:
        d=json.loads([l for l in open("gpurun_out/s2v_bench_%s.json"%f) if l.startswith("{")][0]); print(f, round(d["value"],1), d["ms_per_step"], d["roofline"].get("phases_ms"), round(d["roofline"]["frac"],3))
    except Exception as e: print(f, "ERR", e)
P

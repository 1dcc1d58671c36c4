set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 --maxfail=10 -p no:cacheprovider 2>&1 | tail -8
python bench.py --steps 200 --warmup 20 --no-cpu > gpurun_out/s2u_bench_c2.json 2> gpurun_out/s2u_bench_c2.err; tail -2 gpurun_out/s2u_bench_c2.err
python bench.py --config c3 --steps 300 --warmup 20 --no-cpu --no-e2e > gpurun_out/s2u_bench_c3.json 2> gpurun_out/s2u_bench_c3.err
python - <<'P'
import json
for f in ["c2","c3"]:
    try:
        d=json.loads([l for l in open("gpurun_out/s2u_bench_%s.json"%f) if l.startswith("{")][0]); print(f, round(d["value"],1), d["ms_per_step"], d.get("e2e") and round(d["e2e"]["value"],1), d["roofline"].get("phases_ms"), round(d["roofline"]["frac"],3))
    except Exception as e: print(f, "ERR", e)
P

# session 2, call L: W-map-warp pipeline: suite, C2/C4/C3(forced fuse)/C5 bench lines
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 --maxfail=15 -p no:cacheprovider 2>&1 | tail -12
python bench.py --steps 200 --warmup 20 --no-cpu > gpurun_out/s2l_bench_c2.json 2> gpurun_out/s2l_bench_c2.err; python - <<'P'
import json
for f in ["gpurun_out/s2l_bench_c2.json"]:
    d=json.loads([l for l in open(f) if l.startswith("{")][0]); print(f, round(d["value"],1), d["ms_per_step"], d["e2e"], d["roofline"]["phases_ms"], d["roofline"]["frac"], d["clocks"])
P
tail -3 gpurun_out/s2l_bench_c2.err
for cfg in c4 c3; do python bench.py --config $cfg --steps 300 --warmup 20 --no-cpu > gpurun_out/s2l_bench_$cfg.json 2> gpurun_out/s2l_bench_$cfg.err; tail -3 gpurun_out/s2l_bench_$cfg.err; done
POGS_B200_FORCE_FUSE=1 python bench.py --config c3 --steps 300 --warmup 20 --no-cpu --no-e2e > gpurun_out/s2l_bench_c3_forced.json 2> gpurun_out/s2l_bench_c3_forced.err; tail -3 gpurun_out/s2l_bench_c3_forced.err
timeout 600 python bench.py --config c5 --steps 50 --warmup 5 > gpurun_out/s2l_bench_c5.json 2> gpurun_out/s2l_bench_c5.err; tail -3 gpurun_out/s2l_bench_c5.err
python - <<'P'
import json
for f in ["c4","c3","c3_forced","c5"]:
    try:
        d=json.loads([l for l in open("gpurun_out/s2l_bench_%s.json"%f) if l.startswith("{")][0]); print(f, round(d["value"],1), d["ms_per_step"], d.get("e2e") and round(d["e2e"]["value"],1), d["roofline"].get("phases_ms"), round(d["roofline"]["frac"],3))
    except Exception as e: print(f, "ERR", e)
P

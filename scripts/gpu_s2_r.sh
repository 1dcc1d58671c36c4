# session 2, call R: final validation of the round: suite, blocked sparse layout (opt-in), bench lines, launch list
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 --maxfail=15 -p no:cacheprovider 2>&1 | tail -8
POGS_B200_SPMV=blocked timeout 600 python -m pytest tests/test_gpu_sparse.py -m gpu -q --timeout=300 -p no:cacheprovider 2>&1 | tail -8
timeout 400 python bench.py --config c5 --steps 50 --warmup 5 > gpurun_out/s2r_bench_c5_plain.json 2> gpurun_out/s2r_bench_c5_plain.err; tail -2 gpurun_out/s2r_bench_c5_plain.err
POGS_B200_SPMV=blocked timeout 400 python bench.py --config c5 --steps 50 --warmup 5 > gpurun_out/s2r_bench_c5_blocked.json 2> gpurun_out/s2r_bench_c5_blocked.err; tail -2 gpurun_out/s2r_bench_c5_blocked.err
python - <<'P'
import json
for f in ["c5_plain","c5_blocked"]:
    try:
        d=json.loads([l for l in open("gpurun_out/s2r_bench_%s.json"%f) if l.startswith("{")][0]); print(f, round(d["value"],1), d["ms_per_step"], d["cgls_inner_per_iteration"], round(d["roofline"]["frac"],3), d["setup_ms"], d["converged_run"])
    except Exception as e: print(f, "ERR", e)
P
python bench.py --steps 200 --warmup 20 > gpurun_out/s2r_bench_c2.json 2> gpurun_out/s2r_bench_c2.err; tail -c 3400 gpurun_out/s2r_bench_c2.json; tail -3 gpurun_out/s2r_bench_c2.err
python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/s2r_bench_ref.json 2> gpurun_out/s2r_bench_ref.err; tail -c 900 gpurun_out/s2r_bench_ref.json
POGS_B200_NO_COND=1 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_fused_pass|k_rowdot|k_colacc|k_prox|k_control|k_gram|k_split' -s 110 -c 60 --csv --log-file gpurun_out/s2r_launches_c2.csv python bench.py --steps 6 --warmup 4 --no-e2e --no-cpu > gpurun_out/s2r_ncu_list.log 2>&1; tail -2 gpurun_out/s2r_ncu_list.log

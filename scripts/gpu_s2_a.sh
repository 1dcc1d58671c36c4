# session 2, call A: full GPU test suite, C2 bench with and without the single-pass kernel,
# launch list and one full ncu capture of the single-pass kernel
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 --maxfail=15 -p no:cacheprovider 2>&1 | tail -25
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
python bench.py --steps 200 --warmup 20 > gpurun_out/s2a_bench_c2.json 2> gpurun_out/s2a_bench_c2.err; tail -c 3500 gpurun_out/s2a_bench_c2.json; tail -3 gpurun_out/s2a_bench_c2.err
POGS_B200_NO_FUSE=1 python bench.py --steps 200 --warmup 20 --no-cpu --no-e2e > gpurun_out/s2a_bench_c2_twopass.json 2> gpurun_out/s2a_bench_c2_twopass.err; tail -c 2500 gpurun_out/s2a_bench_c2_twopass.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_' -s 300 -c 150 --csv --log-file gpurun_out/s2a_launches_c2.csv python bench.py --steps 6 --warmup 4 --no-e2e --no-cpu > gpurun_out/s2a_ncu_list.log 2>&1; tail -3 gpurun_out/s2a_ncu_list.log
ncu --set full --clock-control none --import-source on -k regex:k_fused_pass -s 6 -c 1 -o gpurun_out/s2a_prof_fused_c2 -f python bench.py --steps 6 --warmup 4 --no-e2e --no-cpu > gpurun_out/s2a_ncu_fused.log 2>&1; tail -3 gpurun_out/s2a_ncu_fused.log
ls -la gpurun_out/

set -x
python -m pytest tests/test_gpu_units.py -m gpu -q --timeout=300 -x -p no:cacheprovider -k "gemv" 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
python bench.py --steps 100 --warmup 10 > gpurun_out/bench_r1_a.json 2> gpurun_out/bench_r1_a.err; tail -c 3000 gpurun_out/bench_r1_a.json; tail -5 gpurun_out/bench_r1_a.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_rowdot|k_colacc|k_prox|k_control' -s 200 -c 120 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_list.log 2>&1; tail -3 gpurun_out/ncu_list.log
ncu --set full --clock-control none --import-source on -k regex:k_colacc -s 100 -c 3 -o gpurun_out/prof_colacc_r1 -f python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_colacc.log 2>&1; tail -3 gpurun_out/ncu_colacc.log
ncu --set full --clock-control none --import-source on -k regex:k_rowdot -s 101 -c 5 -o gpurun_out/prof_rowdot_r1 -f python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_rowdot.log 2>&1; tail -3 gpurun_out/ncu_rowdot.log
ls -la gpurun_out/

set -x
mkdir -p gpurun_out
CFG=${1:-c3}
ncu --set full --clock-control none -k regex:k_fused_pass -s 2 -c 1 -o gpurun_out/prof_fused_$CFG -f python bench.py --config $CFG --steps 6 --warmup 4 --no-e2e --no-cpu > gpurun_out/ncu_fused_$CFG.log 2>&1; tail -2 gpurun_out/ncu_fused_$CFG.log
ls -la gpurun_out

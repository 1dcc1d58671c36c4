set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_fused_pass -s 1 -c 4 -o gpurun_out/prof_fused_c2 -f python bench.py --steps 6 --warmup 4 --no-e2e --no-cpu > gpurun_out/ncu_fused_c2.log 2>&1; tail -2 gpurun_out/ncu_fused_c2.log
ncu --set full --clock-control none --import-source on -k regex:k_fused_pass -s 1 -c 4 -o gpurun_out/prof_fused_c3 -f python bench.py --config c3 --steps 6 --warmup 4 --no-e2e --no-cpu > gpurun_out/ncu_fused_c3.log 2>&1; tail -2 gpurun_out/ncu_fused_c3.log

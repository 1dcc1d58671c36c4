set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_solve.py tests/test_gpu_units.py -m gpu -q --timeout=600 --maxfail=15 -p no:cacheprovider -x 2>&1 | tail -15
for cfg in c2 c3 c4; do
  python bench.py --config $cfg --steps 200 --warmup 20 --no-cpu --no-e2e 2> gpurun_out/bench_f_$cfg.err | grep '^{' > gpurun_out/bench_f_$cfg.json
  python -c "
import json; d=json.load(open('gpurun_out/bench_f_$cfg.json')); print('$cfg fused', round(d['value'],1), d['ms_per_step'], {k:round(v*1e3,1) for k,v in d['roofline']['phases_ms'].items()}, d['gpu_launches'], d['roofline']['single_pass_iterations'], d['roofline']['frac'])"
  tail -3 gpurun_out/bench_f_$cfg.err
done

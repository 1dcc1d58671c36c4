set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=600 --maxfail=10 -p no:cacheprovider 2>&1 | tail -15
python bench.py --steps 200 --warmup 20 > gpurun_out/bench_r1_c2.json 2> gpurun_out/bench_r1_c2.err; tail -c 1500 gpurun_out/bench_r1_c2.json; tail -3 gpurun_out/bench_r1_c2.err
python bench.py --config c3 --steps 500 --warmup 20 --no-cpu > gpurun_out/bench_r1_c3.json 2> gpurun_out/bench_r1_c3.err; tail -c 1800 gpurun_out/bench_r1_c3.json; tail -3 gpurun_out/bench_r1_c3.err
python bench.py --config c4 --steps 200 --warmup 20 --no-cpu > gpurun_out/bench_r1_c4.json 2> gpurun_out/bench_r1_c4.err; tail -c 1800 gpurun_out/bench_r1_c4.json; tail -3 gpurun_out/bench_r1_c4.err
python bench.py --config c5s --steps 50 --warmup 5 > gpurun_out/bench_r1_c5s.json 2> gpurun_out/bench_r1_c5s.err; tail -c 1800 gpurun_out/bench_r1_c5s.json; tail -3 gpurun_out/bench_r1_c5s.err
timeout 900 python bench.py --config c5 --steps 20 --warmup 3 > gpurun_out/bench_r1_c5.json 2> gpurun_out/bench_r1_c5.err; tail -c 1800 gpurun_out/bench_r1_c5.json; tail -3 gpurun_out/bench_r1_c5.err
python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/bench_r1_ref.json 2> gpurun_out/bench_r1_ref.err; tail -c 1200 gpurun_out/bench_r1_ref.json

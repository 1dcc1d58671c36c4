#!/bin/bash
# round 2, call Z (2 GPUs): full GPU suite incl. the row-block tests after the sparse re-layout and the rho-action
# prediction, 2-GPU bench line, per-solve host timeline of the lambda path
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi -L
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2z_pytest.log 2>&1
tail -4 gpurun_out/r2z_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 20 --no-cpu > gpurun_out/r2z_bench_c2_n2.json 2> gpurun_out/r2z_bench_c2_n2.err
tail -c 300 gpurun_out/r2z_bench_c2_n2.err
POGS_B200_TRACE=1 timeout 600 python bench.py --config c3path --no-cpu > gpurun_out/r2z_bench_c3path.json 2> gpurun_out/r2z_bench_c3path.err
grep trace gpurun_out/r2z_bench_c3path.err | sed -n 60,80p
python - <<'PY'
import json
for f in ("r2z_bench_c2_n2","r2z_bench_c3path"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"],1), round(d["ms_per_step"]*1e3,1), "e2e", d.get("e2e"), "conv", d.get("converged"), d["sanity"].get("k_then_k"))
    except Exception as e:
        print(f, "ERR", e)
PY

#!/bin/bash
# Round 2, third GPU visit (2 GPUs): the module's own pixel-tile path over NCCL — parity against one GPU, then strong scaling of C3.
tag=${1:-r2c}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${tag}_gpus.txt
( time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x ) > gpurun_out/${tag}_multi_tests.log 2>&1
tail -5 gpurun_out/${tag}_multi_tests.log
show() {
  python - <<P
import json
try:
    d = json.load(open("gpurun_out/${tag}_$1.json"))
    print("$1", round(d["value"], 3), "Msamples/s e2e", round(d["e2e"]["value"], 3), "ms/step", round(d["ms_per_step"], 2), d["config"].get("collective_ms_per_iteration"))
except Exception as e:
    print("$1 failed", e)
P
}
timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_n1.json 2> gpurun_out/${tag}_n1.err; show n1
for lanes in 4 2; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 --lanes $lanes --no-cpu-baseline > gpurun_out/${tag}_n2_tile_l$lanes.json 2> gpurun_out/${tag}_n2_tile_l$lanes.err; show n2_tile_l$lanes
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 8 --warmup 3 --parallelism iteration --no-cpu-baseline > gpurun_out/${tag}_n2_iter.json 2> gpurun_out/${tag}_n2_iter.err; show n2_iter
tail -3 gpurun_out/${tag}_n2_tile_l4.err
exit 0

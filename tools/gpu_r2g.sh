#!/bin/bash
# Round 2, seventh GPU visit (8 GPUs): the driver's scaling run in small — C3, steps 20 / warmup 5, N = 8 (both multi-GPU modes) and N = 1 on the same box.
tag=${1:-r2g}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${tag}_gpus.txt
show() {
  python - <<P
import json
try:
    d = json.load(open("gpurun_out/${tag}_$1.json"))
    print("$1", d["config"]["mode"], round(d["value"], 3), "Msamples/s e2e", round(d["e2e"]["value"], 3), "ms/step", round(d["ms_per_step"], 2), d["modes"])
except Exception as e:
    print("$1 failed", e)
P
}
timeout 500 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_n1.json 2> gpurun_out/${tag}_n1.err; show n1
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_n8.json 2> gpurun_out/${tag}_n8.err; show n8
tail -3 gpurun_out/${tag}_n8.err
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_n4.json 2> gpurun_out/${tag}_n4.err; show n4
exit 0

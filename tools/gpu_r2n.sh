#!/bin/bash
# Round 2, fourteenth GPU visit (2 GPUs): the multi-GPU tests on the final tree (tiles, lanes, replicas with a camera-split left-over, the path tracer's
# replicas) and N = 2 end to end as the driver launches it.
tag=${1:-r2n}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q ) > gpurun_out/${tag}_multi_tests.log 2>&1
tail -4 gpurun_out/${tag}_multi_tests.log
show() {
  python - <<P
import json
try:
    d = json.load(open("gpurun_out/${tag}_$1.json"))
    print("$1", d["config"]["mode"], round(d["value"], 3), "Msamples/s e2e", round(d["e2e"]["value"], 3), "ms/step", round(d["ms_per_step"], 2), {k: (round(v["value"], 2), round(v["e2e"], 2)) for k, v in d["modes"].items()})
except Exception as e:
    print("$1 failed", e)
P
}
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_n2.json 2> gpurun_out/${tag}_n2.err; show n2
tail -2 gpurun_out/${tag}_n2.err
exit 0

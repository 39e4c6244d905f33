#!/bin/bash
# Round 2, ninth GPU visit (2 GPUs): the 2-GPU tests (tiles, lanes, replicas with a camera-split left-over) and N = 2 end to end, K = 20 and K = 21.
tag=${1:-r2i}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q ) > gpurun_out/${tag}_multi_tests.log 2>&1
tail -4 gpurun_out/${tag}_multi_tests.log
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
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_n1.json 2> gpurun_out/${tag}_n1.err; show n1
for k in 20 21; do
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$((k-20)) bench.py --gpus 2 --steps $k --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_n2_k$k.json 2> gpurun_out/${tag}_n2_k$k.err; show n2_k$k
done
tail -2 gpurun_out/${tag}_n2_k21.err
exit 0

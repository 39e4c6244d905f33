#!/bin/bash
# Round 2, eighth GPU visit (2 GPUs): -m gpu suite after the wide-tree arithmetic change, register A/B of the bounce / connection kernels, N = 2 end to end.
tag=${1:-r2h}
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q -s ) > gpurun_out/${tag}_gpu_tests.log 2>&1
tail -4 gpurun_out/${tag}_gpu_tests.log
grep -h "statistical parity\] C3\|FAILED" gpurun_out/${tag}_gpu_tests.log | cut -c1-330
run() { # name, workload, steps, env...
  local name=$1 wl=$2 steps=$3; shift 3
  env "$@" timeout 400 python bench.py --workload $wl --steps $steps --warmup 3 --lanes 1 --no-cpu-baseline > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
  python - <<P
import json
try:
    d = json.load(open("gpurun_out/${tag}_${name}.json"))
    print("${name}", round(d["value"], 3), "Msamples/s", {k: v for k, v in list(d["roofline"]["kernel_ms_per_iteration"].items())[:10]})
except Exception as e:
    print("${name} failed", e)
P
}
run c3 C3 4 X=1
run c3_mb3 C3 4 ETXB_LIB_FAST=$PWD/etx_tracer_b200/libetx_b200_mb3.so
run c3_mb2 C3 4 ETXB_LIB_FAST=$PWD/etx_tracer_b200/libetx_b200_mb2.so
run c2 C2 8 X=1
run c2_mb3 C2 8 ETXB_LIB_FAST=$PWD/etx_tracer_b200/libetx_b200_mb3.so
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
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_n2.json 2> gpurun_out/${tag}_n2.err; show n2; fi
exit 0

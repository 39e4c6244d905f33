#!/bin/bash
# Second A/B + parity tests + ncu evidence in one visit.
tag=${1:-ab2}
mkdir -p gpurun_out
run() { # name, workload, steps, lanes, env...
  local name=$1 wl=$2 steps=$3 lanes=$4; shift 4
  env "$@" timeout 300 python bench.py --workload $wl --steps $steps --warmup 3 --lanes $lanes --no-cpu-baseline > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
}
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/${tag}_gpu_tests.log 2>&1
tail -3 gpurun_out/${tag}_gpu_tests.log
run c3_perquery_l1 C3 4 1 ETXB_MERGE_BATCHED=0
run c3_batched_l1 C3 4 1 X=1
run c3_batched_l2 C3 6 2 X=1
run c3_batched_l4 C3 8 4 X=1
run c2_l1 C2 12 1 X=1
run c2_l2 C2 12 2 X=1
run c2_l4 C2 12 4 X=1
run c2_l6 C2 18 6 X=1
run c2_l8 C2 24 8 X=1
run c4_l4 C4 8 4 X=1
run c5_l4 C5 8 4 X=1
for f in gpurun_out/${tag}_c*.json; do echo $f; cut -c1-110 $f; done
bash tools/gpu_profile.sh ${tag}

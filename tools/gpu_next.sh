#!/bin/bash
# First GPU visit of the next round: the two default-off experiments of round 1 (written after its GPU budget ended).
#   ETXB_MERGE_TILED=1           cell-tiled Lambert gather (C2-dominant kernel, L2-bandwidth bound today)
#   ETXB_MERGE_MATERIAL_MAJOR=1  material-major order of the gather queue (C3-dominant kernel, instruction-fetch bound today)
# Each: the product-build parts of the parity suite with the switch on (tolerance tests), then bench.py --lanes 1 on / off.
tag=${1:-next}
mkdir -p gpurun_out
run() { # name, workload, steps, env...
  local name=$1 wl=$2 steps=$3; shift 3
  env "$@" timeout 300 python bench.py --workload $wl --steps $steps --warmup 3 --lanes 1 --no-cpu-baseline > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
}
( time ETXB_MERGE_TILED=1 timeout 900 python -m pytest tests -m gpu -q -k "product or tolerance or full_size or material or lanes or in_flight or config" ) > gpurun_out/${tag}_tests_tiled.log 2>&1
tail -3 gpurun_out/${tag}_tests_tiled.log
( time ETXB_MERGE_MATERIAL_MAJOR=1 timeout 900 python -m pytest tests -m gpu -q -k "product or tolerance or material or million or config" ) > gpurun_out/${tag}_tests_matmajor.log 2>&1
tail -3 gpurun_out/${tag}_tests_matmajor.log
run c2_base C2 12 X=1
run c2_tiled C2 12 ETXB_MERGE_TILED=1
run c1_base C1 16 X=1
run c1_tiled C1 16 ETXB_MERGE_TILED=1
run c3_base C3 4 X=1
run c3_matmajor C3 4 ETXB_MERGE_MATERIAL_MAJOR=1
run c3_both C3 4 ETXB_MERGE_MATERIAL_MAJOR=1 ETXB_MERGE_TILED=1
for f in gpurun_out/${tag}_c*.json; do echo $f; cut -c1-110 $f; done
exit 0

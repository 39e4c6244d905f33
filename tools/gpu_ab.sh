#!/bin/bash
# A/B of experimental builds on one box (ETXB_LIB_* override the library the ctypes layer loads; env switches select code paths).
tag=${1:-ab}
mkdir -p gpurun_out
E=$PWD/etx_tracer_b200
run() { # name, workload, steps, env...
  local name=$1 wl=$2 steps=$3; shift 3
  env "$@" timeout 300 python bench.py --workload $wl --steps $steps --warmup 3 --lanes 1 --no-cpu-baseline > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
}
run c2_base C2 12 X=1
run c2_e1 C2 12 ETXB_LIB_FAST=$E/exp_e135.so ETXB_CONNECT_DEFERRED=0
run c2_e13 C2 12 ETXB_LIB_FAST=$E/exp_e135.so
run c2_e13_mb2 C2 12 ETXB_LIB_FAST=$E/exp_e13_mb2.so
run c2_e13_mb3 C2 12 ETXB_LIB_FAST=$E/exp_e13_mb3.so
run c2_e13_plain C2 12 ETXB_LIB_FAST=$E/exp_e13_plain.so
run c3_e1 C3 4 ETXB_LIB_FAST=$E/exp_e135.so ETXB_SORT_MATERIAL=0
run c3_e15 C3 4 ETXB_LIB_FAST=$E/exp_e135.so
run c3_mb2 C3 4 ETXB_LIB_FAST=$E/exp_e13_mb2.so
run c3_mb3 C3 4 ETXB_LIB_FAST=$E/exp_e13_mb3.so
run c3_plain C3 4 ETXB_LIB_FAST=$E/exp_e13_plain.so
run c4_e15 C4 4 ETXB_LIB_FAST=$E/exp_e135.so
env ETXB_LIB_FAST=$E/exp_e135.so timeout 300 python bench.py --steps 12 --warmup 3 --lanes 4 --no-cpu-baseline > gpurun_out/${tag}_c2_e13_l4.json 2>/dev/null
( time ETXB_LIB_FAST=$E/exp_e135.so ETXB_LIB_PARITY=$E/exp_e135_parity.so timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/${tag}_gpu_tests_e135.log 2>&1
tail -3 gpurun_out/${tag}_gpu_tests_e135.log
for f in gpurun_out/${tag}_c*.json; do echo $f; cut -c1-120 $f; done

#!/bin/bash
# Fourth A/B: bounce kernels specialised for plain scenes (template PLAIN) vs the general instantiation, + the parity tests on that build.
tag=${1:-ab4}
mkdir -p gpurun_out
E=$PWD/etx_tracer_b200
run() { # name, workload, steps, lanes, env...
  local name=$1 wl=$2 steps=$3 lanes=$4; shift 4
  env "$@" timeout 300 python bench.py --workload $wl --steps $steps --warmup 3 --lanes $lanes --no-cpu-baseline > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
}
( time ETXB_LIB_FAST=$E/exp_plain.so ETXB_LIB_PARITY=$E/exp_plain_parity.so timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/${tag}_gpu_tests_plain.log 2>&1
tail -3 gpurun_out/${tag}_gpu_tests_plain.log
run c2_general C2 12 1 ETXB_LIB_FAST=$E/exp_plain.so ETXB_PLAIN_KERNELS=0
run c2_plain C2 12 1 ETXB_LIB_FAST=$E/exp_plain.so
run c2_plain_l4 C2 12 4 ETXB_LIB_FAST=$E/exp_plain.so
run c3_general C3 4 1 ETXB_LIB_FAST=$E/exp_plain.so ETXB_PLAIN_KERNELS=0
run c3_plain C3 4 1 ETXB_LIB_FAST=$E/exp_plain.so
run c3_plain_l4 C3 8 4 ETXB_LIB_FAST=$E/exp_plain.so
for f in gpurun_out/${tag}_c*.json; do echo $f; cut -c1-110 $f; done
exit 0

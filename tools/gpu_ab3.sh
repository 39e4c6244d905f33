#!/bin/bash
# Third A/B: launch bounds (registers / occupancy) of the bounce and connection kernels.
tag=${1:-ab3}
mkdir -p gpurun_out
E=$PWD/etx_tracer_b200
run() { # name, workload, steps, env...
  local name=$1 wl=$2 steps=$3; shift 3
  env "$@" timeout 300 python bench.py --workload $wl --steps $steps --warmup 3 --lanes 1 --no-cpu-baseline > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
}
run c2_base C2 12 X=1
for v in b4 c4 b4c4; do
  [ -f $E/exp_$v.so ] && run c2_$v C2 12 ETXB_LIB_FAST=$E/exp_$v.so
done
run c3_base C3 4 X=1
for v in b4 c4 b4c4; do
  [ -f $E/exp_$v.so ] && run c3_$v C3 4 ETXB_LIB_FAST=$E/exp_$v.so
done
for f in gpurun_out/${tag}_c*.json; do echo $f; cut -c1-110 $f; done
for f in gpurun_out/${tag}_c*.err; do [ -s $f ] && (echo $f; tail -5 $f); done

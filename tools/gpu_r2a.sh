#!/bin/bash
# Round 2, first GPU visit: full -m gpu suite (bit-exact parity + the new statistical parity of the product build), A/B of the product-build
# arithmetic (SFU transcendentals + approximate division vs round 1's libm / IEEE division), of the two default-off experiments of round 1,
# the default bench line (C3), and ncu evidence for C3's four largest kernels.
tag=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q -s ) > gpurun_out/${tag}_gpu_tests.log 2>&1
tail -5 gpurun_out/${tag}_gpu_tests.log
grep -h "statistical parity\|tier B" gpurun_out/${tag}_gpu_tests.log
run() { # name, workload, steps, env...
  local name=$1 wl=$2 steps=$3; shift 3
  env "$@" timeout 400 python bench.py --workload $wl --steps $steps --warmup 3 --lanes 1 --no-cpu-baseline > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
  python - <<P
import json
try:
    d = json.load(open("gpurun_out/${tag}_${name}.json"))
    print("${name}", round(d["value"], 3), "Msamples/s", {k: v for k, v in list(d["roofline"]["kernel_ms_per_iteration"].items())[:8]})
except Exception as e:
    print("${name} failed", e)
P
}
run c3_fast C3 4 X=1
run c3_precise C3 4 ETXB_LIB_FAST=$PWD/etx_tracer_b200/libetx_b200_precise.so
run c3_matmajor C3 4 ETXB_MERGE_MATERIAL_MAJOR=1
run c2_fast C2 8 X=1
run c2_precise C2 8 ETXB_LIB_FAST=$PWD/etx_tracer_b200/libetx_b200_precise.so
run c2_tiled C2 8 ETXB_MERGE_TILED=1
# the default line
timeout 600 python bench.py > gpurun_out/${tag}_bench_default.json 2> gpurun_out/${tag}_bench_default.err
cut -c1-400 gpurun_out/${tag}_bench_default.json
# ncu: launch list of one C3 iteration, then --set full of the head launches of its four largest kernels
M=gpu__time_duration.sum,launch__grid_size,smsp__thread_inst_executed_per_inst_executed.ratio
timeout 500 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/${tag}_c3_launches.csv python tools/profile_run.py C3 1 > gpurun_out/${tag}_ncu_c3.log 2>&1
for k in k_camera_merge_generic_batched k_camera_connect k_camera_shade k_light_bounce k_trace_closest; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o /tmp/p3_$k python tools/profile_run.py C3 1 > gpurun_out/${tag}_ncu_full_c3_$k.log 2>&1
  if [ -f /tmp/p3_$k.ncu-rep ]; then
    ncu -i /tmp/p3_$k.ncu-rep --page raw --csv > gpurun_out/${tag}_c3_$k.raw.csv 2>/dev/null
    python tools/ncu_summary.py gpurun_out/${tag}_c3_$k.raw.csv | tr '\n' ' ' | cut -c1-1500; echo
  fi
done
ls gpurun_out | grep ${tag} | head -50
exit 0

#!/bin/bash
# Round 2, sixth GPU visit (1 GPU): the 4-wide quantised BVH in the product build's traversal kernels — whole -m gpu suite, A/B on C3 / C4.
tag=${1:-r2f}
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q -s ) > gpurun_out/${tag}_gpu_tests.log 2>&1
tail -6 gpurun_out/${tag}_gpu_tests.log
grep -h "statistical parity\|tier B" gpurun_out/${tag}_gpu_tests.log | cut -c1-330
run() { # name, workload, steps, env...
  local name=$1 wl=$2 steps=$3; shift 3
  env "$@" timeout 400 python bench.py --workload $wl --steps $steps --warmup 3 --lanes 1 --no-cpu-baseline > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
  python - <<P
import json
try:
    d = json.load(open("gpurun_out/${tag}_${name}.json"))
    print("${name}", round(d["value"], 3), "Msamples/s", {k: v for k, v in list(d["roofline"]["kernel_ms_per_iteration"].items())[:10]}, "dominant", d["roofline"]["kernel"], round(d["roofline"]["frac"], 3))
except Exception as e:
    print("${name} failed", e)
P
}
run c3 C3 4 X=1
run c3_bvh2 C3 4 ETXB_WIDE_BVH=0
run c4 C4 6 X=1
run c4_bvh2 C4 6 ETXB_WIDE_BVH=0
timeout 600 python bench.py > gpurun_out/${tag}_bench_default.json 2> gpurun_out/${tag}_bench_default.err
python - <<P
import json
d = json.load(open("gpurun_out/${tag}_bench_default.json"))
print("default:", d["value"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"], d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"]["bvh"])
P
for k in k_trace_closest_wide k_shadow_resolve k_camera_shade k_camera_connect; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o /tmp/p3_$k python tools/profile_run.py C3 1 > gpurun_out/${tag}_ncu_full_c3_$k.log 2>&1
  if [ -f /tmp/p3_$k.ncu-rep ]; then
    ncu -i /tmp/p3_$k.ncu-rep --page raw --csv > gpurun_out/${tag}_c3_$k.raw.csv 2>/dev/null
    python tools/ncu_summary.py gpurun_out/${tag}_c3_$k.raw.csv | tr '\n' ' ' | cut -c1-1800; echo
  fi
done
M=gpu__time_duration.sum,launch__grid_size,smsp__thread_inst_executed_per_inst_executed.ratio
timeout 500 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/${tag}_c3_launches.csv python tools/profile_run.py C3 1 > gpurun_out/${tag}_ncu_c3.log 2>&1
exit 0

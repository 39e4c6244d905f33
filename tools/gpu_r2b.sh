#!/bin/bash
# Round 2, second GPU visit: the whole -m gpu suite (no -x: every test reports), A/B of the closure gather / connections, default bench, ncu of
# the closure gather and the connection kernel.
tag=${1:-r2b}
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q -s ) > gpurun_out/${tag}_gpu_tests.log 2>&1
tail -15 gpurun_out/${tag}_gpu_tests.log
grep -h "statistical parity\|tier B" gpurun_out/${tag}_gpu_tests.log
run() { # name, workload, steps, env...
  local name=$1 wl=$2 steps=$3; shift 3
  env "$@" timeout 400 python bench.py --workload $wl --steps $steps --warmup 3 --lanes 1 --no-cpu-baseline > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
  python - <<P
import json
try:
    d = json.load(open("gpurun_out/${tag}_${name}.json"))
    print("${name}", round(d["value"], 3), "Msamples/s", {k: v for k, v in list(d["roofline"]["kernel_ms_per_iteration"].items())[:9]})
except Exception as e:
    print("${name} failed", e)
P
}
run c3_closure C3 4 X=1
run c3_cl3 C3 4 ETXB_LIB_FAST=$PWD/etx_tracer_b200/libetx_b200_cl3.so
run c3_noclosure C3 4 ETXB_MERGE_CLOSURE=0
run c3_nomatmajor C3 4 ETXB_MERGE_MATERIAL_MAJOR=0
run c2 C2 8 X=1
run c4 C4 6 X=1
timeout 600 python bench.py > gpurun_out/${tag}_bench_default.json 2> gpurun_out/${tag}_bench_default.err
cut -c1-300 gpurun_out/${tag}_bench_default.json
python - <<P
import json
d = json.load(open("gpurun_out/${tag}_bench_default.json"))
print("default:", d["value"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"], "bvh", d["roofline"].get("bvh"))
P
for k in k_camera_merge_closure k_camera_connect; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o /tmp/p3_$k python tools/profile_run.py C3 1 > gpurun_out/${tag}_ncu_full_c3_$k.log 2>&1
  if [ -f /tmp/p3_$k.ncu-rep ]; then
    ncu -i /tmp/p3_$k.ncu-rep --page raw --csv > gpurun_out/${tag}_c3_$k.raw.csv 2>/dev/null
    python tools/ncu_summary.py gpurun_out/${tag}_c3_$k.raw.csv | tr '\n' ' ' | cut -c1-2500; echo
  fi
done
exit 0

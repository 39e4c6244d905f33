#!/bin/bash
# Round 2, thirteenth GPU visit (1 GPU): A/B of the two ordering experiments (ETXB_SHADOW_SORT=1: shadow list by origin; ETXB_QUEUE_SORT_SPATIAL=1: path
# queues by (material, hit point)), VCM and path tracer, on C3.
tag=${1:-r2m}
mkdir -p gpurun_out
run() { # name, lanes, env...
  local name=$1 lanes=$2; shift 2
  env "$@" timeout 400 python bench.py --workload C3 --steps 8 --warmup 3 --lanes $lanes --no-cpu-baseline --no-path-tracer > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
  python - <<P
import json
try:
    d = json.load(open("gpurun_out/${tag}_${name}.json"))
    k = d["roofline"]["kernel_ms_per_iteration"]
    print("${name}", round(d["value"], 3), "Msamples/s; shadow_trace", k.get("shadow_trace"), "trace", k.get("trace_closest(camera)"), k.get("trace_closest(light)"), "sum", round(sum(k.values()), 1))
except Exception as e:
    print("${name} failed", e)
P
}
run c3_l1 1 X=1
run c3_l1_sort 1 ETXB_SHADOW_SORT=1
run c3_l1_spatial 1 ETXB_QUEUE_SORT_SPATIAL=1
run c3_l1_both 1 ETXB_QUEUE_SORT_SPATIAL=1 ETXB_SHADOW_SORT=1
run c3_l4 4 X=1
run c3_l4_sort 4 ETXB_SHADOW_SORT=1
run c3_l4_spatial 4 ETXB_QUEUE_SORT_SPATIAL=1
run c3_l4_both 4 ETXB_QUEUE_SORT_SPATIAL=1 ETXB_SHADOW_SORT=1
timeout 200 python tools/pt_throughput.py C3 8 0 > gpurun_out/${tag}_pt.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/${tag}_pt.json')); print('pt', round(d['value'],2), d['kernel_ms_per_iteration'], d['in_flight_4'])"
ETXB_SHADOW_SORT=1 timeout 200 python tools/pt_throughput.py C3 8 0 > gpurun_out/${tag}_pt_sort.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/${tag}_pt_sort.json')); print('pt sort', round(d['value'],2), d['kernel_ms_per_iteration'], d['in_flight_4'])"
# parity of the sorted orders: the statistical C3 test with both switches on
ETXB_QUEUE_SORT_SPATIAL=1 ETXB_SHADOW_SORT=1 timeout 600 python -m pytest tests/test_gpu_statistical.py -m gpu -q -s -k "C3" 2>&1 | tail -4
exit 0

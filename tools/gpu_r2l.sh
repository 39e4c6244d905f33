#!/bin/bash
# Round 2, twelfth GPU visit (1 GPU): the whole -m gpu suite on the final tree, smoke(), the default bench line (with its path_tracer key), the path tracer's
# launch list and one --set full capture of k_pt_shade.
tag=${1:-r2l}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -s ) > gpurun_out/${tag}_gpu_tests.log 2>&1
tail -5 gpurun_out/${tag}_gpu_tests.log
grep -h "FAILED\|product parity\|statistical parity\] C3" gpurun_out/${tag}_gpu_tests.log | cut -c1-300 | head -30
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${tag}_smoke.log 2>&1; tail -2 gpurun_out/${tag}_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_c3.json 2> gpurun_out/${tag}_bench_c3.err
python - <<P
import json
try:
    d = json.load(open("gpurun_out/${tag}_bench_c3.json"))
    print("bench", round(d["value"], 3), "Msamples/s e2e", round(d["e2e"]["value"], 3), "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"], 4), "frac", round(d["roofline"]["frac"], 3))
    print("path_tracer", d.get("path_tracer"))
except Exception as e:
    print("bench failed", e)
P
tail -3 gpurun_out/${tag}_bench_c3.err
for l in 6 8; do
  timeout 300 python bench.py --steps 20 --warmup 5 --lanes $l --no-cpu-baseline --no-path-tracer > gpurun_out/${tag}_bench_c3_lanes$l.json 2> /dev/null
  python -c "
import json
try:
    d = json.load(open('gpurun_out/${tag}_bench_c3_lanes$l.json')); print('lanes $l', round(d['value'], 3), 'e2e', round(d['e2e']['value'], 3))
except Exception as e: print('lanes $l failed', e)"
done
timeout 300 python tools/pt_throughput.py C3 8 0 > gpurun_out/${tag}_pt_c3.json 2> gpurun_out/${tag}_pt_c3.err; cut -c1-600 gpurun_out/${tag}_pt_c3.json
timeout 300 ncu --metrics gpu__time_duration.sum,launch__grid_size,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none --csv --log-file gpurun_out/${tag}_pt_c3_launches.csv python tools/profile_run_pt.py C3 1 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_pt_shade -s 1 -c 1 -o /tmp/ptshade python tools/profile_run_pt.py C3 1 > /dev/null 2>&1
ncu -i /tmp/ptshade.ncu-rep --page raw --csv > gpurun_out/${tag}_c3_k_pt_shade.raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${tag}_c3_k_pt_shade.raw.csv 2>&1 | tail -12
exit 0

#!/bin/bash
# One GPU-box visit: parity tests, the bench lines, launch list.  Usage: gpurun -- bash tools/gpu_round.sh [tag]
tag=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/${tag}_gpu_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_gpu_tests.log
timeout 400 python bench.py --steps 12 --warmup 3 > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench_c2.err
timeout 300 python bench.py --steps 12 --warmup 3 --lanes 1 --no-cpu-baseline > gpurun_out/${tag}_bench_c2_lane1.json 2> gpurun_out/${tag}_bench_c2_lane1.err
timeout 500 python bench.py --workload C3 --steps 6 --warmup 3 --cpu-budget 10 > gpurun_out/${tag}_bench_c3.json 2> gpurun_out/${tag}_bench_c3.err
timeout 300 python bench.py --workload C3 --steps 6 --warmup 3 --lanes 1 --no-cpu-baseline > gpurun_out/${tag}_bench_c3_lane1.json 2> gpurun_out/${tag}_bench_c3_lane1.err
timeout 400 python bench.py --workload C4 --steps 4 --warmup 3 --cpu-budget 8 > gpurun_out/${tag}_bench_c4.json 2> gpurun_out/${tag}_bench_c4.err
timeout 400 python bench.py --workload C5 --steps 6 --warmup 3 --cpu-budget 8 > gpurun_out/${tag}_bench_c5.json 2> gpurun_out/${tag}_bench_c5.err
timeout 300 ncu --metrics gpu__time_duration.sum,launch__grid_size,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none --csv --log-file gpurun_out/${tag}_c2_launches.csv python tools/profile_run.py C2 2 > gpurun_out/${tag}_ncu_c2.log 2>&1
# A/B of experimental builds (ETXB_LIB_* override the library the ctypes layer loads)
if [ -f etx_tracer_b200/exp_e13.so ]; then
  for lanes in 1 4; do
    timeout 200 python bench.py --steps 12 --warmup 3 --lanes $lanes --no-cpu-baseline > gpurun_out/${tag}_ab_base_l${lanes}.json 2>/dev/null
    ETXB_CONNECT_DEFERRED=0 ETXB_LIB_FAST=$PWD/etx_tracer_b200/exp_e13.so timeout 200 python bench.py --steps 12 --warmup 3 --lanes $lanes --no-cpu-baseline > gpurun_out/${tag}_ab_e1_l${lanes}.json 2>/dev/null
    ETXB_LIB_FAST=$PWD/etx_tracer_b200/exp_e13.so timeout 200 python bench.py --steps 12 --warmup 3 --lanes $lanes --no-cpu-baseline > gpurun_out/${tag}_ab_e13_l${lanes}.json 2>/dev/null
  done
  for v in mb2 mb3 plain; do
    if [ -f etx_tracer_b200/exp_e13_$v.so ]; then
      ETXB_LIB_FAST=$PWD/etx_tracer_b200/exp_e13_$v.so timeout 200 python bench.py --steps 12 --warmup 3 --lanes 1 --no-cpu-baseline > gpurun_out/${tag}_ab_e13_${v}_l1.json 2>/dev/null
      ETXB_LIB_FAST=$PWD/etx_tracer_b200/exp_e13_$v.so timeout 200 python bench.py --steps 12 --warmup 3 --lanes 4 --no-cpu-baseline > gpurun_out/${tag}_ab_e13_${v}_l4.json 2>/dev/null
      ETXB_LIB_FAST=$PWD/etx_tracer_b200/exp_e13_$v.so timeout 300 python bench.py --workload C3 --steps 6 --warmup 3 --lanes 1 --no-cpu-baseline > gpurun_out/${tag}_ab_e13_${v}_c3_l1.json 2>/dev/null
    fi
  done
  ETXB_LIB_FAST=$PWD/etx_tracer_b200/exp_e13.so timeout 300 python bench.py --workload C3 --steps 6 --warmup 3 --lanes 1 --no-cpu-baseline > gpurun_out/${tag}_ab_e13_c3_l1.json 2>/dev/null
  ( time ETXB_LIB_FAST=$PWD/etx_tracer_b200/exp_e13.so ETXB_LIB_PARITY=$PWD/etx_tracer_b200/exp_e13_parity.so timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/${tag}_gpu_tests_e13.log 2>&1
  tail -3 gpurun_out/${tag}_gpu_tests_e13.log
fi
tail -3 gpurun_out/${tag}_gpu_tests.log
cut -c1-300 gpurun_out/${tag}_bench_c2.json

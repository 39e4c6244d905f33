#!/bin/bash
# One GPU-box visit: parity tests, the bench lines, per-kernel breakdown.  Usage: gpurun -- bash tools/gpu_round.sh [tag]
tag=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_gpu_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_gpu_tests.log
timeout 400 python bench.py --steps 12 --warmup 3 > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench_c2.err
timeout 300 python bench.py --steps 12 --warmup 3 --lanes 1 --no-cpu-baseline > gpurun_out/${tag}_bench_c2_lane1.json 2> gpurun_out/${tag}_bench_c2_lane1.err
timeout 500 python bench.py --workload C3 --steps 6 --warmup 3 --cpu-budget 10 > gpurun_out/${tag}_bench_c3.json 2> gpurun_out/${tag}_bench_c3.err
timeout 300 python bench.py --workload C3 --steps 6 --warmup 3 --lanes 1 --no-cpu-baseline > gpurun_out/${tag}_bench_c3_lane1.json 2> gpurun_out/${tag}_bench_c3_lane1.err
tail -3 gpurun_out/${tag}_gpu_tests.log
cat gpurun_out/${tag}_bench_c2.json | cut -c1-600

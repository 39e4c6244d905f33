#!/bin/bash
# Round-end style visit: parity tests, smoke, the bench lines of every config, the reference arm.
tag=${1:-final}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/${tag}_gpu_tests.log 2>&1
tail -3 gpurun_out/${tag}_gpu_tests.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${tag}_smoke.log 2>&1
tail -2 gpurun_out/${tag}_smoke.log
timeout 400 python bench.py > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench_c2.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err
timeout 300 python bench.py --workload C1 --steps 16 --warmup 3 --cpu-budget 6 > gpurun_out/${tag}_bench_c1.json 2> gpurun_out/${tag}_bench_c1.err
timeout 500 python bench.py --workload C3 --steps 8 --warmup 3 --cpu-budget 10 > gpurun_out/${tag}_bench_c3.json 2> gpurun_out/${tag}_bench_c3.err
timeout 400 python bench.py --workload C4 --steps 8 --warmup 3 --cpu-budget 8 > gpurun_out/${tag}_bench_c4.json 2> gpurun_out/${tag}_bench_c4.err
timeout 400 python bench.py --workload C5 --steps 16 --warmup 3 --cpu-budget 8 > gpurun_out/${tag}_bench_c5.json 2> gpurun_out/${tag}_bench_c5.err
# where does NCCL's banner go?  (bench.py sets NCCL_DEBUG_FILE=/dev/stderr so that stdout carries one JSON line)
NCCL_DEBUG=VERSION NCCL_DEBUG_FILE=/dev/stderr timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29519 tools/nccl_banner_check.py > gpurun_out/${tag}_nccl_stdout.txt 2> gpurun_out/${tag}_nccl_stderr.txt
echo "nccl banner lines on stdout: $(grep -c 'NCCL version' gpurun_out/${tag}_nccl_stdout.txt), on stderr: $(grep -c 'NCCL version' gpurun_out/${tag}_nccl_stderr.txt)"
for f in gpurun_out/${tag}_bench_*.json; do echo $f; cut -c1-140 $f; done
exit 0

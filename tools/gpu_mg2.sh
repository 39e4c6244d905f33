#!/bin/bash
# Two-GPU visit (gpurun --gpus 2): the bench line as the driver launches it, the tile-sharded mode, and the film equivalence check.
tag=${1:-mg2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 400 $TR bench.py --gpus 2 --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_c2_n2.json 2> gpurun_out/${tag}_bench_c2_n2.err
timeout 400 $TR bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline --workload C3 > gpurun_out/${tag}_bench_c3_n2.json 2> gpurun_out/${tag}_bench_c3_n2.err
timeout 300 $TR bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline --parallelism tile > gpurun_out/${tag}_bench_c2_n2_tile.json 2> gpurun_out/${tag}_bench_c2_n2_tile.err
timeout 300 $TR tools/multigpu_check.py 256 > gpurun_out/${tag}_check.log 2>&1
for f in gpurun_out/${tag}_bench_*.json; do echo $f; cut -c1-140 $f; done
tail -6 gpurun_out/${tag}_check.log
exit 0

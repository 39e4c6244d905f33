#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/multigpu_profile.py [workload] : wall time of every phase of the tile-sharded iteration
(light pass / light-image all-reduce / photon exchange / grid build / camera pass), max over ranks, next to the single-GPU step."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from etx_tracer_b200 import scenes
from etx_tracer_b200.api import GPUVCM
from etx_tracer_b200.multigpu import ShardedVCM

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
workload = sys.argv[1] if len(sys.argv) > 1 else "C2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
sd = scenes.config(workload)
g = GPUVCM(sd, flavor="fast", device=local)
sh = ShardedVCM(g, dist, rank, world)
g.run(0)
for _ in range(3):
    sh.iterate()
g.wait(); torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(iters):
    sh.iterate()
g.wait(); torch.cuda.synchronize(); dist.barrier()
plain = (time.perf_counter() - t0) / iters
sh.phase_seconds = {}
for _ in range(iters):
    sh.iterate()
names = list(sh.phase_seconds)
t = torch.tensor([sh.phase_seconds[n] / iters for n in names] + [plain], dtype=torch.float64, device="cuda")
tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
tmin = t.clone(); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
if rank == 0:
    out = {"workload": workload, "world": world, "ms_per_iteration": plain * 1e3,
           "phases_ms_max_over_ranks": {n: float(tmax[k]) * 1e3 for k, n in enumerate(names)},
           "phases_ms_min_over_ranks": {n: float(tmin[k]) * 1e3 for k, n in enumerate(names)},
           "light_vertices_rank0": int(g.counters()["light_vertices"])}
    print("MULTIGPU_PROFILE", json.dumps(out), flush=True)
g.close()
dist.destroy_process_group()

#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/multigpu_check.py : the tile-sharded render (NCCL exchanges) against the single-GPU render."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from etx_tracer_b200 import scenes, structs as S
from etx_tracer_b200.api import GPUVCM
from etx_tracer_b200.multigpu import ShardedVCM, InterleavedVCM

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
res, iters = int(sys.argv[1]) if len(sys.argv) > 1 else 256, 3
ok = True
for name, sd in (("C2", scenes.cornell_box(res, res, samples=256, spectral=True, sphere=True)), ("sky", scenes.sky_room(res, res, spectral=True))):
    for merging in (True, False):
        g = GPUVCM(sd, flavor="fast", device=local)
        if not merging:
            g.options["options"] = S.VCM_CONNECT_ONLY
        sh = ShardedVCM(g, dist, rank, world)
        g.run(0)
        for _ in range(iters):
            sh.iterate()
        img = sh.reduce_film()
        light = g.film(S.FILM_LIGHT)
        if rank == 0:
            cam = g.film(S.FILM_CAMERA)
            ref = GPUVCM(sd, flavor="fast", device=local)
            if not merging:
                ref.options["options"] = S.VCM_CONNECT_ONLY
            ref.render(iters)
            def rel(a, b):
                a = a[..., :3].astype(np.float64); b = b[..., :3].astype(np.float64)
                return float(np.sqrt(((a - b) ** 2).sum()) / np.sqrt((b ** 2).sum()))
            rc, rl = rel(cam, ref.film(S.FILM_CAMERA)), rel(light, ref.film(S.FILM_LIGHT))
            exact = bool(np.array_equal(cam[..., :3], ref.film(S.FILM_CAMERA)[..., :3]))
            print(f"{name} merging={merging} world={world}: camera rel-L2 {rc:.3e} (bit-identical: {exact}), light rel-L2 {rl:.3e}", flush=True)
            # merging off: tiles are independent -> exact.  merging on: per-cell photon order differs from the single-GPU pool, so float sums
            # round differently (1e-4); with stochastic BSDFs the per-lane merge streams see photons in another order -> Monte-Carlo level
            stochastic = name != "C2"
            ok &= (rc < ((2e-2 if stochastic else 1e-4) if merging else 1e-6)) and (rl < 1e-4)
            ref.close()
        g.close()
        dist.barrier()
# iteration-interleaved mode (2 iterations in flight per GPU): the union of the ranks' iterations is a set of distinct indices, the
# reduced film is their mean -> compare with a single context rendering the same indices
from etx_tracer_b200.api import GPUVCMGroup
sd = scenes.cornell_box(res, res, samples=256, spectral=True, sphere=True)
grp = GPUVCMGroup(sd, lanes=2, flavor="fast", device=local)
iv = InterleavedVCM(grp, dist, rank, world)
iv.begin()
per_rank = 3
iv.enqueue(per_rank)
iv.wait()
combined = iv.reduce_film()
if rank == 0:
    ref = GPUVCM(sd, flavor="fast", device=local)
    ref.render(per_rank * world)  # indices 0 .. per_rank * world - 1 = the union of r + j * world
    a = combined.cpu().numpy().reshape(res, res, 4)[..., :3].astype(np.float64)
    b = ref.film(S.FILM_RESULT)[..., :3].astype(np.float64)
    r = float(np.sqrt(((a - b) ** 2).sum()) / np.sqrt((b ** 2).sum()))
    print(f"interleaved world={world}: {per_rank * world} iterations, result rel-L2 vs single GPU {r:.3e}", flush=True)
    ok &= r < 1e-5
    ref.close()
grp.close()
dist.barrier()
if rank == 0:
    print("MULTIGPU_CHECK", "OK" if ok else "FAILED", flush=True)
dist.destroy_process_group()

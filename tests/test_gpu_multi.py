"""Two GPUs, one process each: the module's own pixel-tile path (etxb_comm_init / etxb_group_comm_init — NCCL all-reduce of the light image,
all-gather of the photon records, reduce of the film inside the C ABI) against one GPU rendering the whole frame.  Run with -m gpu on a box
with at least two devices (gpurun --gpus 2); skipped on one."""
import os
import socket

import numpy as np
import pytest

from conftest import rel_l2
from etx_tracer_b200 import scenes, structs as S

pytestmark = pytest.mark.gpu


PT_REPLICAS = 1000


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir, lanes, flavor):
    import torch
    import torch.distributed as dist
    from etx_tracer_b200 import api
    from etx_tracer_b200.multigpu import distribute_comm_ids
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        sd = scenes.cornell_box(96, 80, samples=16, spectral=True, sphere=True)
        n = 5
        if lanes == PT_REPLICAS:  # the path tracer's iterations dealt to the ranks (adaptive sampling off), one count-weighted film reduce
            g = api.GPUVCMGroup(sd, lanes=2, flavor=flavor, device=rank)
            g.set_integrator(S.INTEGRATOR_PT)
            g.comm_init_replicas(world, rank, distribute_comm_ids(dist, rank, 1, lambda c: api.comm_unique_ids(c, flavor), device=torch.device("cuda", rank)))
            g.render(6)
            assert g.status()["completed_iterations"] == 3
            film = g.comm_reduce_film(S.FILM_CAMERA)
            if rank == 0:
                np.savez(os.path.join(out_dir, "sharded.npz"), camera=film)
            g.close()
            return
        if lanes == 0:
            g = api.GPUVCM(sd, flavor=flavor, device=rank)
            g.comm_init(world, rank, distribute_comm_ids(dist, rank, 1, lambda c: api.comm_unique_ids(c, flavor), device=torch.device("cuda", rank)))
            g.render(n)
        elif lanes < 0:  # whole-frame iterations dealt to the ranks, one count-weighted film reduce
            g = api.GPUVCMGroup(sd, lanes=-lanes + 1, flavor=flavor, device=rank)
            g.comm_init_replicas(world, rank, distribute_comm_ids(dist, rank, 1, lambda c: api.comm_unique_ids(c, flavor), device=torch.device("cuda", rank)), split_lane=True)
            g.render(n)
            # indices 0, 2 | 1, 3 whole; index 4 split by camera tile over both ranks (it counts where part 0 ran)
            assert g.status()["completed_iterations"] == (3 if rank == 0 else 2)
        else:
            g = api.GPUVCMGroup(sd, lanes=lanes, flavor=flavor, device=rank)
            g.comm_init(world, rank, distribute_comm_ids(dist, rank, lanes + 1, lambda c: api.comm_unique_ids(c, flavor), device=torch.device("cuda", rank)))
            g.render(n)
        out = {}
        for name, layer in (("result", S.FILM_RESULT), ("camera", S.FILM_CAMERA), ("light", S.FILM_LIGHT)):
            film = g.comm_reduce_film(layer)
            if rank == 0:
                out[name] = film
        if rank == 0:
            np.savez(os.path.join(out_dir, "sharded.npz"), **out)
        g.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("lanes,flavor", [(0, "parity"), (0, "fast"), (2, "fast"), (-2, "fast")])
def test_two_gpu_tiles_render_the_single_gpu_frame(tmp_path, lanes, flavor):
    import torch
    import torch.multiprocessing as mp
    from etx_tracer_b200 import api
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path), lanes, flavor), nprocs=2, join=True)
    got = np.load(tmp_path / "sharded.npz")
    sd = scenes.cornell_box(96, 80, samples=16, spectral=True, sphere=True)
    ref = api.GPUVCM(sd, flavor=flavor)
    ref.render(5)
    for name, layer in (("result", S.FILM_RESULT), ("camera", S.FILM_CAMERA), ("light", S.FILM_LIGHT)):
        a, b = got[name][..., :3], ref.film(layer)[..., :3]
        assert np.isfinite(a).all()
        # the gathered photon records arrive rank-major instead of path-major: the merge sums the same photons in another order
        assert rel_l2(a, b) < (2e-5 if flavor == "parity" else 2e-3), f"{name}: {rel_l2(a, b):.3e}"
    ref.close()


def test_two_gpu_path_tracer_replicas_render_the_single_gpu_frame(tmp_path):
    """The second device integrator over two GPUs: whole iterations dealt to the ranks (index j on rank j % 2), nothing exchanged inside an iteration,
    one count-weighted ncclReduce of the camera films — the mean over the six iterations, as one GPU renders it."""
    import torch
    import torch.multiprocessing as mp
    from etx_tracer_b200 import api
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path), PT_REPLICAS, "fast"), nprocs=2, join=True)
    got = np.load(tmp_path / "sharded.npz")["camera"][..., :3]
    sd = scenes.cornell_box(96, 80, samples=16, spectral=True, sphere=True)
    ref = api.GPUPathTracing(sd, flavor="fast")
    ref.set_scene_settings(0.0, 0.0)
    ref.render(6)
    want = ref.film(S.FILM_CAMERA)[..., :3]
    assert np.isfinite(got).all() and rel_l2(got, want) < 1e-5, rel_l2(got, want)
    ref.close()

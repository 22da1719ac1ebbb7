"""CPU-only, world size 2 over gloo: the host-side logic of the multi-GPU path — which iterations a rank renders
(bench.py's `1 + k*S*world + rank` with stride `world`) and the radiance all-reduce — checked with the oracle's
renderer standing in for the GPU (no GPU compute in the -m "not gpu" suite)."""
import ctypes
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def shard_iterations(step, spp_per_step, world, rank):
    """the iteration indices bench.py gives kdpt_render(begin, count, stride) for one step"""
    begin = 1 + step * spp_per_step * world + rank
    return [begin + q * world for q in range(spp_per_step)]


def test_shards_partition_the_sample_indices():
    for world in (1, 2, 4, 8):
        for S in (1, 3, 16):
            seen = []
            for step in range(3):
                for r in range(world):
                    seen += shard_iterations(step, S, world, r)
            assert sorted(seen) == list(range(1, 3 * S * world + 1))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import kdtreepathtraceroptimization_b200 as kdpt
    from conftest import P, ko_params, ko_scene
    O = ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
    O.ko_render_iteration.restype = ctypes.c_long
    sc = kdpt.Scene.from_npz(os.path.join(ROOT, "tests", "golden", "cornell_sphere_low_1.npz")).with_resolution(48, 48)
    cam = sc.camera.copy(); O.ko_camera_runcuda(P(cam)); sc.camera = cam
    p = kdpt.default_params(mode=kdpt.FULLSTACK, sortByMaterial=0)
    ks, kp = ko_scene(sc), ko_params(p, sc.trace_depth)
    img = np.zeros((48 * 48, 3), 'f4')
    for it in shard_iterations(0, 3, world, rank):
        O.ko_render_iteration(ctypes.byref(ks), ctypes.byref(kp), it, P(img), None, 1)
    t = torch.from_numpy(img)
    dist.all_reduce(t)                       # the one collective of the path: sum of the radiance buffers
    if rank == 0:
        ref = np.zeros((48 * 48, 3), 'f4')
        for it in range(1, 3 * world + 1):
            O.ko_render_iteration(ctypes.byref(ks), ctypes.byref(kp), it, P(ref), None, 1)
        np.save(out, np.stack([t.numpy(), ref]))
    dist.barrier()
    dist.destroy_process_group()


def test_sample_sharded_render_plus_allreduce_equals_single_rank(oracle, tmp_path):
    out = str(tmp_path / "imgs.npy")
    mp.spawn(_worker, args=(2, 29500 + os.getpid() % 2000, out), nprocs=2, join=True)
    got, ref = np.load(out)
    assert ref.sum() > 0
    # same samples, different float summation order only
    assert np.allclose(got, ref, rtol=1e-5, atol=1e-6)

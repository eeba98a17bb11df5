"""Host logic of the hash-range-sharded map on CPU, world_size 2 over gloo: owner function,
stable partition contract, the all-to-all exchange and the ordering guarantee the bit-exactness of the sharded
map rests on. (The routing kernels themselves run in tests/test_gpu_parity.py.)"""
import ctypes as C
import multiprocessing as mp
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from vina_slam_b200 import capi, sharded  # noqa: E402


def voxel_keys(pw: np.ndarray, voxel_size: float) -> np.ndarray:
    """voxel_map.cpp:246-253: double divide -> float -> '-1 if negative' in float -> truncate."""
    f = (pw / voxel_size).astype(np.float32)
    f = np.where(f < 0, f - np.float32(1.0), f)
    return f.astype(np.int64)


def owners(keys: np.ndarray, world: int) -> np.ndarray:
    lib = capi.load()
    return np.array([lib.vina_shard_owner(C.c_int64(int(k[0])), C.c_int64(int(k[1])), C.c_int64(int(k[2])), world)
                     for k in keys], dtype=np.int64)


def make_records(n: int, seed: int):
    rng = np.random.default_rng(seed)
    pw = rng.uniform(-40, 40, size=(n, 3))
    rec = np.zeros((n, sharded.REC))
    rec[:, 0:3] = rng.normal(size=(n, 3))
    rec[:, 3:9] = rng.uniform(size=(n, 6))
    rec[:, 9:12] = pw
    rec[:, 12] = np.arange(n, dtype=np.int64).view(np.float64)  # scan index, bit-cast like the kernel does
    return rec, voxel_keys(pw, 1.0)


def route_numpy(rec, keys, first, count, world):
    """What vina_shard_route produces for the slice [first, first+count): stable partition by owner."""
    ow = owners(keys[first:first + count], world)
    order = np.argsort(ow, kind="stable")
    return rec[first:first + count][order], np.bincount(ow, minlength=world)


def test_owner_is_a_partition_and_balanced():
    _, keys = make_records(20000, 1)
    for world in (1, 2, 4, 8):
        ow = owners(keys, world)
        assert ow.min() >= 0 and ow.max() < world
        frac = np.bincount(ow, minlength=world) / len(ow)
        assert np.all(np.abs(frac - 1.0 / world) < 0.05), frac
    # same voxel -> same owner, also for negative coordinates and exactly-integer negatives (Appendix A.2)
    # (-1 + 1e-9 rounds to -1.0f before the "-1 if negative": the float quirk moves it one cell down as well)
    k = voxel_keys(np.array([[-1.0, -0.5, 0.5], [-1.0 + 1e-3, -0.25, 0.75], [-1.0 + 1e-9, 0.0, -0.0]]), 1.0)
    assert k[0].tolist() == [-2, -1, 0] and k[1].tolist() == [-1, -1, 0] and k[2].tolist() == [-2, 0, 0]
    assert capi.load().vina_shard_owner(C.c_int64(1 << 21), C.c_int64(0), C.c_int64(0), 2) == -1  # out of key range


def test_slices_cover_the_scan_in_order():
    for n in (0, 1, 7, 60001):
        for world in (1, 2, 3, 8):
            parts = [sharded.slice_of(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            for (f0, c0), (f1, _) in zip(parts, parts[1:]):
                assert f1 == f0 + c0


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 5003
    rec, keys = make_records(n, seed=7)  # every rank sees the same scan ...
    first, count = sharded.slice_of(n, rank, world)  # ... and routes its ascending slice
    send, counts = route_numpy(rec, keys, first, count, world)
    recv = sharded.exchange_records(torch.from_numpy(np.ascontiguousarray(send)), counts).numpy()
    gidx = recv[:, 12].copy().view(np.int64)
    ow = owners(voxel_keys(recv[:, 9:12], 1.0), world)
    tot = torch.tensor([recv.shape[0]], dtype=torch.int64)
    dist.all_reduce(tot)
    # an empty slice must not wedge the collective
    e = sharded.exchange_records(torch.zeros((0, sharded.REC), dtype=torch.float64), [0] * world)
    q.put((rank, bool(np.all(ow == rank)), bool(np.all(np.diff(gidx) > 0)), int(tot[0]), gidx.tolist(),
           recv.tobytes(), int(e.shape[0])))
    dist.destroy_process_group()


def test_exchange_world_size_2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n = 5003
    seen = []
    for rank, all_mine, ascending, tot, gidx, raw, n_empty in res:
        assert all_mine          # every record landed on the owner of its voxel
        assert ascending         # ... in ascending scan order (source ranks hold ascending slices, routing is stable)
        assert tot == n and n_empty == 0
        seen += gidx
    assert sorted(seen) == list(range(n))  # nothing lost, nothing duplicated
    # the in-process permutation used by the single-GPU emulation test is the same permutation
    rec, keys = make_records(n, seed=7)
    routed = [route_numpy(rec, keys, *sharded.slice_of(n, r, world), world) for r in range(world)]
    local = sharded.local_exchange([s for s, _ in routed], [c for _, c in routed])
    for (rank, *_rest), loc in zip(res, local):
        assert loc.tobytes() == _rest[4]

"""Synthetic generator contract and the multi-GPU (replica) host logic on world_size-2 gloo."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vina_slam_b200 import replicas, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_scan_contract():
    for name in ("mid360", "velodyne32", "robosense128", "hilti_xt32"):
        cfg = synth.small_sensor(name, min(synth.SENSORS[name].n_beams, 8), 300)
        a, b = synth.Sequence(cfg), synth.Sequence(cfg)
        s1, s2 = a.next_scan(), b.next_scan()
        assert np.array_equal(s1.xyzt, s2.xyzt) and np.array_equal(s1.imu, s2.imu)  # seeded
        x = s1.xyzt
        assert x.dtype == np.float32 and x.shape[1] == 4
        assert np.all(np.diff(x[:, 3]) >= 0) and x[-1, 3] < 0.11  # sorted by time, <= 0.11 s (lidar_decoder.cpp:30-35)
        assert np.all(np.einsum("ij,ij->i", x[:, :3], x[:, :3]) > cfg.blind ** 2)
        n_imu = s1.imu.shape[0]
        assert n_imu > 4 and np.all(np.diff(s1.imu[:, 0]) > 0) and s1.imu[-1, 0] <= s1.end_time + 1e-9
        s3 = a.next_scan()
        assert s3.imu[0, 0] > s1.imu[-1, 0]  # IMU samples are handed out once (sync.cpp:63-72)
        assert abs(np.linalg.norm(s1.imu[:, 4:7], axis=1).mean() - 9.8) < 0.5  # specific force ~ g
    assert synth.SENSORS["robosense128"].n_points == 240000 and synth.SENSORS["velodyne32"].n_points == 57600


def test_full_size_shapes():
    assert synth.SENSORS["mid360"].n_points == 20000 and synth.SENSORS["hilti_xt32"].n_points == 64000
    assert synth.SENSORS["hilti_xt32"].imu_rate == 400.0
    assert synth.SENSORS["mid360"].voxel_size == 0.5 and synth.SENSORS["velodyne32"].max_layer == 3


def _clearance(world, pts):
    """Distance from each position to the nearest rectangle of the world."""
    other = {0: (1, 2), 1: (0, 2), 2: (0, 1)}
    best = np.full(pts.shape[0], np.inf)
    for axis, off, lo0, hi0, lo1, hi1 in world.rects:
        a = int(axis)
        b0, b1 = other[a]
        d0 = np.maximum(np.maximum(lo0 - pts[:, b0], pts[:, b0] - hi0), 0.0)
        d1 = np.maximum(np.maximum(lo1 - pts[:, b1], pts[:, b1] - hi1), 0.0)
        best = np.minimum(best, np.sqrt((pts[:, a] - off) ** 2 + d0 * d0 + d1 * d1))
    return best


def test_every_rank_seed_yields_full_scans():
    """bench.py --gpus N gives rank r the sequence of seed base + r (replicas.sequence_seed). Every one of them
    has to deliver scans of the nominal size: (i) over ten minutes of every path the sensor keeps more than the
    blind zone between itself and the nearest vertical surface of the world (the reason a scan loses returns),
    (ii) bootstrap + 25 scans of a reduced-density copy of every sensor (same field of view, same blind zone,
    same path: the kept FRACTION does not depend on the beam density) keep >= 98 % of their points, and
    (iii) a full-size scan per sensor does."""
    t = np.arange(0.0, 600.0, 0.05)
    for name, base in synth.SENSORS.items():
        for r in range(8):
            seq = synth.Sequence(synth.small_sensor(name, min(base.n_beams, 8), 150), seed=replicas.sequence_seed(base.seed, r))
            walls = synth.World()
            walls.rects = walls.rects[walls.rects[:, 0] != 2]  # floor / ceiling are met by the elevation limits
            assert _clearance(walls, seq.traj.pos(seq.t0 + t)).min() > base.blind + 0.3, (name, r)
            for k in range(base.win_size + 25):
                sc = seq.next_scan(deskewed=k < base.win_size)
                assert sc.xyzt.shape[0] >= 0.98 * seq.cfg.n_points, (name, r, k, sc.xyzt.shape[0])
        sc = synth.Sequence(base, seed=replicas.sequence_seed(base.seed, 2)).next_scan()
        assert sc.xyzt.shape[0] >= 0.98 * base.n_points


def test_empty_scan_is_a_clear_error():
    cfg = synth.small_sensor("robosense128", 4, 50)
    seq = synth.Sequence(dataclasses_replace(cfg, blind=500.0))
    with pytest.raises(RuntimeError, match="no returns outside"):
        seq.next_scan()


def dataclasses_replace(cfg, **kw):
    import dataclasses

    return dataclasses.replace(cfg, **kw)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from oracle import oracle_py as op

    cfg = synth.small_sensor("robosense128", 8, 200)
    seed = replicas.sequence_seed(cfg.seed, rank)
    seq = synth.Sequence(cfg, seed=seed)
    od = op.Odom(cfg)
    sc = None
    for _ in range(3):
        sc = seq.next_scan(deskewed=True)
        od.bootstrap(sc.xyzt, op.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
    pts = float(sc.xyzt.shape[0])
    t_local = 0.1 * (rank + 1)
    tot, (tmax,) = replicas.reduce_throughput(pts, [t_local])
    digest = float(np.abs(sc.xyzt).sum())
    lst = [None] * world
    dist.all_gather_object(lst, (seed, digest, pts))
    q.put((rank, tot, tmax, lst))
    dist.destroy_process_group()


def test_replicas_world_size_2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, tot, tmax, lst in res:
        assert tmax == pytest.approx(0.2)  # max over ranks
        assert tot == pytest.approx(sum(x[2] for x in lst))  # points summed over ranks
        seeds = [x[0] for x in lst]
        assert seeds[1] == seeds[0] + 1  # rank r replays sequence base + r
        assert lst[0][1] != lst[1][1]  # ... which are different sequences


def test_single_process_reduce_is_identity():
    tot, ts = replicas.reduce_throughput(10.0, [1.0, 2.0])
    assert tot == 10.0 and ts == [1.0, 2.0]


def test_replay_tum_format_and_npz_reader(tmp_path):
    """The file-replay front end's host logic (no GPU): TUM lines like FileReaderWriter::save_pose_tum
    (io.cpp:67-77) and the recording reader."""
    from vina_slam_b200 import replay

    rng = np.random.default_rng(0)
    for _ in range(20):
        R = synth.rot_exp(rng.normal(0, 2.0, 3))
        q = replay.quat_xyzw(R)
        x, y, z, w = q
        Rq = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                       [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                       [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        assert abs(np.linalg.norm(q) - 1) < 1e-12 and w >= 0 and np.max(np.abs(Rq - R)) < 1e-12
    line = replay.tum_line(12.5, np.array([1.0, -2.0, 0.25]), np.eye(3))
    assert line == "12.500000000 1.000000000 -2.000000000 0.250000000 0.000000000 0.000000000 0.000000000 1.000000000\n"
    cfg = synth.small_sensor("robosense128", 4, 50)
    seq = synth.Sequence(cfg)
    boots = [seq.next_scan(deskewed=True) for _ in range(cfg.win_size)]
    scans = [seq.next_scan() for _ in range(2)]
    d = {}
    for k, f in enumerate(boots + scans):
        d[f"xyzt_{k}"], d[f"beg_{k}"], d[f"imu_{k}"] = f.xyzt, f.beg_time, f.imu
    d["boot_R"] = np.array([f.gt_R for f in boots])
    d["boot_p"] = np.array([f.gt_p for f in boots])
    d["boot_v"] = np.array([f.gt_v for f in boots])
    path = str(tmp_path / "rec.npz")
    np.savez(path, **d)
    b2, s2 = replay.npz_frames(path, cfg.win_size)
    assert len(b2) == cfg.win_size and len(s2) == 2
    assert np.array_equal(s2[0].xyzt, scans[0].xyzt) and s2[1].beg_time == scans[1].beg_time
    assert np.array_equal(b2[3].gt_R, boots[3].gt_R) and abs(b2[-1].end_time - boots[-1].end_time) < 1e-6

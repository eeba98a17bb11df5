"""File replay front end (the part around the hot path that needs no ROS): feed a
recorded or synthetic sequence through `vina_odom_step` and write the trajectory in the reference's TUM format
(`FileReaderWriter::save_pose_tum`, src/platform/ros2/io.cpp:67-77, written after every scan at
local_mapping.cpp:429-430: `t x y z qx qy qz qw`, 9 decimals).

Two ways to start: `--cold-start` runs the reference's own start-up phase (`initialization()`, node.cpp:293-366:
IMU initialisation, kd-tree IEKF, motion_init with the gravity BA) on the first scans - nothing but scans and IMU
samples is needed; without it the window is bootstrapped from `win_size` scans at known poses, which a recording
then has to provide (`boot_R`, `boot_p`, `boot_v`) - the synthetic generator does.

    python -m vina_slam_b200.replay --workload mid360 --scans 50 --ba --out traj.txt
    python -m vina_slam_b200.replay --npz recording.npz --config robosense128 --out traj.txt

npz layout: `xyzt_<k>` (n, 4) float32 = x, y, z, time offset within the scan (sorted), `beg_<k>` scalar, `imu_<k>`
(m, 7) = t, gyro, accel for k = 0 .. K-1; the first `win_size` scans are the bootstrap (already deskewed) with
`boot_R` (w, 3, 3), `boot_p` (w, 3), `boot_v` (w, 3); `anchor_imu` (7,) = the last IMU sample before scan `win_size`.
"""
from __future__ import annotations

import argparse
import sys
import time

import numpy as np

from . import capi, synth


def quat_xyzw(R: np.ndarray) -> np.ndarray:
    """Rotation matrix -> unit quaternion (x, y, z, w), w >= 0 (Eigen::Quaterniond(R) up to the sign convention)."""
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
        q = np.zeros(4)
        q[i] = 0.25 * s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
        q[3] = (R[k, j] - R[j, k]) / s
    if q[3] < 0:
        q = -q
    return q / np.linalg.norm(q)


def tum_line(t: float, p: np.ndarray, R: np.ndarray) -> str:
    q = quat_xyzw(R)
    return "%.9f %.9f %.9f %.9f %.9f %.9f %.9f %.9f\n" % (t, p[0], p[1], p[2], q[0], q[1], q[2], q[3])


def synthetic_frames(cfg, n_scans: int, seed=None):
    seq = synth.Sequence(cfg, seed=cfg.seed if seed is None else seed)
    boots = [seq.next_scan(deskewed=True) for _ in range(cfg.win_size)]
    scans = [seq.next_scan() for _ in range(n_scans)]
    return boots, scans


def _undistorted_first(cfg, n_scans: int):
    """A synthetic sequence for a cold start: every scan raw (distorted by the motion), none at a known state."""
    seq = synth.Sequence(cfg)
    return [seq.next_scan() for _ in range(n_scans)]


def npz_frames(path: str, win_size: int):
    d = np.load(path)

    class F:
        pass

    k, frames = 0, []
    while f"xyzt_{k}" in d:
        f = F()
        f.xyzt = np.ascontiguousarray(d[f"xyzt_{k}"], dtype=np.float32)
        f.beg_time = float(d[f"beg_{k}"])
        f.end_time = f.beg_time + float(f.xyzt[-1, 3])
        f.imu = np.ascontiguousarray(d[f"imu_{k}"], dtype=np.float64)
        f.gt_R = f.gt_p = f.gt_v = None
        frames.append(f)
        k += 1
    if k <= win_size:
        raise SystemExit(f"{path}: {k} scans, need more than win_size = {win_size}")
    for i in range(win_size):
        frames[i].gt_R, frames[i].gt_p, frames[i].gt_v = d["boot_R"][i], d["boot_p"][i], d["boot_v"][i]
    boots, scans = frames[:win_size], frames[win_size:]
    if "anchor_imu" in d:
        boots[-1].imu = np.asarray(d["anchor_imu"], dtype=np.float64).reshape(1, 7)
    return boots, scans


def stream_packages(scans, point_notime: int = 0, t_last_of=None):
    """The scans and their IMU samples as the two message streams a live system sees - every IMU sample arrives at
    its stamp, a scan when it is complete - paired by `vina_sync` (sync_packages, src/sensor/sync.cpp:18-96; host
    only). Yields (index into scans, pcl_beg_time, pcl_end_time, imu (m, 7)) for every package; the last scan of the
    list stays pending (its package closes with the first IMU sample after its end, sync.cpp:60-63).
    t_last_of(k) = back().curvature of scan k once prepared (default: of the scan as given)."""
    sync = capi.Sync(point_notime)
    msgs = []
    for k, f in enumerate(scans):
        for row in f.imu:
            msgs.append((float(row[0]), 0, k, row))
        msgs.append((f.end_time, 1, k, None))
    msgs.sort(key=lambda m: (m[0], m[1]))
    try:
        for _, kind, k, row in msgs:
            if kind == 0:
                sync.push_imu(row)
            else:
                tl = float(scans[k].xyzt[-1, 3]) if t_last_of is None else float(t_last_of(k))
                sync.push_scan(scans[k].beg_time, tl, k)
            while True:
                r, tag, beg, end, imu = sync.next()
                if r < 0:
                    raise capi.VinaError(r, "vina_sync_next")
                if r == 0:
                    break
                if r == 1:
                    yield tag, beg, end, imu
    finally:
        sync.close()


def replay_stream(cfg, boots, scans, out=None, max_iter=4, caps=None, point_filter_num=1, shuffle_seed=0,
                  prune_horizon=700):
    """Like `replay`, but from RAW message streams: every scan is handed over the way a driver delivers it (points
    in arbitrary time order - here shuffled), goes through the device front end (`vina_scan_prepare`: decoder keep
    rule, time sort, 0.11 s cut) and is paired with its IMU samples by `vina_sync`; the step is
    `vina_odom_step_prepared`. Returns (trajectory rows, seconds per scan, worst position error or None)."""
    caps = caps or dict(max_scan_points=max(300000, max(f.xyzt.shape[0] for f in boots + scans) + 1024))
    gx = capi.Ctx(cfg, **caps)
    for f in boots:
        gx.bootstrap(f.xyzt, capi.make_state(f.gt_R, f.gt_p, f.gt_v, t=f.end_time))
    gx.set_imu_anchor(boots[-1].end_time, boots[-1].imu[-1])
    rng = np.random.default_rng(shuffle_seed)
    raws = [f.xyzt[rng.permutation(f.xyzt.shape[0])] for f in scans]
    blind2 = float(cfg.blind) ** 2  # node.cpp:210
    prepared = {"k": -1, "t_last": 0.0}

    def prepare(k):
        if prepared["k"] != k:
            _, prepared["t_last"] = gx.scan_prepare(raws[k], point_filter_num, blind2)
            prepared["k"] = k
        return prepared["t_last"]

    rows, worst = [], None
    fh = open(out, "w") if out else None
    t0 = time.perf_counter()
    for k, beg, end, imu in stream_packages(scans, 0, t_last_of=prepare):
        prepare(k)
        prepared["k"] = -1  # the step consumes the prepared scan
        s = capi.state_arrays(gx.step_prepared(beg, imu, True, max_iter))
        rows.append(np.concatenate([[s["t"]], s["p"], quat_xyzw(s["R"])]))
        if fh:
            fh.write(tum_line(s["t"], s["p"], s["R"]))
        if scans[k].gt_p is not None:
            e = float(np.linalg.norm(s["p"] - scans[k].gt_p))
            worst = e if worst is None else max(worst, e)
        if prune_horizon > 0:
            gx.idle(prune_horizon)
    gx.sync()
    dt = (time.perf_counter() - t0) / max(len(rows), 1)
    if fh:
        fh.close()
    gx.close()
    return np.array(rows), dt, worst


def replay(cfg, boots, scans, out=None, ba=False, max_iter=4, caps=None, prune_horizon=700, stats=None):
    """Returns (trajectory rows (K, 8): t, p, q_xyzw; seconds per scan; worst position error vs ground truth or None).

    After every scan the idle path of the reference's loop runs (local_mapping.cpp:303-341: a live system is idle
    between two scans): when the vehicle has moved on, root voxels last marginalised `prune_horizon` metres of
    travel ago are erased (700 in the reference; 0 = never). `stats`, if a dict, receives the totals."""
    caps = caps or dict(max_scan_points=max(300000, max(f.xyzt.shape[0] for f in boots + scans) + 1024))
    gx = capi.Ctx(cfg, **caps)  # raises without a CUDA device: there is no CPU path
    if ba:
        gx.set_ba(True)
    for f in boots:
        gx.bootstrap(f.xyzt, capi.make_state(f.gt_R, f.gt_p, f.gt_v, t=f.end_time))
    gx.set_imu_anchor(boots[-1].end_time, boots[-1].imu[-1])
    rows, worst = [], None
    erased = freed = 0
    fh = open(out, "w") if out else None
    t0 = time.perf_counter()
    for f in scans:
        s = capi.state_arrays(gx.step(f.xyzt, f.beg_time, f.imu, True, max_iter))
        rows.append(np.concatenate([[s["t"]], s["p"], quat_xyzw(s["R"])]))
        if fh:
            fh.write(tum_line(s["t"], s["p"], s["R"]))
        if f.gt_p is not None:
            e = float(np.linalg.norm(s["p"] - f.gt_p))
            worst = e if worst is None else max(worst, e)
        if prune_horizon > 0:
            a, b = gx.idle(prune_horizon)
            erased, freed = erased + a, freed + b
    gx.sync()
    if stats is not None:
        stats.update(roots_erased=erased, nodes_freed=freed, journey=gx.journey()[0], map_count=gx.map_count())
    dt = (time.perf_counter() - t0) / max(len(scans), 1)
    if fh:
        fh.close()
    gx.close()
    return np.array(rows), dt, worst


def replay_cold_start(cfg, scans, out=None, ba=False, max_iter=4, caps=None, prune_horizon=700, max_init_scans=200):
    """No known states at all: the reference's own start-up phase (VINA_SLAM::initialization, node.cpp:293-366) through
    vina_odom_cold_start / vina_odom_init_scan - IMU initialisation, the kd-tree IEKF over win_size scans,
    Initialization::motion_init (gravity BA, gravity alignment) - then the per-scan loop. The trajectory lives in the
    gravity-aligned frame of the first window frame, like the reference's. Returns (rows, seconds per scan, scans the
    start-up consumed)."""
    caps = caps or dict(max_scan_points=max(300000, max(f.xyzt.shape[0] for f in scans) + 1024))
    gx = capi.Ctx(cfg, **caps)
    if ba:
        gx.set_ba(True)
    gx.cold_start()
    rows, used, started = [], 0, False
    fh = open(out, "w") if out else None
    t0 = time.perf_counter()
    for f in scans:
        if not started:
            st, s = gx.init_scan(f.xyzt, f.beg_time, f.imu)
            used += 1
            started = st == 1
            if used > max_init_scans and not started:
                raise RuntimeError(f"the start-up phase did not converge within {max_init_scans} scans")
            if not started:
                continue
            s = capi.state_arrays(s)
        else:
            s = capi.state_arrays(gx.step(f.xyzt, f.beg_time, f.imu, True, max_iter))
            if prune_horizon > 0:
                gx.idle(prune_horizon)
        rows.append(np.concatenate([[s["t"]], s["p"], quat_xyzw(s["R"])]))
        if fh:
            fh.write(tum_line(s["t"], s["p"], s["R"]))
    gx.sync()
    dt = (time.perf_counter() - t0) / max(len(scans), 1)
    if fh:
        fh.close()
    gx.close()
    return np.array(rows), dt, used


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--workload", default="robosense128", choices=sorted(synth.SENSORS), help="synthetic sensor shape")
    ap.add_argument("--config", default=None, choices=sorted(synth.SENSORS), help="yaml-equivalent parameter set for --npz")
    ap.add_argument("--npz", default=None, help="recorded sequence (see the module docstring)")
    ap.add_argument("--scans", type=int, default=30)
    ap.add_argument("--ba", action="store_true", help="LocalBA.if_BA: 1")
    ap.add_argument("--out", default=None, help="TUM trajectory file")
    ap.add_argument("--raw", action="store_true",
                    help="feed raw message streams: shuffled scans through vina_scan_prepare, pairing by vina_sync")
    ap.add_argument("--cold-start", action="store_true",
                    help="no bootstrap states: start with the reference's own initialisation (IMU init, kd-tree IEKF, "
                         "motion_init) on the sequence's scans")
    ap.add_argument("--prune-horizon", type=int, default=700,
                    help="metres of travel after which unvisited root voxels are erased (reference: 700; 0 = never)")
    a = ap.parse_args(argv)
    cfg = synth.SENSORS[a.config or a.workload]
    boots, scans = npz_frames(a.npz, cfg.win_size) if a.npz else synthetic_frames(cfg, a.scans)
    st = {}
    if a.cold_start:
        rows, dt, used = replay_cold_start(cfg, boots + scans if a.npz else
                                           [f for f in _undistorted_first(cfg, a.scans)], out=a.out, ba=a.ba,
                                           prune_horizon=a.prune_horizon)
        print(f"{len(rows)} poses after a start-up phase of {used} scans, {1e3 * dt:.3f} ms/scan (wall clock)"
              + (f", trajectory -> {a.out}" if a.out else ""))
        return 0
    if a.raw:
        rows, dt, worst = replay_stream(cfg, boots, scans, out=a.out, prune_horizon=a.prune_horizon)
    else:
        rows, dt, worst = replay(cfg, boots, scans, out=a.out, ba=a.ba, prune_horizon=a.prune_horizon, stats=st)
    msg = f"{len(rows)} scans, {1e3 * dt:.3f} ms/scan (wall clock, Python loop included)"
    if st:
        msg += f", journey {st['journey']:.1f} m, {st['roots_erased']} root voxels pruned"
    if worst is not None:
        msg += f", max |p - p_gt| = {worst:.4f} m"
    if a.out:
        msg += f", trajectory -> {a.out}"
    print(msg)
    return 0


if __name__ == "__main__":
    sys.exit(main())

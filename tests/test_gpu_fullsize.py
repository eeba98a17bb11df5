"""GPU parity against the oracle at the FULL sizes of BASELINE.json's configs (20 k / 57.6 k / 240 k / 64 k points per
scan) and with a large resident map - the same bars as tests/test_gpu_parity.py, which runs reduced-size scans:
voxel keys / match flags / associated leaves bit for bit, H / b within 1e-4 (asserted at 1e-7), the whole map state
after insert / recut / margi, the trajectory within 1 mm / 0.01 deg.

Reference semantics: src/mapping/voxel_map.cpp:241-266, src/mapping/octree.cpp:551-595 (association),
src/pipeline/odometry.cpp:98-148 (sums), voxel_map.cpp:47-135 + octree.cpp:151-495 (map update).
"""
import copy

import numpy as np
import pytest

from helpers import bootstrap_pair, compare_maps, cov_blocks, iekf_compare, sort_nodes
from vina_slam_b200 import synth

pytestmark = pytest.mark.gpu


def full_caps(cfg):
    return dict(max_scan_points=cfg.n_points + 1024, max_nodes=400000, hash_capacity_log2=20,
                fix_pool_points=8 << 20, win_pool_points=1 << 20)


@pytest.mark.parametrize("name", ["mid360", "velodyne32", "robosense128", "hilti_xt32"])
def test_full_size_parity_vs_oracle(oracle_lib, gpu_lib, name):
    """One sequence per config, full-size scans, oracle and CUDA path side by side:
    (1) map after the bootstrap and after 3 more scans with the window sliding (same down-sampled input on both
        sides): structure, cluster sums, eigen-decompositions, planes bit for bit;
    (2) IEKF association / sums on the next scan from a perturbed start, every iteration;
    (3) 5 steps of the whole path (each side with its own deskew and down-sampling): 1 mm / 0.01 deg."""
    cfg = synth.SENSORS[name]
    seq, od, gx, last = bootstrap_pair(oracle_lib, gpu_lib, cfg, caps=full_caps(cfg))
    try:
        mo, _ = compare_maps(od.map_export(), gx.map_export())
        assert mo.shape[0] > 3000 and (mo["is_plane"] > 0).sum() > 500
        for _ in range(3):
            sc = seq.next_scan(deskewed=True)
            assert sc.xyzt.shape[0] >= 0.98 * cfg.n_points
            od.bootstrap(sc.xyzt, oracle_lib.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
            gx.set_state(gpu_lib.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
            gx.down_upload(od.last_down())
            gx.var_init(1)
            gx.odom_map_update()
        compare_maps(od.map_export(), gx.map_export())
        assert od.map_count()[2] == gx.map_count()[2]
        od.set_imu_anchor(sc.end_time, sc.imu[-1])
        gx.set_imu_anchor(sc.end_time, sc.imu[-1])

        # (on a copy of the generator: the scan the comparison consumes is not inserted, the sequence itself goes on)
        matched = iekf_compare(oracle_lib, gpu_lib, cfg, pair=(copy.deepcopy(seq), od, gx), min_match=0.5)
        assert matched > cfg.n_points  # at least two iterations' worth of gate passes were compared

        worst_p = worst_r = 0.0
        od.set_state(oracle_lib.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
        gx.set_state(gpu_lib.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
        for k in range(5):
            s = seq.next_scan()
            r, _ = od.step(s.xyzt, s.beg_time, s.imu, iekf_on_full=True, max_iter=4)
            assert r == 0
            sg = gpu_lib.state_arrays(gx.step(s.xyzt, s.beg_time, s.imu, iekf_on_full=True, max_iter=4))
            so = oracle_lib.state_arrays(od.get_state())
            worst_p = max(worst_p, float(np.linalg.norm(sg["p"] - so["p"])))
            worst_r = max(worst_r, synth.rot_err_deg(sg["R"], so["R"]))
            tol = 0.05 if cfg.handheld else 0.02
            assert np.linalg.norm(so["p"] - s.gt_p) < tol and np.linalg.norm(sg["p"] - s.gt_p) < tol
        gx.sync()
        assert worst_p < 1e-3 and worst_r < 0.01, (worst_p, worst_r)
    finally:
        od.close()
        gx.close()


def _fill_scans(n_vox, per_scan, rng):
    """Body-frame clouds that put ONE point into each of `per_scan` distinct 1 m voxels (a 3-D block per scan) and the
    poses that lay the blocks side by side: many root voxels, no planes - what a long run leaves behind in surf_map."""
    side = int(round(per_scan ** (1 / 3)))
    g = np.stack(np.meshgrid(np.arange(side), np.arange(side), np.arange(side), indexing="ij"), -1).reshape(-1, 3)
    out = []
    for b in range(int(np.ceil(n_vox / g.shape[0]))):
        pts = np.zeros((g.shape[0], 4), dtype=np.float32)
        pts[:, :3] = g + rng.uniform(0.2, 0.8, g.shape)
        out.append((pts, np.array([(b % 8) * side, ((b // 8) % 8) * side, (b // 64) * side], dtype=np.float64)))
    return out


def test_large_map_association_parity(oracle_lib, gpu_lib):
    """Association against a map of > 10^6 root voxels on both sides: hash table at load factor 0.5 (2^21 slots, long
    linear-probe runs that the building's keys have to walk through), the resident voxels placed at the limits of
    the packed key (+2^20 on one axis, -2^20 on the other two), the building itself on negative coordinates. The
    whole map node for node, then keys / flags / associated leaves of a scan bit for bit, sums to tolerance.
    (The planes stay near the origin: plane_var carries the squared lever arm of the plane centre, 10^6 m away the
    reference's own sigma_l has no significant digits left - nothing to compare.)"""
    cfg = synth.small_sensor("hilti_xt32", 32, 1000)
    world = synth.World(offset=(-70.0, -40.0, -9.5))
    caps = dict(max_scan_points=140000, max_nodes=1250000, hash_capacity_log2=21, fix_pool_points=40 << 20,
                win_pool_points=1 << 20)
    od = oracle_lib.Odom(cfg)
    gx = gpu_lib.Ctx(cfg, **caps)
    try:
        rng = np.random.default_rng(77)
        lim = 1048575.0  # keys in [-lim, lim] are representable
        base = np.array([lim - 8 * 50 - 2.0, -lim + 1.0, -lim + 1.0])
        for pts, shift in _fill_scans(1_040_000, 125000, rng):
            p = base + shift
            od.bootstrap(pts, oracle_lib.make_state(np.eye(3), p, np.zeros(3)))
            gx.set_state(gpu_lib.make_state(np.eye(3), p, np.zeros(3)))
            gx.down_upload(od.last_down())
            gx.var_init(1)
            gx.odom_map_update()
        gx.sync()
        n_nodes, n_roots, _ = gx.map_count()
        assert n_roots > 1_000_000 and od.map_count()[1] == n_roots
        seq = synth.Sequence(cfg, world=world)
        sc = None
        for _ in range(cfg.win_size):
            sc = seq.next_scan(deskewed=True)
            od.bootstrap(sc.xyzt, oracle_lib.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
            gx.set_state(gpu_lib.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
            gx.down_upload(od.last_down())
            gx.var_init(1)
            gx.odom_map_update()
        gx.sync()
        assert od.map_count()[0] == gx.map_count()[0] and od.map_count()[1] == gx.map_count()[1]
        mo, mg = compare_maps(od.map_export(), gx.map_export())
        assert mo["key"][:, 0].max() >= lim - 3 and mo["key"][:, 1].min() < -lim + 3 and mo["key"][:, 2].min() < -lim + 3
        assert (mo["is_plane"] > 0).sum() > 300
        del mo, mg
        matched = iekf_compare(oracle_lib, gpu_lib, cfg, pair=(seq, od, gx), min_match=0.3)
        assert matched > 10000
    finally:
        od.close()
        gx.close()


def test_sigma_gate_ten_million_evaluations(oracle_lib, gpu_lib):
    """The second gate of OctoTree::match, dis_to_plane < 3 sqrt(sigma_l) (octree.cpp:564-569): the kernel evaluates
    sigma_l from per-plane hoisted terms (relative difference to the reference's evaluation order ~1e-9), so the
    decision is exact unless the two sides of the comparison agree to that level. More than 10^7 accepted gate
    evaluations on full-size scans (many starts around the ground truth, so that points sweep across their gates)
    must give exactly the oracle's flags and leaves."""
    cfg = synth.SENSORS["robosense128"]
    seq, od, gx, last = bootstrap_pair(oracle_lib, gpu_lib, cfg, caps=full_caps(cfg))
    try:
        sc = seq.next_scan(deskewed=True)
        pnt, var = oracle_lib.var_init(sc.xyzt, cfg)
        n = pnt.shape[0]
        cov = oracle_lib.state_arrays(oracle_lib.make_state())["cov"]
        rot_var, tsl_var = cov_blocks(cov)
        gx.pvec_upload(0, pnt, var)
        rng = np.random.default_rng(2024)
        evals = 0
        od.set_dump(True)
        rep = 0
        while evals < 10_000_000:
            # displacements from 1 mm to 6 cm / 0.01 to 0.5 deg: the gate of many points is close to its limit
            scale = 10.0 ** rng.uniform(-1.5, 0.3)
            R0 = sc.gt_R @ oracle_lib.exp_so3(rng.normal(0, 4e-3 * scale, 3))
            p0 = sc.gt_p + rng.normal(0, 0.03 * scale, 3)
            od.set_state(oracle_lib.make_state(R0, p0, sc.gt_v, t=sc.end_time))
            od.iekf(pnt, var, 1)
            d = od.iter_dump(0, n)
            gx.iekf_begin(0, rot_var, tsl_var)  # fresh leaf cache: every point is looked up and gated
            g = gx.iekf_accumulate(d["R_col"], d["p"], debug=True)
            a = gx.iekf_debug_assoc(n)
            assert np.array_equal(a["keys"], d["keys"]), rep
            assert np.array_equal(a["flags"], d["flags"]), (rep, int((a["flags"] != d["flags"]).sum()))
            assert np.array_equal(a["codes"], d["codes"]), rep
            assert g["match_num"] == d["match_num"]
            evals += int(d["match_num"])  # (a lower bound: the points that failed the second gate also evaluated it)
            rep += 1
        assert rep >= 40
    finally:
        od.close()
        gx.close()


def test_raw_front_to_back_path_matches_oracle(oracle_lib, gpu_lib):
    """The whole front of the loop, raw messages in: shuffled scans -> decoder keep rule + pcl_handler (filter, time
    sort, 0.11 s cut; lidar_decoder.cpp:8-43) -> sync_packages (sync.cpp:18-96) -> the per-scan step, through
    replay.replay_stream (vina_scan_prepare + vina_sync + vina_odom_step_prepared) against the ORACLE's restatements
    of the same three stages driven by the same message streams: same packages (scan, pcl_beg / pcl_end, IMU
    samples), prepared scans bit for bit, trajectory within 1 mm / 0.01 deg."""
    from vina_slam_b200 import replay

    cfg = synth.small_sensor("robosense128", 64, 900)  # 57 600 points per scan
    pfn = 2  # point_filter_num: every second point, like the reference's yaml files use
    boots, scans = replay.synthetic_frames(cfg, 9)
    caps = dict(max_scan_points=cfg.n_points + 1024, max_nodes=300000, hash_capacity_log2=19)
    rows, _, worst = replay.replay_stream(cfg, boots, scans, caps=caps, point_filter_num=pfn, shuffle_seed=3, prune_horizon=0)
    assert rows.shape[0] == len(scans) - 1 and worst < 0.02  # (the last scan's package stays open)

    # the oracle side: its own pcl_handler on the same shuffled raw scans, its own sync_packages on the same streams
    od = oracle_lib.Odom(cfg)
    for f in boots:
        od.bootstrap(f.xyzt, oracle_lib.make_state(f.gt_R, f.gt_p, f.gt_v, t=f.end_time))
    od.set_imu_anchor(boots[-1].end_time, boots[-1].imu[-1])
    rng = np.random.default_rng(3)
    raws = [f.xyzt[rng.permutation(f.xyzt.shape[0])] for f in scans]
    blind2 = float(cfg.blind) ** 2
    prepared = [oracle_lib.scan_prepare(r, pfn, blind2) for r in raws]
    assert all(p is not None and p.shape[0] > 0.45 * cfg.n_points for p in prepared)
    sync = oracle_lib.Sync(0)
    msgs = []
    for k, f in enumerate(scans):
        for row in f.imu:
            msgs.append((float(row[0]), 0, k, row))
        msgs.append((f.end_time, 1, k, None))
    msgs.sort(key=lambda m: (m[0], m[1]))
    got = []
    try:
        for _, kind, k, row in msgs:
            if kind == 0:
                sync.push_imu(row)
            else:
                sync.push_scan(scans[k].beg_time, float(prepared[k][-1, 3]), k)
            while True:
                r, tag, beg, end, imu = sync.next()
                assert r >= 0
                if r == 0:
                    break
                if r == 1:
                    rr, _ = od.step(prepared[tag], beg, imu, iekf_on_full=True, max_iter=4)
                    assert rr == 0
                    s = oracle_lib.state_arrays(od.get_state())
                    got.append((tag, s["t"], s["p"].copy(), s["R"].copy()))
    finally:
        sync.close()
        od.close()
    assert [g[0] for g in got] == list(range(len(scans) - 1))
    for (tag, t, p, R), row in zip(got, rows):
        assert abs(row[0] - t) < 1e-9
        assert np.linalg.norm(row[1:4] - p) < 1e-3, (tag, np.linalg.norm(row[1:4] - p))
        q = replay.quat_xyzw(R)
        ang = 2 * np.degrees(np.arccos(min(1.0, abs(float(np.dot(q, row[4:8]))))))
        assert ang < 0.01, (tag, ang)

    # and the prepared scans themselves, bit for bit (order of equal stamps included)
    gx = gpu_lib.Ctx(cfg, **caps)
    try:
        for k in (0, 4):
            n, t_last = gx.scan_prepare(raws[k], pfn, blind2)
            dev = gx.scan_download(n)
            assert n == prepared[k].shape[0] and np.array_equal(dev, prepared[k]) and t_last == prepared[k][-1, 3]
    finally:
        gx.close()

"""CPU tests of the oracle (the parity checker itself): analytic checks, invariants, golden vectors.

The reference ships no tests or fixtures and cannot be compiled here, so the oracle is validated by
(i) independent numerics (numpy eigh / inv), (ii) the derivation in docs/VNCLio_formulation.md
(numerical Jacobian of the point-to-plane residual), (iii) invariants of the reference's own data
structures, (iv) synthetic scenes with known planes and poses, (v) golden vectors of the oracle.
"""
import os

import numpy as np
import pytest

from vina_slam_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_small.npz")


def test_eig3_matches_numpy(oracle_lib):
    rng = np.random.default_rng(1)
    for k in range(200):
        A = rng.normal(size=(3, 3))
        S = A @ A.T * 10.0 ** rng.uniform(-6, 2)
        if k % 5 == 0:  # plane-like: one tiny eigenvalue
            S = S + np.outer(A[0], A[0]) * 1e4
        vals, vecs = oracle_lib.eig3(S)
        w, _ = np.linalg.eigh(S)
        assert np.all(np.diff(vals) >= 0)
        assert np.allclose(vals, w, rtol=1e-9, atol=1e-12 * abs(w).max())
        assert np.allclose(vecs.T @ vecs, np.eye(3), atol=1e-12)
        assert np.allclose(S @ vecs, vecs * vals, atol=1e-9 * abs(w).max())
    # only the lower triangle is read (Eigen's SelfAdjointView<Lower>)
    S = np.array([[2.0, 99.0, 99.0], [0.5, 3.0, 99.0], [0.1, 0.2, 4.0]])
    vals, _ = oracle_lib.eig3(S)
    L = np.tril(S) + np.tril(S, -1).T
    assert np.allclose(vals, np.linalg.eigvalsh(L), rtol=1e-12)
    # diagonal input: no rotation at all
    vals, vecs = oracle_lib.eig3(np.diag([3.0, 1.0, 2.0]))
    assert np.array_equal(vals, [1.0, 2.0, 3.0]) and np.allclose(np.abs(vecs), np.eye(3)[:, [1, 2, 0]])


def test_inverse15_matches_numpy(oracle_lib):
    rng = np.random.default_rng(2)
    A = rng.normal(size=(15, 15))
    P = A @ A.T + np.eye(15) * 1e-3
    assert np.allclose(oracle_lib.inverse15(P) @ P, np.eye(15), atol=1e-9)
    G = rng.normal(size=(15, 15))  # general (needs pivoting)
    G[0, 0] = 0.0
    assert np.allclose(oracle_lib.inverse15(G), np.linalg.inv(G), rtol=1e-8, atol=1e-10)


def test_so3_exp_log(oracle_lib):
    rng = np.random.default_rng(3)
    for _ in range(50):
        w = rng.normal(size=3)
        w = w / np.linalg.norm(w) * rng.uniform(0.002, 3.0)  # |w| < pi
        R = oracle_lib.exp_so3(w)
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-13) and abs(np.linalg.det(R) - 1) < 1e-13
        assert np.allclose(oracle_lib.log_so3(R), w, atol=1e-9)
    assert np.array_equal(oracle_lib.exp_so3(np.array([1e-10, 0, 0])), np.eye(3))  # math.hpp:15 threshold
    # Exp(w, dt) uses the 1e-7 threshold on |w| (math.hpp:29), not on |w| dt
    assert np.array_equal(oracle_lib.exp_so3(np.array([5e-8, 0, 0]), 1e6), np.eye(3))
    assert not np.array_equal(oracle_lib.exp_so3(np.array([2e-7, 0, 0]), 1e6), np.eye(3))


def test_voxel_key_semantics(oracle_lib):
    """voxel_map.cpp:246-253: double divide -> float -> '-1 if negative' -> truncate."""
    pw = np.array([[0.0, 0.99, 1.0], [-0.0, -0.01, -1.0], [-1.5, 2.5, -2.0], [1e-9, -1e-9, 16777217.0],
                   [0.3, 0.6, 0.9]])
    k = oracle_lib.voxel_keys(pw, 1.0)
    assert k.tolist() == [[0, 0, 1], [0, -1, -2], [-2, 2, -3], [0, -1, 16777216], [0, 0, 0]]
    # exactly-integer negatives fall one cell lower; -0.0 is not negative
    k = oracle_lib.voxel_keys(np.array([[0.3, 0.6, 0.9]]), 0.3)
    f = np.float32(0.9 / 0.3)  # 3.0000002 -> 3 after float rounding? the float decides, not the double
    assert k[0, 2] == int(f)
    k5 = oracle_lib.voxel_keys(np.array([[-0.25, 0.25, 1.75]]), 0.5)
    assert k5.tolist() == [[-1, 0, 3]]


def test_calc_body_var_properties(oracle_lib):
    cfg = synth.SENSORS["velodyne32"]
    xyz = np.array([[10.0, 0.0, 0.0, 0], [3.0, -4.0, 0.0, 0], [1.0, 2.0, 2.0, 0], [0.0, 0.0, 5.0, 0]], dtype=np.float32)
    pnt, var = oracle_lib.var_init(xyz, cfg)
    R = np.asarray(cfg.ext_R).reshape(3, 3)
    t = np.asarray(cfg.ext_t)
    for i in range(4):
        p = xyz[i, :3].astype(np.float64)
        if p[2] == 0:
            p[2] = 1e-4  # point_utils.cpp:5-8 mutates the point
        assert np.allclose(pnt[i], R @ p + t, atol=1e-12)
        V = var[i].reshape(3, 3).T
        assert np.allclose(V, V.T, atol=1e-15)
        w = np.linalg.eigvalsh(V)
        rng_ = np.linalg.norm(p)
        # one eigenvalue = range variance, two = (range * sin(beam_err))^2
        assert np.isclose(w.max() if cfg.dept_err ** 2 > (rng_ * np.sin(np.radians(cfg.beam_err))) ** 2 else w.min(),
                          np.float32(cfg.dept_err) ** 2, rtol=1e-5)
        ang = (np.float32(rng_) * np.sin(np.radians(np.float32(cfg.beam_err)))) ** 2
        assert np.isclose(np.sort(w)[1], ang, rtol=1e-5)


def test_downsample_voxel_mean(oracle_lib):
    rng = np.random.default_rng(4)
    pts = np.zeros((500, 4), dtype=np.float32)
    pts[:, :3] = rng.uniform(-2, 2, (500, 3))
    out = oracle_lib.down_sampling_voxel(pts, 0.5)
    key = lambda a: np.floor(a[:, :3].astype(np.float64) / 0.5).astype(np.int64)
    ko = {tuple(k) for k in key(out)}
    assert ko == {tuple(k) for k in key(pts)} and len(ko) == out.shape[0]
    assert out[:, 3].sum() == 500  # curvature carries the per-voxel count (point_utils.hpp:28-38)
    for row in out[:20]:
        sel = np.all(key(pts) == key(row[None]), axis=1)
        assert np.allclose(pts[sel, :3].mean(axis=0), row[:3], atol=1e-5)
    assert oracle_lib.down_sampling_voxel(pts, 0.0005).shape[0] == 500  # voxel_size < 0.001: untouched


def _plane_world_pair(oracle_lib, cfg, n_boot=None):
    seq = synth.Sequence(cfg)
    od = oracle_lib.Odom(cfg)
    sc = None
    for _ in range(cfg.win_size if n_boot is None else n_boot):
        sc = seq.next_scan(deskewed=True)
        od.bootstrap(sc.xyzt, oracle_lib.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
    od.set_imu_anchor(sc.end_time, sc.imu[-1])
    return seq, od, sc


def test_map_invariants_after_bootstrap(oracle_lib):
    cfg = synth.small_sensor("robosense128", 16, 400)
    seq, od, sc = _plane_world_pair(oracle_lib, cfg)
    m = od.map_export()
    leaf = m["octo_state"] == 0
    planes = m[leaf & (m["is_plane"] > 0)]
    assert planes.shape[0] > 100
    ev = planes["eig_value"]
    assert np.all(np.diff(ev, axis=1) >= 0)  # ascending
    assert np.all(ev[:, 0] < cfg.min_eigen_value)  # plane_judge (octree.cpp:198-201)
    assert np.all(ev[:, 0] / ev[:, 2] < 1.0 / np.array(cfg.plane_thre)[planes["layer"]])
    assert np.all(planes["N_add"] > np.array([20, 20, 15, 10])[planes["layer"]])
    # planes that have been through plane_update: unit normal, centre = v / N, the world is axis-aligned
    upd = planes[np.linalg.norm(planes["normal"], axis=1) > 0]
    assert upd.shape[0] > 50
    assert np.allclose(np.linalg.norm(upd["normal"], axis=1), 1.0, atol=1e-12)
    assert np.mean(np.max(np.abs(upd["normal"]), axis=1) > 0.98) > 0.9
    # every leaf: cov_add symmetric PSD-ish, interior nodes carry no window
    C = m["cov_add"].reshape(-1, 9, 9)
    assert np.allclose(C, np.transpose(C, (0, 2, 1)), rtol=1e-9, atol=1e-18)
    assert np.all(m["has_sw"][m["octo_state"] == 1] == 0)
    assert np.all(m["layer"] <= cfg.max_layer)
    # children centres: parent +- quater_length (octree.cpp:219-223)
    roots = m[m["layer"] == 0]
    vs = cfg.voxel_size
    assert np.allclose(roots["voxel_center"], (roots["key"] + 0.5) * vs)
    assert np.all(roots["quater_length"] == np.float32(vs / 4))


def test_point_to_plane_jacobian_and_convergence(oracle_lib):
    """docs/VNCLio_formulation.md §2.1: b = -sum R^-1 J r must be the (weighted) gradient of
    0.5 sum R^-1 r^2 w.r.t. a right-perturbation of R and a translation; and the IEKF recovers a
    perturbed pose on a scene of known planes."""
    cfg = synth.small_sensor("robosense128", 16, 400)
    seq, od, sc = _plane_world_pair(oracle_lib, cfg)
    nxt = seq.next_scan(deskewed=True)
    pnt, var = oracle_lib.var_init(nxt.xyzt, cfg)
    n = pnt.shape[0]
    od.set_dump(True)

    def sums(R, p):
        od.set_state(oracle_lib.make_state(R, p, nxt.gt_v, t=nxt.end_time))
        od.iekf(pnt, var, 1)
        return od.iter_dump(0, n)

    R0 = nxt.gt_R @ oracle_lib.exp_so3(np.array([0.002, -0.001, 0.0015]))
    p0 = nxt.gt_p + np.array([0.01, -0.02, 0.01])
    d0 = sums(R0, p0)
    assert d0["match_num"] > 0.4 * n  # sparse 16-beam scan: voxels further than ~10 m stay below min_point
    H, b = d0["HTH"], d0["HTz"]
    assert np.allclose(H, H.T, rtol=1e-12) and np.all(np.linalg.eigvalsh(H) > 0)
    # Gauss-Newton step from the sums moves the pose towards the ground truth
    dx = np.linalg.solve(H, b)
    R1 = R0 @ oracle_lib.exp_so3(dx[:3])
    p1 = p0 + dx[3:]
    assert np.linalg.norm(p1 - nxt.gt_p) < 0.5 * np.linalg.norm(p0 - nxt.gt_p)
    assert synth.rot_err_deg(R1, nxt.gt_R) < 0.5 * synth.rot_err_deg(R0, nxt.gt_R)
    # full IEKF from the perturbed pose
    od.set_state(oracle_lib.make_state(R0, p0, nxt.gt_v, t=nxt.end_time))
    ok = od.iekf(pnt, var, 4)
    s = oracle_lib.state_arrays(od.get_state())
    assert ok == 1 and np.linalg.norm(s["p"] - nxt.gt_p) < 1e-2 and synth.rot_err_deg(s["R"], nxt.gt_R) < 0.05
    # posterior covariance shrinks (odometry.cpp:223)
    assert np.trace(s["cov"][:6, :6]) < np.trace(oracle_lib.state_arrays(oracle_lib.make_state())["cov"][:6, :6])


def test_tracks_ground_truth_sequence(oracle_lib):
    cfg = synth.small_sensor("velodyne32", 16, 450)
    seq, od, sc = _plane_world_pair(oracle_lib, cfg)
    for _ in range(8):
        sc = seq.next_scan()
        r, desk = od.step(sc.xyzt, sc.beg_time, sc.imu, iekf_on_full=True, max_iter=4)
        assert r == 0
        s = oracle_lib.state_arrays(od.get_state())
        assert np.linalg.norm(s["p"] - sc.gt_p) < 0.01 and synth.rot_err_deg(s["R"], sc.gt_R) < 0.05
    assert od.window()[0] == cfg.win_size - 1
    assert sorted(od.window()[1].tolist()) == list(range(cfg.win_size))


def test_lidar_time_regress_is_reported(oracle_lib):
    """imu_ekf.cpp:19-24: the reference exit(0)s; the restatement returns an error code."""
    cfg = synth.small_sensor("robosense128", 8, 200)
    seq, od, sc = _plane_world_pair(oracle_lib, cfg, n_boot=1)
    nxt = seq.next_scan()
    od.set_imu_anchor(nxt.beg_time + 0.05, nxt.imu[0])  # last scan "ended" after this one begins
    r, _ = od.step(nxt.xyzt, nxt.beg_time, nxt.imu)
    assert r == -1


def test_golden_vectors(oracle_lib):
    g = np.load(GOLD)
    cfg = synth.small_sensor("robosense128", 16, 300)
    pnt, var = oracle_lib.var_init(g["scan_xyzt"][:64], cfg)
    assert np.array_equal(pnt, g["var_init_pnt"]) and np.array_equal(var, g["var_init_var"])
    import importlib.util
    import sys

    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(GOLD), "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["make_golden"] = mod
    spec.loader.exec_module(mod)
    o = mod.build()
    assert np.array_equal(o["scan_xyzt"], g["scan_xyzt"])  # the generator is seeded
    for k in ("it0_keys", "it0_codes", "it0_flags", "it0_match", "map_key", "map_code", "map_N", "map_plane",
              "map_state"):
        assert np.array_equal(o[k], g[k]), k
    for k in ("deskewed", "it0_HTH", "it0_HTz", "it0_nnt", "state_R", "state_p", "traj", "map_eig"):
        assert np.allclose(o[k], g[k], rtol=1e-9, atol=1e-12), k


def _scan_prepare_numpy(a, pfn, blind2):
    a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 4)
    x, y, z = a[:, 0], a[:, 1], a[:, 2]
    r2 = (x * x + y * y) + z * z  # float32 products and sums, left to right
    keep = (np.arange(a.shape[0]) % pfn == 0) & (r2.astype(np.float64) > blind2)
    b = a[keep]
    if b.shape[0] == 0:
        b = np.array([[0, 0, 0, 0], [0, 0, 0, 0.09]], dtype=np.float32)
    b = b[np.argsort(b[:, 3], kind="stable")]
    b = b[~(b[:, 3].astype(np.float64) > 0.11)]
    return b if b.shape[0] else None


def test_scan_front_end_restatement(oracle_lib):
    """Decoder keep rule + pcl_handler (lidar_pointcloud_decoder.cpp:70; lidar_decoder.cpp:16-34): the C++ restatement
    against an independent numpy one - decimation, blind zone, stable time sort with many equal stamps, the 0.11 s
    cut, the two-point stand-in for an empty cloud, and the case the reference cannot survive."""
    rng = np.random.default_rng(11)
    n = 50000
    a = np.zeros((n, 4), dtype=np.float32)
    a[:, :3] = rng.uniform(-30, 30, (n, 3))
    a[::7, :3] *= 0.01  # inside the blind zone
    a[:, 3] = rng.integers(0, 1300, n).astype(np.float32) * np.float32(1e-4)  # many ties, some beyond 0.11 s
    for pfn, blind2 in ((1, 0.01), (3, 0.01), (2, 4.0)):
        o, g = oracle_lib.scan_prepare(a, pfn, blind2), _scan_prepare_numpy(a, pfn, blind2)
        assert o.shape == g.shape and o.shape[0] > 1000 and np.array_equal(o, g)
        assert np.all(np.diff(o[:, 3]) >= 0) and o[-1, 3] <= np.float32(0.11)
    near = a.copy()
    near[:, :3] *= 1e-4
    o = oracle_lib.scan_prepare(near, 1, 0.01)  # everything in the blind zone -> lidar_decoder.cpp:16-27
    assert np.array_equal(o, np.array([[0, 0, 0, 0], [0, 0, 0, 0.09]], dtype=np.float32))
    assert np.array_equal(oracle_lib.scan_prepare(np.zeros((0, 4), dtype=np.float32), 1, 0.01), o)
    late = a.copy()
    late[:, 3] += np.float32(0.2)
    assert oracle_lib.scan_prepare(late, 1, 0.01) is None and _scan_prepare_numpy(late, 1, 0.01) is None

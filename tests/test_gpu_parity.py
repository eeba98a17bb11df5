"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): voxel keys and point-to-voxel associations bit-exact; H/b within 1e-4
relative; trajectory within 1 mm / 0.01 deg. Tighter bars are asserted where the design guarantees them
(bit-exact body points, cluster sums, eigen-decompositions and plane decisions).
"""
import numpy as np
import pytest

from helpers import (SMALL_CAPS, bootstrap_pair, col, compare_maps as _compare_maps, cov_blocks, iekf_compare as _iekf_compare,
                     rel_err, small_cfg, sort_nodes, ulp_diff_f32)
from vina_slam_b200 import synth

pytestmark = pytest.mark.gpu


def test_ctx_create_and_empty_map(gpu_lib):
    cfg = small_cfg()
    gx = gpu_lib.Ctx(cfg, **SMALL_CAPS)
    n, nr, ns = gx.map_count()
    assert (n, nr, ns) == (0, 0, 0)
    gx.close()


def test_var_init_bit_exact(oracle_lib, gpu_lib):
    """a3: body point and the upper triangle of the body covariance, bit for bit (incl. the z == 0 mutation)."""
    for base in ("robosense128", "velodyne32", "mid360"):
        cfg = small_cfg(base, 16, 500) if base != "mid360" else synth.small_sensor(base, 1, 8000)
        seq = synth.Sequence(cfg)
        sc = seq.next_scan(deskewed=True)
        xyzt = sc.xyzt.copy()
        xyzt[5, 2] = 0.0  # calcBodyVar's pb[2] == 0 branch (point_utils.cpp:5-8)
        pnt_o, var_o = oracle_lib.var_init(xyzt, cfg)
        gx = gpu_lib.Ctx(cfg, **SMALL_CAPS)
        gx.down_upload(xyzt)
        gx.var_init(1)
        pnt_g, var_g = gx.pvec_download(1, xyzt.shape[0])
        assert np.array_equal(pnt_g, pnt_o)
        for (i, j) in ((0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)):
            assert np.array_equal(var_g[:, i + 3 * j], var_o[:, i + 3 * j]), (base, i, j)
        gx.close()


def test_imu_propagation_and_deskew(oracle_lib, gpu_lib):
    """a1 (host) within 1e-12; a2 deskewed float32 xyz within 1 ulp and >= 99.99 % bit-equal
    (device sin/cos vs glibc differ in the last ulp of a double now and then)."""
    cfg = small_cfg()
    seq, od, gx, last = bootstrap_pair(oracle_lib, gpu_lib, cfg, n_boot=2)
    sc = seq.next_scan()
    end = sc.beg_time + float(sc.xyzt[-1, 3])
    st = oracle_lib.make_state(last.gt_R, last.gt_p, last.gt_v, t=last.end_time)
    od.set_state(st)
    gx.set_state(gpu_lib.make_state(last.gt_R, last.gt_p, last.gt_v, t=last.end_time))
    assert od.propagate(sc.beg_time, end, sc.imu) == 0
    poses_o = od.imu_poses()
    poses_g = gx.propagate(sc.beg_time, end, sc.imu)
    assert poses_g.shape[0] == poses_o.shape[0] and poses_o.shape[0] >= 15
    flat_g = np.concatenate([poses_g[k].reshape(len(poses_g), -1) for k in ("t", "R", "p", "v", "w", "a")], axis=1)
    assert np.max(np.abs(flat_g - poses_o)) < 1e-12
    so = oracle_lib.state_arrays(od.get_state())
    sg = gpu_lib.state_arrays(gx.get_state())
    assert np.max(np.abs(so["R"] - sg["R"])) < 1e-13 and np.max(np.abs(so["p"] - sg["p"])) < 1e-12
    assert rel_err(sg["cov"], so["cov"]) < 1e-12
    # deskew with the ORACLE's pose table on both sides
    desk_o = od.deskew(sc.xyzt)
    ps = np.zeros(poses_o.shape[0], dtype=gpu_lib.IMU_POSE_DTYPE)
    ps["t"], ps["R"], ps["p"], ps["v"], ps["w"], ps["a"] = (poses_o[:, 0], poses_o[:, 1:10], poses_o[:, 10:13],
                                                           poses_o[:, 13:16], poses_o[:, 16:19], poses_o[:, 19:22])
    gx.scan_upload(sc.xyzt)
    gx.deskew(ps, col(so["R"]), so["p"])
    desk_g = gx.scan_download(sc.xyzt.shape[0])
    assert np.array_equal(desk_g[:, 3], sc.xyzt[:, 3])
    ulps = ulp_diff_f32(desk_g[:, :3], desk_o[:, :3])
    assert ulps.max() <= 1
    assert (ulps > 0).mean() < 1e-4
    # points at or before the first pose stay untouched (imu_ekf.cpp:124)
    head = sc.xyzt[:, 3] <= poses_o[0, 0]
    assert np.array_equal(desk_g[head, :3], sc.xyzt[head, :3])
    moved = np.abs(desk_o[:, :3] - sc.xyzt[:, :3]).max()
    assert moved > 1e-3  # the scan really was distorted
    gx.close()


def test_deskew_rejects_unsorted_scan(oracle_lib, gpu_lib):
    cfg = small_cfg()
    seq, od, gx, last = bootstrap_pair(oracle_lib, gpu_lib, cfg, n_boot=1)
    sc = seq.next_scan()
    bad = sc.xyzt.copy()
    bad[[10, 2000]] = bad[[2000, 10]]
    poses = gx.propagate(sc.beg_time, sc.end_time, sc.imu)
    gx.scan_upload(bad)
    s = gpu_lib.state_arrays(gx.get_state())
    gx.deskew(poses, col(s["R"]), s["p"])
    with pytest.raises(gpu_lib.VinaError) as ei:
        gx.scan_download(bad.shape[0])
    assert ei.value.code == -4
    gx.close()


def test_iekf_association_and_sums_robosense(oracle_lib, gpu_lib):
    _iekf_compare(oracle_lib, gpu_lib, small_cfg("robosense128", 32, 600))


def test_iekf_association_and_sums_velodyne_layer3(oracle_lib, gpu_lib):
    _iekf_compare(oracle_lib, gpu_lib, small_cfg("velodyne32", 32, 500))


def test_iekf_association_and_sums_mid360_negative_keys(oracle_lib, gpu_lib):
    """voxel_size 0.5, non-repetitive pattern, world shifted so that keys are negative on all three axes."""
    cfg = synth.small_sensor("mid360", 1, 12000)
    total = _iekf_compare(oracle_lib, gpu_lib, cfg, world=synth.World(offset=(-70.0, -40.0, -9.5)))
    assert total > 0


@pytest.mark.parametrize("base,beams,steps", [("robosense128", 32, 600), ("velodyne32", 32, 500)])
def test_map_insert_recut_margi_parity(oracle_lib, gpu_lib, base, beams, steps):
    """a8-a12 over bootstrap + sliding steps with identical inputs: the whole map state is compared."""
    cfg = small_cfg(base, beams, steps)
    seq, od, gx, last = bootstrap_pair(oracle_lib, gpu_lib, cfg)
    mo, mg = _compare_maps(od.map_export(), gx.map_export())
    assert mo.shape[0] > 2000 and (mo["is_plane"] > 0).sum() > 300 and (mo["octo_state"] == 1).sum() > 100
    assert od.map_count()[2] == gx.map_count()[2]  # surf_map_slide size
    # more scans with the window sliding (marginalisation each scan), same poses on both sides
    for k in range(6):
        sc = seq.next_scan(deskewed=True)
        od.bootstrap(sc.xyzt, oracle_lib.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
        gx.set_state(gpu_lib.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
        gx.down_upload(od.last_down())
        gx.var_init(1)
        gx.odom_map_update()
    _compare_maps(od.map_export(), gx.map_export())
    assert od.window()[0] == gx.window()[0] and np.array_equal(od.window()[1], gx.window()[1])
    assert od.map_count()[2] == gx.map_count()[2]
    gx.close()


def test_insert_dropped_when_fewer_roots_than_threads(oracle_lib, gpu_lib):
    """voxel_map.cpp:96-97: a scan touching fewer root voxels than thread_num is not inserted at all."""
    cfg = small_cfg()
    od = oracle_lib.Odom(cfg)
    gx = gpu_lib.Ctx(cfg, **SMALL_CAPS)
    rng = np.random.default_rng(5)
    pts = np.zeros((300, 4), dtype=np.float32)
    pts[:, :3] = rng.uniform(0.1, 0.9, (300, 3)) + np.array([3.0, 3.0, 0.0])  # a single 1 m voxel ... plus 2 more
    pts[100:200, 0] += 1.0
    pts[200:, 1] += 1.0
    st_o = oracle_lib.make_state(np.eye(3), np.zeros(3), np.zeros(3))
    od.bootstrap(pts, st_o)
    gx.set_state(gpu_lib.make_state(np.eye(3), np.zeros(3), np.zeros(3)))
    gx.down_upload(od.last_down())
    gx.var_init(1)
    gx.odom_map_update()
    mo, mg = sort_nodes(od.map_export()), sort_nodes(gx.map_export())
    assert mo.shape[0] == mg.shape[0] == 3
    assert (mo["N_add"] == 0).all() and (mg["N_add"] == 0).all()
    assert np.array_equal(mo["isexist"], mg["isexist"]) and np.array_equal(mo["key"], mg["key"])
    gx.close()


def test_downsample_matches_reference_voxel_set(oracle_lib, gpu_lib):
    """f1: same voxels and counts as the reference; means agree to fp32 rounding (order-dependent running
    mean in the reference vs exact sum here)."""
    cfg = small_cfg()
    seq = synth.Sequence(cfg)
    sc = seq.next_scan(deskewed=True)
    ref = oracle_lib.down_sampling_voxel(sc.xyzt, cfg.down_size)
    gx = gpu_lib.Ctx(cfg, **SMALL_CAPS)
    gx.scan_upload(sc.xyzt)
    gx.downsample()
    got = gx.down_download(sc.xyzt.shape[0])
    assert got.shape[0] == ref.shape[0]

    def keyed(a):
        k = np.floor(a[:, :3].astype(np.float64) / cfg.down_size).astype(np.int64)
        o = np.lexsort((k[:, 2], k[:, 1], k[:, 0]))
        return k[o], a[o]

    kr, r = keyed(ref)
    kg, g = keyed(got)
    assert np.array_equal(kr, kg)
    assert np.array_equal(r[:, 3], g[:, 3])  # per-voxel point counts
    assert np.max(np.abs(r[:, :3] - g[:, :3])) < 2e-5
    # deterministic: voxels come out in order of their first point
    gx.scan_upload(sc.xyzt)
    gx.downsample()
    again = gx.down_download(sc.xyzt.shape[0])
    assert np.array_equal(again, got)
    gx.close()


def test_end_to_end_trajectory(oracle_lib, gpu_lib):
    """Full path through vina_odom_step from host buffers vs the oracle's step: 1 mm / 0.01 deg."""
    cfg = small_cfg("robosense128", 32, 600)
    seq, od, gx, last = bootstrap_pair(oracle_lib, gpu_lib, cfg, gpu_own_downsample=True)
    worst_p, worst_r = 0.0, 0.0
    for k in range(12):
        sc = seq.next_scan()
        r, _ = od.step(sc.xyzt, sc.beg_time, sc.imu, iekf_on_full=True, max_iter=4)
        assert r == 0
        sg = gpu_lib.state_arrays(gx.step(sc.xyzt, sc.beg_time, sc.imu, iekf_on_full=True, max_iter=4))
        so = oracle_lib.state_arrays(od.get_state())
        worst_p = max(worst_p, float(np.linalg.norm(sg["p"] - so["p"])))
        worst_r = max(worst_r, synth.rot_err_deg(sg["R"], so["R"]))
        # and both track the ground truth of the synthetic trajectory
        assert np.linalg.norm(sg["p"] - sc.gt_p) < 0.02
    gx.sync()
    assert worst_p < 1e-3, worst_p
    assert worst_r < 0.01, worst_r
    gx.close()


def test_error_codes_and_edge_cases(oracle_lib, gpu_lib):
    """Error behaviour of the boundary: codes instead of exit() / exceptions across the ABI."""
    import ctypes as C

    cfg = small_cfg()
    gx = gpu_lib.Ctx(cfg, **SMALL_CAPS)
    lib = gx.lib
    # accumulate before begin
    z9, z3 = np.zeros(9), np.zeros(3)
    with pytest.raises(gpu_lib.VinaError) as ei:
        gx.iekf_accumulate(np.eye(3).reshape(-1), z3)
    assert ei.value.code == -6
    # scan larger than max_scan_points
    big = np.zeros((SMALL_CAPS["max_scan_points"] + 1, 4), dtype=np.float32)
    assert lib.vina_scan_upload(gx.h, big.ctypes.data_as(C.c_void_p), C.c_int(big.shape[0])) == -3
    assert b"max_scan_points" in lib.vina_last_error(gx.h)
    # bad arguments
    assert lib.vina_scan_upload(gx.h, None, C.c_int(5)) == -1
    assert lib.vina_var_init(gx.h, C.c_int(7)) == -1
    assert lib.vina_map_recut(gx.h, C.c_int(0), None) == -1
    # empty clouds are fine for the per-stage entries
    gx.down_upload(np.zeros((0, 4), dtype=np.float32))
    gx.var_init(1)
    assert gx.pvec_download(1, 1)[0].shape[0] == 0
    # IEKF on an empty map: zero matches, zero sums
    seq = synth.Sequence(cfg)
    sc = seq.next_scan(deskewed=True)
    gx.scan_upload(sc.xyzt)
    gx.var_init(0)
    gx.iekf_begin(0, np.eye(3).reshape(-1) * 1e-4, np.eye(3).reshape(-1) * 1e-4)
    r = gx.iekf_accumulate(col(sc.gt_R), sc.gt_p)
    assert r["match_num"] == 0 and not r["HTH"].any() and not r["HTz"].any()
    # LiDAR time regress (imu_ekf.cpp:19-24 exit(0)s) -> VINA_E_TIME
    gx.set_imu_anchor(sc.beg_time + 0.05, sc.imu[0])
    with pytest.raises(gpu_lib.VinaError) as ei:
        gx.step(sc.xyzt, sc.beg_time, sc.imu)
    assert ei.value.code == -5
    gx.close()
    # node pool exhaustion is reported as VINA_E_CAPACITY at the next sync, never a crash
    tiny = dict(SMALL_CAPS)
    tiny["max_nodes"] = 64
    gy = gpu_lib.Ctx(cfg, **tiny)
    with pytest.raises(gpu_lib.VinaError) as ei:
        gy.bootstrap(sc.xyzt, gpu_lib.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
    assert ei.value.code == -3
    gy.close()
    # voxel keys outside the 21-bit range
    gz = gpu_lib.Ctx(cfg, **SMALL_CAPS)
    far = sc.xyzt.copy()
    far[:, 0] += 3.0e6
    with pytest.raises(gpu_lib.VinaError) as ei:
        gz.bootstrap(far, gpu_lib.make_state(np.eye(3), np.zeros(3), np.zeros(3)))
    assert ei.value.code == -3
    gz.close()


def test_iekf_is_deterministic(oracle_lib, gpu_lib):
    """Same inputs -> bit-identical sums (fixed reduction order, no floating-point atomics)."""
    cfg = small_cfg()
    seq, od, gx, last = bootstrap_pair(oracle_lib, gpu_lib, cfg)
    sc = seq.next_scan(deskewed=True)
    gx.scan_upload(sc.xyzt)
    gx.var_init(0)
    rv = np.eye(3).reshape(-1) * 1e-4
    outs = []
    for _ in range(3):
        gx.iekf_begin(0, rv, rv)
        outs.append(gx.iekf_accumulate(col(sc.gt_R), sc.gt_p))
    for o in outs[1:]:
        assert np.array_equal(o["HTH"], outs[0]["HTH"]) and np.array_equal(o["HTz"], outs[0]["HTz"])
        assert np.array_equal(o["nnt"], outs[0]["nnt"]) and o["match_num"] == outs[0]["match_num"]
    assert outs[0]["match_num"] > 0.5 * sc.xyzt.shape[0]
    gx.close()


def test_gpu_vs_reference_build(oracle_lib, gpu_lib):
    """The CUDA path directly against oracle/_ref (the reference's own sources compiled against the header shims):
    same map (structure, cluster sums, planes) after bootstrap + sliding steps, same trajectory within 1 mm / 0.01 deg."""
    if not oracle_lib.have_ref():
        pytest.skip("oracle/_ref is not built here (needs /root/reference)")
    cfg = small_cfg("robosense128", 32, 500, seed=11)
    seq = synth.Sequence(cfg)
    rf = oracle_lib.Odom(cfg, ref=True)
    gx = gpu_lib.Ctx(cfg, **SMALL_CAPS)
    try:
        sc = None
        for _ in range(cfg.win_size):
            sc = seq.next_scan(deskewed=True)
            rf.bootstrap(sc.xyzt, oracle_lib.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
            gx.set_state(gpu_lib.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
            gx.down_upload(rf.last_down())
            gx.var_init(1)
            gx.odom_map_update()
        _compare_maps(rf.map_export(), gx.map_export(), eig_planes_only=True)
        q = lambda imu: np.column_stack([np.round(imu[:, 0] * 1e9) * 1e-9, imu[:, 1:]])  # rclcpp::Time keeps ns
        rf.set_imu_anchor(sc.end_time, q(sc.imu)[-1])
        gx.set_imu_anchor(sc.end_time, q(sc.imu)[-1])
        for k in range(8):
            sc = seq.next_scan()
            r, _ = rf.step(sc.xyzt, sc.beg_time, q(sc.imu), True, 4)
            assert r == 0
            sg = gpu_lib.state_arrays(gx.step(sc.xyzt, sc.beg_time, q(sc.imu), True, 4))
            sr = oracle_lib.state_arrays(rf.get_state())
            assert np.linalg.norm(sg["p"] - sr["p"]) < 1e-3 and synth.rot_err_deg(sg["R"], sr["R"]) < 0.01
    finally:
        rf.close()
        gx.close()


@pytest.mark.parametrize("name", ["mid360", "velodyne32", "robosense128", "hilti_xt32"])
def test_full_size_configs_properties(gpu_lib, name):
    """BASELINE.json configs at their full sizes (20 k / 57.6 k / 240 k / 64 k pts per scan), checked through
    size-independent properties: the odometry tracks the synthetic ground truth, the run is reproducible bit for
    bit (two contexts, same inputs -> identical states and maps), and the map keeps its structural invariants."""
    cfg = synth.SENSORS[name]
    seq = synth.Sequence(cfg)
    boots = [seq.next_scan(deskewed=True) for _ in range(cfg.win_size)]
    scans = [seq.next_scan() for _ in range(4)]
    caps = dict(max_scan_points=cfg.n_points + 1024, max_nodes=400000, hash_capacity_log2=20)
    states, maps = [], []
    for rep in range(2):
        gx = gpu_lib.Ctx(cfg, **caps)
        for sc in boots:
            gx.bootstrap(sc.xyzt, gpu_lib.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
        gx.set_imu_anchor(boots[-1].end_time, boots[-1].imu[-1])
        st = []
        for sc in scans:
            s = gpu_lib.state_arrays(gx.step(sc.xyzt, sc.beg_time, sc.imu, True, 4))
            st.append(np.concatenate([s["p"], s["R"].reshape(-1), s["cov"].reshape(-1)]))
            tol_p, tol_r = (0.03, 0.2) if cfg.handheld else (0.01, 0.05)
            assert np.linalg.norm(s["p"] - sc.gt_p) < tol_p, (name, np.linalg.norm(s["p"] - sc.gt_p))
            assert synth.rot_err_deg(s["R"], sc.gt_R) < tol_r
        states.append(np.array(st))
        maps.append(sort_nodes(gx.map_export()))
        gx.close()
    assert np.array_equal(states[0], states[1]), "same inputs must give bit-identical states"
    m0, m1 = maps
    assert m0.shape == m1.shape
    for f in m0.dtype.names:
        assert np.array_equal(m0[f], m1[f]), f"map field {f} is not reproducible"
    leaf = m0["octo_state"] == 0
    planes = m0[leaf & (m0["is_plane"] > 0)]
    assert planes.shape[0] > 200
    ev = planes["eig_value"]
    assert np.all(np.diff(ev, axis=1) >= 0) and np.all(ev[:, 0] < cfg.min_eigen_value)
    assert np.all(ev[:, 0] / ev[:, 2] < 1.0 / np.array(cfg.plane_thre)[planes["layer"]])
    assert np.all(m0["layer"] <= cfg.max_layer) and np.all(m0["has_sw"][m0["octo_state"] == 1] == 0)
    roots = m0[m0["layer"] == 0]
    assert np.array_equal(roots["voxel_center"], (roots["key"] + 0.5) * cfg.voxel_size)
    upd = planes[np.linalg.norm(planes["normal"], axis=1) > 0]
    assert np.allclose(np.linalg.norm(upd["normal"], axis=1), 1.0, atol=1e-12)


def test_device_iekf_loop_matches_host_solve(oracle_lib, gpu_lib):
    """a7 on the device (k_iekf's last block: K(:,0:6) = P(:,0:6)(I + H P66)^-1, boxplus, convergence / rematch
    logic, P = (I - G)P) against the same loop with the reference's 15x15 update restated on the host
    (odometry.cpp:192-230): same iteration count, states and covariances to 1e-9."""
    cfg = small_cfg("robosense128", 32, 600)
    seq, od, gx, last = bootstrap_pair(oracle_lib, gpu_lib, cfg, gpu_own_downsample=True)
    for k in range(4):
        sc = seq.next_scan(deskewed=True)
        # start both variants from the same perturbed state on the same scan
        st = gpu_lib.make_state(sc.gt_R @ synth.rot_exp(np.array([2e-3, -1e-3, 1.5e-3])),
                                sc.gt_p + np.array([0.02, -0.015, 0.01]), sc.gt_v, t=sc.end_time)
        gx.scan_upload(sc.xyzt)
        gx.var_init(0)
        res = []
        for host in (False, True):
            gx.set_state(st)
            it, ok = gx.odom_iekf(0, 4, host_solve=host)
            res.append((it, ok, gpu_lib.state_arrays(gx.get_state())))
        (it_d, ok_d, sd), (it_h, ok_h, sh) = res
        assert it_d == it_h and ok_d == ok_h, (it_d, it_h, ok_d, ok_h)
        assert 1 <= it_d <= 4
        for f in ("R", "p", "v", "bg", "ba"):
            assert np.max(np.abs(sd[f] - sh[f])) < 1e-9, (f, np.max(np.abs(sd[f] - sh[f])))
        assert rel_err(sd["cov"], sh["cov"]) < 1e-7
        # and the update really moved the perturbed state back to the ground truth
        assert np.linalg.norm(sd["p"] - sc.gt_p) < 5e-3
    gx.close()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_map_equals_single_gpu_map(oracle_lib, gpu_lib, world):
    """the map partitioned by voxel-hash range. `world` shard contexts (here on one GPU, records
    exchanged in-process with the permutation tests/test_sharded_cpu.py checks against the real all-to-all) are
    fed ascending slices of every scan; the union of the shards must equal, byte for byte, the map one context
    builds from the same scans - through bootstrap, subdivision and the sliding-window marginalisation."""
    import torch

    from vina_slam_b200 import sharded

    cfg = small_cfg("robosense128", 32, 600)
    seq = synth.Sequence(cfg)
    od = oracle_lib.Odom(cfg)  # only supplies the down-sampled clouds
    single = gpu_lib.Ctx(cfg, **SMALL_CAPS)
    shards = [sharded.MapShard(gpu_lib.Ctx(cfg, **SMALL_CAPS), r, world) for r in range(world)]
    routed_to = np.zeros(world, dtype=np.int64)
    for k in range(cfg.win_size + 5):
        sc = seq.next_scan(deskewed=True)
        od.bootstrap(sc.xyzt, oracle_lib.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
        down = od.last_down()
        st = gpu_lib.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time)
        single.set_state(st)
        single.down_upload(down)
        single.var_init(1)
        single.odom_map_update()
        sa = gpu_lib.state_arrays(st)
        Rc, p = col(sa["R"]), sa["p"]
        rv, tv = cov_blocks(sa["cov"])
        sends, counts = [], []
        for sh in shards:  # every rank holds the scan, routes its slice
            sh.ctx.down_upload(down)
            sh.ctx.var_init(1)
            sh.push_pose(Rc, p)
            first, cnt = sharded.slice_of(down.shape[0], sh.rank, world)
            s_, c_ = sh.route(first, cnt, 0, Rc, p, rv, tv)
            assert int(c_.sum()) == cnt
            sends.append(s_)
            counts.append(c_)
        recvs = sharded.local_exchange(sends, counts)
        torch.cuda.synchronize()
        loc = [sh.insert_begin(rv_) for sh, rv_ in zip(shards, recvs)]
        g_roots, g_slide = sum(a for a, _ in loc), sum(b for _, b in loc)
        for sh, rv_ in zip(shards, recvs):
            routed_to[sh.rank] += rv_.shape[0]
            # ascending scan order at the owner: the property the bit-exactness rests on
            gi = rv_[:, 12].contiguous().view(torch.int64).cpu().numpy()
            assert np.all(np.diff(gi) > 0)
            sh.insert_finish(g_roots, g_slide)
            sh.recut_margi()
        if k == cfg.win_size + 2:
            # the idle path's pruning (1 m horizon), every rank on its own shard with the replicated journey: the
            # shards erase, together, exactly the roots the single map erases; two more scans follow on the pruned maps
            assert all((sh.jour, sh.release_flag) == single.journey() for sh in shards) and single.journey()[1]
            es = single.idle(1)
            ep = [sh.idle(1) for sh in shards]
            assert es[0] > 20 and (sum(a for a, _ in ep), sum(b for _, b in ep)) == es, (es, ep)
    ms = sort_nodes(single.map_export())
    parts = [sh.ctx.map_export() for sh in shards]
    lib = gpu_lib.load()
    import ctypes as C

    for r, part in enumerate(parts):  # every node sits on the owner of its root voxel
        ow = [lib.vina_shard_owner(C.c_int64(int(k[0])), C.c_int64(int(k[1])), C.c_int64(int(k[2])), world)
              for k in part["key"]]
        assert part.shape[0] > 100 and all(o == r for o in ow)
    mu = sort_nodes(np.concatenate(parts))
    assert mu.shape[0] == ms.shape[0] > 2000
    for f in ms.dtype.names:
        assert np.array_equal(ms[f], mu[f]), f"sharded map differs from the single-GPU map in {f}"
    assert (ms["octo_state"] == 1).sum() > 100 and (ms["N_fix"] > 0).sum() > 100  # subdivision and margi happened
    assert sum(sh.ctx.map_count()[2] for sh in shards) == single.map_count()[2]  # surf_map_slide
    assert routed_to.min() > 0.5 * routed_to.mean()  # balanced
    single.close()
    for sh in shards:
        sh.ctx.close()


def test_batch_replay_matches_separate_contexts(oracle_lib, gpu_lib):
    """vina_batch (config 5): three different sequences advanced in lock step with one k_iekf launch per
    iteration for all of them, against the same sequences stepped one context at a time. Same kernels; only the
    number of blocks per sequence (hence the order of the final sum) differs -> states to 1e-9, identical maps."""
    import torch

    cfg = small_cfg("robosense128", 32, 600)
    B = 3
    seqs = [synth.Sequence(cfg, seed=cfg.seed + 11 * b) for b in range(B)]
    solo = [gpu_lib.Ctx(cfg, **SMALL_CAPS) for _ in range(B)]
    grp = [gpu_lib.Ctx(cfg, **SMALL_CAPS) for _ in range(B)]
    last = [None] * B
    for b in range(B):
        for _ in range(cfg.win_size):
            sc = seqs[b].next_scan(deskewed=True)
            st = gpu_lib.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time)
            solo[b].bootstrap(sc.xyzt, st)
            grp[b].bootstrap(sc.xyzt, st)
            last[b] = sc
        for g in (solo[b], grp[b]):
            g.set_imu_anchor(last[b].end_time, last[b].imu[-1])
    batch = gpu_lib.Batch(grp)
    dev = torch.device("cuda", 0)
    for k in range(5):
        scs = [s.next_scan() for s in seqs]
        d = [torch.from_numpy(sc.xyzt).to(dev) for sc in scs]
        torch.cuda.synchronize()
        outs = batch.step_resident([t.data_ptr() for t in d], [sc.xyzt.shape[0] for sc in scs],
                                   [sc.beg_time for sc in scs], [sc.end_time for sc in scs], [sc.imu for sc in scs])
        for b in range(B):
            so = gpu_lib.state_arrays(solo[b].step_resident(d[b].data_ptr(), scs[b].xyzt.shape[0], scs[b].beg_time,
                                                            scs[b].end_time, scs[b].imu))
            sb = gpu_lib.state_arrays(outs[b])
            for f in ("R", "p", "v", "bg", "ba"):
                assert np.max(np.abs(so[f] - sb[f])) < 1e-9, (k, b, f)
            assert rel_err(sb["cov"], so["cov"]) < 1e-7
            assert np.linalg.norm(sb["p"] - scs[b].gt_p) < 0.02
            assert solo[b].timings().iekf_iters == grp[b].timings().iekf_iters
    batch.sync()
    ms, launches = batch.iekf_time()
    assert launches == 4 and len(ms) == 4 and ms[0] > 0
    for b in range(B):
        ma, mb = sort_nodes(solo[b].map_export()), sort_nodes(grp[b].map_export())
        assert ma.shape[0] == mb.shape[0]
        for f in ("key", "code", "N_add", "N_fix", "is_plane", "octo_state"):
            assert np.array_equal(ma[f], mb[f]), f
    batch.close()
    for g in solo + grp:
        g.close()


def test_sharded_association_matches_single_gpu(oracle_lib, gpu_lib):
    """SURVEY §8e, the query side: the IEKF against a map sharded over 2 contexts (queries routed to the owners,
    per-shard sums added - here in-process, the permutation is the one the gloo test checks) against the same
    loop on the single-GPU map with the same host update: identical match counts, sums to 1e-10, states to 1e-10."""
    import torch

    from vina_slam_b200 import sharded

    world = 2
    cfg = small_cfg("robosense128", 32, 600)
    seq = synth.Sequence(cfg)
    od = oracle_lib.Odom(cfg)
    single = gpu_lib.Ctx(cfg, **SMALL_CAPS)
    shards = [sharded.MapShard(gpu_lib.Ctx(cfg, **SMALL_CAPS), r, world) for r in range(world)]
    for k in range(cfg.win_size):
        sc = seq.next_scan(deskewed=True)
        od.bootstrap(sc.xyzt, oracle_lib.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
        down = od.last_down()
        st = gpu_lib.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time)
        single.set_state(st)
        single.down_upload(down)
        single.var_init(1)
        single.odom_map_update()
        sa = gpu_lib.state_arrays(st)
        Rc, p = col(sa["R"]), sa["p"]
        rv, tv = cov_blocks(sa["cov"])
        sends, counts = [], []
        for sh in shards:
            sh.ctx.down_upload(down)
            sh.ctx.var_init(1)
            sh.push_pose(Rc, p)
            s_, c_ = sh.route(*sharded.slice_of(down.shape[0], sh.rank, world), 0, Rc, p, rv, tv)
            sends.append(s_)
            counts.append(c_)
        recvs = sharded.local_exchange(sends, counts)
        torch.cuda.synchronize()
        loc = [sh.insert_begin(r_) for sh, r_ in zip(shards, recvs)]
        for sh in shards:
            sh.insert_finish(sum(a for a, _ in loc), sum(b for _, b in loc))
            sh.recut_margi()
    # a new scan, a perturbed start
    sc = seq.next_scan(deskewed=True)
    st = gpu_lib.make_state(sc.gt_R @ synth.rot_exp(np.array([2e-3, -1e-3, 1.5e-3])),
                            sc.gt_p + np.array([0.02, -0.015, 0.01]), sc.gt_v, t=sc.end_time)
    n = sc.xyzt.shape[0]
    single.set_state(st)
    single.scan_upload(sc.xyzt)
    single.var_init(0)
    sa = gpu_lib.state_arrays(st)
    rv, tv = cov_blocks(sa["cov"])
    single.iekf_begin(0, rv, tv)
    g0 = single.iekf_accumulate(col(sa["R"]), sa["p"])
    it_single, _ = single.odom_iekf(0, 4, host_solve=True)
    s_single = gpu_lib.state_arrays(single.get_state())

    iek = [sharded.ShardedIekf(sh) for sh in shards]
    for sh in shards:
        sh.ctx.set_state(st)
        sh.ctx.scan_upload(sc.xyzt)
        sh.ctx.var_init(0)
        sh.ctx.odom_iekf_host_begin(4)
    iters = 0
    for it in range(4):
        cur = gpu_lib.state_arrays(shards[0].ctx.get_state())
        Rc, p = col(cur["R"]), cur["p"]
        routed = [q.route(*sharded.slice_of(n, q.sh.rank, world), Rc, p) for q in iek]
        assert sum(int(c.sum()) for _, c in routed) == n
        recvs = sharded.local_exchange([s_ for s_, _ in routed], [c_ for _, c_ in routed])
        torch.cuda.synchronize()
        tot = np.zeros(34)
        for q, r_ in zip(iek, recvs):
            assert r_.shape[1] == sharded.QREC
            tot += q.accumulate(r_, Rc, p, rv, tv).cpu().numpy()
        if it == 0:
            # the same sums as the single-GPU accumulate kernel: count exactly, the rest to rounding
            assert int(round(tot[33])) == g0["match_num"] > 0.3 * n
            iu = np.triu_indices(6)
            assert rel_err(tot[:21], g0["HTH"][iu]) < 1e-10 and rel_err(tot[21:27], g0["HTz"]) < 1e-10
        done = [sh.ctx.odom_iekf_host_update(tot) for sh in shards]
        iters = it + 1
        assert done[0] == done[1]
        if done[0]:
            break
    assert iters == it_single
    for sh in shards:
        s_sh = gpu_lib.state_arrays(sh.ctx.get_state())
        for f in ("R", "p", "v", "bg", "ba"):
            # (sums in another order, and a point within float rounding of a voxel face keeps its cached leaf on one
            # GPU but is looked up by key here: nanometres)
            assert np.max(np.abs(s_sh[f] - s_single[f])) < 5e-9, f
        assert rel_err(s_sh["cov"], s_single["cov"]) < 1e-9
    assert np.linalg.norm(s_single["p"] - sc.gt_p) < 5e-3
    single.close()
    for sh in shards:
        sh.ctx.close()


def test_small_scan_retry_and_new_entry_points_edges(oracle_lib, gpu_lib):
    """(i) a scan that down-samples to < 2000 points takes the "down_size / 2" retry of local_mapping.cpp:396-403
    inside the full step, like the oracle; (ii) the device IEKF loop on a map without any plane leaves the state
    where the host variant leaves it; (iii) argument / state errors of the batch, shard and host-IEKF entry points."""
    import ctypes as C

    import torch

    # (i) ---------------------------------------------------------------------------------------------
    cfg = small_cfg("robosense128", 8, 150)  # 1200 points per scan: always below the 2000-point threshold
    seq, od, gx, last = bootstrap_pair(oracle_lib, gpu_lib, cfg, gpu_own_downsample=True)
    for k in range(4):
        sc = seq.next_scan()
        r, _ = od.step(sc.xyzt, sc.beg_time, sc.imu, iekf_on_full=True, max_iter=4)
        assert r == 0
        sg = gpu_lib.state_arrays(gx.step(sc.xyzt, sc.beg_time, sc.imu, iekf_on_full=True, max_iter=4))
        so = oracle_lib.state_arrays(od.get_state())
        n_d = gx.n_down()
        assert 0 < n_d < 2000 and n_d == od.last_down().shape[0]  # same voxel set after the retry
        assert np.linalg.norm(sg["p"] - so["p"]) < 1e-3 and synth.rot_err_deg(sg["R"], so["R"]) < 0.01
    assert gx.map_count()[0] == od.map_count()[0]
    gx.close()

    # (ii) --------------------------------------------------------------------------------------------
    cfg = small_cfg()
    gx = gpu_lib.Ctx(cfg, **SMALL_CAPS)
    seq = synth.Sequence(cfg)
    sc = seq.next_scan(deskewed=True)
    st = gpu_lib.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time)
    gx.scan_upload(sc.xyzt)
    gx.var_init(0)
    res = []
    for host in (False, True):
        gx.set_state(st)
        it, ok = gx.odom_iekf(0, 4, host_solve=host)
        res.append((it, ok, gpu_lib.state_arrays(gx.get_state())))
    assert res[0][0] == res[1][0] == 2 and res[0][1] == res[1][1] == 0  # no matches: converged at once, degenerate
    for f in ("R", "p", "v", "bg", "ba"):
        assert np.array_equal(res[0][2][f], gpu_lib.state_arrays(st)[f]) and np.array_equal(res[1][2][f], res[0][2][f])
    assert rel_err(res[0][2]["cov"], res[1][2]["cov"]) < 1e-9

    # (iii) -------------------------------------------------------------------------------------------
    lib = gx.lib
    h = C.c_void_p()
    arr = (C.c_void_p * 1)(gx.h)
    assert lib.vina_batch_create(arr, C.c_int(0), C.byref(h)) == -1
    assert lib.vina_batch_create(arr, C.c_int(17), C.byref(h)) == -1
    assert lib.vina_batch_create(None, C.c_int(1), C.byref(h)) == -1
    one = gpu_lib.Batch([gx])  # a batch of one is legal
    one.close()
    cnt = (C.c_int32 * 16)()
    z9, z3 = np.zeros(9), np.zeros(3)
    dp = gpu_lib._dp
    buf = torch.empty((sc.xyzt.shape[0], 13), dtype=torch.float64, device="cuda")
    assert lib.vina_shard_route(gx.h, C.c_int(0), C.c_int(0), C.c_int(1), C.c_int64(0), dp(z9), dp(z3), dp(z9), dp(z9),
                                C.c_void_p(buf.data_ptr()), cnt) == -1      # world < 1
    assert lib.vina_shard_route(gx.h, C.c_int(17), C.c_int(0), C.c_int(1), C.c_int64(0), dp(z9), dp(z3), dp(z9), dp(z9),
                                C.c_void_p(buf.data_ptr()), cnt) == -1      # world > VINA_MAX_WORLD
    assert lib.vina_shard_route(gx.h, C.c_int(2), C.c_int(0), C.c_int(10 ** 6), C.c_int64(0), dp(z9), dp(z3), dp(z9),
                                dp(z9), C.c_void_p(buf.data_ptr()), cnt) == -1  # slice beyond the point set
    assert b"exceeds" in lib.vina_last_error(gx.h)
    # an empty slice routes nothing; world = 1 routes everything to rank 0, in order
    gx.down_upload(sc.xyzt[:500])
    gx.var_init(1)
    c0 = gx.shard_route(4, 0, 0, 0, col(np.eye(3)), z3, z9, z9, buf.data_ptr())
    assert c0.tolist() == [0, 0, 0, 0]
    c1 = gx.shard_route(1, 0, 500, 7, col(np.eye(3)), z3, z9, z9, buf.data_ptr())
    assert c1.tolist() == [500]
    assert np.array_equal(buf[:500, 12].contiguous().view(torch.int64).cpu().numpy(), 7 + np.arange(500))
    # host IEKF update without begin
    assert lib.vina_odom_iekf_host_update(gx.h, dp(np.zeros(34))) == -6
    gx.odom_iekf_host_begin(4)
    assert gx.odom_iekf_host_update(np.zeros(34)) is False  # zero sums: no step, converged once
    assert gx.odom_iekf_host_update(np.zeros(34)) is True   # ... twice: finished
    assert lib.vina_odom_iekf_host_update(gx.h, dp(np.zeros(34))) == -6
    gx.close()


def test_sharded_map_p2p_exchange_equals_single_gpu_map(oracle_lib, gpu_lib):
    """The exchange fused into the routing kernel: three shard contexts on this GPU connected through each
    other's inbox / control-block pointers (vina_shard_p2p_connect_local; across processes the same pointers come
    from CUDA IPC). Records are stored by the routing kernel at their final position in the owner's inbox; counts
    and completion flags travel the same way; nothing synchronises with the host between routing and arrival.
    The union of the shards must again equal the single-context map byte for byte."""
    from vina_slam_b200 import sharded

    world = 3
    cfg = small_cfg("robosense128", 32, 600)
    seq = synth.Sequence(cfg)
    od = oracle_lib.Odom(cfg)
    single = gpu_lib.Ctx(cfg, **SMALL_CAPS)
    shards = [sharded.MapShard(gpu_lib.Ctx(cfg, **SMALL_CAPS), r, world, own_stream=True) for r in range(world)]
    for sh in shards:
        sh.ctx.shard_p2p_create(sh.rank, world, SMALL_CAPS["max_scan_points"])
    ptrs = [sh.ctx.shard_p2p_pointers() for sh in shards]
    for sh in shards:
        sh.ctx.shard_p2p_connect_local([a for a, _ in ptrs], [b for _, b in ptrs])
    received = np.zeros(world, dtype=np.int64)
    for k in range(cfg.win_size + 4):
        sc = seq.next_scan(deskewed=True)
        od.bootstrap(sc.xyzt, oracle_lib.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
        down = od.last_down()
        st = gpu_lib.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time)
        single.set_state(st)
        single.down_upload(down)
        single.var_init(1)
        single.odom_map_update()
        sa = gpu_lib.state_arrays(st)
        Rc, p = col(sa["R"]), sa["p"]
        rv, tv = cov_blocks(sa["cov"])
        for sh in shards:
            sh.ctx.down_upload(down)
            sh.ctx.var_init(1)
            sh.push_pose(Rc, p)
        # several ranks on one GPU, one host thread: first everything that never waits, then the waiting parts
        # (CUDA may put two streams into one hardware queue; a kernel must not wait for work enqueued behind it)
        for phase in (1, 2):
            for sh in shards:
                sh.ctx.shard_route_p2p(*sharded.slice_of(down.shape[0], sh.rank, world), 0, Rc, p, rv, tv, phase)
        loc = [sh.ctx.shard_insert_begin_p2p(sh.win_count - 1) for sh in shards]
        assert sum(n for n, _, _ in loc) == down.shape[0]
        for sh, (n, _, _) in zip(shards, loc):
            received[sh.rank] += n
            sh.insert_finish(sum(a for _, a, _ in loc), sum(b for _, _, b in loc))
            sh.recut_margi()
    ms = sort_nodes(single.map_export())
    mu = sort_nodes(np.concatenate([sh.ctx.map_export() for sh in shards]))
    assert mu.shape[0] == ms.shape[0] > 2000
    for f in ms.dtype.names:
        assert np.array_equal(ms[f], mu[f]), f"p2p-sharded map differs from the single-GPU map in {f}"
    assert received.min() > 0.5 * received.mean()
    single.close()
    for sh in shards:
        sh.ctx.close()


def test_sharded_iekf_fused_exchange_matches_single_gpu(oracle_lib, gpu_lib):
    """The IEKF against the sharded map with everything on the device (vina_odom_iekf_sharded_p2p): queries stored
    into the owners' inboxes by the routing kernel, the shards' 34 sums stored into every peer's control block,
    rank-ordered total and update on every rank's own device iterate. Three ranks on this GPU (connect_local),
    driven phase by phase from one host thread. Against the single-GPU loop with the host update: same iteration
    count, every rank ends with bitwise the same state, and that state equals the single-GPU one to 1e-9
    (device 6x6 push-through solve vs the reference's 15x15 route, no leaf cache across ranks)."""
    import torch

    from vina_slam_b200 import sharded

    world = 3
    cfg = small_cfg("robosense128", 32, 600)
    seq = synth.Sequence(cfg)
    od = oracle_lib.Odom(cfg)
    single = gpu_lib.Ctx(cfg, **SMALL_CAPS)
    shards = [sharded.MapShard(gpu_lib.Ctx(cfg, **SMALL_CAPS), r, world, own_stream=True) for r in range(world)]
    for sh in shards:
        sh.ctx.shard_p2p_create(sh.rank, world, SMALL_CAPS["max_scan_points"])
    ptrs = [sh.ctx.shard_p2p_pointers() for sh in shards]
    for sh in shards:
        sh.ctx.shard_p2p_connect_local([a for a, _ in ptrs], [b for _, b in ptrs])
    for k in range(cfg.win_size):
        sc = seq.next_scan(deskewed=True)
        od.bootstrap(sc.xyzt, oracle_lib.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
        down = od.last_down()
        st = gpu_lib.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time)
        single.set_state(st)
        single.down_upload(down)
        single.var_init(1)
        single.odom_map_update()
        sa = gpu_lib.state_arrays(st)
        Rc, p = col(sa["R"]), sa["p"]
        rv, tv = cov_blocks(sa["cov"])
        for sh in shards:
            sh.ctx.down_upload(down)
            sh.ctx.var_init(1)
            sh.push_pose(Rc, p)
        for phase in (1, 2):
            for sh in shards:
                sh.ctx.shard_route_p2p(*sharded.slice_of(down.shape[0], sh.rank, world), 0, Rc, p, rv, tv, phase)
        loc = [sh.ctx.shard_insert_begin_p2p(sh.win_count - 1) for sh in shards]
        for sh in shards:
            sh.insert_finish(sum(a for _, a, _ in loc), sum(b for _, _, b in loc))
            sh.recut_margi()
    for trial, (dth, dp) in enumerate([((2e-3, -1e-3, 1.5e-3), (0.02, -0.015, 0.01)), ((0, 0, 0), (0, 0, 0))]):
        sc = seq.next_scan(deskewed=True)
        st = gpu_lib.make_state(sc.gt_R @ synth.rot_exp(np.array(dth)), sc.gt_p + np.array(dp), sc.gt_v, t=sc.end_time)
        n = sc.xyzt.shape[0]
        single.set_state(st)
        single.scan_upload(sc.xyzt)
        single.var_init(0)
        it_single, _ = single.odom_iekf(0, 4, host_solve=True)
        s_single = gpu_lib.state_arrays(single.get_state())
        for sh in shards:
            sh.ctx.set_state(st)
            sh.ctx.scan_upload(sc.xyzt)
            sh.ctx.var_init(0)
        slices = [sharded.slice_of(n, sh.rank, world) for sh in shards]
        for sh, (f, c) in zip(shards, slices):
            sh.ctx.odom_iekf_sharded_p2p(f, c, 4, gpu_lib.SHARD_IEKF_STAGE)
        for it in range(4):
            for phase in (gpu_lib.SHARD_IEKF_ROUTE, gpu_lib.SHARD_IEKF_SEND, gpu_lib.SHARD_IEKF_EVAL, gpu_lib.SHARD_IEKF_SOLVE):
                for sh, (f, c) in zip(shards, slices):
                    sh.ctx.odom_iekf_sharded_p2p(f, c, 4, phase)
        res = [sh.ctx.odom_iekf_sharded_p2p(f, c, 4, gpu_lib.SHARD_IEKF_FINISH) for sh, (f, c) in zip(shards, slices)]
        assert all(r[0] == it_single for r in res), (res, it_single)
        states = [gpu_lib.state_arrays(sh.ctx.get_state()) for sh in shards]
        # the same sharded loop with the host in it (staged records, in-process permutation, host update)
        iek = [sharded.ShardedIekf(sh) for sh in shards]
        sa = gpu_lib.state_arrays(st)
        rv, tv = cov_blocks(sa["cov"])
        for sh in shards:
            sh.ctx.set_state(st)
            sh.ctx.odom_iekf_host_begin(4)
        it_host = 0
        for it in range(4):
            cur = gpu_lib.state_arrays(shards[0].ctx.get_state())
            Rc, p = col(cur["R"]), cur["p"]
            routed = [q.route(f, c, Rc, p) for q, (f, c) in zip(iek, slices)]
            recvs = sharded.local_exchange([s_ for s_, _ in routed], [c_ for _, c_ in routed])
            torch.cuda.synchronize()
            tot = np.zeros(34)
            for q, r_ in zip(iek, recvs):
                sums = q.accumulate(r_, Rc, p, rv, tv)
                q.sh.ctx.sync()  # the contexts run on private streams here: torch's copy below does not wait for them
                tot += sums.cpu().numpy()
            done = [sh.ctx.odom_iekf_host_update(tot) for sh in shards]
            it_host = it + 1
            if done[0]:
                break
        s_host = gpu_lib.state_arrays(shards[0].ctx.get_state())
        assert it_host == res[0][0]
        for s_sh in states:
            for f in ("R", "p", "v", "bg", "ba", "cov"):
                assert np.array_equal(s_sh[f], states[0][f]), f"ranks diverged in {f}"
                tol = 1e-9 if f != "cov" else 1e-7 * np.max(np.abs(s_single["cov"]))
                assert np.max(np.abs(s_sh[f] - s_host[f])) < tol, (trial, f)
                # vs one GPU: the per-point leaf cache (odometry.cpp:124-127) does not cross GPUs, a point within
                # float rounding of a voxel face is looked up by key instead
                assert np.max(np.abs(s_sh[f] - s_single[f])) < 1000 * tol, (trial, f)
        assert np.linalg.norm(s_single["p"] - sc.gt_p) < 5e-3
        # (when the loop converges early the remaining enqueued iterations are no-ops on every rank)
    single.close()
    for sh in shards:
        sh.ctx.close()


def test_overlapped_step_is_bitwise_the_serial_step(oracle_lib, gpu_lib):
    """vina_set_overlap: down-sampling on the side stream and the map update enqueued behind the IEKF loop with the
    pose read from the device iterate must give bitwise the states and the map of the serial schedule (host in
    between) - over enough scans for the window to slide, leaves to split and the early scans to be marginalised."""
    cfg = small_cfg("robosense128", 32, 600)
    seq = synth.Sequence(cfg)
    ctxs = [gpu_lib.Ctx(cfg, **SMALL_CAPS) for _ in range(2)]
    ctxs[0].set_overlap(True)
    ctxs[1].set_overlap(False)
    for k in range(cfg.win_size):
        sc = seq.next_scan(deskewed=True)
        for gx in ctxs:
            gx.bootstrap(sc.xyzt, gpu_lib.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
    for gx in ctxs:
        gx.set_imu_anchor(sc.end_time, sc.imu[-1])
    for k in range(25):
        sc = seq.next_scan()
        sa, sb = [gpu_lib.state_arrays(gx.step(sc.xyzt, sc.beg_time, sc.imu, iekf_on_full=True, max_iter=4)) for gx in ctxs]
        for f in ("R", "p", "v", "bg", "ba", "cov"):
            assert np.array_equal(sa[f], sb[f]), (k, f)
        assert np.linalg.norm(sa["p"] - sc.gt_p) < 0.02
    for gx in ctxs:
        gx.sync()
    ma, mb = [sort_nodes(gx.map_export()) for gx in ctxs]
    assert ma.shape[0] == mb.shape[0] > 2000
    for f in ma.dtype.names:
        assert np.array_equal(ma[f], mb[f]), f
    assert ctxs[0].window()[0] == ctxs[1].window()[0]
    for gx in ctxs:
        gx.close()


def test_iekf_loop_kernel_matches_per_iteration_launches_and_the_oracle(oracle_lib, gpu_lib):
    """vina_set_iekf_loop: the whole iteration loop of LioStateEstimation (odometry.cpp:98-231) as one persistent
    cooperative launch (k_iekf_loop: scan in shared memory through 1-D TMA bulk copies, grid barrier + the update
    between iterations, every block solving redundantly) with the two fused front launches, against the default
    schedule (one k_iekf launch per iteration) and against the oracle: same iteration count on every scan, states within
    1e-9 of each other (the sums are added in another fixed order) and within 1 mm / 0.01 deg of the oracle, the same
    number of map nodes - over enough scans for the window to slide and leaves to split. The overlapped and the serial
    step of the loop schedule must again be bitwise identical."""
    cfg = small_cfg("robosense128", 32, 600)
    seq = synth.Sequence(cfg)
    od = oracle_lib.Odom(cfg)
    ctxs = [gpu_lib.Ctx(cfg, **SMALL_CAPS) for _ in range(3)]
    ctxs[0].set_iekf_loop(True)
    ctxs[1].set_iekf_loop(False)
    ctxs[2].set_iekf_loop(True)
    ctxs[2].set_overlap(False)
    for k in range(cfg.win_size):
        sc = seq.next_scan(deskewed=True)
        od.bootstrap(sc.xyzt, oracle_lib.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
        for gx in ctxs:
            gx.bootstrap(sc.xyzt, gpu_lib.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
    od.set_imu_anchor(sc.end_time, sc.imu[-1])
    for gx in ctxs:
        gx.set_imu_anchor(sc.end_time, sc.imu[-1])
    for k in range(16):
        sc = seq.next_scan()
        r, _ = od.step(sc.xyzt, sc.beg_time, sc.imu, iekf_on_full=True, max_iter=4)
        assert r == 0
        so = oracle_lib.state_arrays(od.get_state())
        sa, sb, scx = [gpu_lib.state_arrays(gx.step(sc.xyzt, sc.beg_time, sc.imu, iekf_on_full=True, max_iter=4)) for gx in ctxs]
        assert ctxs[0].timings().iekf_iters == ctxs[1].timings().iekf_iters == od.last_iters(), k
        for f in ("R", "p", "v", "bg", "ba"):
            assert np.max(np.abs(sa[f] - sb[f])) < 1e-9, (k, f)
            assert np.array_equal(sa[f], scx[f]), (k, f)
        assert np.max(np.abs(sa["cov"] - sb["cov"])) < 1e-9 * max(1.0, float(np.max(np.abs(sb["cov"]))))
        assert np.array_equal(sa["cov"], scx["cov"]), k
        assert np.linalg.norm(sa["p"] - so["p"]) < 1e-3
        assert synth.rot_err_deg(sa["R"], so["R"]) < 0.01
    for gx in ctxs:
        gx.sync()
    na, nb, nc = [gx.map_count()[0] for gx in ctxs]
    assert na == nb == nc > 2000
    ma, mc = sort_nodes(ctxs[0].map_export()), sort_nodes(ctxs[2].map_export())
    for f in ma.dtype.names:
        assert np.array_equal(ma[f], mc[f]), f
    for gx in ctxs:
        gx.close()


def test_host_buffer_step_is_bitwise_the_resident_step(gpu_lib):
    """vina_odom_step uploads the scan in two chunks on the copy stream and lets the fused deskew follow them
    (vn_scan_upload_chunked; with vina_set_upload_ordered the copy starts behind the work already on the stream). Same
    states and the same map, bit for bit, as vina_odom_step_resident on a scan that is already in HBM - pinned and
    pageable host buffers, scans above and below the chunking threshold."""
    import torch

    for beams, steps in ((64, 700), (16, 400)):  # 44 800 points (two chunks) / 6 400 points (one)
        cfg = small_cfg("robosense128", beams, steps)
        seq = synth.Sequence(cfg)
        caps = dict(SMALL_CAPS, max_scan_points=max(SMALL_CAPS.get("max_scan_points", 0), beams * steps + 64))
        ctxs = [gpu_lib.Ctx(cfg, **caps) for _ in range(3)]
        ctxs[1].set_upload_ordered(True)
        for k in range(cfg.win_size):
            sc = seq.next_scan(deskewed=True)
            for gx in ctxs:
                gx.bootstrap(sc.xyzt, gpu_lib.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
        for gx in ctxs:
            gx.set_imu_anchor(sc.end_time, sc.imu[-1])
        for k in range(12):
            sc = seq.next_scan()
            pinned = torch.from_numpy(sc.xyzt).pin_memory()
            dev = torch.from_numpy(sc.xyzt).cuda()
            sa = gpu_lib.state_arrays(ctxs[0].step(sc.xyzt, sc.beg_time, sc.imu, iekf_on_full=True, max_iter=4))
            sb = gpu_lib.state_arrays(ctxs[1].step(pinned.numpy(), sc.beg_time, sc.imu, iekf_on_full=True, max_iter=4))
            end_time = sc.beg_time + float(sc.xyzt[-1, 3])
            scx = gpu_lib.state_arrays(ctxs[2].step_resident(dev.data_ptr(), sc.xyzt.shape[0], sc.beg_time, end_time, sc.imu,
                                                              True, 4))
            for f in ("R", "p", "v", "bg", "ba", "cov"):
                assert np.array_equal(sa[f], scx[f]) and np.array_equal(sb[f], scx[f]), (beams, k, f)
            assert np.linalg.norm(sa["p"] - sc.gt_p) < 0.02
        for gx in ctxs:
            gx.sync()
        ma, mb, mc = [sort_nodes(gx.map_export()) for gx in ctxs]
        for f in ma.dtype.names:
            assert np.array_equal(ma[f], mc[f]) and np.array_equal(mb[f], mc[f]), f
        for gx in ctxs:
            gx.close()


def test_iekf_loop_schedule_small_scan_retry(oracle_lib, gpu_lib):
    """The fused front of the loop schedule publishes the down-sampled count through mapped memory; a scan that
    down-samples to < 2000 points must still take the "down_size / 2" retry of local_mapping.cpp:396-403 (separate
    kernels, then var_init again) and land on the oracle's voxel set and trajectory."""
    cfg = small_cfg("robosense128", 8, 150)  # 1200 points per scan: always below the 2000-point threshold
    seq, od, gx, last = bootstrap_pair(oracle_lib, gpu_lib, cfg, gpu_own_downsample=True)
    gx.set_iekf_loop(True)
    for k in range(4):
        sc = seq.next_scan()
        r, _ = od.step(sc.xyzt, sc.beg_time, sc.imu, iekf_on_full=True, max_iter=4)
        assert r == 0
        sg = gpu_lib.state_arrays(gx.step(sc.xyzt, sc.beg_time, sc.imu, iekf_on_full=True, max_iter=4))
        so = oracle_lib.state_arrays(od.get_state())
        n_d = gx.n_down()
        assert 0 < n_d < 2000 and n_d == od.last_down().shape[0]
        assert np.linalg.norm(sg["p"] - so["p"]) < 1e-3 and synth.rot_err_deg(sg["R"], so["R"]) < 0.01
    assert gx.map_count()[0] == od.map_count()[0]
    gx.close()


def test_ba_lidar_factor_matches_oracle(oracle_lib, gpu_lib):
    """SURVEY section 8f rank 3, the data-parallel part of the sliding-window BA: the device factor store
    (tras_opt) and LidarFactor::acc_evaluate2 / evaluate_only_residual (factors.cpp:22-158) against the oracle,
    which reproduces the reference's own factors.cpp bit for bit (tests/test_oracle_vs_ref.py). Same factor set;
    Hessian / gradient / residual to 1e-9 of the largest entry (the sum over ~10^3 factors runs in another order);
    every factor's lambda_0 after evaluate_only_residual bit for bit; the overwrite of the stored eig / pcr_add by
    evaluate_only_residual is visible in the next Hessian exactly like in the reference's container."""
    import importlib.util
    import os

    spec = importlib.util.spec_from_file_location(
        "make_ref_golden", os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_ref_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)

    cfg = synth.small_sensor("robosense128", 24, 400, seed=31)  # the scenario of tests/golden/ref_ba.npz
    seq = synth.Sequence(cfg)
    od = oracle_lib.Odom(cfg)
    gx = gpu_lib.Ctx(cfg, **SMALL_CAPS)
    od.ba_probe(True)
    gx.ba_set_capture(True)
    for _ in range(cfg.win_size + 3):
        sc = seq.next_scan(deskewed=True)
        od.bootstrap(sc.xyzt, oracle_lib.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
        gx.set_state(gpu_lib.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
        gx.down_upload(od.last_down())
        gx.var_init(1)
        gx.odom_map_update()
    n = od.ba_count()
    assert gx.ba_count() == n > 1000
    poses = od.ba_poses()
    pert = mod.perturbed_poses(poses)

    def close(a, b, tol=1e-9):
        return np.max(np.abs(np.asarray(a) - np.asarray(b))) <= tol * max(np.max(np.abs(np.asarray(b))), 1e-300)

    for ps in (poses, pert):
        Ho, Jo, ro = od.ba_hess(ps)
        Hg, Jg, rg = gx.ba_hess(ps)
        assert close(Hg, Ho) and close(Jg, Jo) and abs(rg - ro) <= 1e-12 * abs(ro), (np.max(np.abs(Hg - Ho)), np.max(np.abs(Ho)))
        assert np.array_equal(Hg[6:12, 0:6], Hg[0:6, 6:12].T)  # lower blocks mirrored (factors.cpp:123-125)
    ro, lo = od.ba_residual(pert)
    rg, lg = gx.ba_residual(pert)
    assert abs(rg - ro) <= 1e-12 * abs(ro) and ro > 0.2  # the perturbed poses are visibly worse (0.10 at the captured ones)
    assert np.array_equal(np.sort(lg), np.sort(lo)), "per-factor eigenvalues differ"
    # the stored factors now carry the perturbed eig / pcr_add on both sides
    Ho, Jo, ro2 = od.ba_hess(pert)
    Hg, Jg, rg2 = gx.ba_hess(pert)
    assert close(Hg, Ho) and close(Jg, Jo) and abs(rg2 - ro2) <= 1e-12 * abs(ro2)
    assert abs(ro2 - ro) <= 1e-12 * ro  # acc_evaluate2 reports the stored lambda_0 sum
    # and against the reference's golden vectors directly
    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_ba.npz")))
    assert int(g["n_factors"][0]) == n and close(Hg, g["H2"]) and close(Jg, g["J2"])
    gx.close()
    od.close()


@pytest.mark.parametrize("base,beams,steps", [("robosense128", 32, 600), ("velodyne32", 32, 500)])
def test_ba_odometry_matches_oracle(oracle_lib, gpu_lib, base, beams, steps):
    """The per-scan loop with LocalBA.if_BA: 1 (mid360.yaml / velodyne.yaml; local_mapping.cpp:437-441, 492-497,
    541-546) through vina_odom_step: IMU pre-integration factors and the LM loop on the host, the LiDAR factor on the
    device, margi taking the re-evaluated factors back - against the oracle, whose BA reproduces the reference's own
    imu_preintegration.cpp / optimizers.cpp / factors.cpp bit for bit (tests/test_oracle_vs_ref.py). Same number of
    BA runs and LM iterations, trajectory within the north-star's 1 mm / 0.01 deg at every scan, and BA visibly
    changes the result (it is not a no-op). velodyne32: 3 octree layers, non-identity extrinsic (its yaml has if_BA: 1)."""
    cfg = small_cfg(base, beams, steps)
    seq = synth.Sequence(cfg)
    od = oracle_lib.Odom(cfg)
    gx = gpu_lib.Ctx(cfg, **SMALL_CAPS)
    g0 = gpu_lib.Ctx(cfg, **SMALL_CAPS)  # the same sequence without BA
    od.set_ba(True)
    gx.set_ba(True)
    for _ in range(cfg.win_size):
        sc = seq.next_scan(deskewed=True)
        od.bootstrap(sc.xyzt, oracle_lib.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
        for g in (gx, g0):
            g.bootstrap(sc.xyzt, gpu_lib.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
    od.set_imu_anchor(sc.end_time, sc.imu[-1])
    for g in (gx, g0):
        g.set_imu_anchor(sc.end_time, sc.imu[-1])
    worst_p, worst_r, moved = 0.0, 0.0, 0.0
    for k in range(14):
        sc = seq.next_scan()
        imu = sc.imu.copy()
        imu[:, 0] = np.round(imu[:, 0] * 1e9) * 1e-9  # the reference keeps IMU stamps as integer nanoseconds
        r, _ = od.step(sc.xyzt, sc.beg_time, imu, iekf_on_full=True, max_iter=4)
        assert r == 0
        sg = gpu_lib.state_arrays(gx.step(sc.xyzt, sc.beg_time, imu, iekf_on_full=True, max_iter=4))
        s0 = gpu_lib.state_arrays(g0.step(sc.xyzt, sc.beg_time, imu, iekf_on_full=True, max_iter=4))
        so = oracle_lib.state_arrays(od.get_state())
        worst_p = max(worst_p, float(np.linalg.norm(sg["p"] - so["p"])))
        worst_r = max(worst_r, synth.rot_err_deg(sg["R"], so["R"]))
        moved = max(moved, float(np.linalg.norm(sg["p"] - s0["p"])))
        assert np.linalg.norm(sg["p"] - sc.gt_p) < 0.02
        assert gx.ba_stats() == od.ba_stats(), (k, gx.ba_stats(), od.ba_stats())
    gx.sync()
    assert od.ba_stats()[0] >= 5
    assert worst_p < 1e-3 and worst_r < 0.01, (worst_p, worst_r)
    assert moved > 1e-6, "BA changed nothing"
    ng, no = gx.map_count(), od.map_count()
    assert ng[0] == no[0] and ng[2] == no[2], (ng, no)
    for g in (gx, g0):
        g.close()
    od.close()


def test_ba_and_fused_loop_entry_points_edges(oracle_lib, gpu_lib):
    """Error behaviour of the newer entry points: codes, never exit() / exceptions across the ABI; a BA over an empty
    factor store (multi_recut's early-out leaves voxhess empty) gives zero Hessian / gradient / residual."""
    import ctypes as C

    cfg = small_cfg()
    gx = gpu_lib.Ctx(cfg, **SMALL_CAPS)
    lib, dp = gx.lib, gpu_lib._dp
    w = cfg.win_size
    poses = np.zeros((w, 12))
    poses[:, 0] = poses[:, 4] = poses[:, 8] = 1.0
    H, J, r, n = np.zeros(36 * w * w), np.zeros(6 * w), C.c_double(0), C.c_int32(0)
    vp = poses.ctypes.data_as(C.c_void_p)
    # nothing collected yet
    assert lib.vina_ba_lidar_hessian(gx.h, vp, C.c_int(w), dp(H), dp(J), C.byref(r)) == -6
    assert lib.vina_ba_lidar_residual(gx.h, vp, C.c_int(w), C.byref(r), None, C.c_int(0)) == -6
    assert lib.vina_ba_count(gx.h, C.byref(n)) == -6
    assert b"no BA factors" in lib.vina_last_error(gx.h)
    # bad arguments
    assert lib.vina_ba_lidar_hessian(gx.h, None, C.c_int(w), dp(H), dp(J), C.byref(r)) == -1
    assert lib.vina_ba_lidar_hessian(gx.h, vp, C.c_int(11), dp(H), dp(J), C.byref(r)) == -1
    assert lib.vina_ba_collect(None, C.byref(n)) == -1
    assert lib.vina_odom_set_ba(None, C.c_int(1), C.c_double(0)) == -1
    assert lib.vina_set_overlap(None, C.c_int(1)) == -1
    # an empty map: collect finds nothing, the evaluations are all zero
    assert gx.ba_collect() == 0
    Hh, Jj, rr = gx.ba_hess(poses)
    assert not Hh.any() and not Jj.any() and rr == 0.0
    rr, lam = gx.ba_residual(poses)
    assert rr == 0.0 and lam.shape == (0,)
    assert lib.vina_ba_lidar_hessian(gx.h, vp, C.c_int(w - 1), dp(H), dp(J), C.byref(r)) == -1  # win != LocalBA.win_size
    # if_BA can only be chosen before frames enter the window
    seq = synth.Sequence(cfg)
    sc = seq.next_scan(deskewed=True)
    gx.bootstrap(sc.xyzt, gpu_lib.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
    assert lib.vina_odom_set_ba(gx.h, C.c_int(1), C.c_double(0)) == -6
    assert lib.vina_odom_set_ba(gx.h, C.c_int(0), C.c_double(1e-3)) == 0  # unchanged mode: fine
    assert gx.ba_stats() == (0, 0)
    # the fused sharded IEKF needs connected peers and a valid slice
    it, ok = C.c_int(0), C.c_int(0)
    assert lib.vina_odom_iekf_sharded_p2p(gx.h, C.c_int(0), C.c_int(10), C.c_int(4), C.c_int(0), C.byref(it), C.byref(ok)) == -6
    assert lib.vina_odom_iekf_sharded_p2p(gx.h, C.c_int(0), C.c_int(10), C.c_int(4), C.c_int(7), C.byref(it), C.byref(ok)) == -1
    gx.shard_p2p_create(0, 1, SMALL_CAPS["max_scan_points"])  # a world of one is connected to itself
    gx.scan_upload(sc.xyzt)
    gx.var_init(0)
    assert lib.vina_odom_iekf_sharded_p2p(gx.h, C.c_int(0), C.c_int(10 ** 7), C.c_int(4), C.c_int(0), C.byref(it), C.byref(ok)) == -1
    # world = 1 through the fused loop = the single-GPU loop without a leaf cache (one frame in the map: no planes
    # yet, so nothing matches and the state stays where it was)
    st = gpu_lib.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time)
    gx.set_state(st)
    iters, _ = gx.odom_iekf_sharded_p2p(0, sc.xyzt.shape[0], 4)
    assert iters == 2
    assert np.array_equal(gpu_lib.state_arrays(gx.get_state())["p"], gpu_lib.state_arrays(st)["p"])
    gx.close()


def test_ba_full_size_mid360_properties(gpu_lib):
    """BASELINE configs[0] at full size (Mid-360 shape, 20 000 points per scan, mid360.yaml has if_BA: 1), too
    large for the oracle to follow in the GPU suite: size-independent properties of the BA instead. The LiDAR
    Hessian has mirrored off-diagonal blocks and is finite, the per-factor eigenvalues are non-negative, a
    perturbation of the window poses raises the residual, BA runs every scan once the window has IMU factors,
    and the trajectory stays on the synthetic ground truth."""
    cfg = synth.SENSORS["mid360"]
    caps = dict(max_scan_points=32768, max_nodes=1 << 18, hash_capacity_log2=19)
    seq = synth.Sequence(cfg)
    gx = gpu_lib.Ctx(cfg, **caps)
    gx.set_ba(True)
    gx.ba_set_capture(True)
    for _ in range(cfg.win_size):
        sc = seq.next_scan(deskewed=True)
        gx.bootstrap(sc.xyzt, gpu_lib.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
    gx.set_imu_anchor(sc.end_time, sc.imu[-1])
    poses_gt = []
    for k in range(13):
        sc = seq.next_scan()
        st = gpu_lib.state_arrays(gx.step(sc.xyzt, sc.beg_time, sc.imu, iekf_on_full=True, max_iter=4))
        assert np.linalg.norm(st["p"] - sc.gt_p) < 0.02, k
        poses_gt.append(np.concatenate([sc.gt_R.T.reshape(-1), sc.gt_p]))
    gx.sync()
    runs, iters = gx.ba_stats()
    assert runs == 13 - 8 and 1 <= iters <= 10
    n = gx.ba_count()
    assert n > 500
    w = cfg.win_size
    ps = np.array(poses_gt[-w:])  # the window the last BA worked on (ground-truth poses)
    H, J, r0 = gx.ba_hess(ps)
    for i in range(w):
        for j in range(i + 1, w):
            assert np.array_equal(H[6 * i:6 * i + 6, 6 * j:6 * j + 6], H[6 * j:6 * j + 6, 6 * i:6 * i + 6].T)
    assert np.isfinite(H).all() and np.isfinite(J).all() and np.linalg.eigvalsh(0.5 * (H + H.T))[-1] > 0
    res_gt, lam = gx.ba_residual(ps)
    assert lam.shape == (n,) and (lam >= -1e-12).all()
    rng = np.random.default_rng(3)
    pert = ps.copy()
    pert[:, 9:] += rng.normal(0, 0.03, (w, 3))
    res_pert, _ = gx.ba_residual(pert)
    assert res_pert > 1.5 * res_gt, (res_gt, res_pert)
    gx.close()


def test_long_run_stays_on_the_oracle_trajectory(oracle_lib, gpu_lib):
    """150 scans through the full per-scan path (the window slides 150 times, leaves saturate, point_fix lists
    are dropped and re-created, the slide map turns over): trajectory within 1 mm / 0.01 deg of the oracle at
    every scan, same number of octree nodes at the end, pools within their capacity (no error raised)."""
    cfg = small_cfg("hilti_xt32", 16, 300)  # the handheld path stays in free space for the whole run
    seq, od, gx, last = bootstrap_pair(oracle_lib, gpu_lib, cfg, gpu_own_downsample=True)
    worst_p, worst_r, worst_gt = 0.0, 0.0, 0.0
    for k in range(150):
        sc = seq.next_scan()
        r, _ = od.step(sc.xyzt, sc.beg_time, sc.imu, iekf_on_full=True, max_iter=4)
        assert r == 0
        sg = gpu_lib.state_arrays(gx.step(sc.xyzt, sc.beg_time, sc.imu, iekf_on_full=True, max_iter=4))
        so = oracle_lib.state_arrays(od.get_state())
        worst_p = max(worst_p, float(np.linalg.norm(sg["p"] - so["p"])))
        worst_r = max(worst_r, synth.rot_err_deg(sg["R"], so["R"]))
        worst_gt = max(worst_gt, float(np.linalg.norm(sg["p"] - sc.gt_p)))
    gx.sync()  # raises if any pool overflowed on the way
    assert worst_p < 1e-3 and worst_r < 0.01, (worst_p, worst_r)
    assert worst_gt < 0.03, worst_gt
    ng, no = gx.map_count(), od.map_count()
    assert ng[0] == no[0] and ng[2] == no[2], (ng, no)
    gx.close()


def test_long_run_with_ba_stays_on_the_oracle_trajectory(oracle_lib, gpu_lib):
    """60 scans with the sliding-window BA every scan (52 BA runs, the window turning over six times, factor store
    refilled every scan, IMU factors created and retired): trajectory within 1 mm / 0.01 deg of the oracle at every
    scan, same BA / LM iteration counts throughout, same number of octree nodes at the end."""
    cfg = small_cfg("hilti_xt32", 16, 300)
    seq = synth.Sequence(cfg)
    od = oracle_lib.Odom(cfg)
    gx = gpu_lib.Ctx(cfg, **SMALL_CAPS)
    od.set_ba(True)
    gx.set_ba(True)
    for _ in range(cfg.win_size):
        sc = seq.next_scan(deskewed=True)
        od.bootstrap(sc.xyzt, oracle_lib.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
        gx.bootstrap(sc.xyzt, gpu_lib.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
    od.set_imu_anchor(sc.end_time, sc.imu[-1])
    gx.set_imu_anchor(sc.end_time, sc.imu[-1])
    worst_p, worst_r, worst_gt = 0.0, 0.0, 0.0
    for k in range(60):
        sc = seq.next_scan()
        imu = sc.imu.copy()
        imu[:, 0] = np.round(imu[:, 0] * 1e9) * 1e-9
        r, _ = od.step(sc.xyzt, sc.beg_time, imu, iekf_on_full=True, max_iter=4)
        assert r == 0
        sg = gpu_lib.state_arrays(gx.step(sc.xyzt, sc.beg_time, imu, iekf_on_full=True, max_iter=4))
        so = oracle_lib.state_arrays(od.get_state())
        worst_p = max(worst_p, float(np.linalg.norm(sg["p"] - so["p"])))
        worst_r = max(worst_r, synth.rot_err_deg(sg["R"], so["R"]))
        worst_gt = max(worst_gt, float(np.linalg.norm(sg["p"] - sc.gt_p)))
        assert gx.ba_stats() == od.ba_stats(), (k, gx.ba_stats(), od.ba_stats())
    gx.sync()
    assert od.ba_stats()[0] == 60 - 8
    assert worst_p < 1e-3 and worst_r < 0.01, (worst_p, worst_r)
    assert worst_gt < 0.03, worst_gt
    ng, no = gx.map_count(), od.map_count()
    assert ng[0] == no[0] and ng[2] == no[2], (ng, no)
    gx.close()
    od.close()


def test_map_pruning_matches_oracle(oracle_lib, gpu_lib):
    """The idle path's map pruning (local_mapping.cpp:317-341, 509-519) with the 700 m horizon shrunk to 1 m so
    that a 40-scan walk exercises it: same poses on both sides, the journey bookkeeping is compared exactly, every
    pruning erases the same roots / nodes, and the WHOLE map is compared after each one and again after further
    scans have re-used the released node ids, chain blocks and the compacted fixed-point pool."""
    cfg = small_cfg("robosense128", 32, 600)
    seq, od, gx, last = bootstrap_pair(oracle_lib, gpu_lib, cfg)
    assert od.journey() == gx.journey()
    prunings, erased = 0, 0
    for k in range(40):
        sc = seq.next_scan(deskewed=True)
        od.bootstrap(sc.xyzt, oracle_lib.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
        gx.set_state(gpu_lib.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
        gx.down_upload(od.last_down())
        gx.var_init(1)
        gx.odom_map_update()
        jo, jg = od.journey(), gx.journey()
        assert jo == jg, (k, jo, jg)
        flagged = jo[1]
        ro, rg = od.idle(1), gx.idle(1)
        assert ro == rg, (k, ro, rg)
        assert od.journey() == gx.journey() and not gx.journey()[1]
        if not flagged:
            assert rg == (0, 0)
        if rg[0] > 0:
            prunings += 1
            erased += rg[0]
            co, cg = od.map_count(), gx.map_count()
            assert co == cg, (k, co, cg)
            _compare_maps(od.map_export(), gx.map_export())
    assert prunings >= 2 and erased > 200, (prunings, erased)
    _compare_maps(od.map_export(), gx.map_export())
    assert od.map_count() == gx.map_count()
    # a pruning with nothing stale leaves the map as it is; the reference's own horizon never fires on this walk
    assert gx.map_prune(gx.journey()[0], 700) == (0, 0)
    _compare_maps(od.map_export(), gx.map_export())
    gx.close()


def test_odometry_with_map_pruning_stays_on_the_oracle_trajectory(oracle_lib, gpu_lib):
    """Full per-scan path (vina_odom_step) with the idle path called after every scan, 1 m horizon: the IEKF then
    matches against a map whose hash table was rebuilt and whose nodes were recycled. Trajectory within 1 mm /
    0.01 deg of the oracle, same prunings, same node / root / slide counts at the end."""
    cfg = small_cfg("robosense128", 32, 600)
    seq, od, gx, last = bootstrap_pair(oracle_lib, gpu_lib, cfg, gpu_own_downsample=True)
    worst_p, worst_r, erased = 0.0, 0.0, 0
    for k in range(45):
        sc = seq.next_scan()
        r, _ = od.step(sc.xyzt, sc.beg_time, sc.imu, iekf_on_full=True, max_iter=4)
        assert r == 0
        sg = gpu_lib.state_arrays(gx.step(sc.xyzt, sc.beg_time, sc.imu, iekf_on_full=True, max_iter=4))
        so = oracle_lib.state_arrays(od.get_state())
        worst_p = max(worst_p, float(np.linalg.norm(sg["p"] - so["p"])))
        worst_r = max(worst_r, synth.rot_err_deg(sg["R"], so["R"]))
        jo, jg = od.journey(), gx.journey()
        assert jo[1] == jg[1] and abs(jo[0] - jg[0]) < 1e-3, (k, jo, jg)
        ro, rg = od.idle(1), gx.idle(1)
        assert ro == rg, (k, ro, rg)
        erased += rg[0]
    gx.sync()
    assert erased > 200, erased
    assert worst_p < 1e-3 and worst_r < 0.01, (worst_p, worst_r)
    assert od.map_count() == gx.map_count()
    gx.close()


def test_scan_front_end_matches_oracle(oracle_lib, gpu_lib):
    """vina_scan_prepare (decoder keep rule + pcl_handler on the device: filter, stable sort by time offset - one
    bucket pass by time and a shared-memory sort per bucket, the radix passes when a bucket overflows -
    cut at 0.11 s) against the oracle's restatement, bit for bit including the order of equal stamps: many ties,
    blind-zone points, decimation, stamps beyond the cut, negative stamps, sizes around the sort's tile, the empty
    cloud's two-point stand-in, the error the reference cannot survive, and the device-pointer entry."""
    import torch

    cfg = small_cfg()
    gx = gpu_lib.Ctx(cfg, **dict(SMALL_CAPS, max_scan_points=300000))
    rng = np.random.default_rng(11)

    def cloud(n, ties=True):
        a = np.zeros((n, 4), dtype=np.float32)
        a[:, :3] = rng.uniform(-30, 30, (n, 3))
        a[::7, :3] *= 0.01
        if ties:
            a[:, 3] = rng.integers(0, 1300, n).astype(np.float32) * np.float32(1e-4)
        else:
            a[:, 3] = rng.uniform(-0.01, 0.12, n).astype(np.float32)
        return a

    for n, pfn, blind2, ties in ((50000, 1, 0.01, True), (50000, 3, 0.01, True), (240000, 2, 4.0, False),
                                 (2048, 1, 0.01, True), (2049, 1, 0.01, False), (5, 1, 0.01, True), (1, 1, 0.01, True)):
        a = cloud(n, ties)
        if n <= 5:
            a[:, :3] = 5.0
            a[:, 3] = np.float32(0.05)
        o = oracle_lib.scan_prepare(a, pfn, blind2)
        k, tl = gx.scan_prepare(a, pfn, blind2)
        g = gx.scan_download(k)
        assert k == o.shape[0] and np.array_equal(g, o), (n, pfn, k, o.shape)
        assert tl == o[-1, 3]
    # one stamp for the whole scan (and a scan with 3 distinct stamps): the time buckets of the fast path overflow
    # shared memory and the radix passes take over - same result
    for stamps in (1, 3):
        a = cloud(60000, True)
        a[:, 3] = (rng.integers(0, stamps, a.shape[0]).astype(np.float32) * np.float32(0.03) + np.float32(0.02))
        o = oracle_lib.scan_prepare(a, 1, 0.01)
        k, tl = gx.scan_prepare(a, 1, 0.01)
        assert k == o.shape[0] and np.array_equal(gx.scan_download(k), o) and tl == o[-1, 3]
    a = cloud(70000, False)
    o = oracle_lib.scan_prepare(a, 2, 0.25)
    d = torch.from_numpy(a).cuda()
    k, tl = gx.scan_prepare(a.shape[0], 2, 0.25, d_ptr=d.data_ptr())
    assert k == o.shape[0] and np.array_equal(gx.scan_download(k), o) and tl == o[-1, 3]
    dummy = np.array([[0, 0, 0, 0], [0, 0, 0, 0.09]], dtype=np.float32)
    near = cloud(3000)
    near[:, :3] *= 1e-4
    for b in (near, np.zeros((0, 4), dtype=np.float32)):
        k, tl = gx.scan_prepare(b, 1, 0.01)
        assert k == 2 and tl == np.float32(0.09) and np.array_equal(gx.scan_download(2), dummy)
    late = cloud(3000)
    late[:, 3] += np.float32(0.2)
    with pytest.raises(gpu_lib.VinaError) as e:
        gx.scan_prepare(late, 1, 0.01)
    assert e.value.code == -1  # VINA_E_ARG
    with pytest.raises(gpu_lib.VinaError):
        gx.step_prepared(0.0, np.zeros((5, 7)))  # no prepared scan
    gx.close()


def test_odometry_from_unsorted_scans(oracle_lib, gpu_lib):
    """Raw scans as a driver delivers them - shuffled in time, with blind-zone returns and stragglers beyond 0.11 s -
    through vina_scan_prepare + vina_odom_step_prepared vs the oracle's front end + step: 1 mm / 0.01 deg."""
    cfg = small_cfg("robosense128", 32, 600)
    seq, od, gx, last = bootstrap_pair(oracle_lib, gpu_lib, cfg, gpu_own_downsample=True)
    rng = np.random.default_rng(3)
    worst_p, worst_r = 0.0, 0.0
    for k in range(6):
        sc = seq.next_scan()
        junk = np.zeros((500, 4), dtype=np.float32)
        junk[:250, :3] = rng.uniform(-0.05, 0.05, (250, 3))          # inside the blind zone
        junk[:250, 3] = rng.uniform(0, 0.09, 250)
        junk[250:, :3] = rng.uniform(5, 20, (250, 3))                # stragglers
        junk[250:, 3] = rng.uniform(0.111, 0.2, 250)
        raw = np.concatenate([sc.xyzt, junk])[rng.permutation(sc.xyzt.shape[0] + 500)]
        prep = oracle_lib.scan_prepare(raw, 1, 0.01)
        assert prep.shape[0] == sc.xyzt.shape[0]
        r, _ = od.step(prep, sc.beg_time, sc.imu, iekf_on_full=True, max_iter=4)
        assert r == 0
        n, tl = gx.scan_prepare(raw, 1, 0.01)
        assert n == prep.shape[0] and tl == prep[-1, 3]
        sg = gpu_lib.state_arrays(gx.step_prepared(sc.beg_time, sc.imu, iekf_on_full=True, max_iter=4))
        so = oracle_lib.state_arrays(od.get_state())
        worst_p = max(worst_p, float(np.linalg.norm(sg["p"] - so["p"])))
        worst_r = max(worst_r, synth.rot_err_deg(sg["R"], so["R"]))
        assert np.linalg.norm(sg["p"] - sc.gt_p) < 0.02
    gx.sync()
    assert worst_p < 1e-3 and worst_r < 0.01, (worst_p, worst_r)
    gx.close()


def test_replay_front_end(gpu_lib, tmp_path):
    """vina_slam_b200.replay: a sequence through vina_odom_step, trajectory written in the reference's TUM format
    (io.cpp:67-77), with and without the sliding-window BA."""
    from vina_slam_b200 import replay

    cfg = small_cfg("robosense128", 16, 300)
    for ba in (False, True):
        boots, scans = replay.synthetic_frames(cfg, 12)
        out = str(tmp_path / f"traj_{int(ba)}.txt")
        rows, dt, worst = replay.replay(cfg, boots, scans, out=out, ba=ba, caps=SMALL_CAPS)
        assert rows.shape == (12, 8) and worst < 0.02 and dt > 0
        lines = open(out).read().splitlines()
        assert len(lines) == 12
        vals = np.array([[float(v) for v in l.split()] for l in lines])
        assert vals.shape == (12, 8) and np.max(np.abs(vals - rows)) < 1e-8
        assert np.all(np.diff(vals[:, 0]) > 0) and np.allclose(np.linalg.norm(vals[:, 4:], axis=1), 1.0, atol=1e-8)

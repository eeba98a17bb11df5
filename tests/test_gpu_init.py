"""GPU parity of the start-up phase: a COLD START through vina_odom_cold_start /
vina_odom_init_scan - IMU initialisation, win_size scans of the kd-tree IEKF (odometry.cpp:267-439), then
Initialization::motion_init (initialization.cpp:158-367) and the window's first marginalisation - against the oracle,
whose restatement of the same code is pinned bit for bit to the reference's own files (tests/test_oracle_vs_ref.py).
"""
import numpy as np
import pytest

from vina_slam_b200 import synth

pytestmark = pytest.mark.gpu


def _q(imu):
    return np.column_stack([np.round(imu[:, 0] * 1e9) * 1e-9, imu[:, 1:]])  # rclcpp::Time keeps integer nanoseconds


@pytest.mark.parametrize("base,beams,steps", [("robosense128", 32, 600), ("velodyne32", 32, 900)])
def test_cold_start_matches_oracle(oracle_lib, gpu_lib, base, beams, steps):
    cfg = synth.small_sensor(base, beams, steps)
    seq = synth.Sequence(cfg)
    od = oracle_lib.Odom(cfg)
    od.cold_start()
    gx = gpu_lib.Ctx(cfg, max_scan_points=cfg.n_points + 1024, max_nodes=200000, hash_capacity_log2=19)
    gx.cold_start()
    try:
        codes = []
        for k in range(40):
            sc = seq.next_scan()
            ro = od.init_scan(sc.xyzt, sc.beg_time, _q(sc.imu))
            rg, sg = gx.init_scan(sc.xyzt, sc.beg_time, _q(sc.imu))
            so, sg = oracle_lib.state_arrays(od.get_state()), gpu_lib.state_arrays(sg)
            assert rg == ro, (k, rg, ro)
            codes.append(ro)
            # the kd-tree phase is an estimator of its own (down-sampled means to fp32 rounding, plane fits through five
            # neighbours): close, not to the micrometre; gravity and the state after motion_init are the contract
            assert np.linalg.norm(sg["p"] - so["p"]) < 2e-3 and synth.rot_err_deg(sg["R"], so["R"]) < 0.02, (k, ro)
            assert np.linalg.norm(sg["g"] - so["g"]) < 1e-3
            if ro == 1:
                break
        assert codes[-1] == 1 and codes.count(0) >= cfg.win_size, codes
        so, sg = oracle_lib.state_arrays(od.get_state()), gpu_lib.state_arrays(gx.get_state())
        assert np.linalg.norm(sg["p"] - so["p"]) < 1e-3 and synth.rot_err_deg(sg["R"], so["R"]) < 0.01
        assert np.linalg.norm(sg["v"] - so["v"]) < 5e-3 and np.linalg.norm(sg["g"] - so["g"]) < 5e-3
        assert abs(np.linalg.norm(sg["g"]) - 9.8) < 0.2
        # same map in structure after the first marginalisation (node / root / slide counts)
        co, cg = od.map_count(), gx.map_count()
        assert abs(co[0] - cg[0]) <= 0.01 * co[0] and abs(co[1] - cg[1]) <= 0.01 * co[1], (co, cg)
        # ... and the per-scan loop carries on from there: 1 mm / 0.01 deg against the oracle, and the motion between
        # scans is the ground truth's (the start-up frame is gravity-aligned, not the synthetic world frame)
        prev_o = so["p"].copy()
        prev_gt = None
        for k in range(6):
            sc = seq.next_scan()
            r, _ = od.step(sc.xyzt, sc.beg_time, _q(sc.imu), True, 4)
            assert r == 0
            sg = gpu_lib.state_arrays(gx.step(sc.xyzt, sc.beg_time, _q(sc.imu), True, 4))
            so = oracle_lib.state_arrays(od.get_state())
            assert np.linalg.norm(sg["p"] - so["p"]) < 1e-3, (k, np.linalg.norm(sg["p"] - so["p"]))
            assert synth.rot_err_deg(sg["R"], so["R"]) < 0.01
            if prev_gt is not None:
                assert abs(np.linalg.norm(so["p"] - prev_o) - np.linalg.norm(sc.gt_p - prev_gt)) < 5e-3
            prev_o, prev_gt = so["p"].copy(), sc.gt_p.copy()
        gx.sync()
    finally:
        od.close()
        gx.close()


def test_init_scan_needs_cold_start(gpu_lib):
    cfg = synth.small_sensor("robosense128", 8, 100)
    gx = gpu_lib.Ctx(cfg, max_scan_points=4096, max_nodes=20000, hash_capacity_log2=14)
    sc = synth.Sequence(cfg).next_scan()
    with pytest.raises(gpu_lib.VinaError) as ei:
        gx.init_scan(sc.xyzt, sc.beg_time, sc.imu)
    assert ei.value.code == -6
    gx.close()

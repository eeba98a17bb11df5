"""Regenerates tests/golden/ref_small.npz from oracle/_ref/libvina_ref.so, i.e. from the REFERENCE'S OWN
SOURCES (point_utils.cpp, octree.cpp, voxel_map.cpp, imu_ekf.cpp, odometry.cpp compiled unmodified against
oracle/ref_shim). Only possible where /root/reference is mounted; the vectors travel with the repo so that
the oracle stays pinned on machines without the reference tree.  Run: python tests/golden/make_ref_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import oracle_py as op  # noqa: E402
from vina_slam_b200 import synth  # noqa: E402


def quantise_imu(imu):
    """The reference keeps IMU stamps as integer nanoseconds (rclcpp::Time): feed both sides that value."""
    q = imu.copy()
    q[:, 0] = np.round(q[:, 0] * 1e9) * 1e-9
    return q


def sorted_map(od):
    m = od.map_export()
    return m[np.lexsort((m["code"], m["key"][:, 2], m["key"][:, 1], m["key"][:, 0]))]


def run(make_odom, ref_stateless):
    """The scenario both implementations are driven through (also used by tests/test_oracle_vs_ref.py)."""
    out = {}
    cfg = synth.small_sensor("velodyne32", 16, 300)  # non-identity extrinsic, max_layer 3
    seq = synth.Sequence(cfg)
    sc = seq.next_scan(deskewed=True)
    xyzt = sc.xyzt[:256].copy()
    xyzt[3, 2] = 0.0  # calcBodyVar's z == 0 branch
    out["vi_pnt"], out["vi_var"] = op.var_init(xyzt, cfg, ref=ref_stateless)
    out["down"] = op.down_sampling_voxel(sc.xyzt, cfg.down_size, ref=ref_stateless)
    out["exp"] = op.exp_so3(np.array([0.3, -0.2, 0.5]), ref=ref_stateless)
    out["exp_dt"] = op.exp_so3(np.array([0.3, -0.2, 0.5]), 0.01, ref=ref_stateless)
    out["log"] = op.log_so3(out["exp"], ref=ref_stateless)
    cov = np.eye(15) * 1e-4
    cov[0, 1] = cov[1, 0] = 2e-5
    out["pu_var"], out["pu_pw"] = op.pvec_update(out["vi_pnt"], out["vi_var"], sc.gt_R.T.reshape(-1).copy(), sc.gt_p,
                                                 cov.T.reshape(-1).copy(), ref=ref_stateless)

    od = make_odom(cfg)
    seq = synth.Sequence(cfg)
    last = None
    for _ in range(cfg.win_size):
        last = seq.next_scan(deskewed=True)
        od.bootstrap(last.xyzt, op.make_state(last.gt_R, last.gt_p, last.gt_v, t=last.end_time))
    od.set_imu_anchor(last.end_time, quantise_imu(last.imu)[-1])
    m = sorted_map(od)
    for f in ("key", "code", "octo_state", "is_plane", "isexist", "has_sw", "N_add", "N_fix", "n_point_fix", "last_num",
              "opt_state", "P_add", "v_add", "P_fix", "center", "normal", "radius"):
        out["boot_" + f] = m[f]
    # the big per-node matrices as row digests (every entry still has to match for the digest to match bitwise)
    w36, w81 = np.cos(np.arange(36.0)), np.cos(np.arange(81.0))
    out["boot_plane_var_digest"] = m["plane_var"] @ w36
    out["boot_cov_add_digest"] = m["cov_add"] @ w81
    # match() for the points of the next scan placed with its ground-truth pose
    nxt = seq.next_scan(deskewed=True)
    pnt, var = op.var_init(nxt.xyzt, cfg, ref=ref_stateless)
    varw, pw = op.pvec_update(pnt, var, nxt.gt_R.T.reshape(-1).copy(), nxt.gt_p, (np.eye(15) * 1e-4).reshape(-1),
                              ref=ref_stateless)
    out["match_flags"], out["match_sigma"], out["match_center"] = od.match(pw, varw)
    # motion_blur + full steps
    traj, desk = [], None
    for k in range(4):
        sc = seq.next_scan() if k else None
        if k == 0:
            # the scan generated above was "deskewed"; start the odometry from the following one
            sc = seq.next_scan()
            od.set_imu_anchor(nxt.end_time, quantise_imu(nxt.imu)[-1])
            od.set_state(op.make_state(nxt.gt_R, nxt.gt_p, nxt.gt_v, t=nxt.end_time))
        r, desk = od.step(sc.xyzt, sc.beg_time, quantise_imu(sc.imu), True, 4)
        assert r == 0
        s = op.state_arrays(od.get_state())
        traj.append(np.concatenate([s["p"], s["v"], s["R"].reshape(-1), s["cov"].reshape(-1)]))
    out["traj"] = np.array(traj)
    out["deskewed_last"] = desk
    m = sorted_map(od)
    for f in ("key", "code", "octo_state", "is_plane", "N_add", "P_add", "v_add", "center", "normal"):
        out["end_" + f] = m[f]
    out["end_plane_var_digest"] = m["plane_var"] @ w36
    out["end_cov_add_digest"] = m["cov_add"] @ w81
    od.close()
    return out


def perturbed_poses(poses12, seed=7, rot=2e-3, tsl=0.02):
    """The captured window poses with a small seeded perturbation per frame (frame 0 included)."""
    rng = np.random.default_rng(seed)
    out = poses12.copy()
    for i in range(out.shape[0]):
        R = out[i, :9].reshape(3, 3).T
        R = R @ synth.rot_exp(rng.normal(0, rot, 3))
        out[i, :9] = R.T.reshape(-1)
        out[i, 9:] += rng.normal(0, tsl, 3)
    return out


def run_ba(make_odom):
    """The BA-probe scenario (LiDAR factor only): LidarFactor::acc_evaluate2 and
    evaluate_only_residual (factors.cpp:22-158) on the factors tras_opt collected in the last map update."""
    out = {}
    cfg = synth.small_sensor("robosense128", 24, 400, seed=31)
    seq = synth.Sequence(cfg)
    od = make_odom(cfg)
    od.ba_probe(True)
    for _ in range(cfg.win_size + 3):
        sc = seq.next_scan(deskewed=True)
        od.bootstrap(sc.xyzt, op.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
    out["n_factors"] = np.array([od.ba_count()])
    poses = od.ba_poses()
    out["poses"] = poses
    out["H0"], out["J0"], r0 = od.ba_hess(poses)
    pert = perturbed_poses(poses)
    out["poses_pert"] = pert
    out["H1"], out["J1"], r1 = od.ba_hess(pert)
    r2, lam = od.ba_residual(pert)  # overwrites the factors' eig / pcr_add like the reference's container
    out["lam0_sorted"] = np.sort(lam)
    out["H2"], out["J2"], r3 = od.ba_hess(pert)
    out["residuals"] = np.array([r0, r1, r2, r3])
    od.close()
    return out


def run_ba_odometry(make_odom, steps=14, sensor=("robosense128", 24, 350, 77)):
    """The per-scan loop with if_BA: 1 (local_mapping.cpp:437-441, 492-497, 541-546): IMU pre-integration factors,
    LI_BA_Optimizer::damping_iter (LM, 10 frames x 15 states), margi taking the re-evaluated factors back."""
    cfg = synth.small_sensor(sensor[0], sensor[1], sensor[2], seed=sensor[3])
    seq = synth.Sequence(cfg)
    od = make_odom(cfg)
    od.set_ba(True)
    for _ in range(cfg.win_size):
        a = seq.next_scan(deskewed=True)
        od.bootstrap(a.xyzt, op.make_state(a.gt_R, a.gt_p, a.gt_v, t=a.end_time))
    od.set_imu_anchor(a.end_time, quantise_imu(a.imu)[-1])
    traj = []
    for k in range(steps):
        sc = seq.next_scan()
        r, _ = od.step(sc.xyzt, sc.beg_time, quantise_imu(sc.imu), True, 4)
        assert r == 0
        s = op.state_arrays(od.get_state())
        traj.append(np.concatenate([s["p"], s["v"], s["R"].reshape(-1), s["cov"].reshape(-1)]))
    out = {"ba_traj": np.array(traj), "ba_runs": np.array([od.ba_stats()[0]])}
    m = sorted_map(od)
    for f in ("key", "code", "octo_state", "is_plane", "N_add", "P_add", "v_add", "center", "normal"):
        out["ba_end_" + f] = m[f]
    od.close()
    return out


if __name__ == "__main__":
    assert op.build_ref(), "oracle/_ref needs /root/reference"
    here = os.path.dirname(os.path.abspath(__file__))
    if "--ba-only" not in sys.argv:
        o = run(lambda cfg: op.Odom(cfg, ref=True), True)
        path = os.path.join(here, "ref_small.npz")
        np.savez_compressed(path, **o)
        print(path, os.path.getsize(path), "bytes")
    o = run_ba(lambda cfg: op.Odom(cfg, ref=True))
    o.update(run_ba_odometry(lambda cfg: op.Odom(cfg, ref=True)))
    path = os.path.join(here, "ref_ba.npz")
    np.savez_compressed(path, **o)
    print(path, os.path.getsize(path), "bytes", "factors:", int(o["n_factors"][0]), "residuals:", o["residuals"])

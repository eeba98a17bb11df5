"""Regenerates tests/golden/ref_init.npz from oracle/_ref/libvina_ref.so: a COLD START through the reference's own
start-up code (IMUEKF::IMU_init / process, lio_state_estimation_kdtree, Initialization::motion_init with
LI_BA_OptimizerGravity, align_gravity - initialization.cpp, odometry.cpp:267-439, optimizers.cpp:624-826 compiled
unmodified; ref_harness.cpp restates node.cpp:293-408 around them) followed by three ordinary steps.
Only possible where /root/reference is mounted.  Run: python tests/golden/make_ref_init_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import oracle_py as op  # noqa: E402
from vina_slam_b200 import synth  # noqa: E402


def quantise_imu(imu):
    q = imu.copy()
    q[:, 0] = np.round(q[:, 0] * 1e9) * 1e-9  # rclcpp::Time keeps integer nanoseconds
    return q


def scenario_config():
    return synth.small_sensor("robosense128", 16, 300)


def run(make_odom, n_steps=3):
    """(rows, map): one row per scan = return code, p, R, v, g, bg, ba, cov; the exported map at the end."""
    cfg = scenario_config()
    seq = synth.Sequence(cfg)
    od = make_odom(cfg)
    od.cold_start()
    rows = []
    for _ in range(40):
        sc = seq.next_scan()
        r = od.init_scan(sc.xyzt, sc.beg_time, quantise_imu(sc.imu))
        s = op.state_arrays(od.get_state())
        rows.append(np.concatenate([[r], s["p"], s["R"].reshape(-1), s["v"], s["g"], s["bg"], s["ba"], s["cov"].reshape(-1)]))
        if r != 0:
            break
    for _ in range(n_steps):
        sc = seq.next_scan()
        r, _ = od.step(sc.xyzt, sc.beg_time, quantise_imu(sc.imu), True, 4)
        s = op.state_arrays(od.get_state())
        rows.append(np.concatenate([[r], s["p"], s["R"].reshape(-1), s["v"], s["g"], s["bg"], s["ba"], s["cov"].reshape(-1)]))
    m = od.map_export()
    m = m[np.lexsort((m["code"], m["key"][:, 2], m["key"][:, 1], m["key"][:, 0]))]
    out = {"rows": np.array(rows), "counts": np.array(od.map_count())}
    for f in ("key", "code", "octo_state", "is_plane", "N_add", "N_fix", "P_add", "v_add", "center", "normal", "radius"):
        out["map_" + f] = m[f]
    od.close()
    return out


if __name__ == "__main__":
    if not op.have_ref():
        raise SystemExit("oracle/_ref is not built (needs /root/reference): python -c 'from oracle import oracle_py as o; o.build_ref()'")
    g = run(lambda cfg: op.Odom(cfg, ref=True))
    assert g["rows"][:, 0].max() == 1, "the reference's motion_init did not succeed on the scenario"
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_init.npz"), **g)
    print("rows", g["rows"].shape, "map nodes", g["map_key"].shape[0])

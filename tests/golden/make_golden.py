"""Regenerates tests/golden/oracle_small.npz from the strict-IEEE oracle.

The reference has no golden vectors and cannot be built in this image ("parity
unpinned"), so these vectors pin the ORACLE against accidental change; they are produced by the oracle
itself from a seeded synthetic sequence. Run: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import oracle_py as op  # noqa: E402
from vina_slam_b200 import synth  # noqa: E402


def build():
    cfg = synth.small_sensor("robosense128", 16, 300)
    seq = synth.Sequence(cfg)
    od = op.Odom(cfg)
    sc = None
    for _ in range(cfg.win_size):
        sc = seq.next_scan(deskewed=True)
        od.bootstrap(sc.xyzt, op.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
    od.set_imu_anchor(sc.end_time, sc.imu[-1])
    out = {}
    first = seq.next_scan()
    out["scan_xyzt"] = first.xyzt
    out["scan_imu"] = first.imu
    out["scan_beg"] = np.array([first.beg_time])
    pnt, var = op.var_init(first.xyzt[:64], cfg)
    out["var_init_pnt"], out["var_init_var"] = pnt, var
    od.set_dump(True)
    r, desk = od.step(first.xyzt, first.beg_time, first.imu, iekf_on_full=True, max_iter=4)
    assert r == 0
    out["deskewed"] = desk
    n = first.xyzt.shape[0]
    d0 = od.iter_dump(0, n)
    out["it0_HTH"], out["it0_HTz"], out["it0_nnt"] = d0["HTH"], d0["HTz"], d0["nnt"]
    out["it0_match"] = np.array([d0["match_num"]])
    out["it0_keys"], out["it0_codes"], out["it0_flags"] = d0["keys"], d0["codes"], d0["flags"]
    s = op.state_arrays(od.get_state())
    out["state_R"], out["state_p"], out["state_cov"] = s["R"], s["p"], s["cov"]
    traj = []
    for _ in range(4):
        sc = seq.next_scan()
        od.step(sc.xyzt, sc.beg_time, sc.imu, iekf_on_full=True, max_iter=4)
        s = op.state_arrays(od.get_state())
        traj.append(np.concatenate([s["p"], s["R"].reshape(-1)]))
    out["traj"] = np.array(traj)
    m = od.map_export()
    order = np.lexsort((m["code"], m["key"][:, 2], m["key"][:, 1], m["key"][:, 0]))
    m = m[order]
    out["map_key"], out["map_code"], out["map_N"] = m["key"], m["code"], m["N_add"]
    out["map_plane"], out["map_state"] = m["is_plane"], m["octo_state"]
    out["map_eig"] = m["eig_value"]
    return out


if __name__ == "__main__":
    o = build()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_small.npz")
    np.savez_compressed(path, **o)
    print(path, os.path.getsize(path), "bytes")

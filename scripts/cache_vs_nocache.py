"""Development tool: does the per-point leaf cache (odometry.cpp:124-127) change the association? Runs the same
IEKF iterations with the cache kept and with the cache reset before every iteration and compares flags / sums."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vina_slam_b200 import capi, synth  # noqa: E402

cfg = synth.SENSORS[sys.argv[1] if len(sys.argv) > 1 else "robosense128"]
seq = synth.Sequence(cfg)
gx = capi.Ctx(cfg, max_scan_points=max(300000, cfg.n_points + 1024))
for _ in range(cfg.win_size):
    sc = seq.next_scan(deskewed=True)
    gx.bootstrap(sc.xyzt, capi.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
sc = seq.next_scan(deskewed=True)
n = sc.xyzt.shape[0]
gx.scan_upload(sc.xyzt)
gx.var_init(0)
cov = np.eye(3).reshape(-1) * 1e-4
poses = []
R0 = sc.gt_R @ synth.rot_exp(np.array([1e-3, -1e-3, 1e-3]))
p0 = sc.gt_p + np.array([0.01, -0.01, 0.005])
for it, f in enumerate((1.0, 0.15, 0.02, 0.0)):  # a converging sequence of iterates
    R = sc.gt_R @ synth.rot_exp(f * np.array([1e-3, -1e-3, 1e-3]))
    poses.append((np.ascontiguousarray(R.T.reshape(-1)), sc.gt_p + f * np.array([0.01, -0.01, 0.005])))
res = {}
for mode in ("cache", "nocache"):
    gx.iekf_begin(0, cov, cov)
    out = []
    for it, (Rc, p) in enumerate(poses):
        if mode == "nocache" and it > 0:
            gx.iekf_begin(0, cov, cov)
        g = gx.iekf_accumulate(Rc, p, debug=True)
        a = gx.iekf_debug_assoc(n)
        out.append((g, a["flags"].copy(), a["codes"].copy(), a["keys"].copy()))
    res[mode] = out
for it in range(len(poses)):
    (g1, f1, c1, k1), (g2, f2, c2, k2) = res["cache"][it], res["nocache"][it]
    df = np.nonzero(f1 != f2)[0]
    dc = np.nonzero((c1 != c2) & (f1 == f2))[0]
    print(f"it{it}: match {g1['match_num']} / {g2['match_num']}  flags differ {df.size}  codes differ {dc.size}  "
          f"HTH rel diff {np.abs(g1['HTH'] - g2['HTH']).max() / np.abs(g1['HTH']).max():.2e}")
    for i in df[:5]:
        print("   pt", i, "flags", f1[i], f2[i], "codes", c1[i], c2[i], "key", k1[i], k2[i])

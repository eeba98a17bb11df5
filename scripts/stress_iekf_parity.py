"""Development tool: repeat the IEKF association / sigma comparison of tests/test_gpu_parity.py (rare-race hunt)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import bootstrap_pair, cov_blocks, small_cfg  # noqa: E402
from oracle import oracle_py as op  # noqa: E402
from vina_slam_b200 import capi, synth  # noqa: E402

op.build()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
cfg = small_cfg("velodyne32", 32, 500)
seq, od, gx, last = bootstrap_pair(op, capi, cfg)
sc = seq.next_scan(deskewed=True)
pnt, var = op.var_init(sc.xyzt, cfg)
n = pnt.shape[0]
R0 = sc.gt_R @ op.exp_so3(np.array([0.004, -0.003, 0.005]))
p0 = sc.gt_p + np.array([0.03, -0.02, 0.015])
od.set_state(op.make_state(R0, p0, sc.gt_v, t=sc.end_time))
od.set_dump(True)
od.iekf(pnt, var, 4)
iters = od.last_iters()
cov = op.state_arrays(op.make_state())["cov"]
rot_var, tsl_var = cov_blocks(cov)
dumps = [od.iter_dump(it, n) for it in range(iters)]
for rep in range(reps):
    gx.pvec_upload(0, pnt, var)
    gx.iekf_begin(0, rot_var, tsl_var)
    line = []
    for it in range(iters):
        d = dumps[it]
        for dbg in (True, False, True):
            g = gx.iekf_accumulate(d["R_col"], d["p"], debug=dbg)
            if dbg:
                a = gx.iekf_debug_assoc(n)
                m = d["flags"] > 0
                e = np.abs(a["sigma"][m] - d["sigma"][m]) / d["sigma"][m]
                bad = int((e > 1e-6).sum())
                line.append(f"it{it} sig {e.max():.1e} bad {bad} flags {int((a['flags'] != d['flags']).sum())} codes {int((a['codes'] != d['codes']).sum())}")
                if bad:
                    w = np.nonzero(m)[0][np.argmax(e)]
                    line.append(f"[pt {w} code {a['codes'][w]} gpu {a['sigma'][w]:.6e} ref {d['sigma'][w]:.6e}]")
            hh = np.abs(g["HTH"] - d["HTH"]).max() / np.abs(d["HTH"]).max()
            if hh > 1e-7:
                line.append(f"HTH {hh:.1e} (dbg={dbg})")
    print(f"rep {rep}: " + " | ".join(line), flush=True)

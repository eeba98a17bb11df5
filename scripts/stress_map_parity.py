"""Development tool: repeat the bootstrap + map compare of tests/test_gpu_parity.py many times and report
the worst deviation of every covariance-derived field (hunting rare races)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import bootstrap_pair, small_cfg, sort_nodes  # noqa: E402
from oracle import oracle_py as op  # noqa: E402
from vina_slam_b200 import capi, synth  # noqa: E402

op.build()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
for base, beams, steps in (("velodyne32", 32, 500), ("robosense128", 32, 600)):
    cfg = small_cfg(base, beams, steps)
    for rep in range(reps):
        seq, od, gx, last = bootstrap_pair(op, capi, cfg)
        for k in range(3):
            sc = seq.next_scan(deskewed=True)
            od.bootstrap(sc.xyzt, op.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
            gx.set_state(capi.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
            gx.down_upload(od.last_down())
            gx.var_init(1)
            gx.odom_map_update()
        mo, mg = sort_nodes(od.map_export()), sort_nodes(gx.map_export())
        same = mo.shape[0] == mg.shape[0] and np.array_equal(mo["key"], mg["key"]) and np.array_equal(mo["code"], mg["code"])
        msg = [f"{base} rep {rep}: nodes {mo.shape[0]}/{mg.shape[0]} structure {'ok' if same else 'DIFFERS'}"]
        if same:
            leaf = mo["octo_state"] == 0
            for f, sel in (("cov_add", slice(None)), ("plane_var", leaf), ("P_add", leaf), ("v_add", leaf), ("N_add", slice(None)),
                           ("N_fix", slice(None)), ("center", slice(None)), ("eig_value", leaf)):
                a, b = mo[f][sel].astype(np.float64), mg[f][sel].astype(np.float64)
                a, b = a.reshape(a.shape[0], -1), b.reshape(b.shape[0], -1)
                den = np.maximum(np.abs(a).max(axis=1, keepdims=True), 1e-300)
                e = np.abs(a - b) / den
                w = int(np.argmax(e.max(axis=1)))
                msg.append(f"{f} {e.max():.1e}" + (f" @node layer={mo['layer'][sel][w]} N_add={mo['N_add'][sel][w]} N_fix={mo['N_fix'][sel][w]}" if e.max() > 1e-9 else ""))
        print(" | ".join(msg), flush=True)
        gx.close()

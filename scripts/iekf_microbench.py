"""Times k_iekf alone on a full-size bootstrapped map (development tool, not part of the product)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vina_slam_b200 import capi, synth  # noqa: E402

cfg = synth.SENSORS[sys.argv[1] if len(sys.argv) > 1 else "robosense128"]
seq = synth.Sequence(cfg)
gx = capi.Ctx(cfg, max_scan_points=max(300000, cfg.n_points + 1024))
sc = None
for _ in range(cfg.win_size):
    sc = seq.next_scan(deskewed=True)
    gx.bootstrap(sc.xyzt, capi.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
sc = seq.next_scan(deskewed=True)
gx.scan_upload(sc.xyzt)
gx.var_init(0)
cov = np.eye(3).reshape(-1) * 1e-4
gx.iekf_begin(0, cov, cov)
R = np.ascontiguousarray(sc.gt_R.T.reshape(-1))
p = np.ascontiguousarray(sc.gt_p)
r = gx.iekf_accumulate(R, p)
print("match", r["match_num"], "of", sc.xyzt.shape[0])
ms = C.c_float(0)
import os
BPS = int(os.environ.get("VINA_IEKF_BLOCKS_PER_SM", "1"))
B18 = (18 * BPS) << 8  # the share of one sequence in a batch of 8 on 148 SMs
for variant, name in [(0, "product"), (32, "no prefetch"), (1, "no final reduce"), (8, "no DMMA reduction"),
                      (4, "no gate math"), (12, "no gate, no reduction"), (16, "stream loads only"),
                      (17, "stream loads, no final"), (B18, "18 blocks"), (B18 | 32, "18 blocks, no prefetch"),
                      (B18 | 8, "18 blocks, no DMMA"), (B18 | 4, "18 blocks, no gate"), (B18 | 16, "18 blocks, loads only"),
                      (((37 * BPS) << 8), "37 blocks")]:
    for reset, tag in [(0, "cached"), (1, "cold cache (hash+descent)")]:
        gx.lib.vina_iekf_time_kernel(gx.h, capi._dp(R), capi._dp(p), C.c_int(50), C.c_int(variant), C.c_int(reset),
                                     C.byref(ms))
        print(f"variant {variant:2d} {name:28s} {tag:28s} {ms.value * 1e3:8.2f} us/launch")

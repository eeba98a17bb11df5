"""Development tool: ShardedIekf at world = 1 (no leaf cache, routing through the record format) against
vina_odom_iekf_host on the same context / map / scan: isolates the effect of dropping the per-point cache."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vina_slam_b200 import capi, sharded, synth  # noqa: E402

cfg = synth.SENSORS[sys.argv[1] if len(sys.argv) > 1 else "robosense128"]
seq = synth.Sequence(cfg)
gx = capi.Ctx(cfg, max_scan_points=max(300000, cfg.n_points + 1024))
sh = sharded.MapShard(gx, 0, 1)
for _ in range(cfg.win_size):
    sc = seq.next_scan(deskewed=True)
    gx.bootstrap(sc.xyzt, capi.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
iek = sharded.ShardedIekf(sh)
for k in range(4):
    sc = seq.next_scan(deskewed=True)
    n = sc.xyzt.shape[0]
    pert = capi.make_state(sc.gt_R @ synth.rot_exp(np.array([1e-3, -1e-3, 1e-3])), sc.gt_p + np.array([0.01, -0.01, 0.005]),
                           sc.gt_v, t=sc.end_time)
    gx.scan_upload(sc.xyzt)
    gx.var_init(0)
    gx.set_state(pert)
    it_a, _ = gx.odom_iekf(0, 4, host_solve=True)
    a = capi.state_arrays(gx.get_state())
    gx.set_state(pert)
    it_d, _ = gx.odom_iekf(0, 4, host_solve=False)
    d = capi.state_arrays(gx.get_state())
    gx.set_state(pert)
    it_b = iek.run(0, n, 4)
    b = capi.state_arrays(gx.get_state())
    print(f"scan {k}: iters host {it_a} device {it_d} world1 {it_b} | |p_world1 - p_host| {np.linalg.norm(a['p'] - b['p']):.3e} "
          f"|p_device - p_host| {np.linalg.norm(a['p'] - d['p']):.3e} | err vs gt {np.linalg.norm(a['p'] - sc.gt_p):.3e}")
    gx.set_state(capi.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
    gx.downsample()
    gx.n_down()
    gx.var_init(1)
    gx.odom_map_update()

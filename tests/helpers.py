"""Shared helpers of the parity tests: drive the CPU oracle and the CUDA path with identical inputs."""
from __future__ import annotations

import numpy as np

from vina_slam_b200 import synth

SMALL_CAPS = dict(max_scan_points=40000, max_nodes=60000, hash_capacity_log2=17, fix_pool_points=1 << 20,
                  win_pool_points=200000)


def small_cfg(base="robosense128", beams=32, steps=600, seed=None):
    return synth.small_sensor(base, beams, steps, seed)


def col(R_row):
    return np.asarray(R_row, dtype=np.float64).reshape(3, 3).T.reshape(-1).copy()


def cov_blocks(cov_row):
    """rot_var / tsl_var (column-major 9) of a row-major 15x15 covariance."""
    c = np.asarray(cov_row)
    return c[0:3, 0:3].T.reshape(-1).copy(), c[3:6, 3:6].T.reshape(-1).copy()


def bootstrap_pair(op, capi, cfg, n_boot=None, caps=None, gpu_own_downsample=False, world=None):
    """Bootstrap the oracle and the GPU context with the same deskewed scans at ground-truth states.

    The GPU gets the oracle's down-sampled cloud (bit-identical map inputs) unless gpu_own_downsample.
    Returns (seq, oracle_odom, gpu_ctx, last_scan).
    """
    seq = synth.Sequence(cfg, world=world)
    od = op.Odom(cfg)
    gx = capi.Ctx(cfg, **(caps or SMALL_CAPS))
    n_boot = cfg.win_size if n_boot is None else n_boot
    sc = None
    for _ in range(n_boot):
        sc = seq.next_scan(deskewed=True)
        st_o = op.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time)
        od.bootstrap(sc.xyzt, st_o)
        st_g = capi.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time)
        if gpu_own_downsample:
            gx.bootstrap(sc.xyzt, st_g)
        else:
            gx.set_state(st_g)
            gx.down_upload(od.last_down())
            gx.var_init(1)
            gx.odom_map_update()
            gx.sync()
    od.set_imu_anchor(sc.end_time, sc.imu[-1])
    gx.set_imu_anchor(sc.end_time, sc.imu[-1])
    return seq, od, gx, sc


def sort_nodes(rec):
    order = np.lexsort((rec["code"], rec["key"][:, 2], rec["key"][:, 1], rec["key"][:, 0]))
    return rec[order]


def ulp_diff_f32(a, b):
    ai = np.asarray(a, dtype=np.float32).view(np.int32).astype(np.int64)
    bi = np.asarray(b, dtype=np.float32).view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, -(ai & 0x7FFFFFFF), ai)
    bi = np.where(bi < 0, -(bi & 0x7FFFFFFF), bi)
    return np.abs(ai - bi)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = max(np.max(np.abs(b)), 1e-300)
    return float(np.max(np.abs(a - b)) / den)

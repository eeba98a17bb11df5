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


def iekf_compare(oracle_lib, gpu_lib, cfg, n_iter=4, world=None, caps=None, pair=None, min_match=0.3):
    """a4-a6 on one scan from a perturbed start: voxel keys, match flags and associated leaves of every point bit
    for bit per iteration, sigma_l / H / b / nn^T to tolerance (odometry.cpp:98-148, voxel_map.cpp:241-266,
    octree.cpp:551-595). `pair` = an existing (seq, oracle odometry, GPU context) to run on (kept open)."""
    seq, od, gx = pair if pair is not None else bootstrap_pair(oracle_lib, gpu_lib, cfg, world=world, caps=caps)[:3]
    sc = seq.next_scan(deskewed=True)
    pnt, var = oracle_lib.var_init(sc.xyzt, cfg)
    n = pnt.shape[0]
    # start from a perturbed state so that several iterations and re-associations happen
    R0 = sc.gt_R @ oracle_lib.exp_so3(np.array([0.004, -0.003, 0.005]))
    p0 = sc.gt_p + np.array([0.03, -0.02, 0.015])
    od.set_state(oracle_lib.make_state(R0, p0, sc.gt_v, t=sc.end_time))
    od.set_dump(True)
    od.iekf(pnt, var, n_iter)
    iters = od.last_iters()
    assert iters >= 2
    cov = oracle_lib.state_arrays(oracle_lib.make_state())["cov"]
    rot_var, tsl_var = cov_blocks(cov)
    gx.pvec_upload(0, pnt, var)
    gx.iekf_begin(0, rot_var, tsl_var)
    total = 0
    for it in range(iters):
        d = od.iter_dump(it, n)
        g = gx.iekf_accumulate(d["R_col"], d["p"], debug=True)
        a = gx.iekf_debug_assoc(n)
        assert np.array_equal(a["keys"], d["keys"]), f"voxel keys differ at iteration {it}"
        assert np.array_equal(a["flags"], d["flags"]), f"match flags differ at iteration {it}"
        assert np.array_equal(a["codes"], d["codes"]), f"associated leaves differ at iteration {it}"
        assert g["match_num"] == d["match_num"] and d["match_num"] > min_match * n
        m = d["flags"] > 0
        # sigma_l = J plane_var J^T + n^T var n cancels ~7 digits (plane_var carries the lever arm of a
        # centre tens of metres from the origin), so rounding-level differences show up at ~1e-9
        sig_err = np.abs(a["sigma"][m] - d["sigma"][m]) / d["sigma"][m]
        assert np.max(sig_err) < 1e-6, (it, float(np.max(sig_err)), int((sig_err > 1e-6).sum()), int(m.sum()))
        assert rel_err(g["HTH"], d["HTH"]) < 1e-4 and rel_err(g["HTz"], d["HTz"]) < 1e-4
        assert rel_err(g["nnt"], d["nnt"]) < 1e-4
        # the design is far tighter than the contract
        assert rel_err(g["HTH"], d["HTH"]) < 1e-7 and rel_err(g["HTz"], d["HTz"]) < 1e-7
        assert rel_err(g["nnt"], d["nnt"]) < 1e-12
        # same sums from the non-debug kernel
        g2 = gx.iekf_accumulate(d["R_col"], d["p"], debug=False)
        assert np.array_equal(g2["HTH"], g["HTH"]) and g2["match_num"] == g["match_num"]
        total += d["match_num"]
    od.set_dump(False)
    if pair is None:
        gx.close()
    return total


def compare_maps(mo, mg, exact_cov=False, eig_planes_only=False):
    mo, mg = sort_nodes(mo), sort_nodes(mg)
    assert mo.shape[0] == mg.shape[0], "different number of octree nodes"
    for f in ("key", "code", "layer", "octo_state", "isexist", "has_sw", "is_plane", "last_num", "opt_state",
              "N_add", "N_fix", "n_point_fix", "n_win_points", "N_local"):
        assert np.array_equal(mo[f], mg[f]), f"map field {f} differs"
    assert np.array_equal(mo["voxel_center"], mg["voxel_center"])
    assert np.array_equal(mo["quater_length"], mg["quater_length"])
    leaf = mo["octo_state"] == 0
    # cluster sums: lower triangle + v, bit for bit (same summation order as the reference)
    low = [0, 1, 2, 4, 5, 8]
    for f in ("P_add", "P_fix"):
        assert np.array_equal(mo[f][leaf][:, low], mg[f][leaf][:, low]), f"{f} differs"
    assert np.array_equal(mo["v_add"][leaf], mg["v_add"][leaf])
    assert np.array_equal(mo["v_fix"][leaf], mg["v_fix"][leaf])
    # eigen-decomposition and plane parameters derived from them: bit for bit
    # (the reference build leaves eig_* uninitialised until a leaf has been judged)
    em = leaf & (mo["is_plane"] > 0) if eig_planes_only else leaf
    assert np.array_equal(mo["eig_value"][em], mg["eig_value"][em])
    assert np.array_equal(mo["eig_vector"][em], mg["eig_vector"][em])
    assert np.array_equal(mo["center"], mg["center"]) and np.array_equal(mo["normal"], mg["normal"])
    assert np.array_equal(mo["radius"], mg["radius"])
    # covariance-derived quantities: the device stores point covariances symmetric -> tolerance
    # (plane_var = u_c cov_add u_c^T cancels several digits, like sigma_l)
    # (the plane of an interior node is dead state: the device reuses its storage for the children index)
    for f, tol, sel in (("cov_add", 1e-12, slice(None)), ("plane_var", 1e-7, leaf)):
        den = np.maximum(np.abs(mo[f][sel]).max(axis=1, keepdims=True), 1e-300)
        assert np.max(np.abs(mo[f][sel] - mg[f][sel]) / den) < tol, f"{f} differs"
    return mo, mg

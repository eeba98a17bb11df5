"""CPU checks of the host side of the sliding-window BA (vina_slam_b200/host/vina_ba.cpp) - the part of the product
that needs no device: the IMU pre-integration factor (IMU_PRE::push_imu / give_evaluate,
src/estimation/imu_preintegration.cpp:32-163) against the oracle, whose IMU_PRE reproduces the reference's own
source file bit for bit (tests/test_oracle_vs_ref.py), and the symmetric solve of the LM step
(optimizers.cpp:466) against numpy."""
import ctypes as C

import numpy as np

from vina_slam_b200 import synth


def _states(capi_like, sc0, sc1, rng, bg, ba):
    out = []
    for sc in (sc0, sc1):
        s = capi_like.make_state(sc.gt_R @ synth.rot_exp(rng.normal(0, 2e-3, 3)), sc.gt_p + rng.normal(0, 0.01, 3),
                                 sc.gt_v + rng.normal(0, 0.01, 3), t=sc.end_time)
        for k in range(3):
            s.bg[k] = bg[k] + rng.normal(0, 1e-4)
            s.ba[k] = ba[k] + rng.normal(0, 1e-3)
        out.append(s)
    return out


def test_imu_factor_matches_oracle(oracle_lib, gpu_lib):
    lib, olib = gpu_lib.load(), oracle_lib.load()
    olib.vo_ba_imu_evaluate.restype = C.c_double
    cfg = synth.small_sensor("robosense128", 8, 100)
    seq = synth.Sequence(cfg)
    sc0 = seq.next_scan()
    sc1 = seq.next_scan()
    imu = sc1.imu.copy()
    imu[:, 0] = np.round(imu[:, 0] * 1e9) * 1e-9
    assert imu.shape[0] >= 10
    bg, ba = np.array([1e-3, -2e-3, 5e-4]), np.array([0.02, -0.01, 0.03])
    gcfg, ocfg = gpu_lib.make_config(cfg), oracle_lib.make_config(cfg)
    ia = gpu_lib.imu_array(imu)
    worst = 0.0
    for trial in range(3):
        rng = np.random.default_rng(10 + trial)
        sg = _states(gpu_lib, sc0, sc1, rng, bg, ba)
        rng = np.random.default_rng(10 + trial)
        so = _states(oracle_lib, sc0, sc1, rng, bg, ba)
        Jg, gg, rg = np.zeros(900), np.zeros(30), C.c_double(0)
        rc = lib.vina_ba_imu_evaluate(C.byref(gcfg), gpu_lib._dp(bg), gpu_lib._dp(ba), ia.ctypes.data_as(C.c_void_p),
                                      C.c_int(ia.shape[0]), C.c_double(1.0), C.byref(sg[0]), C.byref(sg[1]), C.byref(rg),
                                      gpu_lib._dp(Jg), gpu_lib._dp(gg))
        assert rc == 0
        Jo, go = np.zeros(900), np.zeros(30)
        ro = olib.vo_ba_imu_evaluate(C.byref(ocfg), gpu_lib._dp(bg), gpu_lib._dp(ba), gpu_lib._dp(np.ascontiguousarray(imu)),
                                     C.c_int(imu.shape[0]), C.c_double(1.0), C.byref(so[0]), C.byref(so[1]),
                                     gpu_lib._dp(Jo), gpu_lib._dp(go))
        assert ro > 0 and np.abs(Jo).max() > 0
        # same expressions, same order: agreement to rounding (the two builds differ in compiler and flags only)
        assert abs(rg.value - ro) <= 1e-9 * abs(ro)
        assert np.max(np.abs(Jg - Jo)) <= 1e-9 * np.abs(Jo).max()
        assert np.max(np.abs(gg - go)) <= 1e-9 * np.abs(go).max()
        # residual-only evaluation returns the same value
        r2 = C.c_double(0)
        assert lib.vina_ba_imu_evaluate(C.byref(gcfg), gpu_lib._dp(bg), gpu_lib._dp(ba), ia.ctypes.data_as(C.c_void_p),
                                        C.c_int(ia.shape[0]), C.c_double(1.0), C.byref(sg[0]), C.byref(sg[1]), C.byref(r2),
                                        None, None) == 0
        assert r2.value == rg.value
        J = Jg.reshape(30, 30).T
        assert np.max(np.abs(J - J.T)) <= 1e-9 * np.abs(J).max()  # J^T cov^-1 J is symmetric
        worst = max(worst, abs(rg.value - ro) / abs(ro))
    # bad arguments: codes, no crash
    assert lib.vina_ba_imu_evaluate(None, gpu_lib._dp(bg), gpu_lib._dp(ba), ia.ctypes.data_as(C.c_void_p), C.c_int(5),
                                    C.c_double(1.0), C.byref(sg[0]), C.byref(sg[1]), C.byref(rg), None, None) == -1
    assert lib.vina_ba_imu_evaluate(C.byref(gcfg), gpu_lib._dp(bg), gpu_lib._dp(ba), ia.ctypes.data_as(C.c_void_p),
                                    C.c_int(1), C.c_double(1.0), C.byref(sg[0]), C.byref(sg[1]), C.byref(rg), None, None) == -1


def test_lm_solve_matches_numpy(gpu_lib):
    """The LDL^T with diagonal pivoting and delayed panel updates behind dxi = (Hess + u D).ldlt().solve(-JacT):
    sizes around the panel width, the 135 x 135 system of a 10-frame window, an indefinite matrix (pivoting), and a
    matrix whose upper triangle holds garbage (only the lower triangle is read)."""
    lib = gpu_lib.load()
    rng = np.random.default_rng(5)
    for n in (1, 2, 7, 8, 9, 16, 17, 135, 150):
        B = rng.normal(size=(n, n))
        scale = 1.0 + 3.0 * (np.arange(n) % 5)
        A = (B @ B.T + 1e-3 * np.eye(n)) * np.outer(scale, scale)
        if n >= 7:
            A[::3, :] *= -1  # symmetric sign flips: indefinite, still well conditioned enough
            A[:, ::3] *= -1
            A -= 2.0 * np.diag(np.diag(A)) * (np.arange(n) % 4 == 0)
            A = 0.5 * (A + A.T)
        b = rng.normal(size=n)
        x = np.zeros(n)
        Acm = np.ascontiguousarray(A.T.reshape(-1))
        assert lib.vina_ba_solve(gpu_lib._dp(Acm), C.c_int(n), gpu_lib._dp(b), gpu_lib._dp(x)) == 0
        ref = np.linalg.solve(A, b)
        assert np.max(np.abs(x - ref)) <= 1e-7 * max(np.max(np.abs(ref)), 1e-300) * max(np.linalg.cond(A) * 1e-9, 1.0), n
        # garbage above the diagonal must not matter
        G = A.copy()
        G[np.triu_indices(n, 1)] = 1e30
        x2 = np.zeros(n)
        assert lib.vina_ba_solve(gpu_lib._dp(np.ascontiguousarray(G.T.reshape(-1))), C.c_int(n), gpu_lib._dp(b),
                                 gpu_lib._dp(x2)) == 0
        assert np.array_equal(x, x2), n
    assert lib.vina_ba_solve(None, C.c_int(3), gpu_lib._dp(np.zeros(3)), gpu_lib._dp(np.zeros(3))) == -1

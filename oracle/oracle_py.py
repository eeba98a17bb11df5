"""ORACLE — TEST INFRASTRUCTURE ONLY (pinned against oracle/_ref, see vina_oracle.hpp).

ctypes view of oracle/liboracle.so (strict IEEE build, the parity checker) and
oracle/liboracle_fast.so (-O3 -ffast-math, the timed CPU baseline).  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class VoConfig(C.Structure):
    _fields_ = [
        ("voxel_size", C.c_double), ("min_eigen_value", C.c_double), ("plane_eigen_value_thre", C.c_double * 4),
        ("min_point", C.c_double * 4), ("dept_err", C.c_double), ("beam_err", C.c_double), ("down_size", C.c_double),
        ("ext_R", C.c_double * 9), ("ext_t", C.c_double * 3), ("cov_gyr", C.c_double), ("cov_acc", C.c_double),
        ("rdw_gyr", C.c_double), ("rdw_acc", C.c_double), ("max_layer", C.c_int32), ("max_points", C.c_int32),
        ("win_size", C.c_int32), ("thread_num", C.c_int32),
    ]


class VoState(C.Structure):
    _fields_ = [
        ("t", C.c_double), ("R", C.c_double * 9), ("p", C.c_double * 3), ("v", C.c_double * 3),
        ("bg", C.c_double * 3), ("ba", C.c_double * 3), ("g", C.c_double * 3), ("cov", C.c_double * 225),
    ]


NODE_DTYPE = np.dtype([
    ("key", "<i8", 3), ("code", "<i4"), ("layer", "<i4"), ("octo_state", "<i4"), ("isexist", "<i4"),
    ("has_sw", "<i4"), ("is_plane", "<i4"), ("last_num", "<i4"), ("opt_state", "<i4"), ("N_add", "<i4"),
    ("N_fix", "<i4"), ("n_point_fix", "<i4"), ("n_win_points", "<i4"), ("N_local", "<i4", 16),
    ("P_add", "<f8", 9), ("v_add", "<f8", 3), ("P_fix", "<f8", 9), ("v_fix", "<f8", 3), ("eig_value", "<f8", 3),
    ("eig_vector", "<f8", 9), ("center", "<f8", 3), ("normal", "<f8", 3), ("plane_var", "<f8", 36),
    ("radius", "<f8"), ("cov_add", "<f8", 81), ("voxel_center", "<f8", 3), ("quater_length", "<f8"),
], align=True)


def build(force: bool = False) -> None:
    need = force or not all(os.path.exists(os.path.join(_HERE, f)) for f in ("liboracle.so", "liboracle_fast.so"))
    if not need:
        srcs = ["vina_oracle.cpp", "oracle_capi.cpp", "omat.hpp", "vina_oracle.hpp", "oracle_capi.h", "Makefile"]
        so_t = min(os.path.getmtime(os.path.join(_HERE, f)) for f in ("liboracle.so", "liboracle_fast.so"))
        need = any(os.path.getmtime(os.path.join(_HERE, s)) > so_t for s in srcs)
    if need:
        subprocess.check_call(["make", "-C", _HERE, "-s"])


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _dp(a):
    return _ptr(a, C.c_double)


def _fp(a):
    return _ptr(a, C.c_float)


_LIBS = {}


def have_ref() -> bool:
    """oracle/_ref/libvina_ref.so = the reference's own sources compiled against oracle/ref_shim (see Makefile)."""
    return os.path.exists(os.path.join(_HERE, "_ref", "libvina_ref.so"))


def build_ref(reference="/root/reference") -> bool:
    """Only possible where the reference tree is mounted (this container); the GPU box gets the prebuilt files."""
    if not os.path.isdir(os.path.join(reference, "src")):
        return have_ref()
    subprocess.check_call(["make", "-C", _HERE, "-s", "ref", f"REF={reference}"])
    return have_ref()


def load(fast: bool = False, ref: bool = False):
    if ref:
        name = os.path.join("_ref", "libvina_ref_fast.so" if fast else "libvina_ref.so")
    else:
        name = "liboracle_fast.so" if fast else "liboracle.so"
    if name in _LIBS:
        return _LIBS[name]
    path = os.path.join(_HERE, name)
    if not os.path.exists(path):
        if ref:
            raise FileNotFoundError(path)
        build()
    lib = C.CDLL(path)
    lib.vo_odom_create.restype = C.c_void_p
    lib.vo_odom_create.argtypes = [C.POINTER(VoConfig)]
    for fn in ("vo_odom_destroy", "vo_odom_set_state", "vo_odom_get_state", "vo_odom_set_imu_anchor",
               "vo_odom_bootstrap", "vo_odom_stage_times", "vo_odom_deskew", "vo_odom_set_dump",
               "vo_odom_map_update", "vo_odom_set_ba", "vo_odom_ba_stats", "vo_odom_ba_probe", "vo_odom_cold_start"):
        if hasattr(lib, fn):
            getattr(lib, fn).restype = None
    if hasattr(lib, "vo_sync_create"):
        lib.vo_sync_create.restype = C.c_void_p
        lib.vo_sync_destroy.restype = None
        lib.vo_sync_destroy.argtypes = [C.c_void_p]
        lib.vo_sync_push_imu.restype = None
        lib.vo_sync_push_scan.restype = None
    lib.vo_odom_map_count.restype = C.c_int64
    lib.vo_odom_map_export.restype = C.c_int64
    _LIBS[name] = lib
    return lib


def make_config(cfg) -> VoConfig:
    """cfg: vina_slam_b200.synth.SensorConfig (duck-typed)."""
    c = VoConfig()
    c.voxel_size = cfg.voxel_size
    c.min_eigen_value = cfg.min_eigen_value
    for i in range(4):
        c.plane_eigen_value_thre[i] = cfg.plane_thre[i]
        c.min_point[i] = (20, 20, 15, 10)[i]  # node.cpp:219
    c.dept_err, c.beam_err, c.down_size = cfg.dept_err, cfg.beam_err, cfg.down_size
    R = cfg.ext_R_colmajor()
    for i in range(9):
        c.ext_R[i] = R[i]
    for i in range(3):
        c.ext_t[i] = cfg.ext_t[i]
    c.cov_gyr, c.cov_acc, c.rdw_gyr, c.rdw_acc = cfg.cov_gyr, cfg.cov_acc, cfg.rdw_gyr, cfg.rdw_acc
    c.max_layer, c.max_points, c.win_size, c.thread_num = cfg.max_layer, cfg.max_points, cfg.win_size, cfg.thread_num
    return c


def make_state(R_rowmajor=None, p=None, v=None, t=0.0, cov=None, g=(0.0, 0.0, -9.8)) -> VoState:
    s = VoState()
    s.t = t
    R = np.eye(3) if R_rowmajor is None else np.asarray(R_rowmajor, dtype=np.float64).reshape(3, 3)
    Rc = R.T.reshape(-1)  # column-major
    for i in range(9):
        s.R[i] = Rc[i]
    for i in range(3):
        s.p[i] = 0.0 if p is None else float(p[i])
        s.v[i] = 0.0 if v is None else float(v[i])
        s.g[i] = g[i]
    if cov is None:  # IMUST::setZero, types.hpp:101-112
        cov = np.eye(15) * 1e-4
        cov[9:, 9:] = np.eye(6) * 1e-5
    cc = np.asarray(cov, dtype=np.float64).T.reshape(-1)
    for i in range(225):
        s.cov[i] = cc[i]
    return s


def state_arrays(s: VoState):
    R = np.array(s.R[:]).reshape(3, 3).T  # back to row-major numpy
    cov = np.array(s.cov[:]).reshape(15, 15).T
    return dict(t=s.t, R=R, p=np.array(s.p[:]), v=np.array(s.v[:]), bg=np.array(s.bg[:]), ba=np.array(s.ba[:]),
                g=np.array(s.g[:]), cov=cov)


class Odom:
    """One sequence's oracle context (vo::Odom)."""

    def __init__(self, cfg, fast: bool = False, ref: bool = False):
        self.lib = load(fast, ref)
        self.is_ref = ref
        self.cfg = cfg
        self._c = make_config(cfg)
        self.h = C.c_void_p(self.lib.vo_odom_create(C.byref(self._c)))
        if not self.h:
            raise RuntimeError("the reference build keeps its configuration in globals: one instance per process")

    def close(self):
        if self.h:
            self.lib.vo_odom_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_state(self, s: VoState):
        self.lib.vo_odom_set_state(self.h, C.byref(s))

    def get_state(self) -> VoState:
        s = VoState()
        self.lib.vo_odom_get_state(self.h, C.byref(s))
        return s

    def set_imu_anchor(self, last_end: float, last_imu7, scale_gravity: float = 1.0):
        a = np.ascontiguousarray(last_imu7, dtype=np.float64)
        self.lib.vo_odom_set_imu_anchor(self.h, C.c_double(last_end), _dp(a), C.c_double(scale_gravity))

    def bootstrap(self, xyz4: np.ndarray, state: VoState):
        a = np.ascontiguousarray(xyz4, dtype=np.float32)
        self.lib.vo_odom_bootstrap(self.h, _fp(a), C.c_int(a.shape[0]), C.byref(state))

    def step(self, xyz4: np.ndarray, beg_time: float, imu7: np.ndarray, iekf_on_full: bool = True, max_iter: int = 4):
        a = np.ascontiguousarray(xyz4, dtype=np.float32).copy()
        im = np.ascontiguousarray(imu7, dtype=np.float64)
        r = self.lib.vo_odom_step(self.h, _fp(a), C.c_int(a.shape[0]), C.c_double(beg_time), _dp(im),
                                  C.c_int(im.shape[0]), C.c_int(1 if iekf_on_full else 0), C.c_int(max_iter))
        return r, a

    def cold_start(self):
        """Forget the bootstrap: the next scans go through init_scan (VINA_SLAM::initialization)."""
        self.lib.vo_odom_cold_start(self.h)

    def init_scan(self, xyz4: np.ndarray, beg_time: float, imu7: np.ndarray) -> int:
        """One scan of the start-up phase (node.cpp:293-366 + local_mapping.cpp:362-388): 0 = collecting,
        1 = initialised (go on with step), -1 = motion_init failed, system reset."""
        a = np.ascontiguousarray(xyz4, dtype=np.float32)
        im = np.ascontiguousarray(imu7, dtype=np.float64)
        return self.lib.vo_odom_init_scan(self.h, _fp(a), C.c_int(a.shape[0]), C.c_double(beg_time), _dp(im),
                                          C.c_int(im.shape[0]))

    def stage_times(self):
        t = np.zeros(4)
        self.lib.vo_odom_stage_times(self.h, _dp(t))
        return t

    def last_iters(self) -> int:
        return self.lib.vo_odom_last_iters(self.h)

    def last_down(self) -> np.ndarray:
        n = self.lib.vo_odom_last_down(self.h, None, C.c_int(0))
        a = np.zeros((n, 4), dtype=np.float32)
        self.lib.vo_odom_last_down(self.h, _fp(a), C.c_int(n))
        return a

    def propagate(self, beg: float, end: float, imu7: np.ndarray) -> int:
        im = np.ascontiguousarray(imu7, dtype=np.float64)
        return self.lib.vo_odom_propagate(self.h, C.c_double(beg), C.c_double(end), _dp(im), C.c_int(im.shape[0]))

    def imu_poses(self) -> np.ndarray:
        n = self.lib.vo_odom_imu_poses(self.h, None, C.c_int(0))
        a = np.zeros((n, 22))
        self.lib.vo_odom_imu_poses(self.h, _dp(a), C.c_int(n))
        return a

    def deskew(self, xyz4: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(xyz4, dtype=np.float32).copy()
        self.lib.vo_odom_deskew(self.h, _fp(a), C.c_int(a.shape[0]))
        return a

    def motion_blur(self, xyz4: np.ndarray, beg: float, end: float, imu7: np.ndarray):
        a = np.ascontiguousarray(xyz4, dtype=np.float32).copy()
        im = np.ascontiguousarray(imu7, dtype=np.float64)
        r = self.lib.vo_odom_motion_blur(self.h, _fp(a), C.c_int(a.shape[0]), C.c_double(beg), C.c_double(end), _dp(im),
                                         C.c_int(im.shape[0]))
        return r, a

    def match(self, wld: np.ndarray, var: np.ndarray):
        w = np.ascontiguousarray(wld, dtype=np.float64)
        v = np.ascontiguousarray(var, dtype=np.float64)
        n = w.shape[0]
        flags = np.zeros(n, dtype=np.uint8)
        sigma = np.zeros(n)
        centers = np.zeros((n, 3))
        self.lib.vo_odom_match(self.h, C.c_int(n), _dp(w), _dp(v), _ptr(flags, C.c_uint8), _dp(sigma), _dp(centers))
        return flags, sigma, centers

    def set_dump(self, on: bool):
        self.lib.vo_odom_set_dump(self.h, C.c_int(1 if on else 0))

    def iekf(self, pnt: np.ndarray, var: np.ndarray, max_iter: int = 4) -> int:
        p = np.ascontiguousarray(pnt, dtype=np.float64)
        v = np.ascontiguousarray(var, dtype=np.float64)
        return self.lib.vo_odom_iekf(self.h, C.c_int(p.shape[0]), _dp(p), _dp(v), C.c_int(max_iter))

    def iter_dump(self, it: int, n: int):
        HTH, HTz, nnt = np.zeros(36), np.zeros(6), np.zeros(9)
        mn = C.c_int32(0)
        keys = np.zeros((n, 3), dtype=np.int64)
        codes = np.zeros(n, dtype=np.int32)
        flags = np.zeros(n, dtype=np.uint8)
        sigma = np.zeros(n)
        R, p = np.zeros(9), np.zeros(3)
        r = self.lib.vo_odom_iter_dump(self.h, C.c_int(it), _dp(HTH), _dp(HTz), _dp(nnt), C.byref(mn),
                                       _ptr(keys, C.c_int64), _ptr(codes, C.c_int32), _ptr(flags, C.c_uint8),
                                       _dp(sigma), _dp(R), _dp(p))
        if r < 0:
            return None
        return dict(HTH=HTH.reshape(6, 6).T, HTz=HTz, nnt=nnt.reshape(3, 3).T, match_num=mn.value, keys=keys,
                    codes=codes, flags=flags, sigma=sigma, R_col=R, p=p)

    def map_update(self, pnt: np.ndarray, var: np.ndarray):
        p = np.ascontiguousarray(pnt, dtype=np.float64)
        v = np.ascontiguousarray(var, dtype=np.float64)
        self.lib.vo_odom_map_update(self.h, C.c_int(p.shape[0]), _dp(p), _dp(v))

    def map_count(self):
        nr, ns = C.c_int64(0), C.c_int64(0)
        n = self.lib.vo_odom_map_count(self.h, C.byref(nr), C.byref(ns))
        return n, nr.value, ns.value

    def map_export(self) -> np.ndarray:
        n, _, _ = self.map_count()
        out = np.zeros(n, dtype=NODE_DTYPE)
        self.lib.vo_odom_map_export(self.h, out.ctypes.data_as(C.c_void_p), C.c_int64(n))
        return out

    def window(self):
        wc = C.c_int(0)
        mp = np.zeros(16, dtype=np.int32)
        ws = self.lib.vo_odom_window(self.h, C.byref(wc), _ptr(mp, C.c_int), C.c_int(16))
        return wc.value, mp[:ws].copy()

    def journey(self):
        """(jour, release_flag) of the per-scan loop (local_mapping.cpp:509-519)."""
        j, f = C.c_double(0), C.c_int(0)
        self.lib.vo_odom_journey(self.h, C.byref(j), C.byref(f))
        return j.value, bool(f.value)

    def idle(self, horizon: int = 700):
        """The `release_flag` branch of the idle path (local_mapping.cpp:317-341); (roots erased, nodes freed)."""
        nf = C.c_int(0)
        r = self.lib.vo_odom_idle(self.h, C.c_int(horizon), C.byref(nf))
        return int(r), nf.value

    def set_ba(self, on: bool = True, imu_coef: float = 0.0):
        """if_BA (local_mapping.cpp:492-497): LI_BA_Optimizer after every recut with a full window."""
        self.lib.vo_odom_set_ba(self.h, C.c_int(1 if on else 0), C.c_double(imu_coef))

    def ba_stats(self):
        a, b = C.c_int(0), C.c_int(0)
        self.lib.vo_odom_ba_stats(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    # ---- BA probe: LidarFactor (factors.cpp:22-158) on the factors captured by the last full-window map update
    def ba_probe(self, on: bool = True):
        self.lib.vo_odom_ba_probe(self.h, C.c_int(1 if on else 0))

    def ba_count(self) -> int:
        return int(self.lib.vo_odom_ba_count(self.h))

    def ba_poses(self) -> np.ndarray:
        """(win, 12): R column-major (9), p (3) of every window frame at capture time."""
        out = np.zeros((16, 12), dtype=np.float64)
        n = self.lib.vo_odom_ba_poses(self.h, _ptr(out, C.c_double), C.c_int(16))
        return out[:n].copy()

    def ba_hess(self, poses12: np.ndarray):
        """acc_evaluate2 over all factors: (Hess (6w,6w), JacT (6w,), residual)."""
        ps = np.ascontiguousarray(poses12, dtype=np.float64)
        w = ps.shape[0]
        H = np.zeros(36 * w * w, dtype=np.float64)
        J = np.zeros(6 * w, dtype=np.float64)
        r = C.c_double(0)
        rc = self.lib.vo_odom_ba_hess(self.h, _ptr(ps, C.c_double), C.c_int(w), _ptr(H, C.c_double), _ptr(J, C.c_double),
                                      C.byref(r))
        assert rc == 0
        return H.reshape(6 * w, 6 * w).T.copy(), J, r.value

    def ba_residual(self, poses12: np.ndarray):
        """evaluate_only_residual over all factors (updates the captured factors like the reference's container):
        (residual, lambda_0 of every factor)."""
        ps = np.ascontiguousarray(poses12, dtype=np.float64)
        n = self.ba_count()
        lam = np.zeros(max(n, 1), dtype=np.float64)
        r = C.c_double(0)
        rc = self.lib.vo_odom_ba_residual(self.h, _ptr(ps, C.c_double), C.c_int(ps.shape[0]), C.byref(r),
                                          _ptr(lam, C.c_double), C.c_int(n))
        assert rc == 0
        return r.value, lam[:n]


# ---- stateless helpers -----------------------------------------------------
def eig3(A: np.ndarray):
    lib = load()
    a = np.ascontiguousarray(np.asarray(A, dtype=np.float64).T.reshape(-1))
    vals, vecs = np.zeros(3), np.zeros(9)
    lib.vo_eig3(_dp(a), _dp(vals), _dp(vecs))
    return vals, vecs.reshape(3, 3).T


def inverse15(A: np.ndarray) -> np.ndarray:
    lib = load()
    a = np.ascontiguousarray(np.asarray(A, dtype=np.float64).T.reshape(-1))
    out = np.zeros(225)
    lib.vo_inverse15(_dp(a), _dp(out))
    return out.reshape(15, 15).T


def var_init(xyz4: np.ndarray, cfg, ref: bool = False):
    lib = load(ref=ref)
    a = np.ascontiguousarray(xyz4, dtype=np.float32)
    n = a.shape[0]
    pnt, var = np.zeros((n, 3)), np.zeros((n, 9))
    R = cfg.ext_R_colmajor()
    t = np.asarray(cfg.ext_t, dtype=np.float64)
    lib.vo_var_init(C.c_int(n), _fp(a), _dp(R), _dp(t), C.c_double(cfg.dept_err), C.c_double(cfg.beam_err),
                    _dp(pnt), _dp(var))
    return pnt, var


def pvec_update(pnt, var, R_col, p, cov_col, ref: bool = False):
    lib = load(ref=ref)
    pn = np.ascontiguousarray(pnt, dtype=np.float64)
    vr = np.ascontiguousarray(var, dtype=np.float64).copy()
    pw = np.zeros_like(pn)
    lib.vo_pvec_update(C.c_int(pn.shape[0]), _dp(pn), _dp(vr), _dp(np.ascontiguousarray(R_col)),
                       _dp(np.ascontiguousarray(p)), _dp(np.ascontiguousarray(cov_col)), _dp(pw))
    return vr, pw


def voxel_keys(pw: np.ndarray, voxel_size: float) -> np.ndarray:
    lib = load()
    a = np.ascontiguousarray(pw, dtype=np.float64)
    k = np.zeros((a.shape[0], 3), dtype=np.int64)
    lib.vo_voxel_keys(C.c_int(a.shape[0]), _dp(a), C.c_double(voxel_size), _ptr(k, C.c_int64))
    return k


def down_sampling_voxel(xyz4: np.ndarray, voxel_size: float, ref: bool = False) -> np.ndarray:
    lib = load(ref=ref)
    a = np.ascontiguousarray(xyz4, dtype=np.float32)
    out = np.zeros_like(a)
    n = lib.vo_down_sampling_voxel(C.c_int(a.shape[0]), _fp(a), C.c_double(voxel_size), _fp(out))
    return out[:n].copy()


class Sync:
    """sync_packages and its buffers (src/sensor/sync.cpp:5-96)."""

    def __init__(self, point_notime: int = 0, ref: bool = False):
        """ref: the reference's own sync.cpp (oracle/_ref) - its state is global, ONE instance per process."""
        self.lib = load(ref=ref)
        self.h = C.c_void_p(self.lib.vo_sync_create(C.c_int(point_notime)))
        if not self.h:
            raise RuntimeError("the reference's sync_packages keeps its state in globals: one instance per process")

    def close(self):
        if self.h:
            self.lib.vo_sync_destroy(self.h)
            self.h = None

    def push_imu(self, imu7):
        a = np.ascontiguousarray(imu7, dtype=np.float64)
        self.lib.vo_sync_push_imu(self.h, _dp(a))

    def push_scan(self, t_start: float, t_last: float, tag: int):
        self.lib.vo_sync_push_scan(self.h, C.c_double(t_start), C.c_double(t_last), C.c_int64(tag))

    def next(self, cap: int = 256):
        """(code, tag, beg, end, imu7[m, 7])"""
        tag, beg, end, m = C.c_int64(-1), C.c_double(0), C.c_double(0), C.c_int(0)
        buf = np.zeros((cap, 7), dtype=np.float64)
        r = self.lib.vo_sync_next(self.h, C.byref(tag), C.byref(beg), C.byref(end), _dp(buf), C.c_int(cap), C.byref(m))
        return r, tag.value, beg.value, end.value, buf[:m.value].copy()


def scan_prepare(xyz4: np.ndarray, point_filter_num: int, blind2: float, fast: bool = False, ref: bool = False) -> np.ndarray:
    """Decoder keep rule + pcl_handler (filter, stable sort by time offset, cut at 0.11 s); None where the reference
    would be left with an empty cloud."""
    lib = load(fast=fast, ref=ref)  # ref: the reference's own pcl_handler + velodyne_handler (oracle/_ref)
    a = np.ascontiguousarray(xyz4, dtype=np.float32).reshape(-1, 4)
    out = np.zeros((max(a.shape[0], 2), 4), dtype=np.float32)
    n = lib.vo_scan_prepare(C.c_int(a.shape[0]), _fp(a), C.c_int(point_filter_num), C.c_double(blind2), _fp(out))
    return None if n < 0 else out[:n].copy()


def exp_so3(w, dt=None, ref: bool = False):
    lib = load(ref=ref)
    R = np.zeros(9)
    w = np.ascontiguousarray(w, dtype=np.float64)
    if dt is None:
        lib.vo_exp(_dp(w), _dp(R))
    else:
        lib.vo_exp_dt(_dp(w), C.c_double(dt), _dp(R))
    return R.reshape(3, 3).T


def log_so3(R, ref: bool = False):
    lib = load(ref=ref)
    a = np.ascontiguousarray(np.asarray(R, dtype=np.float64).T.reshape(-1))
    w = np.zeros(3)
    lib.vo_log(_dp(a), _dp(w))
    return w


# ---- the decoders' handlers (src/sensor/lidar_pointcloud_decoder.cpp:55-240), restated in numpy; pinned against the
# reference's own file compiled into oracle/_ref (decode_handler_ref below; tests/test_oracle_vs_ref.py).
def decode_handler(lidar_type: int, pts: np.ndarray, header_stamp: float, blind2: float, point_filter_num: int,
                   omega_l: float = 3610.0) -> np.ndarray:
    """pts: structured array with float32 x, y, z and the handler's time field `t` (float32 seconds for Velodyne,
    uint32 ns for Ouster / Livox, float64 seconds for Hesai / RoboSense; absent for TartanAir), i.e. what
    pcl::fromROSMsg leaves in the handler's point struct. Returns (n_kept, 4) float32 = x, y, z, curvature."""
    n = pts.shape[0]
    x, y, z = (pts[k].astype(np.float32) for k in ("x", "y", "z"))
    i = np.arange(n)
    r2 = (x * x + y * y) + z * z  # float32, left to right
    dec = (i % point_filter_num) == 0
    if n == 0:
        return np.zeros((0, 4), dtype=np.float32)
    if lidar_type == 5:  # tartanair_handler (:228-239)
        c, keep = np.zeros(n, dtype=np.float32), np.ones(n, dtype=bool)
    elif lidar_type == 0:  # livox_handler (:55-75)
        c = (pts["t"].astype(np.float64) * 1e-9).astype(np.float32)
        keep = dec & (r2.astype(np.float64) > blind2)
    elif lidar_type == 2:  # ouster_handler (:143-165)
        c = (pts["t"].astype(np.float64) / 1e9).astype(np.float32)
        keep = dec & (r2.astype(np.float64) > blind2)
    elif lidar_type == 3:  # hesai_handler (:167-194)
        c = (pts["t"].astype(np.float64) - float(pts["t"][0])).astype(np.float32)
        keep = dec & (r2.astype(np.float64) > blind2)
    elif lidar_type == 4:  # robosense_handler (:196-224): planar blind test
        c = (pts["t"].astype(np.float64) - header_stamp).astype(np.float32)
        keep = dec & ((x * x + y * y).astype(np.float64) > blind2)
    elif lidar_type == 1:  # velodyne_handler (:77-141)
        tl = np.float32(pts["t"][-1])
        if tl > 0.01 and tl < 0.12:
            c = pts["t"].astype(np.float32)
            keep = dec & (r2.astype(np.float64) > blind2)
        else:
            out, first, yaw0, yaw_last, bias, cool = [], True, 0.0, 0.0, 0.0, 0
            for k in range(n):
                if abs(float(x[k])) < 0.1:
                    continue
                yaw = float(np.arctan2(y[k], x[k])) * 57.2957795 - bias
                if first:
                    yaw0 = yaw_last = yaw
                    first = False
                if float(r2[k]) < blind2:
                    continue
                if (yaw - yaw_last) > 180:
                    cool -= 1
                    if cool + 1 <= 0:
                        bias += 360
                        yaw -= 360
                        cool = 1000
                if abs(yaw - yaw_last) > 180:
                    yaw += 360
                cur = np.float32((yaw0 - yaw) / omega_l)
                yaw_last = yaw
                if cur >= 0 and cur < 0.1 and (k % point_filter_num) == 0:
                    out.append((x[k], y[k], z[k], cur))
            return np.array(out, dtype=np.float32).reshape(-1, 4)
    else:
        raise ValueError("Unsupported lidar type")
    return np.stack([x, y, z, c], axis=1)[keep].astype(np.float32)


def decode_handler_ref(lidar_type: int, data: bytes, n_points: int, point_step: int, off_xyz, off_t: int, t_datatype: int,
                       header_stamp: float, blind2: float, point_filter_num: int, omega_l: float = 3610.0) -> np.ndarray:
    """The reference's own LidarPointCloudDecoder::process (oracle/_ref, compiled unmodified; pcl::fromROSMsg from the
    shim) on the bytes of a PointCloud2."""
    lib = load(ref=True)
    lib.vo_decode_handler.restype = C.c_int64
    out = np.zeros((max(n_points, 1), 4), dtype=np.float32)
    buf = (C.c_uint8 * max(len(data), 1)).from_buffer_copy(data if len(data) else b"\0")
    r = lib.vo_decode_handler(C.c_int(lidar_type), buf, C.c_int64(n_points), C.c_int(point_step), C.c_int(off_xyz[0]),
                              C.c_int(off_xyz[1]), C.c_int(off_xyz[2]), C.c_int(off_t), C.c_int(t_datatype),
                              C.c_double(header_stamp), C.c_double(omega_l), C.c_double(blind2), C.c_int(point_filter_num),
                              _fp(out), C.c_int64(out.shape[0]))
    assert r >= 0
    return out[:r].copy()


def decode_livox_ref(offset_time: np.ndarray, xyz: np.ndarray, blind2: float, point_filter_num: int) -> np.ndarray:
    lib = load(ref=True)
    lib.vo_decode_livox.restype = C.c_int64
    t = np.ascontiguousarray(offset_time, dtype=np.uint32)
    p = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
    out = np.zeros((max(t.shape[0], 1), 4), dtype=np.float32)
    r = lib.vo_decode_livox(t.ctypes.data_as(C.c_void_p), _fp(p), C.c_int64(t.shape[0]), C.c_double(blind2),
                            C.c_int(point_filter_num), _fp(out), C.c_int64(out.shape[0]))
    assert r >= 0
    return out[:r].copy()

"""ctypes binding of libvina_b200.so (include/vina_b200.h) — the call a Python user makes.

There is no CPU fallback: importing works anywhere (so the symbol table can be
checked on a CPU box), but creating a context without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvina_b200.so")

VINA_MAX_WIN = 10


class VinaConfig(C.Structure):
    _fields_ = [
        ("voxel_size", C.c_double), ("min_eigen_value", C.c_double), ("plane_eigen_value_thre", C.c_double * 4),
        ("min_point", C.c_double * 4), ("dept_err", C.c_double), ("beam_err", C.c_double), ("down_size", C.c_double),
        ("ext_R", C.c_double * 9), ("ext_t", C.c_double * 3), ("cov_gyr", C.c_double), ("cov_acc", C.c_double),
        ("rdw_gyr", C.c_double), ("rdw_acc", C.c_double), ("max_layer", C.c_int32), ("max_points", C.c_int32),
        ("win_size", C.c_int32), ("thread_num", C.c_int32), ("max_scan_points", C.c_int32), ("max_nodes", C.c_int32),
        ("hash_capacity_log2", C.c_int32), ("device", C.c_int32), ("fix_pool_points", C.c_int64),
        ("win_pool_points", C.c_int64),
    ]


class VinaState(C.Structure):
    _fields_ = [
        ("t", C.c_double), ("R", C.c_double * 9), ("p", C.c_double * 3), ("v", C.c_double * 3),
        ("bg", C.c_double * 3), ("ba", C.c_double * 3), ("g", C.c_double * 3), ("cov", C.c_double * 225),
    ]


class VinaTimings(C.Structure):
    _fields_ = [
        ("deskew_ms", C.c_float), ("downsample_ms", C.c_float), ("var_init_ms", C.c_float), ("iekf_ms", C.c_float),
        ("insert_ms", C.c_float), ("recut_ms", C.c_float), ("margi_ms", C.c_float), ("total_ms", C.c_float),
        ("iekf_kernel_ms", C.c_float), ("iekf_iters", C.c_int32), ("kernel_launches", C.c_int32),
    ]


IMU_DTYPE = np.dtype([("t", "<f8"), ("gyr", "<f8", 3), ("acc", "<f8", 3)])
IMU_POSE_DTYPE = np.dtype([("t", "<f8"), ("R", "<f8", 9), ("p", "<f8", 3), ("v", "<f8", 3), ("w", "<f8", 3),
                           ("a", "<f8", 3)])
POSE_DTYPE = np.dtype([("R", "<f8", 9), ("p", "<f8", 3)])
NODE_DTYPE = np.dtype([
    ("key", "<i8", 3), ("code", "<i4"), ("layer", "<i4"), ("octo_state", "<i4"), ("isexist", "<i4"),
    ("has_sw", "<i4"), ("is_plane", "<i4"), ("last_num", "<i4"), ("opt_state", "<i4"), ("N_add", "<i4"),
    ("N_fix", "<i4"), ("n_point_fix", "<i4"), ("n_win_points", "<i4"), ("N_local", "<i4", 16),
    ("P_add", "<f8", 9), ("v_add", "<f8", 3), ("P_fix", "<f8", 9), ("v_fix", "<f8", 3), ("eig_value", "<f8", 3),
    ("eig_vector", "<f8", 9), ("center", "<f8", 3), ("normal", "<f8", 3), ("plane_var", "<f8", 36),
    ("radius", "<f8"), ("cov_add", "<f8", 81), ("voxel_center", "<f8", 3), ("quater_length", "<f8"),
], align=True)

# every symbol include/vina_b200.h declares
EXPORTS = [
    "vina_odom_cold_start", "vina_odom_init_scan", "vina_map_last_counts",
    "vina_config_default", "vina_ctx_create", "vina_ctx_destroy", "vina_last_error", "vina_ctx_set_stream",
    "vina_ctx_sync", "vina_scan_upload", "vina_scan_upload_device", "vina_down_count", "vina_deskew", "vina_scan_download", "vina_downsample", "vina_down_upload",
    "vina_down_download", "vina_var_init", "vina_pvec_upload", "vina_pvec_download", "vina_iekf_begin",
    "vina_iekf_accumulate", "vina_iekf_accumulate_debug", "vina_iekf_debug_assoc", "vina_map_insert",
    "vina_map_recut", "vina_map_margi", "vina_map_shift_window", "vina_map_count", "vina_map_export",
    "vina_odom_set_state", "vina_odom_get_state", "vina_odom_set_imu_anchor", "vina_odom_bootstrap",
    "vina_odom_step", "vina_odom_step_resident", "vina_odom_propagate", "vina_odom_iekf", "vina_odom_iekf_host",
    "vina_odom_map_update",
    "vina_odom_window", "vina_get_timings", "vina_set_profiling", "vina_shard_owner", "vina_shard_route",
    "vina_shard_insert_begin", "vina_shard_insert_finish", "vina_batch_create", "vina_batch_destroy",
    "vina_batch_step_resident", "vina_batch_iekf_time", "vina_batch_sync", "vina_shard_query_route",
    "vina_shard_query_accumulate", "vina_odom_iekf_host_begin", "vina_odom_iekf_host_update",
    "vina_shard_p2p_create", "vina_shard_p2p_connect", "vina_shard_p2p_pointers", "vina_shard_p2p_connect_local",
    "vina_shard_route_p2p", "vina_shard_insert_begin_p2p", "vina_odom_iekf_sharded_p2p",
    "vina_set_overlap", "vina_set_upload_ordered", "vina_set_iekf_loop", "vina_ba_set_capture", "vina_ba_collect", "vina_ba_count", "vina_ba_lidar_hessian",
    "vina_ba_lidar_residual", "vina_odom_set_ba", "vina_odom_ba_stats",
    "vina_ba_imu_evaluate", "vina_ba_solve",
    "vina_map_set_journey", "vina_map_prune", "vina_odom_journey", "vina_odom_idle",
    "vina_scan_prepare", "vina_scan_prepare_device", "vina_odom_step_prepared",
    "vina_sync_create", "vina_sync_destroy", "vina_sync_push_imu", "vina_sync_push_scan", "vina_sync_pending",
    "vina_sync_next", "vina_decode_pointcloud2", "vina_decode_livox", "vina_scan_last_stamp",
]
SHARD_IEKF_ALL, SHARD_IEKF_STAGE, SHARD_IEKF_ROUTE, SHARD_IEKF_SEND, SHARD_IEKF_EVAL, SHARD_IEKF_SOLVE, SHARD_IEKF_FINISH = range(7)
SHARD_RECORD_DOUBLES = 13
SHARD_QUERY_DOUBLES = 10


class VinaError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"vina_b200 error {code}: {msg}")
        self.code = code


_LIB = None


def load():
    """dlopen the in-tree library; fails loudly when it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -m vina_slam_b200.build` (nvcc, sm_100a). "
                          "There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.vina_last_error.restype = C.c_char_p
    lib.vina_ctx_destroy.restype = None
    lib.vina_batch_destroy.restype = None
    lib.vina_config_default.restype = None
    lib.vina_decode_pointcloud2.restype = C.c_int64
    lib.vina_decode_livox.restype = C.c_int64
    lib.vina_sync_destroy.restype = None
    lib.vina_sync_destroy.argtypes = [C.c_void_p]
    lib.vina_map_count.restype = C.c_int64
    lib.vina_map_export.restype = C.c_int64
    _LIB = lib
    return lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def make_config(cfg, **caps) -> VinaConfig:
    """cfg: synth.SensorConfig (duck-typed). caps override the device capacities."""
    lib = load()
    c = VinaConfig()
    lib.vina_config_default(C.byref(c))
    c.voxel_size = cfg.voxel_size
    c.min_eigen_value = cfg.min_eigen_value
    for i in range(4):
        c.plane_eigen_value_thre[i] = cfg.plane_thre[i]
        c.min_point[i] = (20, 20, 15, 10)[i]
    c.dept_err, c.beam_err, c.down_size = cfg.dept_err, cfg.beam_err, cfg.down_size
    R = cfg.ext_R_colmajor()
    for i in range(9):
        c.ext_R[i] = R[i]
    for i in range(3):
        c.ext_t[i] = cfg.ext_t[i]
    c.cov_gyr, c.cov_acc, c.rdw_gyr, c.rdw_acc = cfg.cov_gyr, cfg.cov_acc, cfg.rdw_gyr, cfg.rdw_acc
    c.max_layer, c.max_points, c.win_size, c.thread_num = cfg.max_layer, cfg.max_points, cfg.win_size, cfg.thread_num
    for k, v in caps.items():
        setattr(c, k, v)
    return c


def make_state(R_rowmajor=None, p=None, v=None, t=0.0, cov=None, g=(0.0, 0.0, -9.8)) -> VinaState:
    s = VinaState()
    s.t = t
    R = np.eye(3) if R_rowmajor is None else np.asarray(R_rowmajor, dtype=np.float64).reshape(3, 3)
    Rc = R.T.reshape(-1)
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


def state_arrays(s: VinaState):
    return dict(t=s.t, R=np.array(s.R[:]).reshape(3, 3).T, p=np.array(s.p[:]), v=np.array(s.v[:]),
                bg=np.array(s.bg[:]), ba=np.array(s.ba[:]), g=np.array(s.g[:]),
                cov=np.array(s.cov[:]).reshape(15, 15).T)


def imu_array(imu7: np.ndarray) -> np.ndarray:
    a = np.zeros(imu7.shape[0], dtype=IMU_DTYPE)
    a["t"] = imu7[:, 0]
    a["gyr"] = imu7[:, 1:4]
    a["acc"] = imu7[:, 4:7]
    return a


class Ctx:
    """One sequence on one GPU (vina_ctx)."""

    def __init__(self, cfg, **caps):
        self.lib = load()
        self.cfg = cfg
        self._c = make_config(cfg, **caps)
        h = C.c_void_p()
        r = self.lib.vina_ctx_create(C.byref(self._c), C.byref(h))
        self.h = h
        if r != 0:
            msg = self.lib.vina_last_error(h).decode() if h else "no CUDA device (there is no CPU fallback)"
            if h:
                self.lib.vina_ctx_destroy(h)
            self.h = None
            raise VinaError(r, msg)

    def close(self):
        if getattr(self, "h", None):
            self.lib.vina_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, r):
        if r < 0:
            raise VinaError(int(r), self.lib.vina_last_error(self.h).decode())
        return r

    # ---- plumbing
    def set_stream(self, cuda_stream_ptr: int):
        self._ck(self.lib.vina_ctx_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def sync(self):
        self._ck(self.lib.vina_ctx_sync(self.h))

    def set_profiling(self, on: bool):
        self._ck(self.lib.vina_set_profiling(self.h, C.c_int(1 if on else 0)))

    def timings(self) -> VinaTimings:
        t = VinaTimings()
        self._ck(self.lib.vina_get_timings(self.h, C.byref(t)))
        return t

    # ---- scan stages
    def scan_upload(self, xyzt: np.ndarray):
        a = np.ascontiguousarray(xyzt, dtype=np.float32)
        self._keep = a
        self._ck(self.lib.vina_scan_upload(self.h, _fp(a), C.c_int(a.shape[0])))
        self.sync()

    def deskew(self, poses: np.ndarray, R_end_col, p_end):
        ps = np.ascontiguousarray(poses, dtype=IMU_POSE_DTYPE)
        R = np.ascontiguousarray(R_end_col, dtype=np.float64)
        p = np.ascontiguousarray(p_end, dtype=np.float64)
        self._ck(self.lib.vina_deskew(self.h, ps.ctypes.data_as(C.c_void_p), C.c_int(ps.shape[0]), _dp(R), _dp(p)))

    def scan_download(self, n: int) -> np.ndarray:
        a = np.zeros((n, 4), dtype=np.float32)
        k = self._ck(self.lib.vina_scan_download(self.h, _fp(a), C.c_int(n)))
        return a[:k]

    def scan_prepare(self, xyzt: np.ndarray, point_filter_num: int, blind2: float, d_ptr: int = 0):
        """Decoder keep rule + pcl_handler on the device (filter, stable sort by time offset, cut at 0.11 s); the
        result becomes the context's scan. Returns (points, last time offset). d_ptr: raw points already in HBM."""
        n, t = C.c_int(0), C.c_float(0)
        if d_ptr:
            self._ck(self.lib.vina_scan_prepare_device(self.h, C.c_void_p(d_ptr), C.c_int(int(xyzt)), C.c_int(point_filter_num),
                                                       C.c_double(blind2), C.byref(n), C.byref(t)))
        else:
            a = np.ascontiguousarray(xyzt, dtype=np.float32).reshape(-1, 4)
            self._ck(self.lib.vina_scan_prepare(self.h, _fp(a), C.c_int(a.shape[0]), C.c_int(point_filter_num),
                                                C.c_double(blind2), C.byref(n), C.byref(t)))
        return n.value, t.value

    def step_prepared(self, beg_time: float, imu7: np.ndarray, iekf_on_full: bool = True, max_iter: int = 4):
        """vina_odom_step on the scan left by scan_prepare."""
        im = imu_array(np.asarray(imu7, dtype=np.float64))
        out = VinaState()
        self._ck(self.lib.vina_odom_step_prepared(self.h, C.c_double(beg_time), im.ctypes.data_as(C.c_void_p),
                                                  C.c_int(im.shape[0]), C.c_int(1 if iekf_on_full else 0),
                                                  C.c_int(max_iter), C.byref(out)))
        return out

    def scan_upload_device(self, d_ptr: int, n: int):
        self._ck(self.lib.vina_scan_upload_device(self.h, C.c_void_p(d_ptr), C.c_int(n)))

    def downsample(self):
        self._ck(self.lib.vina_downsample(self.h))

    def n_down(self) -> int:
        return self._ck(self.lib.vina_down_count(self.h))

    def down_upload(self, xyzt: np.ndarray):
        a = np.ascontiguousarray(xyzt, dtype=np.float32)
        self._ck(self.lib.vina_down_upload(self.h, _fp(a), C.c_int(a.shape[0])))
        self.sync()

    def down_download(self, cap: int) -> np.ndarray:
        a = np.zeros((cap, 4), dtype=np.float32)
        k = self._ck(self.lib.vina_down_download(self.h, _fp(a), C.c_int(cap)))
        return a[:k]

    def var_init(self, which: int):
        self._ck(self.lib.vina_var_init(self.h, C.c_int(which)))

    def pvec_upload(self, which: int, pnt: np.ndarray, var: np.ndarray):
        p = np.ascontiguousarray(pnt, dtype=np.float64)
        v = np.ascontiguousarray(var, dtype=np.float64)
        self._ck(self.lib.vina_pvec_upload(self.h, C.c_int(which), _dp(p), _dp(v), C.c_int(p.shape[0])))

    def pvec_download(self, which: int, cap: int):
        p = np.zeros((cap, 3))
        v = np.zeros((cap, 9))
        k = self._ck(self.lib.vina_pvec_download(self.h, C.c_int(which), _dp(p), _dp(v), C.c_int(cap)))
        return p[:k], v[:k]

    # ---- IEKF
    def iekf_begin(self, which: int, rot_var_col, tsl_var_col):
        a = np.ascontiguousarray(rot_var_col, dtype=np.float64)
        b = np.ascontiguousarray(tsl_var_col, dtype=np.float64)
        self._ck(self.lib.vina_iekf_begin(self.h, C.c_int(which), _dp(a), _dp(b)))

    def iekf_accumulate(self, R_col, p, debug: bool = False):
        R = np.ascontiguousarray(R_col, dtype=np.float64)
        pp = np.ascontiguousarray(p, dtype=np.float64)
        HTH, HTz, nnt = np.zeros(36), np.zeros(6), np.zeros(9)
        mn = C.c_int32(0)
        fn = self.lib.vina_iekf_accumulate_debug if debug else self.lib.vina_iekf_accumulate
        self._ck(fn(self.h, _dp(R), _dp(pp), _dp(HTH), _dp(HTz), _dp(nnt), C.byref(mn)))
        return dict(HTH=HTH.reshape(6, 6).T, HTz=HTz, nnt=nnt.reshape(3, 3).T, match_num=mn.value)

    def iekf_debug_assoc(self, n: int):
        keys = np.zeros((n, 3), dtype=np.int64)
        codes = np.zeros(n, dtype=np.int32)
        flags = np.zeros(n, dtype=np.uint8)
        sigma = np.zeros(n)
        self._ck(self.lib.vina_iekf_debug_assoc(self.h, keys.ctypes.data_as(C.c_void_p),
                                                codes.ctypes.data_as(C.c_void_p), flags.ctypes.data_as(C.c_void_p),
                                                _dp(sigma), C.c_int(n)))
        return dict(keys=keys, codes=codes, flags=flags, sigma=sigma)

    # ---- map
    def map_count(self):
        nr, ns = C.c_int64(0), C.c_int64(0)
        n = self._ck(self.lib.vina_map_count(self.h, C.byref(nr), C.byref(ns)))
        return int(n), nr.value, ns.value

    def map_export(self) -> np.ndarray:
        n, _, _ = self.map_count()
        out = np.zeros(max(n, 1), dtype=NODE_DTYPE)
        k = self._ck(self.lib.vina_map_export(self.h, out.ctypes.data_as(C.c_void_p), C.c_int64(out.shape[0])))
        return out[:k]

    # ---- map pruning behind the vehicle (local_mapping.cpp:317-341)
    def map_set_journey(self, jour: float):
        """The `jour` argument of multi_margi (local_mapping.cpp:36, 507) for the next map_margi."""
        self._ck(self.lib.vina_map_set_journey(self.h, C.c_double(jour)))

    def map_prune(self, jour: float, horizon: int = 700):
        """Erase every root voxel with (int)(jour - root.jour) >= horizon; returns (roots erased, nodes freed)."""
        a, b = C.c_int64(0), C.c_int64(0)
        self._ck(self.lib.vina_map_prune(self.h, C.c_double(jour), C.c_int(horizon), C.byref(a), C.byref(b)))
        return a.value, b.value

    def journey(self):
        """(jour, release_flag) of the per-scan loop (local_mapping.cpp:509-519)."""
        j, f = C.c_double(0), C.c_int(0)
        self._ck(self.lib.vina_odom_journey(self.h, C.byref(j), C.byref(f)))
        return j.value, bool(f.value)

    def idle(self, horizon: int = 700):
        """The idle path of the loop (local_mapping.cpp:303-341): prune if release_flag is set."""
        a, b = C.c_int64(0), C.c_int64(0)
        self._ck(self.lib.vina_odom_idle(self.h, C.c_int(horizon), C.byref(a), C.byref(b)))
        return a.value, b.value

    def map_recut(self, win_count: int, x_buf: np.ndarray):
        xb = np.ascontiguousarray(x_buf, dtype=POSE_DTYPE)
        self._ck(self.lib.vina_map_recut(self.h, C.c_int(win_count), xb.ctypes.data_as(C.c_void_p)))

    def map_margi(self, win_count: int, x_buf: np.ndarray):
        xb = np.ascontiguousarray(x_buf, dtype=POSE_DTYPE)
        self._ck(self.lib.vina_map_margi(self.h, C.c_int(win_count), xb.ctypes.data_as(C.c_void_p)))
        self._ck(self.lib.vina_map_shift_window(self.h))

    def map_insert(self, win_ord: int, R_col, p, cov_rot_col, cov_tsl_col):
        a = [np.ascontiguousarray(v, dtype=np.float64) for v in (R_col, p, cov_rot_col, cov_tsl_col)]
        self._ck(self.lib.vina_map_insert(self.h, C.c_int(win_ord), _dp(a[0]), _dp(a[1]), _dp(a[2]), _dp(a[3])))

    # ---- map sharded by voxel-hash range (device pointers are caller-owned, e.g. torch tensors)
    def shard_route(self, world: int, first: int, count: int, index_base: int, R_col, p, cov_rot_col, cov_tsl_col,
                    d_send_ptr: int) -> np.ndarray:
        a = [np.ascontiguousarray(v, dtype=np.float64) for v in (R_col, p, cov_rot_col, cov_tsl_col)]
        counts = np.zeros(world, dtype=np.int32)
        self._ck(self.lib.vina_shard_route(self.h, C.c_int(world), C.c_int(first), C.c_int(count),
                                           C.c_int64(index_base), _dp(a[0]), _dp(a[1]), _dp(a[2]), _dp(a[3]),
                                           C.c_void_p(d_send_ptr), counts.ctypes.data_as(C.c_void_p)))
        return counts

    def shard_insert_begin(self, d_recv_ptr: int, n: int, win_ord: int):
        a, b = C.c_int32(0), C.c_int32(0)
        self._ck(self.lib.vina_shard_insert_begin(self.h, C.c_void_p(d_recv_ptr), C.c_int(n), C.c_int(win_ord),
                                                  C.byref(a), C.byref(b)))
        return a.value, b.value

    def shard_insert_finish(self, win_ord: int, global_roots: int, global_slide: int):
        self._ck(self.lib.vina_shard_insert_finish(self.h, C.c_int(win_ord), C.c_int(global_roots),
                                                   C.c_int(global_slide)))

    # ---- fused route + exchange over peer memory
    def shard_p2p_create(self, rank: int, world: int, inbox_records: int) -> bytes:
        buf = C.create_string_buffer(128)
        self._ck(self.lib.vina_shard_p2p_create(self.h, C.c_int(rank), C.c_int(world), C.c_int64(inbox_records), buf))
        return buf.raw

    def shard_p2p_connect(self, all_handles: bytes):
        self._ck(self.lib.vina_shard_p2p_connect(self.h, C.c_char_p(all_handles)))

    def shard_p2p_pointers(self):
        a, b = C.c_void_p(), C.c_void_p()
        self._ck(self.lib.vina_shard_p2p_pointers(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def shard_p2p_connect_local(self, inbox_ptrs, ctrl_ptrs):
        n = len(inbox_ptrs)
        ia = (C.c_void_p * n)(*inbox_ptrs)
        ca = (C.c_void_p * n)(*ctrl_ptrs)
        self._ck(self.lib.vina_shard_p2p_connect_local(self.h, ia, ca))

    def shard_route_p2p(self, first: int, count: int, index_base: int, R_col, p, cov_rot_col, cov_tsl_col, phase: int = 0):
        a = [np.ascontiguousarray(v, dtype=np.float64) for v in (R_col, p, cov_rot_col, cov_tsl_col)]
        self._ck(self.lib.vina_shard_route_p2p(self.h, C.c_int(first), C.c_int(count), C.c_int64(index_base), _dp(a[0]),
                                               _dp(a[1]), _dp(a[2]), _dp(a[3]), C.c_int(phase)))

    def shard_insert_begin_p2p(self, win_ord: int):
        n, a, b = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        self._ck(self.lib.vina_shard_insert_begin_p2p(self.h, C.c_int(win_ord), C.byref(n), C.byref(a), C.byref(b)))
        return n.value, a.value, b.value

    def shard_query_route(self, world: int, first: int, count: int, index_base: int, R_col, p, d_send_ptr: int):
        a = [np.ascontiguousarray(v, dtype=np.float64) for v in (R_col, p)]
        counts = np.zeros(world, dtype=np.int32)
        self._ck(self.lib.vina_shard_query_route(self.h, C.c_int(world), C.c_int(first), C.c_int(count),
                                                 C.c_int64(index_base), _dp(a[0]), _dp(a[1]), C.c_void_p(d_send_ptr),
                                                 counts.ctypes.data_as(C.c_void_p)))
        return counts

    def shard_query_accumulate(self, d_recv_ptr: int, n: int, R_col, p, rot_var_col, tsl_var_col, d_sums_ptr: int):
        a = [np.ascontiguousarray(v, dtype=np.float64) for v in (R_col, p, rot_var_col, tsl_var_col)]
        self._ck(self.lib.vina_shard_query_accumulate(self.h, C.c_void_p(d_recv_ptr), C.c_int(n), _dp(a[0]), _dp(a[1]),
                                                      _dp(a[2]), _dp(a[3]), C.c_void_p(d_sums_ptr)))

    def odom_iekf_host_begin(self, max_iter: int):
        self._ck(self.lib.vina_odom_iekf_host_begin(self.h, C.c_int(max_iter)))

    def odom_iekf_host_update(self, sums34) -> bool:
        a = np.ascontiguousarray(sums34, dtype=np.float64)
        return self._ck(self.lib.vina_odom_iekf_host_update(self.h, _dp(a))) == 1

    def set_ba(self, on: bool = True, imu_coef: float = 0.0):
        self._ck(self.lib.vina_odom_set_ba(self.h, C.c_int(1 if on else 0), C.c_double(imu_coef)))

    def ba_stats(self):
        a, b = C.c_int32(0), C.c_int32(0)
        self._ck(self.lib.vina_odom_ba_stats(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    # ---- BA LiDAR factor (factors.cpp:22-158 on the device factor store)
    def ba_set_capture(self, on: bool = True):
        self._ck(self.lib.vina_ba_set_capture(self.h, C.c_int(1 if on else 0)))

    def ba_collect(self) -> int:
        n = C.c_int32(0)
        self._ck(self.lib.vina_ba_collect(self.h, C.byref(n)))
        return n.value

    def ba_count(self) -> int:
        n = C.c_int32(0)
        self._ck(self.lib.vina_ba_count(self.h, C.byref(n)))
        return n.value

    def ba_hess(self, poses12):
        """poses12: (win, 12) = R column-major + p per window frame. Returns (Hess (6w, 6w), JacT, residual)."""
        ps = np.ascontiguousarray(poses12, dtype=np.float64)
        w = ps.shape[0]
        H = np.zeros(36 * w * w, dtype=np.float64)
        J = np.zeros(6 * w, dtype=np.float64)
        r = C.c_double(0)
        self._ck(self.lib.vina_ba_lidar_hessian(self.h, ps.ctypes.data_as(C.c_void_p), C.c_int(w), _dp(H), _dp(J), C.byref(r)))
        return H.reshape(6 * w, 6 * w).T.copy(), J, r.value

    def ba_residual(self, poses12):
        """Returns (residual, lambda_0 per factor); overwrites the stored factors' eig / pcr_add."""
        ps = np.ascontiguousarray(poses12, dtype=np.float64)
        n = self.ba_count()
        lam = np.zeros(max(n, 1), dtype=np.float64)
        r = C.c_double(0)
        self._ck(self.lib.vina_ba_lidar_residual(self.h, ps.ctypes.data_as(C.c_void_p), C.c_int(ps.shape[0]), C.byref(r),
                                                 _dp(lam), C.c_int(n)))
        return r.value, lam[:n]

    def set_overlap(self, on: bool):
        self._ck(self.lib.vina_set_overlap(self.h, C.c_int(1 if on else 0)))

    def set_upload_ordered(self, on: bool):
        self._ck(self.lib.vina_set_upload_ordered(self.h, C.c_int(1 if on else 0)))

    def set_iekf_loop(self, on: bool):
        self._ck(self.lib.vina_set_iekf_loop(self.h, C.c_int(1 if on else 0)))

    def odom_iekf_sharded_p2p(self, first: int, count: int, max_iter: int, phase: int = 0):
        """vina_odom_iekf_sharded_p2p; returns (iterations, not_degenerate) - meaningful for phase ALL / FINISH."""
        it, ok = C.c_int(0), C.c_int(0)
        self._ck(self.lib.vina_odom_iekf_sharded_p2p(self.h, C.c_int(first), C.c_int(count), C.c_int(max_iter),
                                                     C.c_int(phase), C.byref(it), C.byref(ok)))
        return it.value, ok.value

    # ---- odometry (host pipeline inside the library)
    def set_state(self, s: VinaState):
        self._ck(self.lib.vina_odom_set_state(self.h, C.byref(s)))

    def get_state(self) -> VinaState:
        s = VinaState()
        self._ck(self.lib.vina_odom_get_state(self.h, C.byref(s)))
        return s

    def set_imu_anchor(self, last_end: float, last_imu7, scale_gravity: float = 1.0):
        a = imu_array(np.asarray(last_imu7, dtype=np.float64).reshape(1, 7))
        self._ck(self.lib.vina_odom_set_imu_anchor(self.h, C.c_double(last_end), a.ctypes.data_as(C.c_void_p),
                                                   C.c_double(scale_gravity)))

    def bootstrap(self, xyzt: np.ndarray, state: VinaState):
        a = np.ascontiguousarray(xyzt, dtype=np.float32)
        self._ck(self.lib.vina_odom_bootstrap(self.h, _fp(a), C.c_int(a.shape[0]), C.byref(state)))
        self.sync()

    def map_last_counts(self):
        """(points inserted, leaves touched, [nodes per layer under surf_map_slide], leaves subdivided) of the last update"""
        a = (C.c_int32 * 8)()
        self._ck(self.lib.vina_map_last_counts(self.h, a))
        return a[0], a[1], [a[2], a[3], a[4], a[5]], a[6]

    def cold_start(self):
        """Start-up phase of the reference (VINA_SLAM::initialization): empty map, zero state, IMU not initialised."""
        self._ck(self.lib.vina_odom_cold_start(self.h))

    def init_scan(self, xyzt: np.ndarray, beg_time: float, imu7: np.ndarray):
        """One scan of the start-up phase -> (status, state): 0 = collecting, 1 = initialised (go on with step),
        -1 = motion_init failed, the system was reset."""
        a = np.ascontiguousarray(xyzt, dtype=np.float32)
        im = imu_array(np.asarray(imu7, dtype=np.float64))
        out, st = VinaState(), C.c_int32(0)
        self._ck(self.lib.vina_odom_init_scan(self.h, _fp(a), C.c_int(a.shape[0]), C.c_double(beg_time),
                                              im.ctypes.data_as(C.c_void_p), C.c_int(im.shape[0]), C.byref(out), C.byref(st)))
        return st.value, out

    def step(self, xyzt: np.ndarray, beg_time: float, imu7: np.ndarray, iekf_on_full: bool = True, max_iter: int = 4):
        a = np.ascontiguousarray(xyzt, dtype=np.float32)
        im = imu_array(np.asarray(imu7, dtype=np.float64))
        out = VinaState()
        self._ck(self.lib.vina_odom_step(self.h, _fp(a), C.c_int(a.shape[0]), C.c_double(beg_time),
                                         im.ctypes.data_as(C.c_void_p), C.c_int(im.shape[0]),
                                         C.c_int(1 if iekf_on_full else 0), C.c_int(max_iter), C.byref(out)))
        return out

    def step_resident(self, d_ptr: int, n: int, beg_time: float, end_time: float, imu7: np.ndarray,
                      iekf_on_full: bool = True, max_iter: int = 4):
        """d_ptr: device pointer to n x 4 float32 raw points already in HBM."""
        im = imu_array(np.asarray(imu7, dtype=np.float64))
        out = VinaState()
        self._ck(self.lib.vina_odom_step_resident(self.h, C.c_void_p(d_ptr), C.c_int(n), C.c_double(beg_time),
                                                  C.c_double(end_time), im.ctypes.data_as(C.c_void_p),
                                                  C.c_int(im.shape[0]), C.c_int(1 if iekf_on_full else 0),
                                                  C.c_int(max_iter), C.byref(out)))
        return out

    def propagate(self, beg: float, end: float, imu7: np.ndarray) -> np.ndarray:
        im = imu_array(np.asarray(imu7, dtype=np.float64))
        poses = np.zeros(96, dtype=IMU_POSE_DTYPE)
        k = self._ck(self.lib.vina_odom_propagate(self.h, C.c_double(beg), C.c_double(end),
                                                  im.ctypes.data_as(C.c_void_p), C.c_int(im.shape[0]),
                                                  poses.ctypes.data_as(C.c_void_p), C.c_int(96)))
        return poses[:k]

    def odom_iekf(self, which: int, max_iter: int, host_solve: bool = False):
        it, ok = C.c_int(0), C.c_int(0)
        fn = self.lib.vina_odom_iekf_host if host_solve else self.lib.vina_odom_iekf
        self._ck(fn(self.h, C.c_int(which), C.c_int(max_iter), C.byref(it), C.byref(ok)))
        return it.value, ok.value

    def odom_map_update(self):
        self._ck(self.lib.vina_odom_map_update(self.h))

    def window(self):
        wc = C.c_int(0)
        mp = np.zeros(16, dtype=np.int32)
        ws = self._ck(self.lib.vina_odom_window(self.h, C.byref(wc), mp.ctypes.data_as(C.c_void_p), C.c_int(16)))
        return wc.value, mp[:ws].copy()


LIDAR_LIVOX, LIDAR_VELODYNE, LIDAR_OUSTER, LIDAR_HESAI, LIDAR_ROBOSENSE, LIDAR_TARTANAIR = range(6)
LIVOX_POINT_DTYPE = np.dtype([("offset_time", "<u4"), ("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("reflectivity", "u1"),
                              ("tag", "u1"), ("line", "u1"), ("pad", "u1")])


class Pc2Layout(C.Structure):
    _fields_ = [("point_step", C.c_int32), ("off_x", C.c_int32), ("off_y", C.c_int32), ("off_z", C.c_int32),
                ("off_t", C.c_int32), ("t_datatype", C.c_int32), ("is_bigendian", C.c_int32)]


def decode_pointcloud2(lidar_type: int, data: bytes, n_points: int, point_step: int, off_xyz, off_t: int, t_datatype: int,
                       header_stamp: float, blind2: float, point_filter_num: int, omega_l: float = 3610.0,
                       is_bigendian: bool = False, cap: int | None = None) -> np.ndarray:
    """vina_decode_pointcloud2 (host only): the bytes of a sensor_msgs/PointCloud2 -> (n_kept, 4) float32."""
    lib = load()
    L = Pc2Layout(point_step, off_xyz[0], off_xyz[1], off_xyz[2], off_t, t_datatype, 1 if is_bigendian else 0)
    cap = n_points if cap is None else cap
    out = np.zeros((max(cap, 1), 4), dtype=np.float32)
    buf = (C.c_uint8 * max(len(data), 1)).from_buffer_copy(data if len(data) else b"\0")
    r = lib.vina_decode_pointcloud2(C.c_int(lidar_type), buf, C.c_int64(n_points), C.byref(L), C.c_double(header_stamp),
                                    C.c_double(omega_l), C.c_double(blind2), C.c_int(point_filter_num), _fp(out),
                                    C.c_int64(cap))
    if r < 0:
        raise VinaError(int(r), "vina_decode_pointcloud2")
    return out[:r].copy()


def decode_livox(pts: np.ndarray, blind2: float, point_filter_num: int, cap: int | None = None) -> np.ndarray:
    """vina_decode_livox (host only): livox CustomMsg points (LIVOX_POINT_DTYPE) -> (n_kept, 4) float32."""
    lib = load()
    a = np.ascontiguousarray(pts, dtype=LIVOX_POINT_DTYPE)
    cap = a.shape[0] if cap is None else cap
    out = np.zeros((max(cap, 1), 4), dtype=np.float32)
    r = lib.vina_decode_livox(a.ctypes.data_as(C.c_void_p), C.c_int64(a.shape[0]), C.c_double(blind2),
                              C.c_int(point_filter_num), _fp(out), C.c_int64(cap))
    if r < 0:
        raise VinaError(int(r), "vina_decode_livox")
    return out[:r].copy()


def scan_last_stamp(xyzt: np.ndarray) -> float:
    """vina_scan_last_stamp (host only): the time offset of the last point the prepared scan will have."""
    lib = load()
    a = np.ascontiguousarray(xyzt, dtype=np.float32).reshape(-1, 4)
    t = C.c_float(0)
    r = lib.vina_scan_last_stamp(_fp(a), C.c_int64(a.shape[0]), C.byref(t))
    if r:
        raise VinaError(r, "vina_scan_last_stamp")
    return t.value


class Sync:
    """Pairing of scans and IMU samples (vina_sync_*: sync_packages and its buffers, src/sensor/sync.cpp:5-96).
    Host-only: works without a CUDA device."""

    def __init__(self, point_notime: int = 0):
        self.lib = load()
        h = C.c_void_p()
        r = self.lib.vina_sync_create(C.c_int(point_notime), C.byref(h))
        if r:
            raise VinaError(r, "vina_sync_create")
        self.h = h

    def close(self):
        if self.h:
            self.lib.vina_sync_destroy(self.h)
            self.h = None

    def push_imu(self, imu7):
        im = imu_array(np.asarray(imu7, dtype=np.float64).reshape(1, 7))
        r = self.lib.vina_sync_push_imu(self.h, im.ctypes.data_as(C.c_void_p))
        if r:
            raise VinaError(r, "vina_sync_push_imu")

    def push_scan(self, t_start: float, t_last: float, tag: int):
        r = self.lib.vina_sync_push_scan(self.h, C.c_double(t_start), C.c_double(t_last), C.c_int64(tag))
        if r:
            raise VinaError(r, "vina_sync_push_scan")

    def pending(self):
        a, b = C.c_int32(0), C.c_int32(0)
        self.lib.vina_sync_pending(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def next(self, cap: int = 256):
        """(code, tag, beg, end, imu7[m, 7]); code as vina_sync_next returns it (negative = VINA_E_*)."""
        tag, beg, end, m = C.c_int64(-1), C.c_double(0), C.c_double(0), C.c_int32(0)
        buf = np.zeros(max(cap, 1), dtype=IMU_DTYPE)
        r = self.lib.vina_sync_next(self.h, C.byref(tag), C.byref(beg), C.byref(end), buf.ctypes.data_as(C.c_void_p),
                                    C.c_int(cap), C.byref(m))
        k = m.value if r >= 0 else 0  # (on VINA_E_CAPACITY m is the number of samples the scan is waiting for)
        out = np.zeros((k, 7), dtype=np.float64)
        out[:, 0], out[:, 1:4], out[:, 4:7] = buf["t"][:k], buf["gyr"][:k], buf["acc"][:k]
        return r, tag.value, beg.value, end.value, out


class Batch:
    """B independent sequences on one GPU advancing in lock step (vina_batch): per-sequence stages on the
    contexts' own streams, the IEKF iterations of all sequences as one k_iekf launch per iteration."""

    def __init__(self, ctxs):
        self.lib = load()
        self.ctxs = list(ctxs)
        arr = (C.c_void_p * len(self.ctxs))(*[c.h for c in self.ctxs])
        h = C.c_void_p()
        r = self.lib.vina_batch_create(arr, C.c_int(len(self.ctxs)), C.byref(h))
        if r != 0:
            raise VinaError(r, "vina_batch_create failed (contexts must live on the same device, at most 16)")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.vina_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def step_resident(self, d_ptrs, ns, begs, ends, imus7, iekf_on_full=True, max_iter=4):
        B = len(self.ctxs)
        ims = [imu_array(np.asarray(a, dtype=np.float64)) for a in imus7]
        dp = (C.c_void_p * B)(*[int(p) for p in d_ptrs])
        nn = (C.c_int32 * B)(*[int(v) for v in ns])
        tb = (C.c_double * B)(*[float(v) for v in begs])
        te = (C.c_double * B)(*[float(v) for v in ends])
        ip = (C.c_void_p * B)(*[a.ctypes.data for a in ims])
        mm = (C.c_int32 * B)(*[a.shape[0] for a in ims])
        out = (VinaState * B)()
        r = self.lib.vina_batch_step_resident(self.h, dp, nn, tb, te, ip, mm, C.c_int(1 if iekf_on_full else 0),
                                              C.c_int(max_iter), out)
        if r < 0:
            msgs = [self.lib.vina_last_error(c.h).decode() for c in self.ctxs]
            raise VinaError(int(r), " | ".join(m for m in msgs if m))
        return list(out)

    def iekf_time(self):
        """(per-launch device ms of the last step's batched k_iekf launches, number of launches)"""
        ms = (C.c_float * 32)()
        k = C.c_int32(0)
        self.lib.vina_batch_iekf_time(self.h, ms, C.c_int(32), C.byref(k))
        return [ms[i] for i in range(min(k.value, 32))], k.value

    def sync(self):
        r = self.lib.vina_batch_sync(self.h)
        if r < 0:
            raise VinaError(int(r), "vina_batch_sync")

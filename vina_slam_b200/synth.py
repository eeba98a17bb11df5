"""Seeded synthetic LiDAR + IMU sequences shaped like the reference's sensors.

SURVEY.md §8(d): world = axis-aligned building shell with interior walls and
pillars (every surface planar, so map voxels become planes), sensor patterns
for the four yaml configs of the reference, smooth analytic trajectory with
IMU samples from its derivatives.  Everything is numpy + ``default_rng(seed)``;
there is no dataset and no network in this environment.

The output contract is the one the reference's sensor layer hands to the hot
path (src/sensor/lidar_decoder.cpp:30-35): points as float32 (x, y, z,
curvature = time offset in s), sorted by time, blind zone removed.
"""
from __future__ import annotations

import dataclasses
import math
from typing import List, Tuple

import numpy as np

G_M_S2 = 9.8  # include/vina_slam/core/constants.hpp:13


@dataclasses.dataclass
class SensorConfig:
    """Parameters of one reference yaml (file:line cited in SENSORS below)."""

    name: str
    pattern: str  # "rosette" | "spin"
    n_beams: int
    n_steps: int
    elev_deg: Tuple[float, float]
    scan_period: float
    imu_rate: float
    blind: float
    voxel_size: float
    max_layer: int
    down_size: float
    dept_err: float
    beam_err: float
    min_eigen_value: float
    plane_thre: Tuple[float, float, float, float]
    ext_R: Tuple[float, ...]  # row-major, as in the yaml
    ext_t: Tuple[float, float, float]
    cov_gyr: float = 0.01
    cov_acc: float = 1.0
    rdw_gyr: float = 1e-4
    rdw_acc: float = 1e-4
    win_size: int = 10
    thread_num: int = 5
    max_points: int = 100  # octree.cpp:70 (never configurable)
    seed: int = 0
    handheld: bool = False

    @property
    def n_points(self) -> int:
        return self.n_beams * self.n_steps

    def ext_R_colmajor(self) -> np.ndarray:
        return np.asarray(self.ext_R, dtype=np.float64).reshape(3, 3).T.reshape(-1).copy()


_I3 = (1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0)

SENSORS = {
    # config/mid360.yaml: voxel_size :55, max_layer :68, down_size :49, blind :14, thre :80,
    # min_eigen_value :57, dept/beam_err :51-53, extrinsic :19-21
    "mid360": SensorConfig(
        name="mid360", pattern="rosette", n_beams=1, n_steps=20000, elev_deg=(-7.0, 52.0), scan_period=0.1,
        imu_rate=200.0, blind=3.0, voxel_size=0.5, max_layer=3, down_size=0.1, dept_err=0.02, beam_err=0.05,
        min_eigen_value=0.0025, plane_thre=(4.0, 4.0, 4.0, 4.0), ext_R=_I3, ext_t=(-0.011, -0.02329, 0.04412),
        seed=360),
    # config/velodyne.yaml: voxel_size :82, max_layer :95, blind :14, extrinsic :35-48
    "velodyne32": SensorConfig(
        name="velodyne32", pattern="spin", n_beams=32, n_steps=1800, elev_deg=(-25.0, 15.0), scan_period=0.1,
        imu_rate=200.0, blind=0.0, voxel_size=1.0, max_layer=3, down_size=0.1, dept_err=0.02, beam_err=0.05,
        min_eigen_value=0.0025, plane_thre=(4.0, 4.0, 4.0, 4.0),
        ext_R=(0.999256713, 0.032388537, -0.020904626, -0.032332863, 0.999472666, 0.002995836, 0.020990633,
               -0.002317703, 0.999776986),
        ext_t=(-0.01907, 0.0045, 0.05714), cov_acc=1.0, seed=32),
    # config/robosense.yaml: lidar_type :12, blind :14, voxel_size :53, max_layer :66,
    # dept/beam_err :49-51, thre :78, extrinsic :17-20
    "robosense128": SensorConfig(
        name="robosense128", pattern="spin", n_beams=128, n_steps=1875, elev_deg=(-25.0, 15.0), scan_period=0.1,
        imu_rate=200.0, blind=2.0, voxel_size=1.0, max_layer=2, down_size=0.1, dept_err=0.01, beam_err=0.01,
        min_eigen_value=0.0025, plane_thre=(1.0, 1.0, 1.0, 1.0),
        ext_R=(0.0, -1.0, 0.0, -1.0, 0.0, 0.0, 0.0, 0.0, -1.0), ext_t=(0.00425, 0.00418, -0.00446), seed=128),
    # config/HILTI.yaml: lidar_type :12, blind :14, voxel_size :53, max_layer :66 (Hesai XT32, 400 Hz IMU)
    "hilti_xt32": SensorConfig(
        name="hilti_xt32", pattern="spin", n_beams=32, n_steps=2000, elev_deg=(-16.0, 15.0), scan_period=0.1,
        imu_rate=400.0, blind=0.7, voxel_size=1.0, max_layer=2, down_size=0.1, dept_err=0.02, beam_err=0.05,
        min_eigen_value=0.0025, plane_thre=(1.0, 1.0, 1.0, 1.0), ext_R=_I3, ext_t=(0.0, 0.0, 0.0), seed=2021,
        handheld=True),
}


def small_sensor(base: str, n_beams: int, n_steps: int, seed: int | None = None) -> SensorConfig:
    """A reduced-size copy of a named sensor for CPU-sized tests."""
    s = dataclasses.replace(SENSORS[base], n_beams=n_beams, n_steps=n_steps)
    if seed is not None:
        s = dataclasses.replace(s, seed=seed)
    return s


# --------------------------------------------------------------------------- world
class World:
    """Axis-aligned rectangles: (axis, offset, lo0, hi0, lo1, hi1) over the two other axes."""

    def __init__(self, size=(60.0, 30.0, 8.0), pitch=6.0, tiles=(1, 1), offset=(0.0, 0.0, 0.0)):
        rects = []
        sx, sy, sz = size
        self.offset = np.asarray(offset, dtype=np.float64)
        for tx in range(tiles[0]):
            for ty in range(tiles[1]):
                ox, oy = tx * sx, ty * sy
                # shell
                rects.append((0, ox + 0.0, oy, oy + sy, 0.0, sz))
                rects.append((0, ox + sx, oy, oy + sy, 0.0, sz))
                rects.append((1, oy + 0.0, ox, ox + sx, 0.0, sz))
                rects.append((1, oy + sy, ox, ox + sx, 0.0, sz))
                rects.append((2, 0.0, ox, ox + sx, oy, oy + sy))
                rects.append((2, sz, ox, ox + sx, oy, oy + sy))
                # interior wall stubs and square pillars every `pitch` metres
                nx = int(sx // pitch)
                ny = int(sy // pitch)
                for i in range(1, nx):
                    x = ox + i * pitch
                    # partial walls from both long sides leaving a corridor in the middle
                    rects.append((0, x, oy + 0.0, oy + sy * 0.3, 0.0, sz))
                    rects.append((0, x, oy + sy * 0.7, oy + sy, 0.0, sz))
                    # square pillars one pitch either side of the centre line: the corridor between them (and between
                    # the wall stubs) stays wider than the largest blind zone of the bundled sensors (3 m) around
                    # every trajectory, so each rank's seed yields full scans
                    for y in (oy + sy / 2 - pitch, oy + sy / 2 + pitch):
                        if ny > 2:
                            h = 0.4  # pillar half-size
                            xc = x + pitch / 2
                            rects.append((0, xc - h, y - h, y + h, 0.0, sz))
                            rects.append((0, xc + h, y - h, y + h, 0.0, sz))
                            rects.append((1, y - h, xc - h, xc + h, 0.0, sz))
                            rects.append((1, y + h, xc - h, xc + h, 0.0, sz))
        self.rects = np.asarray(rects, dtype=np.float64)
        # shift the whole world (negative offsets exercise the "-1 if negative" key rule)
        other = {0: (1, 2), 1: (0, 2), 2: (0, 1)}
        for r in self.rects:
            a = int(r[0])
            b0, b1 = other[a]
            r[1] += self.offset[a]
            r[2] += self.offset[b0]
            r[3] += self.offset[b0]
            r[4] += self.offset[b1]
            r[5] += self.offset[b1]
        self.size = size
        self.tiles = tiles

    def cast(self, origins: np.ndarray, dirs: np.ndarray) -> np.ndarray:
        """Range to the nearest rectangle along each ray (inf when nothing is hit)."""
        n = dirs.shape[0]
        best = np.full(n, np.inf)
        other = {0: (1, 2), 1: (0, 2), 2: (0, 1)}
        for axis, off, lo0, hi0, lo1, hi1 in self.rects:
            a = int(axis)
            da = dirs[:, a]
            with np.errstate(divide="ignore", invalid="ignore"):
                t = (off - origins[:, a]) / da
            b0, b1 = other[a]
            h0 = origins[:, b0] + t * dirs[:, b0]
            h1 = origins[:, b1] + t * dirs[:, b1]
            ok = (t > 1e-3) & (t < best) & (h0 >= lo0) & (h0 <= hi0) & (h1 >= lo1) & (h1 <= hi1)
            best = np.where(ok, t, best)
        return best


# --------------------------------------------------------------------------- trajectory
class Trajectory:
    """Smooth Lissajous path with ZYX Euler attitude; analytic derivatives give the IMU."""

    def __init__(self, world: World, handheld: bool = False, seed: int = 0):
        rng = np.random.default_rng(seed + 7919)
        sx, sy, sz = world.size
        self.c = np.array([sx / 2, sy / 2, 1.6]) + world.offset
        # every path stays inside the pillar-free corridor (pillars stand at y = centre +- 6 m, faces at +- 5.6 m):
        # y amplitude 1.8 m (handheld) / 2.1 m leaves >= 3.5 m to the nearest surface beside the path
        self.A = np.array([sx * 0.30, sy * (0.06 if handheld else 0.07), 0.25])
        self.w = np.array([0.045, 0.09, 0.31]) * (2.0 if handheld else 1.0)
        self.ph = rng.uniform(0, 2 * np.pi, 3)
        k = 4.0 if handheld else 1.0
        self.eA = np.array([0.04 * k, 0.05 * k, 0.9])  # roll, pitch, yaw amplitudes (rad)
        self.ew = np.array([0.9, 0.7, 0.21]) * (1.6 if handheld else 1.0)
        self.eph = rng.uniform(0, 2 * np.pi, 3)

    def pos(self, t):
        t = np.asarray(t, dtype=np.float64)[..., None]
        return self.c + self.A * np.sin(self.w * t + self.ph)

    def vel(self, t):
        t = np.asarray(t, dtype=np.float64)[..., None]
        return self.A * self.w * np.cos(self.w * t + self.ph)

    def acc(self, t):
        t = np.asarray(t, dtype=np.float64)[..., None]
        return -self.A * self.w * self.w * np.sin(self.w * t + self.ph)

    def euler(self, t):
        t = np.asarray(t, dtype=np.float64)[..., None]
        e = self.eA * np.sin(self.ew * t + self.eph)
        de = self.eA * self.ew * np.cos(self.ew * t + self.eph)
        return e, de

    def rot(self, t):
        e, _ = self.euler(t)
        r, p, y = e[..., 0], e[..., 1], e[..., 2]
        cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
        R = np.empty(e.shape[:-1] + (3, 3))
        R[..., 0, 0] = cy * cp
        R[..., 0, 1] = cy * sp * sr - sy * cr
        R[..., 0, 2] = cy * sp * cr + sy * sr
        R[..., 1, 0] = sy * cp
        R[..., 1, 1] = sy * sp * sr + cy * cr
        R[..., 1, 2] = sy * sp * cr - cy * sr
        R[..., 2, 0] = -sp
        R[..., 2, 1] = cp * sr
        R[..., 2, 2] = cp * cr
        return R

    def omega_body(self, t):
        e, de = self.euler(t)
        r, p = e[..., 0], e[..., 1]
        dr, dp, dy = de[..., 0], de[..., 1], de[..., 2]
        w = np.empty(e.shape)
        w[..., 0] = dr - dy * np.sin(p)
        w[..., 1] = dp * np.cos(r) + dy * np.cos(p) * np.sin(r)
        w[..., 2] = -dp * np.sin(r) + dy * np.cos(p) * np.cos(r)
        return w

    def specific_force(self, t):
        R = self.rot(t)
        a = self.acc(t) + np.array([0.0, 0.0, G_M_S2])
        return np.einsum("...ji,...j->...i", R, a)


# --------------------------------------------------------------------------- sensor
def beam_directions(cfg: SensorConfig) -> Tuple[np.ndarray, np.ndarray]:
    """Unit directions in the LiDAR frame and per-point time offsets (sorted ascending)."""
    n = cfg.n_points
    T = cfg.scan_period
    if cfg.pattern == "spin":
        step = np.repeat(np.arange(cfg.n_steps), cfg.n_beams)
        beam = np.tile(np.arange(cfg.n_beams), cfg.n_steps)
        az = 2 * np.pi * step / cfg.n_steps
        el = np.deg2rad(cfg.elev_deg[0] + (cfg.elev_deg[1] - cfg.elev_deg[0]) * beam / max(cfg.n_beams - 1, 1))
        toff = (step + beam / cfg.n_beams) * (T / cfg.n_steps)
    else:  # non-repetitive rosette (Livox-like): two incommensurate frequencies
        k = np.arange(n)
        toff = k * (T / n)
        az = 2 * np.pi * (toff / T) * 17.0 + 0.37 * np.sin(2 * np.pi * 41.0 * toff / T)
        mid = 0.5 * (cfg.elev_deg[0] + cfg.elev_deg[1])
        amp = 0.5 * (cfg.elev_deg[1] - cfg.elev_deg[0])
        el = np.deg2rad(mid + amp * np.sin(2 * np.pi * 29.3 * toff / T + 1.1))
    d = np.stack([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)], axis=1)
    return d, toff


@dataclasses.dataclass
class Scan:
    xyzt: np.ndarray  # float32 (n, 4): x, y, z, curvature
    beg_time: float
    imu: np.ndarray  # float64 (m, 7): t, gx, gy, gz, ax, ay, az (stamps in (prev_end, end])
    gt_R: np.ndarray  # ground truth at pcl_end_time, row-major 3x3
    gt_p: np.ndarray
    gt_v: np.ndarray
    end_time: float


class Sequence:
    """A seeded synthetic sequence for one SensorConfig."""

    def __init__(self, cfg: SensorConfig, seed: int | None = None, world: World | None = None,
                 t0: float = 5.0):
        self.cfg = cfg
        self.seed = cfg.seed if seed is None else seed
        self.world = world or World()
        self.traj = Trajectory(self.world, handheld=cfg.handheld, seed=self.seed)
        self.rng = np.random.default_rng(self.seed)
        self.dirs, self.toff = beam_directions(cfg)
        self.R_L = np.asarray(cfg.ext_R, dtype=np.float64).reshape(3, 3)
        self.t_L = np.asarray(cfg.ext_t, dtype=np.float64)
        self.t0 = t0
        self._imu_dt = 1.0 / cfg.imu_rate
        self._imu_next = 0  # index of the next IMU sample not yet handed out
        self._scan = 0

    # IMU sample k has stamp t0 + k * dt
    def _imu_samples(self, k0: int, k1: int) -> np.ndarray:
        k = np.arange(k0, k1)
        t = self.t0 + k * self._imu_dt
        rng = np.random.default_rng([self.seed, 1, k0])
        gyr = self.traj.omega_body(t) + rng.normal(0, 1e-3, (len(k), 3))
        acc = self.traj.specific_force(t) + rng.normal(0, 1e-2, (len(k), 3))
        return np.concatenate([t[:, None], gyr, acc], axis=1)

    def _measure(self, t_pts: np.ndarray, static_at: float | None, rng) -> np.ndarray:
        cfg = self.cfg
        tt = np.full_like(t_pts, static_at) if static_at is not None else t_pts
        R = self.traj.rot(tt)
        p = self.traj.pos(tt)
        origin = p + np.einsum("nij,j->ni", R, self.t_L)
        d_imu = self.dirs @ self.R_L.T
        d_w = np.einsum("nij,nj->ni", R, d_imu)
        rng_true = self.world.cast(origin, d_w)
        ok = np.isfinite(rng_true)
        r = rng_true + rng.normal(0, cfg.dept_err, rng_true.shape)
        # angular noise: small rotation of the beam direction
        sig = math.radians(cfg.beam_err)
        dn = self.dirs + rng.normal(0, sig, self.dirs.shape)
        dn /= np.linalg.norm(dn, axis=1, keepdims=True)
        pts = dn * r[:, None]
        keep = ok & (np.einsum("ni,ni->n", pts, pts) > cfg.blind * cfg.blind) & (r > 0.05)
        return pts, keep

    def next_scan(self, deskewed: bool = False) -> Scan:
        """Scan k covers [t0 + k*T, t0 + (k+1)*T). ``deskewed`` = sensor frozen at the scan-end pose."""
        cfg = self.cfg
        k = self._scan
        self._scan += 1
        beg = self.t0 + k * cfg.scan_period
        rng = np.random.default_rng([self.seed, 2, k])
        toff32 = self.toff.astype(np.float32)
        # pcl_end_time = beg + last curvature (sync.cpp:40), curvature is float32
        t_abs = beg + toff32.astype(np.float64)
        if deskewed:
            pts, keep = self._measure(t_abs, beg + float(toff32[-1]), rng)
        else:
            pts, keep = self._measure(t_abs, None, rng)
        xyzt = np.concatenate([pts[keep], toff32[keep, None].astype(np.float64)], axis=1).astype(np.float32)
        if xyzt.shape[0] == 0:
            raise RuntimeError(f"synthetic scan {k} of seed {self.seed} ({cfg.name}) has no returns outside the "
                               f"{cfg.blind} m blind zone: the sensor is inside or against a surface of the world")
        end_time = beg + float(xyzt[-1, 3])
        # IMU samples with stamp <= end_time not yet handed out (sync.cpp:63-72)
        k1 = int(math.floor((end_time - self.t0) / self._imu_dt + 1e-9)) + 1
        imu = self._imu_samples(self._imu_next, k1)
        self._imu_next = k1
        return Scan(xyzt=xyzt, beg_time=beg, imu=imu, gt_R=self.traj.rot(end_time), gt_p=self.traj.pos(end_time),
                    gt_v=self.traj.vel(end_time), end_time=end_time)


def _gen_job(job):
    cfg, seed, n_boot, n_steps = job
    seq = Sequence(cfg, seed=seed)
    return [seq.next_scan(deskewed=True) for _ in range(n_boot)], [seq.next_scan() for _ in range(n_steps)]


def gen_sequences(cfg: SensorConfig, seeds, n_boot: int, n_steps: int, procs: int | None = None):
    """[(bootstrap scans, scans)] for several seeds, generated in worker processes (a full-size scan costs about a
    second of numpy ray casting; independent sequences are independent jobs)."""
    import multiprocessing as mp
    import os

    jobs = [(cfg, int(sd), n_boot, n_steps) for sd in seeds]
    procs = min(len(jobs), procs or (os.cpu_count() or 1))
    if procs <= 1 or len(jobs) <= 1:
        return [_gen_job(j) for j in jobs]
    with mp.get_context("fork").Pool(procs) as pool:
        return pool.map(_gen_job, jobs)


def rot_err_deg(Ra: np.ndarray, Rb: np.ndarray) -> float:
    c = (np.trace(Ra.T @ Rb) - 1.0) / 2.0
    return math.degrees(math.acos(max(-1.0, min(1.0, c))))


def rot_exp(w: np.ndarray) -> np.ndarray:
    """Rodrigues: rotation matrix of the rotation vector w."""
    w = np.asarray(w, dtype=np.float64)
    th = float(np.linalg.norm(w))
    if th < 1e-12:
        return np.eye(3)
    k = w / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + math.sin(th) * K + (1 - math.cos(th)) * (K @ K)

#!/usr/bin/env python
"""bench.py — pts/s of the per-scan hot path (deskew -> down-sample -> var_init -> IEKF association /
H-b reduction -> pvec_update -> voxel insert -> recut -> margi) on synthetic scans.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one scan of the named sensor shape through the whole path. The default workload is
BASELINE.json configs[2] (RoboSense-128-shaped, 240 k pts/scan), the configuration the north_star's
">= 20x on a 128-beam-shaped scan" target is quoted on; it fits one GPU. N > 1 = N independent
sequences (replicas, one per rank, no data-path collective): "scaling": "weak".

  value  whole-job pts/s with each raw scan already resident in HBM (vina_odom_step_resident)
  e2e    the same metric through vina_odom_step from pinned HOST buffers (H2D of the scan and the
         34-double readback of every IEKF iteration inside the timed region)
  roofline  dominant kernel (k_iekf): algorithmic bytes per launch / CUDA-event launch duration
  cpu_baseline  the oracle's -O3 -ffast-math build on this box's host cores (bounded sample)

--impl reference times the reference's CPU implementation of the path: the oracle port (kind "port"), which
is bit-identical in results to oracle/_ref (the reference's own source files compiled against the header shims
of oracle/ref_shim) and faster than that build, i.e. the conservative baseline; the _ref timing is added as
"reference_build" for information.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from vina_slam_b200 import synth  # noqa: E402

METRIC = "pts/s IEKF point-to-plane update + voxel insert (deskew -> IEKF -> map insert/recut/margi, per scan)"
UNIT = "pts/s"
MAX_ITER = 4  # the VNC_lio budget, odometry.cpp:68 / local_mapping.cpp:413


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        """index: one GPU index or a comma-separated list (rank 0 samples every GPU of the job: one nvidia-smi client
        instead of one per rank - each query takes the driver's lock and shows up in the other ranks' step times)"""
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._halt = threading.Event()
        self._nv = None
        try:  # NVML is initialised here, before the timed region starts (nvmlInit alone takes longer than 20 steps)
            import pynvml as nv

            nv.nvmlInit()
            hs = [nv.nvmlDeviceGetHandleByIndex(int(i)) for i in str(self.index).split(",")]
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(hs[0], nv.NVML_CLOCK_SM))
            self._nv = (nv, hs)
        except Exception:
            self._nv = None

    def _run_nvml(self):
        """In-process NVML (the same counters nvidia-smi prints, without a subprocess that takes the driver's lock for
        tens of milliseconds in the middle of a 10 ms timed region). False if NVML is not usable here."""
        if self._nv is None:
            return False
        nv, hs = self._nv
        bits = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4}
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        self.source = "nvml"
        while True:
            try:
                for h in hs:
                    self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                    r = int(get_reasons(h))
                    for nm, b in bits.items():
                        if r & b:
                            self.reasons.add(nm)
            except Exception:
                pass
            if self._halt.wait(0.002):
                break
        return True

    def run(self):
        if self._run_nvml():
            return
        self.source = "nvidia-smi"
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                for row in out.strip().splitlines():
                    f = [x.strip() for x in row.split(",")]
                    if len(f) >= 6:
                        self.samples.append(float(f[0]))
                        self.max_mhz = float(f[1])
                        for nm, v in zip(names, f[2:6]):
                            if v.lower().startswith("active"):
                                self.reasons.add(nm)
            except Exception:
                pass
            self._halt.wait(0.25)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples),
                "source": getattr(self, "source", None)}


def gen_sequence(cfg, seed, n_boot, n_steps):
    seq = synth.Sequence(cfg, seed=seed)
    boots = [seq.next_scan(deskewed=True) for _ in range(n_boot)]
    scans = [seq.next_scan() for _ in range(n_steps)]
    return boots, scans


# --------------------------------------------------------------------------- reference arm / cpu baseline
BA_WARMUP = 9  # scans until every pair of consecutive window frames carries an IMU factor: BA runs from then on


def run_cpu(cfg, boots, scans, warmup, steps, ref=False, ba=False):
    """The reference's CPU implementation of the path with its own release flags (-O3 -ffast-math,
    CMakeLists.txt:92-96), IEKF single-threaded, insert/recut/margi on thread_num = 5 std::threads, exactly as the
    reference does. ref=False: the oracle port (kind "port") - bit-identical results to the reference build
    (tests/test_oracle_vs_ref.py) and FASTER than it, hence the conservative baseline. ref=True:
    oracle/_ref/libvina_ref_fast.so, the reference's own source files compiled against the header shims of
    oracle/ref_shim (kind "reference"); its eager stand-in for Eigen has no expression templates and it also runs
    the reference's unreachable VNC preprocessing, so it is slower than a build against the real Eigen would be."""
    from oracle import oracle_py as op

    op.build()
    use_ref = ref and op.have_ref() and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libvina_ref_fast.so"))
    od = op.Odom(cfg, fast=True, ref=use_ref)
    if ba:
        od.set_ba(True)
    for sc in boots:
        od.bootstrap(sc.xyzt, op.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
    od.set_imu_anchor(boots[-1].end_time, boots[-1].imu[-1])
    times, stages, pts = [], np.zeros(4), 0
    for k, sc in enumerate(scans[: warmup + steps]):
        t0 = time.perf_counter()
        r, _ = od.step(sc.xyzt, sc.beg_time, sc.imu, iekf_on_full=True, max_iter=MAX_ITER)
        dt = time.perf_counter() - t0
        assert r == 0
        if k >= warmup:
            times.append(dt)
            stages += od.stage_times()
            pts += sc.xyzt.shape[0]
    total = float(np.sum(times))
    od.close()
    return dict(value=pts / total, ms_per_step=1e3 * total / len(times), stage_ms=(1e3 * stages / len(times)).tolist(),
                steps=len(times), cores=max(1, cfg.thread_num), kind="reference" if use_ref else "port")


def main_reference(args, cfg):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    if args.ba:
        args.warmup = max(args.warmup, BA_WARMUP)
    boots, scans = gen_sequence(cfg, cfg.seed, cfg.win_size, args.warmup + args.steps)
    r = run_cpu(cfg, boots, scans, args.warmup, args.steps, ba=args.ba)
    sample = (f"{r['steps']} full scans of {cfg.n_points} pts after {args.warmup} warm-up scans; IEKF 1 thread, "
              f"map ops {cfg.thread_num} threads (reference threading)")
    ref_build = None
    try:  # informational: the reference's own sources (header-shim build), a few scans
        rr = run_cpu(cfg, boots, scans, args.warmup if args.ba else min(args.warmup, 1), min(args.steps, 3), ref=True, ba=args.ba)
        if rr["kind"] == "reference":
            ref_build = {"value": rr["value"], "unit": UNIT, "ms_per_step": rr["ms_per_step"], "steps": rr["steps"],
                         "note": "reference sources compiled against oracle/ref_shim (eager Eigen stand-in, incl. the "
                                 "unreachable VNC preprocessing); slower than the port, so the port is the baseline"}
    except Exception:
        pass
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": r["steps"], "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(cfg), "max_iter": MAX_ITER, "iekf_on": "full scan", "vnc_terms": False,
                   "if_BA": int(args.ba)},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": sample,
                         "stage_ms": dict(zip(["odom", "insert", "recut", "margi"], r["stage_ms"]))},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if ref_build:
        line["reference_build"] = ref_build
    emit(json.dumps(line))


def workload_name(cfg):
    return (f"{cfg.name}: {cfg.n_beams} beams x {cfg.n_steps} az steps = {cfg.n_points} pts/scan, 10 Hz, "
            f"{int(cfg.imu_rate)} Hz IMU, voxel {cfg.voxel_size} m, max_layer {cfg.max_layer}")


# --------------------------------------------------------------------------- our arm
def main_ours(args, cfg):
    import torch
    import torch.distributed as dist

    from vina_slam_b200 import capi

    rank, world, local = dist_env()
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    if args.ba:
        args.warmup = max(args.warmup, BA_WARMUP)
        args.batch = 0
    W, K = args.warmup, args.steps
    from vina_slam_b200 import replicas as _rep

    boots, scans = gen_sequence(cfg, _rep.sequence_seed(cfg.seed, rank), cfg.win_size, W + K)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak, peak_src = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")

    caps = dict(max_scan_points=max(300000, cfg.n_points + 1024), max_nodes=1 << 20, hash_capacity_log2=21,
                device=local)
    stream = torch.cuda.current_stream(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def new_ctx():
        gx = capi.Ctx(cfg, **caps)
        gx.set_stream(stream.cuda_stream)
        if args.ba:
            gx.set_ba(True)
        for sc in boots:
            gx.bootstrap(sc.xyzt, capi.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
        gx.set_imu_anchor(boots[-1].end_time, boots[-1].imu[-1])
        return gx

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- leg 1: inputs resident in HBM -------------------------------------------------------------
    # two passes over the same scans on fresh contexts: pass 0 is the timed one (no per-stage events, nothing
    # but the product path between the step's two CUDA events); pass 1 repeats it with the library's per-stage /
    # per-launch event timers on, for stage_ms and the kernel's launch duration in the roofline.
    d_scans = [torch.from_numpy(sc.xyzt).to(dev) for sc in scans]
    sampler = None
    for profiled in (False, True):
        gx = new_ctx()
        gx.set_profiling(profiled)
        ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(W + K)]
        ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(W + K)]
        tm_rows, traj, traj_p_pass, map_counts = [], 0.0, [], []
        barrier()
        for k, sc in enumerate(scans):
            if k == W:
                barrier()
                if not profiled and rank == 0:
                    sampler = ClockSampler(",".join(str(i) for i in range(world)) if world > 1 else local)
                    sampler.start()
            flush.fill_(k & 0xFF)  # evict L2 between timed iterations (outside the timed segments)
            ev0[k].record(stream)
            st = gx.step_resident(d_scans[k].data_ptr(), sc.xyzt.shape[0], sc.beg_time, sc.end_time, sc.imu, True,
                                  MAX_ITER)
            ev1[k].record(stream)
            if k >= W:
                tm_rows.append(gx.timings())
                traj = max(traj, float(np.linalg.norm(np.array(st.p[:]) - sc.gt_p)))
                traj_p_pass.append(np.array(st.p[:]))
                if profiled:
                    map_counts.append(gx.map_last_counts())
        barrier()
        if not profiled:
            clocks = sampler.stop() if sampler else {}
            step_ms = np.array([ev0[k].elapsed_time(ev1[k]) for k in range(W, W + K)])
            t_res = float(step_ms.sum()) * 1e-3
            launches = int(sum(t.kernel_launches for t in tm_rows))
            traj_err = traj
            traj_p = traj_p_pass
            ba_runs, ba_iters = (gx.ba_stats()[0] - (W - BA_WARMUP + 1), gx.ba_stats()[1]) if args.ba else (0, 0)
            gx.close()
    pts = sum(sc.xyzt.shape[0] for sc in scans[W:])
    iters = int(sum(t.iekf_iters for t in tm_rows))
    iekf_kernel_ms = float(sum(t.iekf_kernel_ms for t in tm_rows))
    stage = {f: float(np.mean([getattr(t, f) for t in tm_rows])) for f in
             ("deskew_ms", "downsample_ms", "var_init_ms", "iekf_ms", "insert_ms", "recut_ms", "margi_ms", "total_ms")}

    # algorithmic bytes of one k_iekf launch (DESIGN.md "Kernels"): 80 B/pt streamed + one 256-B leaf record
    # per distinct associated leaf + one 16-B hash slot per point that misses its cached leaf
    sc = scans[-1]
    n_last = sc.xyzt.shape[0]
    s = gx.get_state()
    gx.iekf_begin(0, np.zeros(9), np.zeros(9))
    gx.iekf_accumulate(np.array(s.R[:]), np.array(s.p[:]), debug=True)
    a = gx.iekf_debug_assoc(n_last)
    m = a["flags"] > 0
    leaf_id = a["keys"][m] @ np.array([1 << 42, 1 << 21, 1], dtype=np.int64) * 4096 + a["codes"][m]
    U = int(np.unique(leaf_id).shape[0])
    match_frac = float(m.mean())
    n_mean = pts / K
    iters_per_step = iters / K
    # iteration 0 misses everywhere; later iterations miss only where no leaf is cached yet
    miss = (n_mean + (iters_per_step - 1) * n_mean * (1 - match_frac)) / max(iters_per_step, 1)
    bytes_per_launch = 80.0 * n_mean + 256.0 * U + 16.0 * miss + 34 * 8
    launch_ms = iekf_kernel_ms / max(iters, 1)
    # VINA_IEKF_LOOP=1 (opt-in schedule): ONE persistent launch runs all iterations of a scan - the units of a launch
    # are (point, iteration) evaluations, its algorithmic bytes the per-iteration figure x the iterations it ran
    loop_mode = os.environ.get("VINA_IEKF_LOOP", "0") not in ("", "0")
    roof_kernel, roof_launches = "k_iekf", iters
    if loop_mode:
        bytes_per_launch *= iters_per_step
        launch_ms = iekf_kernel_ms / max(K, 1)
        roof_kernel, roof_launches = "k_iekf_loop (all iterations of a scan in one launch)", K
    # dram__bytes_read + write per launch from the committed `ncu --set full` capture of this kernel - valid only for the
    # workload it was taken on and for as long as the kernel's source is the one that was profiled
    traffic, traffic_source = None, None
    try:
        import hashlib

        with open(os.path.join(ROOT, "profiles", "r02b_iekf_ncu_full_summary.json")) as f:
            cap = json.load(f)
        with open(os.path.join(ROOT, "vina_slam_b200", "csrc", "iekf_kernel.cu"), "rb") as f:
            sha = hashlib.sha256(f.read()).hexdigest()[:16]
        if cfg.name == cap.get("workload") and sha == cap.get("kernel_source_sha256_16"):
            traffic = cap["dram_traffic_bytes_per_launch"]
            traffic_source = "profiles/r02b_iekf_ncu_full_summary.json (ncu --set full, same kernel source %s)" % sha
        else:
            traffic_source = "none: the committed ncu capture is of another workload or kernel source"
    except Exception:
        traffic_source = "none: no committed ncu capture"
    achieved = bytes_per_launch / (launch_ms * 1e-3) / 1e9 if launch_ms > 0 else 0.0

    # ---- the other stages: algorithmic bytes (SURVEY.md 8d per-unit figures x the units of this workload) over the
    # stage's device time from the instrumented pass - they are latency-bound kernels over ~10^4 tree nodes, the
    # table says by how much
    mc = np.array([[c[0], c[1], sum(c[2]), c[3]] for c in map_counts], dtype=np.float64).mean(axis=0) if map_counts else np.zeros(4)
    n_ins, u_ins, u_fit, n_split = [float(v) for v in mc]
    stage_bytes = {
        "deskew + var_init (k_deskew_var_init)": (28.0 + 72.0 + 4.0) * n_mean,
        "iekf loop (k_iekf x iterations)": bytes_per_launch * iters_per_step,
        "insert (k_insert_*)": (72.0 + 72.0 + 16.0 + 80.0) * n_ins + 1040.0 * u_ins,
        "recut (k_recut_*, k_split)": (80.0 + 96.0) * u_fit,
        "margi (k_margi_*)": (80.0 + 800.0 + 360.0 + 96.0 + 224.0) * u_fit,
    }
    stage_time = {"deskew + var_init (k_deskew_var_init)": stage["deskew_ms"], "iekf loop (k_iekf x iterations)": stage["iekf_ms"],
                  "insert (k_insert_*)": stage["insert_ms"], "recut (k_recut_*, k_split)": stage["recut_ms"],
                  "margi (k_margi_*)": stage["margi_ms"]}
    roofline_stages = []
    for nm, bts in stage_bytes.items():
        ms = stage_time[nm]
        gbs = bts / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        roofline_stages.append({"stage": nm, "algorithmic_bytes": bts, "ms": ms, "achieved_GBs": gbs, "frac": gbs / hbm_peak})
    whole = sum(stage_bytes.values())
    gx.close()
    del d_scans

    # ---- leg 2: end to end from pinned host buffers ---------------------------------------------------
    gx = new_ctx()
    gx.set_upload_ordered(True)  # the step's host-to-device copy starts behind e0 (not while the L2 flush is still running)
    # a ring of two pinned buffers, refilled by the host right before the step (outside the timed region) like a driver /
    # decoder would: the DMA then reads lines the CPU has just written instead of host memory that was last touched
    # seconds ago (a 3.84 MB copy: 76 us vs ~125 us on this box, scripts/h2d_probe.py)
    ring = [torch.empty((max(sc.xyzt.shape[0] for sc in scans), 4), dtype=torch.float32).pin_memory() for _ in range(2)]
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(W + K)]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in range(W + K)]
    iters_e2e = 0
    barrier()
    for k, sc in enumerate(scans):
        if k == W:
            barrier()
        buf = ring[k & 1][:sc.xyzt.shape[0]]
        buf.numpy()[...] = sc.xyzt
        flush.fill_(k & 0xFF)
        e0[k].record(stream)
        gx.step(buf.numpy(), sc.beg_time, sc.imu, True, MAX_ITER)
        e1[k].record(stream)
        if k >= W:
            iters_e2e += gx.timings().iekf_iters
    barrier()
    e2e_ms = np.array([e0[k].elapsed_time(e1[k]) for k in range(W, W + K)])
    t_e2e = float(e2e_ms.sum()) * 1e-3
    gx.close()

    # ---- leg 3: batch replay (BASELINE.json configs[4]): B INDEPENDENT sequences (seeds base + 0 .. B-1, own maps) on
    # one GPU in lock step through vina_batch - per-sequence stages on the contexts' own streams, the IEKF
    # iterations of all B sequences as ONE k_iekf launch per iteration (grid = blocks x B), the bandwidth-shaped form
    # of the kernel. Sequence 0 is the sequence of the legs above: its trajectory must come out the same.
    # (single-GPU runs only, like cpu_baseline: N x B full-size sequences would take minutes to generate)
    batch = None
    if args.batch > 1 and world == 1:
        B = args.batch
        more = synth.gen_sequences(cfg, [cfg.seed + b for b in range(1, B)], cfg.win_size, W + K)
        seqs = [(boots, scans)] + more
        ctxs = []
        for b in range(B):
            g = capi.Ctx(cfg, **caps)  # own CUDA stream each
            for sc in seqs[b][0]:
                g.bootstrap(sc.xyzt, capi.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
            g.set_imu_anchor(seqs[b][0][-1].end_time, seqs[b][0][-1].imu[-1])
            ctxs.append(g)
        bat = capi.Batch(ctxs)
        d_sc = [[torch.from_numpy(sc.xyzt).to(dev) for sc in seqs[b][1]] for b in range(B)]
        torch.cuda.synchronize(dev)
        lt, lb = 0.0, 0.0  # device time / algorithmic bytes of the batched launches with all sequences active
        n_l = 0
        b_pts, b_err, seq0_dev = 0.0, 0.0, 0.0
        for lo, hi in ((0, W), (W, W + K)):
            barrier()
            t0 = time.perf_counter()
            for k in range(lo, hi):
                row = [seqs[b][1][k] for b in range(B)]
                outs = bat.step_resident([d_sc[b][k].data_ptr() for b in range(B)], [sc.xyzt.shape[0] for sc in row],
                                         [sc.beg_time for sc in row], [sc.end_time for sc in row], [sc.imu for sc in row],
                                         True, MAX_ITER)
                if lo == W:
                    ms, _ = bat.iekf_time()
                    its = min(g.timings().iekf_iters for g in ctxs)
                    for j in range(its):  # launches in which every sequence still iterates
                        miss_j = n_mean if j == 0 else n_mean * (1 - match_frac)
                        lb += B * (80.0 * n_mean + 256.0 * U + 16.0 * miss_j + 34 * 8)
                        lt += ms[j] * 1e-3
                        n_l += 1
                    b_pts += sum(sc.xyzt.shape[0] for sc in row)
                    b_err = max(b_err, max(float(np.linalg.norm(np.array(o.p[:]) - sc.gt_p)) for o, sc in zip(outs, row)))
                    seq0_dev = max(seq0_dev, float(np.linalg.norm(np.array(outs[0].p[:]) - traj_p[k - W])))
            bat.sync()
            torch.cuda.synchronize(dev)
            t_batch = time.perf_counter() - t0
        bat.close()
        for g in ctxs:
            g.close()
        del d_sc
        batch = {"sequences_per_gpu": B, "seconds": t_batch, "points": b_pts, "launch_s": lt, "launch_bytes": lb,
                 "launches": n_l, "gt_traj_err_m": b_err, "seq0_dev": seq0_dev}

    # ---- max over ranks, aggregate --------------------------------------------------------------------
    from vina_slam_b200 import replicas

    per_rank = None
    if world > 1:  # every rank's own figures, for the record (the line's value uses the slowest rank's time)
        mine = torch.tensor([1e3 * t_res / K, 1e3 * t_e2e / K, iters / K, pts / K], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        allr = torch.stack(allr).cpu().numpy()
        per_rank = {"ms_per_step": allr[:, 0].tolist(), "e2e_ms_per_step": allr[:, 1].tolist(),
                    "iekf_iters_per_step": allr[:, 2].tolist(), "pts_per_scan": allr[:, 3].tolist()}
    pts_all, (t_res, t_e2e) = replicas.reduce_throughput(pts, [t_res, t_e2e], device=dev)
    if batch:
        bp, (bt,) = replicas.reduce_throughput(batch["points"], [batch["seconds"]], device=dev)
        ach = batch["launch_bytes"] / batch["launch_s"] / 1e9 if batch["launch_s"] > 0 else 0.0
        batch = {"sequences_per_gpu": batch["sequences_per_gpu"], "value": bp / bt, "unit": UNIT,
                 "seeds": [cfg.seed + b for b in range(batch["sequences_per_gpu"])],
                 "gt_traj_err_m": batch["gt_traj_err_m"],
                 "sequence0_max_dev_from_single_context_run_m": batch["seq0_dev"],
                 "ms_per_scan_amortised": 1e3 * bt / (K * batch["sequences_per_gpu"]),
                 "roofline": {"bound": "hbm", "kernel": "k_iekf (one launch for all sequences)", "achieved": ach,
                              "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                              "launch_us": 1e6 * batch["launch_s"] / max(batch["launches"], 1),
                              "bytes_per_launch": batch["launch_bytes"] / max(batch["launches"], 1),
                              "launches_timed": batch["launches"]},
                 "note": "vina_batch: B independent sequences (own seeds, own maps) in lock step on one GPU; "
                         "per-sequence stages on own streams, IEKF iterations batched into one launch; wall "
                         "clock over the K steps; no explicit L2 flush in this leg (the working set of the B "
                         "sequences - B maps, B x 21 MB of points rewritten every step - exceeds what stays "
                         "resident between a sequence's consecutive steps only partly: B maps stay hot as they "
                         "would in production)"}

    # ---- N > 1: the same scans through the map sharded by voxel-hash range (BASELINE.json configs[4]):
    # fused route + exchange over peer memory, IEKF against the sharded map, union checked against one GPU's map
    shard_res = None
    if world > 1 and not args.no_sharded:
        shard_res = run_sharded(cfg, rank, world, local, dev, 3, min(K, 10), p2p=True, query=True, verify=True)

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            n_cpu = min(K, 8)
            r = run_cpu(cfg, boots, scans, W if args.ba else min(W, 2), n_cpu, ba=args.ba)
            cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                   "sample": f"{r['steps']} full scans of the same workload ({cfg.n_points} pts each); IEKF 1 thread, "
                             f"map ops {cfg.thread_num} threads",
                   "ms_per_step": r["ms_per_step"],
                   "stage_ms": dict(zip(["odom", "insert", "recut", "margi"], r["stage_ms"]))}
        line = {
            "metric": METRIC, "value": pts_all / t_res, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * t_res / K, "ms_per_step_median": float(np.median(step_ms)),
            "ms_per_step_max": float(np.max(step_ms)), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(cfg), "max_iter": MAX_ITER, "iekf_on": "full scan",
                       "vnc_terms": False, "if_BA": int(args.ba), "parallelism": f"replicas x{world}",
                       "l2": "256 MiB buffer written between timed steps (L2 flush); steps timed individually "
                             "with CUDA events on the launching stream and summed; stage_ms / roofline launch times "
                             "come from a second, instrumented pass over the same scans",
                       "schedule": {"iekf": "k_iekf_loop: one persistent cooperative launch per scan" if loop_mode else
                                    "one k_iekf launch per iteration (832-thread blocks), update in the last block",
                                    "programmatic_dependent_launch": os.environ.get("VINA_PDL", "1") != "0",
                                    "side_stream": "down-sampling + var_init of the map's point set next to the IEKF launches"
                                    if not loop_mode else "none (two fused front launches in stream order)"},
                       "iekf_iters_per_step": iters_per_step, "gt_traj_err_m": traj_err,
                       **({"ba": {"runs_in_timed_steps": ba_runs, "lm_iters_last": ba_iters,
                                  "note": "LI_BA_Optimizer every scan: IMU factors + LM on the host, LiDAR factor "
                                          "(Hessian / residual over the plane voxels of the window) on the device"}}
                          if args.ba else {})},
            "e2e": {"value": pts_all / t_e2e, "unit": UNIT, "ms_per_step": 1e3 * t_e2e / K,
                    "ms_per_step_median": float(np.median(e2e_ms)), "ms_per_step_max": float(np.max(e2e_ms)),
                    "how": "vina_odom_step from a ring of two pinned host buffers refilled before every step; the "
                           "host-to-device copy is ordered behind the step's first CUDA event (vina_set_upload_ordered), so "
                           "it lies inside the timed region and cannot start while the L2 flush is still running; the "
                           "converged iterate comes back through mapped host memory",
                    # scan (16 B / point) + pose table (DeskewPoses) + the iterate (IekfDev, 2 616 B) up; the
                    # converged iterate + its sequence number + the down-sampled count back
                    "h2d_bytes_per_step": int(16 * n_mean + 17096 + 2616),
                    "d2h_bytes_per_step": 2616 + 8 + 4},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": roof_kernel, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "frac_of_8000_nominal": achieved / 8000.0, "traffic": traffic,
                         "traffic_source": traffic_source,
                         "peak_source": peak_src, "bytes_per_launch": bytes_per_launch, "launch_us": 1e3 * launch_ms,
                         "launches_timed": roof_launches, "unique_leaves": U, "match_frac": match_frac},
            "stage_ms": stage,
            "roofline_stages": {"note": "every stage of the step: algorithmic bytes (SURVEY 8d) / device time of the "
                                        "instrumented pass / measured HBM peak; units: points inserted %.0f, leaves "
                                        "touched %.0f, tree nodes under the window map %.0f, leaves subdivided %.0f per "
                                        "scan" % (n_ins, u_ins, u_fit, n_split),
                                "stages": roofline_stages,
                                "whole_step": {"algorithmic_bytes": whole, "ms": 1e3 * t_res / K,
                                               "achieved_GBs": whole / (t_res / K) / 1e9,
                                               "frac": whole / (t_res / K) / 1e9 / hbm_peak}},
            "clocks": clocks,
        }
        if per_rank:
            line["per_rank"] = per_rank
        if shard_res:
            line["sharded"] = shard_res
        if cpu:
            line["cpu_baseline"] = cpu
        if batch:
            line["batch"] = batch
        emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------- sharded map build
def run_sharded(cfg, rank, world, local, dev, W, K, p2p=True, query=True, verify=True, prefill=0):
    """Map partitioned by voxel-hash range over the ranks (BASELINE.json configs[4]). Every rank holds the scan,
    routes its ascending slice of the down-sampled points to the owners (p2p: the routing kernel stores the records
    straight into the owners' inboxes over peer memory / NVLink; else one NCCL all-to-all of 104-byte records),
    inserts what it received, recut + margi locally; with `query` the IEKF of every scan runs against the sharded
    map first (fused: queries and the 34 sums travel through peer memory, no collective, no host in the loop).
    Strong scaling: the scans are the same at every N. verify: after the timed scans rank 0 builds the whole map on
    one context from the same scans and the digests are compared (bit-exactness of the union) - outside the timed
    region. prefill: root voxels put into every rank's shard beforehand (planar patches far away from the building),
    the regime the split is for: a map that does not fit one GPU. Returns the result dict on rank 0 (None elsewhere).
    The process group must be initialised when world > 1."""
    import torch
    import torch.distributed as dist

    from vina_slam_b200 import capi, sharded

    seq = synth.Sequence(cfg, seed=cfg.seed)
    scans = [seq.next_scan(deskewed=True) for _ in range(cfg.win_size + W + K)]
    n_fill = int(prefill)
    caps = dict(max_scan_points=max(300000, cfg.n_points + 1024), device=local,
                max_nodes=int(n_fill * 1.15) + (1 << 20),
                hash_capacity_log2=max(21, int(np.ceil(np.log2(max(n_fill, 1) * 2.5)))),
                fix_pool_points=int(n_fill * 25 * 1.1) + (16 << 20))
    free0 = torch.cuda.mem_get_info(dev)[0]
    sh = sharded.MapShard(capi.Ctx(cfg, **caps), rank, world, device=dev)
    if p2p:
        sh.p2p_setup(caps["max_scan_points"])  # inboxes + CUDA IPC handles all-gathered over the group
    t_fill = 0.0
    if n_fill:
        # world x n_fill root voxels through the sharded build itself (every rank generates the same clouds and
        # routes its slice; the owners end up with ~n_fill voxels each)
        z9 = np.zeros(9)
        ident = np.eye(3).reshape(-1)

        def feed(gx):
            gx.downsample()
            nd = gx.n_down()
            gx.var_init(1)
            first, cnt = sharded.slice_of(nd, rank, world)
            (sh.update_p2p if p2p else sh.update)(first, cnt, 0, ident, np.zeros(3), z9, z9)

        t_fill = prefill_map(sh.ctx, n_fill * world, dev, feed=feed)
    iek = sharded.ShardedIekf(sh) if query else None
    q_iters, q_err, q_ms = 0, 0.0, 0.0
    stream = torch.cuda.current_stream(dev)
    d_scans = [torch.from_numpy(sc.xyzt).to(dev) for sc in scans]
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    n_down_tot, n_recv_tot = 0, 0
    final_states = []

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    barrier()
    for k, sc in enumerate(scans):
        if k == cfg.win_size + W:
            barrier()
            ev0.record(stream)
        st = capi.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time)
        sa = capi.state_arrays(st)
        Rc = np.ascontiguousarray(sa["R"].T.reshape(-1))
        rv = np.ascontiguousarray(sa["cov"][0:3, 0:3].T.reshape(-1))
        tv = np.ascontiguousarray(sa["cov"][3:6, 3:6].T.reshape(-1))
        # the (already deskewed) scan is resident; down-sample + var_init on every rank, route this rank's slice
        sh.ctx.scan_upload_device(d_scans[k].data_ptr(), sc.xyzt.shape[0])
        if iek is not None and k >= cfg.win_size:
            # association + IEKF against the sharded map from a perturbed start (same on every rank)
            pert = capi.make_state(sc.gt_R @ synth.rot_exp(np.array([1e-3, -1e-3, 1e-3])),
                                   sc.gt_p + np.array([0.01, -0.01, 0.005]), sc.gt_v, t=sc.end_time)
            sh.ctx.set_state(pert)
            sh.ctx.var_init(0)
            torch.cuda.synchronize(dev)
            tq0 = time.perf_counter()
            q_iters += (iek.run_p2p if p2p else iek.run)(*sharded.slice_of(sc.xyzt.shape[0], rank, world), MAX_ITER)
            if k >= cfg.win_size + W:
                q_ms += 1e3 * (time.perf_counter() - tq0)  # the loop ends with the converged state on the host
            got = capi.state_arrays(sh.ctx.get_state())
            q_err = max(q_err, float(np.linalg.norm(got["p"] - sc.gt_p)))
            final_states.append(got["p"].copy())
        sh.ctx.downsample()
        nd = sh.ctx.n_down()
        sh.ctx.var_init(1)
        first, cnt = sharded.slice_of(nd, rank, world)
        got_n = (sh.update_p2p if p2p else sh.update)(first, cnt, 0, Rc, sa["p"], rv, tv)
        if k >= cfg.win_size + W:
            n_down_tot += nd
            n_recv_tot += got_n
    ev1.record(stream)
    barrier()
    t = ev0.elapsed_time(ev1) * 1e-3
    tt = torch.tensor([t], dtype=torch.float64, device=dev)
    exp = sh.ctx.map_export() if not n_fill else None
    if exp is not None and n_fill == 0:
        dig_val, n_nodes = sharded.map_digest(exp), sh.ctx.map_count()[0]
    else:
        dig_val, n_nodes = 0, sh.ctx.map_count()[0]
    dig = torch.tensor([dig_val, n_nodes, n_recv_tot], dtype=torch.int64, device=dev)
    mem_gb = (free0 - torch.cuda.mem_get_info(dev)[0]) / 1e9
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(dig)
    out = None
    if rank == 0:
        pts = sum(sc.xyzt.shape[0] for sc in scans[cfg.win_size + W:])
        out = {"value": pts / float(tt[0]), "unit": UNIT, "ms_per_scan": 1e3 * float(tt[0]) / K, "scaling": "strong",
               "workload": workload_name(cfg), "parallelism": f"hash-range x{world}",
               "what": "per scan: " + ("IEKF against the sharded map, then " if query else "") +
                       "down-sample -> var_init -> route + exchange -> insert -> recut -> margi on every shard",
               "exchange": ("records stored into the owners' inboxes by the routing kernel over peer memory "
                            "(CUDA IPC / NVLink), device-side flags" if p2p else "NCCL all_to_all_single of staged records"),
               "down_points_per_scan": n_down_tot / K, "routed_points_per_scan": int(dig[2]) / K,
               "record_bytes": 8 * sharded.REC, "nodes_all_shards": int(dig[1]),
               "prefilled_root_voxels_per_rank": n_fill, "prefill_seconds": t_fill, "device_memory_GB_rank0": mem_gb}
        if iek is not None:
            out["sharded_iekf"] = {"loop": ("fused: queries and the 34 sums travel through peer memory, update on every "
                                            "rank's device iterate, no host sync inside the loop" if p2p else
                                            "per iteration: route, NCCL all-to-all, accumulate, NCCL all-reduce, host update"),
                                   "loop_ms_per_scan_rank0_wall": q_ms / K, "iters_per_scan": q_iters / (W + K),
                                   "pos_err_vs_ground_truth_m": q_err}
        if verify and n_fill == 0:
            # the same scans through ONE context (outside the timed region): the union of the shards must be this map
            ref = capi.Ctx(cfg, **caps)
            q_ref_err, q_ref_iters = 0.0, 0
            for k, sc in enumerate(scans):
                st = capi.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time)
                ref.scan_upload_device(d_scans[k].data_ptr(), sc.xyzt.shape[0])
                if iek is not None and k >= cfg.win_size:
                    pert = capi.make_state(sc.gt_R @ synth.rot_exp(np.array([1e-3, -1e-3, 1e-3])),
                                           sc.gt_p + np.array([0.01, -0.01, 0.005]), sc.gt_v, t=sc.end_time)
                    ref.set_state(pert)
                    ref.var_init(0)
                    q_ref_iters += ref.odom_iekf(0, MAX_ITER, host_solve=True)[0]
                    q_ref_err = max(q_ref_err, float(np.linalg.norm(
                        final_states[k - cfg.win_size] - capi.state_arrays(ref.get_state())["p"])))
                ref.set_state(st)
                ref.downsample()
                ref.n_down()
                ref.var_init(1)
                ref.odom_map_update()
            out["union_equals_single_gpu_map"] = bool(int(dig[0]) == sharded.map_digest(ref.map_export()))
            out["single_gpu_nodes"] = ref.map_count()[0]
            if iek is not None:
                out["sharded_iekf"]["pos_diff_vs_single_gpu_m"] = q_ref_err
                out["sharded_iekf"]["iters_per_scan_single_gpu"] = q_ref_iters / (W + K)
            ref.close()
    barrier()
    sh.ctx.close()
    return out


def prefill_map(gx, n_vox, dev, seed=1234, origin=(3000.0, 3000.0, 100.0), per_scan_vox=10000, ppv=25, feed=None):
    """n_vox root voxels with a small planar patch each (ppv points per 1 m voxel), inserted through the product path
    (feed = what to do with each uploaded cloud; default: the single-context map update), in a slab `origin` away from
    the building. Returns the seconds it took."""
    import torch

    from vina_slam_b200 import capi

    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    gx.set_state(capi.make_state(np.eye(3), np.zeros(3), np.zeros(3)))
    side = int(np.ceil(np.sqrt(n_vox / 8)))  # X x Y x 8 layers
    t0 = time.perf_counter()
    for v0 in range(0, n_vox, per_scan_vox):
        idx = torch.arange(v0, min(v0 + per_scan_vox, n_vox), device=dev)
        ix, iy, iz = idx % side, (idx // side) % side, idx // (side * side)
        c = torch.stack([ix, iy, iz], 1).double() + 0.5 + torch.tensor(list(origin), device=dev, dtype=torch.float64)
        nrm = torch.randn((idx.numel(), 3), generator=gen, device=dev, dtype=torch.float64)
        nrm = nrm / nrm.norm(dim=1, keepdim=True)
        a = torch.linalg.cross(nrm, torch.tensor([0.3, 0.5, 0.81], device=dev, dtype=torch.float64).expand_as(nrm))
        a = a / a.norm(dim=1, keepdim=True)
        b = torch.linalg.cross(nrm, a)
        uv = (torch.rand((idx.numel(), ppv, 2), generator=gen, device=dev, dtype=torch.float64) - 0.5) * 0.56
        nz = torch.randn((idx.numel(), ppv, 1), generator=gen, device=dev, dtype=torch.float64) * 0.004
        pts = c[:, None, :] + uv[..., :1] * a[:, None, :] + uv[..., 1:] * b[:, None, :] + nz * nrm[:, None, :]
        xyzt = torch.cat([pts.reshape(-1, 3), torch.zeros((idx.numel() * ppv, 1), device=dev, dtype=torch.float64)],
                         1).float().contiguous()
        torch.cuda.synchronize(dev)
        gx.scan_upload_device(xyzt.data_ptr(), xyzt.shape[0])
        if feed is not None:
            feed(gx)
            continue
        gx.downsample()
        gx.n_down()
        gx.var_init(1)
        gx.odom_map_update()
    gx.sync()
    return time.perf_counter() - t0


def main_sharded(args, cfg):
    """--mode sharded-map: the sharded map build (+ IEKF) alone, as its own JSON line."""
    import torch
    import torch.distributed as dist

    rank, world, local = dist_env()
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    r = run_sharded(cfg, rank, world, local, dev, args.warmup, args.steps, p2p=args.p2p or world > 1 and not args.nccl,
                    query=args.query, verify=args.verify, prefill=args.voxels if args.voxels_set else 0)
    if rank == 0:
        line = {"metric": "pts/s voxel-map build sharded by hash range (down-sample -> var_init -> route -> exchange "
                          "-> insert -> recut -> margi, per scan)", "value": r["value"], "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_scan"], "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(cfg), "mode": "sharded-map", **{k: v for k, v in r.items()
                                                                                      if k not in ("value", "unit")}},
                "gpu_launches": None}
        emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------- large resident map (configs[3])
def main_bigmap(args, cfg):
    """BASELINE.json configs[3]: the per-scan path with ~10^7 root voxels resident. The map is pre-filled on the
    device through the product path itself (synthetic planar patches, ~25 points per 1 m voxel, far away from the
    sequence's own world), then the sequence runs as in the headline leg. Reported: ms/scan and pts/s with the
    large map, the same without it, the resident voxel count and the memory footprint."""
    import torch

    from vina_slam_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    W, K = args.warmup, args.steps
    EXTRA = 5  # scans run after the pruning of the pre-filled map
    boots, scans = gen_sequence(cfg, cfg.seed, cfg.win_size, W + K + EXTRA)
    n_vox = int(args.voxels)
    per_scan_vox, ppv = 10000, 25
    caps = dict(max_scan_points=max(300000, cfg.n_points + 1024), max_nodes=int(n_vox * 1.15) + (1 << 20),
                hash_capacity_log2=max(21, int(np.ceil(np.log2(max(n_vox, 1) * 2.5)))),
                fix_pool_points=int(n_vox * ppv * 1.1) + (16 << 20), device=0)
    stream = torch.cuda.current_stream(dev)
    out = {}
    # one pruning on a throw-away context first: the first call in a process pays one-time costs (≈0.9 s measured:
    # lazy loading of the pruning kernels, first cudaMalloc outside the pools) that are not the operation's
    warm = capi.Ctx(cfg, max_scan_points=max(300000, cfg.n_points + 1024), max_nodes=1 << 18, hash_capacity_log2=18,
                    fix_pool_points=4 << 20, device=0)
    for sc in boots:
        warm.bootstrap(sc.xyzt, capi.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
    warm.map_prune(1e6, 700)
    warm.close()
    for label, fill in (("empty", 0), ("filled", n_vox)):
        free0 = torch.cuda.mem_get_info(dev)[0]
        gx = capi.Ctx(cfg, **(caps if fill else dict(caps, max_nodes=1 << 20, hash_capacity_log2=21,
                                                     fix_pool_points=16 << 20)))
        gx.set_stream(stream.cuda_stream)
        t_fill = 0.0
        if fill:
            t_fill = prefill_map(gx, fill, dev)
        nodes, roots, slide = gx.map_count()
        for sc in boots:
            gx.bootstrap(sc.xyzt, capi.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
        gx.set_imu_anchor(boots[-1].end_time, boots[-1].imu[-1])
        d_scans = [torch.from_numpy(sc.xyzt).to(dev) for sc in scans]
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        gx.set_profiling(True)
        ev0 = [torch.cuda.Event(enable_timing=True) for _ in scans]
        ev1 = [torch.cuda.Event(enable_timing=True) for _ in scans]
        err, rows = 0.0, []
        prune = None
        for k, sc in enumerate(scans):
            if k == W + K:
                # the idle path's map pruning (local_mapping.cpp:317-341) with the journey advanced by 1000 m:
                # everything that is not in the sliding window's map is 700 m or more behind and is erased. Wall
                # clock around the call (it synchronises: mark -> sweep + hash rebuild -> pool compaction).
                torch.cuda.synchronize(dev)
                n_b, r_b, s_b = gx.map_count()
                t0 = time.perf_counter()
                er, fr = gx.map_prune(gx.journey()[0] + 1000.0, 700)
                t_pr = time.perf_counter() - t0
                n_a, r_a, _ = gx.map_count()
                prune = {"ms": 1e3 * t_pr, "roots_erased": er, "nodes_freed": fr, "nodes_before": n_b,
                         "roots_before": r_b, "nodes_after": n_a, "roots_after": r_a}
            flush.fill_(k & 0xFF)
            ev0[k].record(stream)
            st = gx.step_resident(d_scans[k].data_ptr(), sc.xyzt.shape[0], sc.beg_time, sc.end_time, sc.imu, True, MAX_ITER)
            ev1[k].record(stream)
            if W <= k < W + K:
                rows.append(gx.timings())
                err = max(err, float(np.linalg.norm(np.array(st.p[:]) - sc.gt_p)))
            elif k >= W + K:
                prune["gt_traj_err_m_after"] = max(prune.get("gt_traj_err_m_after", 0.0),
                                                   float(np.linalg.norm(np.array(st.p[:]) - sc.gt_p)))
        torch.cuda.synchronize(dev)
        prune["ms_per_scan_after"] = float(np.mean([ev0[k].elapsed_time(ev1[k]) for k in range(W + K + 1, W + K + EXTRA)]))
        prune["nodes_end"] = gx.map_count()[0]
        ms = float(np.mean([ev0[k].elapsed_time(ev1[k]) for k in range(W, W + K)]))
        stage = {f: float(np.mean([getattr(t, f) for t in rows])) for f in
                 ("iekf_ms", "insert_ms", "recut_ms", "margi_ms", "total_ms")}
        n2, r2, s2 = n_b, r_b, s_b  # at the end of the timed scans, before the pruning
        out[label] = {"ms_per_scan": ms, "pts_per_s": cfg.n_points / (ms * 1e-3), "prefilled_root_voxels": roots,
                      "prefilled_nodes": nodes, "nodes_after": n2, "roots_after": r2, "slide_roots": s2,
                      "prefill_seconds": t_fill, "gt_traj_err_m": err, "stage_ms": stage,
                      "device_memory_GB": (free0 - torch.cuda.mem_get_info(dev)[0]) / 1e9, "prune": prune}
        gx.close()
        del d_scans, flush
        torch.cuda.empty_cache()
    line = {"metric": METRIC + " with a large resident map", "value": out["filled"]["pts_per_s"], "unit": UNIT, "n_gpus": 1,
            "steps": K, "warmup": W, "ms_per_step": out["filled"]["ms_per_scan"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(cfg), "mode": "bigmap", "requested_voxels": n_vox,
                       "l2": "256 MiB buffer written between timed steps"},
            "with_large_map": out["filled"], "without": out["empty"]}
    emit(json.dumps(line))


_JSON_OUT = None


def emit(text):
    """The one JSON line goes to the process's real stdout; everything else any library writes to file descriptor 1
    (NCCL prints its version banner there) has been sent to stderr by main()."""
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(text + "\n")
    out.flush()


def main():
    global _JSON_OUT
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="robosense128", choices=sorted(synth.SENSORS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--ba", action="store_true", help="LocalBA.if_BA: 1 (mid360.yaml / velodyne.yaml): sliding-window BA every scan")
    ap.add_argument("--batch", type=int, default=8, help="concurrent sequences per GPU in the batch-replay leg (0/1 = off)")
    ap.add_argument("--voxels", type=float, default=None, help="bigmap: root voxels to pre-fill (default 1e7); "
                    "sharded-map: root voxels pre-filled into every rank's shard (default 0)")
    ap.add_argument("--nccl", action="store_true", help="sharded-map: staged records + NCCL all-to-all instead of the fused exchange")
    ap.add_argument("--mode", default="odometry", choices=["odometry", "sharded-map", "bigmap"],
                    help="odometry = the headline per-scan path (default); sharded-map = map build partitioned by "
                         "voxel-hash range over the ranks (SURVEY 8e)")
    ap.add_argument("--no-sharded", action="store_true", help="N > 1: skip the hash-range-sharded leg of the default line")
    ap.add_argument("--verify", action="store_true", help="sharded-map: rank 0 also builds the single-GPU map and compares")
    ap.add_argument("--p2p", action="store_true", help="sharded-map: fused route + exchange over peer memory instead of NCCL")
    ap.add_argument("--query", action="store_true", help="sharded-map: also run the IEKF against the sharded map every scan")
    args = ap.parse_args()
    args.voxels_set = args.voxels is not None
    if args.voxels is None:
        args.voxels = 1e7
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: W >= 3
    cfg = synth.SENSORS[args.workload]
    if args.impl == "reference":
        main_reference(args, cfg)
    elif args.mode == "sharded-map":
        main_sharded(args, cfg)
    elif args.mode == "bigmap":
        main_bigmap(args, cfg)
    else:
        main_ours(args, cfg)


if __name__ == "__main__":
    main()

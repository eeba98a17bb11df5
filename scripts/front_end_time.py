"""Time of the scan front end (vina_scan_prepare_device: decoder keep rule + stable sort by time offset + 0.11 s cut)
on a RoboSense-128-shaped raw scan (240 000 points, shuffled), CUDA events around the call, against the CPU
restatement (oracle -O3 build, std::stable_sort; the reference uses std::sort in pcl_handler). One JSON line."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from oracle import oracle_py as op  # noqa: E402  (CPU baseline leg only)
from vina_slam_b200 import capi, synth  # noqa: E402

cfg = synth.SENSORS["robosense128"]
rng = np.random.default_rng(0)
n = cfg.n_points
a = np.zeros((n, 4), dtype=np.float32)
a[:, :3] = rng.uniform(-40, 40, (n, 3))
a[:, 3] = (np.arange(n) // 128).astype(np.float32) * np.float32(0.1 / (n // 128))  # 128 beams fire together
a = a[rng.permutation(n)]
gx = capi.Ctx(cfg, max_scan_points=300000)
stream = torch.cuda.current_stream()
gx.set_stream(stream.cuda_stream)
d = torch.from_numpy(a).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    k, tl = gx.scan_prepare(n, 1, 0.01, d_ptr=d.data_ptr())
ms = []
for i in range(20):
    flush.fill_(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    k, tl = gx.scan_prepare(n, 1, 0.01, d_ptr=d.data_ptr())
    e1.record(stream)
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
g = gx.scan_download(k)
t0 = time.perf_counter()
for _ in range(5):
    o = op.scan_prepare(a, 1, 0.01, fast=True)
cpu_ms = (time.perf_counter() - t0) / 5 * 1e3
assert np.array_equal(o, g)
# algorithmic bytes: 16 B read + 16 B written per point (the sort's passes are implementation traffic)
print(json.dumps({"what": "vina_scan_prepare_device, 240000 shuffled points", "gpu_ms": float(np.median(ms)),
                  "gpu_ms_min": float(np.min(ms)), "points_out": int(k), "launches": int(gx.timings().kernel_launches) if False else 3,
                  "algorithmic_GBps": 32 * n / (np.median(ms) * 1e-3) / 1e9, "cpu_ms_oracle_O3_1thread": cpu_ms,
                  "l2": "256 MiB buffer written between timed calls"}))
gx.close()

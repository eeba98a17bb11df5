"""Golden vectors of the REFERENCE BUILD (oracle/_ref) for the callers around the hot path: the six message handlers,
pcl_handler, and the map pruning of the idle path. Generator of tests/golden/ref_front.npz; `scenario(impl)` is
shared with tests/test_oracle_vs_ref.py, which replays it on the restatements where /root/reference is not mounted.
sync_packages has its own process-per-mode driver (sync_vs_ref.py) and is not in the file.

    python tests/golden/make_ref_front_golden.py        # needs oracle/_ref (make -C oracle ref)
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import oracle_py as op  # noqa: E402
from vina_slam_b200 import synth  # noqa: E402


def _td():
    spec = importlib.util.spec_from_file_location("test_decode_cpu", os.path.join(HERE, "..", "test_decode_cpu.py"))
    td = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(td)
    return td


def scenario(ref: bool):
    """ref=True: the reference build; False: the restatements (numpy handlers, C++ scan_prepare / Odom)."""
    td = _td()
    out = {}
    stamp = float(1000250000000) * 1e-9
    for lidar_type in (1, 2, 3, 4, 5):
        rng = np.random.default_rng(500 + lidar_type)
        dt, tname, code = td.LAYOUTS[lidar_type]
        off_t = dt.fields[tname][1] if tname else -1
        a = td._cloud(lidar_type, 1500, rng, stamp)
        if ref:
            r = op.decode_handler_ref(lidar_type, a.tobytes(), 1500, dt.itemsize, [0, 4, 8], off_t, code, stamp, 0.04, 2)
        else:
            r = op.decode_handler(lidar_type, td._as_struct(a, tname), stamp, 0.04, 2)
        out[f"decode_{lidar_type}"] = r
    rng = np.random.default_rng(77)
    n = 6000
    a = np.zeros((n, 4), dtype=np.float32)
    a[:, :3] = rng.uniform(-30, 30, (n, 3))
    a[::7, :3] *= 0.01
    a[:, 3] = (rng.permutation(n).astype(np.float64) * (0.125 / n)).astype(np.float32)
    a[-1, 3] = np.float32(0.05)
    r = op.scan_prepare(a, 2, 0.01, ref=ref)
    out["scan_prepare"] = r[r[:, 3] != np.float32(0.05)]
    # pruning: bootstrap + 24 scans at ground-truth poses, idle path with a 1 m horizon after every scan
    cfg = synth.small_sensor("robosense128", 16, 240, seed=5)
    seq = synth.Sequence(cfg)
    od = op.Odom(cfg, ref=ref)
    rows = []
    try:
        for k in range(cfg.win_size + 24):
            sc = seq.next_scan(deskewed=True)
            od.bootstrap(sc.xyzt, op.make_state(sc.gt_R, sc.gt_p, sc.gt_v, t=sc.end_time))
            j, f = od.journey()
            e, fr = od.idle(1)
            rows.append([j, float(f), e, fr, *od.map_count()])
    finally:
        od.close()
    out["prune_rows"] = np.array(rows, dtype=np.float64)
    return out


if __name__ == "__main__":
    g = scenario(True)
    np.savez_compressed(os.path.join(HERE, "ref_front.npz"), **g)
    print({k: v.shape for k, v in g.items()}, "erased", int(g["prune_rows"][:, 2].sum()))

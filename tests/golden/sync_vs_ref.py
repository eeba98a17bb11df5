"""Drives the oracle's restatement of sync_packages and the REFERENCE'S OWN src/sensor/sync.cpp (oracle/_ref) with the
same randomised message stream and compares every return. Run as a script, one process per mode: the reference keeps
the state in globals and a function-local static. Usage: sync_vs_ref.py <point_notime> <seed>; prints `OK <packages>
<dropped>`. Used by tests/test_oracle_vs_ref.py."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle import oracle_py as op  # noqa: E402


def main(point_notime: int, seed: int):
    rng = np.random.default_rng(seed)
    a, b = op.Sync(point_notime), op.Sync(point_notime, ref=True)
    packages = dropped = held = 0
    t_imu_ns, t_scan, tag = 0, 0.05, 0
    rate_ns = 5_000_000
    for step in range(6000):
        if step % 1500 == 0:
            rate_ns = int(rng.choice([5_000_000, 2_500_000, 33_000_000]))  # 200 Hz, 400 Hz, 30 Hz (scans get dropped)
        r = rng.uniform()
        if r < 0.75:
            t_imu_ns += rate_ns
            imu = np.concatenate([[float(t_imu_ns) * 1e-9], rng.normal(size=6)])
            a.push_imu(imu)
            b.push_imu(imu)
        elif r < 0.85:
            tl = float(np.float32(rng.uniform(0.08, 0.1)))
            a.push_scan(t_scan, tl, tag)
            b.push_scan(t_scan, tl, tag)
            t_scan += 0.1
            tag += 1
        else:
            ra, rb = a.next(), b.next()
            assert ra[0] == rb[0], (step, ra[:4], rb[:4])
            if ra[0] in (1, 2):
                assert ra[1] == rb[1], (step, ra[:4], rb[:4])
                if not (point_notime and ra[0] == 2 and ra[4].shape[0] == 0 and ra[2] == 0.0):
                    assert ra[2:4] == rb[2:4], (step, ra[:4], rb[:4])
                assert np.array_equal(ra[4], rb[4]), step
            packages += ra[0] == 1
            dropped += ra[0] == 2
    print("OK", packages, dropped)


if __name__ == "__main__":
    main(int(sys.argv[1]), int(sys.argv[2]))

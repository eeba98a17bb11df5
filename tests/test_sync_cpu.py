"""sync_packages and its buffers (src/sensor/sync.cpp:5-96) behind the C ABI: host-only entry points of
libvina_b200.so (no device needed) against the oracle's restatement on randomised interleavings of IMU and scan
arrivals, plus the scripted corner cases of the reference's logic."""
import numpy as np
import pytest


def _imu(t, rng=None):
    v = np.zeros(7)
    v[0] = t
    if rng is not None:
        v[1:] = rng.normal(size=6)
    return v


def _drive(events, point_notime, gpu_capi, oracle_lib, cap=256):
    a, b = gpu_capi.Sync(point_notime), oracle_lib.Sync(point_notime)
    out = []
    try:
        for ev in events:
            if ev[0] == "imu":
                a.push_imu(ev[1])
                b.push_imu(ev[1])
            elif ev[0] == "scan":
                a.push_scan(*ev[1:])
                b.push_scan(*ev[1:])
            else:
                ra, rb = a.next(cap), b.next(cap)
                assert ra[0] == rb[0], (ra[:4], rb[:4])
                if ra[0] != 0:
                    assert ra[1] == rb[1]
                if ra[0] in (1, 2):
                    assert ra[2:4] == rb[2:4] and np.array_equal(ra[4], rb[4])
                out.append(ra)
    finally:
        a.close()
        b.close()
    return out


@pytest.fixture(scope="module")
def capi_host():
    from vina_slam_b200 import capi

    capi.load()
    return capi


@pytest.mark.parametrize("point_notime", [0, 1])
def test_sync_packages_random_interleavings(capi_host, oracle_lib, point_notime):
    rng = np.random.default_rng(5 + point_notime)
    packages, dropped = 0, 0
    for trial in range(40):
        rate = rng.choice([200.0, 400.0, 30.0])  # 30 Hz IMU: <= 4 samples per 0.1 s scan -> scans are dropped
        t_imu, t_scan, tag = 0.0, 0.05 + rng.uniform(0, 0.01), 0
        ev = []
        for _ in range(rng.integers(50, 400)):
            r = rng.uniform()
            if r < 0.75:
                t_imu += 1.0 / rate
                ev.append(("imu", _imu(t_imu, rng)))
            elif r < 0.85:
                ev.append(("scan", t_scan, float(np.float32(rng.uniform(0.08, 0.1))), tag))
                t_scan += 0.1
                tag += 1
            else:
                ev.append(("next",))
        ev += [("next",)] * 5
        res = _drive(ev, point_notime, capi_host, oracle_lib)
        packages += sum(1 for r in res if r[0] == 1)
        dropped += sum(1 for r in res if r[0] == 2)
        for r in res:
            if r[0] == 1:
                assert r[4].shape[0] > 4 and np.all(np.diff(r[4][:, 0]) > 0) and r[4][-1, 0] <= r[3]
    assert packages > 50 and dropped > 5, (packages, dropped)


def test_sync_packages_corner_cases(capi_host, oracle_lib):
    # nothing buffered; a scan whose end the IMU stream has not passed is held, not consumed
    ev = [("next",), ("scan", 10.0, 0.1, 7), ("next",)] + [("imu", _imu(10.0 + 0.01 * k)) for k in range(10)] + [("next",)]
    res = _drive(ev, 0, capi_host, oracle_lib)
    assert [r[0] for r in res] == [0, 0, 0]
    # an IMU sample stamped exactly pcl_end_time belongs to the package (sync.cpp:68-76); the later one stays
    ev = [("scan", 10.0, 0.1, 7)] + [("imu", _imu(t)) for t in (9.99, 10.0, 10.02, 10.04, 10.06, 10.08, 10.1, 10.12)] + [("next",)]
    r = _drive(ev, 0, capi_host, oracle_lib)[0]
    assert r[0] == 1 and r[1] == 7 and r[2] == 10.0 and r[3] == 10.0 + 0.1 and r[4].shape[0] == 7 and r[4][-1, 0] == 10.1
    # <= 4 samples: sync_packages returns false and the scan is gone (sync.cpp:87-95)
    ev = [("scan", 10.0, 0.1, 3)] + [("imu", _imu(t)) for t in (10.0, 10.05, 10.2)] + [("next",), ("next",)]
    res = _drive(ev, 0, capi_host, oracle_lib)
    assert res[0][0] == 2 and res[0][1] == 3 and res[1][0] == 0
    # the output buffer is too small
    ev = [("scan", 10.0, 0.1, 1)] + [("imu", _imu(10.0 + 0.01 * k)) for k in range(12)] + [("next",)]
    assert _drive(ev, 0, capi_host, oracle_lib, cap=3)[0][0] == -3
    # ... and nothing is lost: the scan stays held, the call can be repeated with a larger buffer
    s = capi_host.Sync()
    s.push_scan(10.0, 0.1, 1)
    for k in range(12):
        s.push_imu(_imu(10.0 + 0.01 * k))
    r = s.next(3)
    assert r[0] == -3 and r[1] == 1 and s.pending() == (1, 12)
    r = s.next(64)
    assert r[0] == 1 and r[1] == 1 and r[4].shape[0] == 11 and s.pending() == (0, 1)
    s.close()
    # point_notime: the first scan only seeds the interval, the second spans [first stamp, second stamp]
    ev = [("scan", 10.0, 0.0, 0), ("scan", 10.1, 0.0, 1)] + [("imu", _imu(10.0 + 0.01 * k)) for k in range(12)] + [("next",), ("next",)]
    res = _drive(ev, 1, capi_host, oracle_lib)
    assert res[0][0] == 2 and res[0][1] == 0
    assert res[1][0] == 1 and res[1][1] == 1 and res[1][2] == 10.0 and res[1][3] == 10.1 and res[1][4].shape[0] == 11
    # argument errors come back as codes
    import ctypes as C

    lib = capi_host.load()
    assert lib.vina_sync_create(C.c_int(0), None) == -1 and lib.vina_sync_push_imu(None, None) == -1
    s = capi_host.Sync()
    assert s.pending() == (0, 0)
    s.push_scan(1.0, 0.1, 9)
    s.push_imu(_imu(0.5))
    assert s.pending() == (1, 1)
    s.close()


def test_stream_packages_reproduce_the_synthetic_pairing(capi_host):
    """replay.stream_packages: scans and IMU samples as two message streams through vina_sync give, for every scan
    but the pending last one, exactly the package the synthetic generator attaches to it (the IMU samples with
    stamps in (previous end, end], sync.cpp:63-76) and pcl_beg / pcl_end of sync.cpp:36-40."""
    from vina_slam_b200 import replay, synth

    for base in ("robosense128", "hilti_xt32"):
        cfg = synth.small_sensor(base, 8, 60)
        seq = synth.Sequence(cfg)
        for _ in range(cfg.win_size):
            seq.next_scan(deskewed=True)
        scans = [seq.next_scan() for _ in range(8)]
        got = list(replay.stream_packages(scans))
        assert [g[0] for g in got] == list(range(7))
        for k, beg, end, imu in got:
            assert beg == scans[k].beg_time and end == scans[k].beg_time + float(scans[k].xyzt[-1, 3])
            assert np.array_equal(imu, scans[k].imu), k

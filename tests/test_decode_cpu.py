"""The decoders' handlers (src/sensor/lidar_pointcloud_decoder.cpp:21-240) behind the C ABI: host-only entry points of
libvina_b200.so on synthetic sensor_msgs/PointCloud2 byte buffers in the drivers' layouts, against the oracle's numpy
restatement of every handler (bit for bit; the azimuth-derived stamps of the Velodyne fallback to 1 float ulp, since
numpy's float arctan2 need not be libm's atan2f)."""
import numpy as np
import pytest

from oracle import oracle_py as op

LAYOUTS = {
    # lidar_type: (numpy dtype of one point as the driver publishes it, time field, PointField datatype)
    1: (np.dtype({"names": ["x", "y", "z", "intensity", "ring", "time"], "formats": ["<f4", "<f4", "<f4", "<f4", "<u2", "<f4"],
                  "offsets": [0, 4, 8, 12, 16, 18], "itemsize": 22}), "time", 7),
    2: (np.dtype({"names": ["x", "y", "z", "intensity", "t", "reflectivity", "ring", "range"],
                  "formats": ["<f4", "<f4", "<f4", "<f4", "<u4", "<u2", "u1", "<u4"],
                  "offsets": [0, 4, 8, 16, 20, 24, 26, 32], "itemsize": 48}), "t", 6),
    3: (np.dtype({"names": ["x", "y", "z", "intensity", "timestamp", "ring"], "formats": ["<f4", "<f4", "<f4", "<f4", "<f8", "<u2"],
                  "offsets": [0, 4, 8, 12, 16, 24], "itemsize": 32}), "timestamp", 8),
    4: (np.dtype({"names": ["x", "y", "z", "intensity", "ring", "timestamp"], "formats": ["<f4", "<f4", "<f4", "<f4", "<u2", "<f8"],
                  "offsets": [0, 4, 8, 12, 16, 18], "itemsize": 26}), "timestamp", 8),
    5: (np.dtype({"names": ["x", "y", "z"], "formats": ["<f4", "<f4", "<f4"], "offsets": [0, 4, 8], "itemsize": 16}), None, 0),
}


@pytest.fixture(scope="module")
def capi_host():
    from vina_slam_b200 import capi

    capi.load()
    return capi


def _cloud(lidar_type, n, rng, stamp, spin=False):
    dt, tname, _ = LAYOUTS[lidar_type]
    a = np.zeros(n, dtype=dt)
    if spin:  # a spinning sensor: azimuth sweeps 0 .. -360 deg (clockwise), a few rings
        az = -np.linspace(0.05, 2 * np.pi - 0.05, n)
        r = rng.uniform(0.3, 40, n)
        a["x"], a["y"], a["z"] = r * np.cos(az), r * np.sin(az), rng.uniform(-2, 2, n)
    else:
        for k in "xyz":
            a[k] = rng.uniform(-30, 30, n)
    a["x"][::9] *= 0.01
    a["y"][::9] *= 0.01  # some returns inside the planar blind zone ...
    a["z"][::18] *= 0.001  # ... half of them inside the 3-D one as well
    frac = np.sort(rng.uniform(0, 0.1, n))
    if tname == "time":
        a[tname] = 0.0 if spin else frac
    elif tname == "t":
        a[tname] = (frac * 1e9).astype(np.uint32)
    elif tname == "timestamp":
        a[tname] = stamp + 0.003 + frac
    return a


def _as_struct(a, tname):
    dt = [("x", "<f4"), ("y", "<f4"), ("z", "<f4")] + ([("t", a.dtype[tname])] if tname else [])
    s = np.zeros(a.shape[0], dtype=dt)
    for k in "xyz":
        s[k] = a[k]
    if tname:
        s["t"] = a[tname]
    return s


@pytest.mark.parametrize("lidar_type", [1, 2, 3, 4, 5])
def test_pointcloud2_handlers(capi_host, lidar_type):
    rng = np.random.default_rng(100 + lidar_type)
    dt, tname, dtype_code = LAYOUTS[lidar_type]
    stamp = 1700000000.25
    for n, pfn, blind2 in ((5000, 1, 0.01), (5000, 3, 0.25), (1, 1, 0.01)):
        a = _cloud(lidar_type, n, rng, stamp)
        off_t = dt.fields[tname][1] if tname else -1
        g = capi_host.decode_pointcloud2(lidar_type, a.tobytes(), n, dt.itemsize, [0, 4, 8], off_t, dtype_code, stamp,
                                         blind2, pfn)
        o = op.decode_handler(lidar_type, _as_struct(a, tname), stamp, blind2, pfn)
        assert g.shape == o.shape and np.array_equal(g, o), (lidar_type, n, pfn, g.shape, o.shape)
        if n > 1 and lidar_type != 5:
            assert 0 < g.shape[0] < n and g[:, 3].min() >= 0 and g[:, 3].max() < 0.11
    # big-endian message: same values after the byte swap (decode_field, lidar_pointcloud_decoder.cpp:4-19)
    a = _cloud(lidar_type, 300, rng, stamp)
    be = a.astype(a.dtype.newbyteorder(">"))
    off_t = dt.fields[tname][1] if tname else -1
    g1 = capi_host.decode_pointcloud2(lidar_type, a.tobytes(), 300, dt.itemsize, [0, 4, 8], off_t, dtype_code, stamp, 0.01, 1)
    g2 = capi_host.decode_pointcloud2(lidar_type, be.tobytes(), 300, dt.itemsize, [0, 4, 8], off_t, dtype_code, stamp, 0.01,
                                      1, is_bigendian=True)
    assert np.array_equal(g1, g2)
    assert capi_host.decode_pointcloud2(lidar_type, b"", 0, dt.itemsize, [0, 4, 8], off_t, dtype_code, stamp, 0.01, 1).shape == (0, 4)


def test_robosense_blind_test_is_planar(capi_host):
    """lidar_pointcloud_decoder.cpp:217: a return right above the sensor (x, y ~ 0, z large) is dropped by the RoboSense
    handler and kept by the others."""
    dt, tname, code = LAYOUTS[4]
    a = np.zeros(2, dtype=dt)
    a["x"], a["y"], a["z"], a["timestamp"] = [0.01, 5.0], [0.01, 0.0], [9.0, 0.0], [10.01, 10.02]
    g = capi_host.decode_pointcloud2(4, a.tobytes(), 2, dt.itemsize, [0, 4, 8], dt.fields[tname][1], code, 10.0, 0.01, 1)
    assert g.shape[0] == 1 and g[0, 0] == 5.0
    dth, tn, ch = LAYOUTS[3]
    b = np.zeros(2, dtype=dth)
    for k in "xyz":
        b[k] = a[k]
    b["timestamp"] = a["timestamp"]
    assert capi_host.decode_pointcloud2(3, b.tobytes(), 2, dth.itemsize, [0, 4, 8], dth.fields[tn][1], ch, 10.0, 0.01, 1).shape[0] == 2


def test_velodyne_without_stamps_uses_the_azimuth(capi_host):
    """velodyne_handler's fallback (lidar_pointcloud_decoder.cpp:99-139): the last point's `time` is 0, so the stamps
    come from the azimuth at omega_l deg/s - including the wrap of atan2 at +-180 deg."""
    rng = np.random.default_rng(9)
    dt, tname, code = LAYOUTS[1]
    a = _cloud(1, 4000, rng, 0.0, spin=True)
    for pfn in (1, 2):
        g = capi_host.decode_pointcloud2(1, a.tobytes(), 4000, dt.itemsize, [0, 4, 8], dt.fields[tname][1], code, 0.0, 0.01, pfn)
        o = op.decode_handler(1, _as_struct(a, tname), 0.0, 0.01, pfn)
        assert g.shape == o.shape and g.shape[0] > 1000
        assert np.array_equal(g[:, :3], o[:, :3])
        assert np.max(np.abs(g[:, 3] - o[:, 3])) < 2e-8
        assert np.all(np.diff(g[:, 3]) > -1e-6) and g[-1, 3] > 0.09  # one revolution = 360 / 3610 s


def test_livox_and_error_codes(capi_host):
    rng = np.random.default_rng(4)
    n = 3000
    p = np.zeros(n, dtype=capi_host.LIVOX_POINT_DTYPE)
    for k in "xyz":
        p[k] = rng.uniform(-20, 20, n)
    p["x"][::5] *= 0.001
    p["y"][::5] *= 0.001
    p["z"][::5] *= 0.001
    p["offset_time"] = np.sort(rng.integers(0, 100_000_000, n)).astype(np.uint32)
    s = np.zeros(n, dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("t", "<u4")])
    for k in "xyz":
        s[k] = p[k]
    s["t"] = p["offset_time"]
    for pfn in (1, 3):
        g, o = capi_host.decode_livox(p, 0.01, pfn), op.decode_handler(0, s, 0.0, 0.01, pfn)
        assert g.shape == o.shape and 0 < g.shape[0] < n and np.array_equal(g, o)
    with pytest.raises(capi_host.VinaError) as e:
        capi_host.decode_livox(p, 0.01, 1, cap=10)
    assert e.value.code == -3
    dt, tname, code = LAYOUTS[3]
    a = _cloud(3, 100, rng, 5.0)
    with pytest.raises(capi_host.VinaError) as e:  # "Unsupported lidar type" (lidar_pointcloud_decoder.cpp:49-51)
        capi_host.decode_pointcloud2(9, a.tobytes(), 100, dt.itemsize, [0, 4, 8], 16, code, 5.0, 0.01, 1)
    assert e.value.code == -1
    with pytest.raises(capi_host.VinaError) as e:
        capi_host.decode_pointcloud2(3, a.tobytes(), 100, dt.itemsize, [0, 4, 8], 16, code, 5.0, 0.01, 1, cap=5)
    assert e.value.code == -3


def test_library_against_the_reference_handlers(capi_host):
    """The library's unpacking against the REFERENCE'S OWN handlers (lidar_pointcloud_decoder.cpp compiled unmodified
    into oracle/_ref): bit for bit for every sensor type, the azimuth-derived Velodyne stamps included (both sides call
    libm's atan2f)."""
    if not op.have_ref():
        pytest.skip("oracle/_ref is not built here (needs /root/reference)")
    stamp = float(1000250000000) * 1e-9
    for lidar_type in (1, 2, 3, 4, 5):
        rng = np.random.default_rng(70 + lidar_type)
        dt, tname, code = LAYOUTS[lidar_type]
        off_t = dt.fields[tname][1] if tname else -1
        for n, pfn, blind2, spin in ((4000, 1, 0.01, False), (4000, 3, 0.25, False), (3000, 1, 0.01, True), (3000, 2, 0.5, True)):
            if spin and lidar_type != 1:
                continue
            a = _cloud(lidar_type, n, rng, stamp, spin=spin)
            g = capi_host.decode_pointcloud2(lidar_type, a.tobytes(), n, dt.itemsize, [0, 4, 8], off_t, code, stamp, blind2, pfn)
            r = op.decode_handler_ref(lidar_type, a.tobytes(), n, dt.itemsize, [0, 4, 8], off_t, code, stamp, blind2, pfn)
            assert g.shape == r.shape and g.shape[0] > 500 and np.array_equal(g, r), (lidar_type, n, pfn, spin)


def test_scan_last_stamp_is_what_the_prepared_scan_ends_with(capi_host):
    rng = np.random.default_rng(2)
    a = np.zeros((5000, 4), dtype=np.float32)
    a[:, :3] = rng.uniform(1, 30, (5000, 3))
    a[:, 3] = rng.uniform(0, 0.13, 5000)
    assert capi_host.scan_last_stamp(a) == op.scan_prepare(a, 1, -1.0)[-1, 3]
    assert capi_host.scan_last_stamp(np.zeros((0, 4), dtype=np.float32)) == np.float32(0.09)
    a[:, 3] += np.float32(0.2)
    with pytest.raises(capi_host.VinaError) as e:
        capi_host.scan_last_stamp(a)
    assert e.value.code == -1

"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header declares,
its POD layouts match the Python mirror, and it fails loudly without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vina_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vina_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(gpu_lib):
    lib = gpu_lib.load()
    names = _declared_functions()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/vina_b200.h but not exported"
    assert sorted(gpu_lib.EXPORTS) == names


def test_pod_layouts_match_header(gpu_lib):
    assert C.sizeof(gpu_lib.VinaState) == 8 * (1 + 9 + 15 + 225)
    assert gpu_lib.IMU_POSE_DTYPE.itemsize == 22 * 8 and gpu_lib.IMU_DTYPE.itemsize == 7 * 8
    assert gpu_lib.POSE_DTYPE.itemsize == 12 * 8
    assert gpu_lib.NODE_DTYPE.itemsize == 1448
    from oracle import oracle_py

    assert oracle_py.NODE_DTYPE == gpu_lib.NODE_DTYPE  # same record on both sides of the parity tests
    c = gpu_lib.make_config(__import__("vina_slam_b200.synth", fromlist=["x"]).SENSORS["robosense128"])
    assert c.win_size == 10 and c.max_layer == 2 and c.thread_num == 5 and c.max_points == 100
    assert abs(c.ext_R[1] - (-1.0)) < 1e-15  # column-major: R(1,0) of the row-major yaml matrix


def test_no_cpu_fallback(gpu_lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from vina_slam_b200 import synth

    with pytest.raises(gpu_lib.VinaError) as ei:
        gpu_lib.Ctx(synth.small_sensor("robosense128", 8, 100))
    assert ei.value.code == -2  # VINA_E_CUDA


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under vina_slam_b200/ may reference it."""
    pkg = os.path.join(ROOT, "vina_slam_b200")
    for dp, _, files in os.walk(pkg):
        if "_build" in dp:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle_py" not in txt and "liboracle" not in txt and "vina_oracle" not in txt, f

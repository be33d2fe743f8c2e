"""CPU tests of the host side: the C ABI library loads and exports every symbol include/uvs.h
declares, fails loudly without a GPU (no CPU fallback), and the window container round-trips."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import uvs_b200
from uvs_b200 import Window
from tools import gen_window as gw

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "uvs.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(uvs_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = uvs_b200.load_library()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert lib.uvs_abi_version() == 2   # version 2: relocalisation fields of UvsWindow
    assert set(uvs_b200.binding.EXPORTS) == set(names)


def test_struct_layouts_match_the_header():
    lib = uvs_b200.load_library()
    o = uvs_b200.UvsOptionsStruct()
    lib.uvs_default_options(C.byref(o))
    d = uvs_b200.default_options()
    for f, _ in uvs_b200.UvsOptionsStruct._fields_:
        a, b = getattr(o, f), getattr(d, f)
        if f == "gravity":
            assert list(a) == list(b)
        else:
            assert a == b, f
    assert C.sizeof(uvs_b200.UvsSummaryStruct) == 16 + 16 + 5 * 8 * 64 + 4 * 64
    assert C.sizeof(uvs_b200.UvsWindowStruct) == 12 * 4 + 40 * 8 + 2 * 4 + 3 * 8   # + the relocalisation fields (ABI version 2)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = uvs_b200.load_library()
    h = C.c_void_p()
    rc = lib.uvs_create(0, C.byref(h))
    assert rc == -2 and not h.value          # UVS_ERR_CUDA
    with pytest.raises(uvs_b200.UvsError):
        uvs_b200.Solver(0)
    assert b"no CPU fallback" in lib.uvs_status_string(rc)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "uv-slam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "liborc" not in txt and "oracle/" not in txt and "tests.orc" not in txt and "from tests" not in txt, f


def test_window_roundtrip_and_sizes():
    w = gw.make_window("tiny")
    w2 = Window.from_bytes(w.to_bytes())
    for n in ("pose", "speed_bias", "inv_depth", "ortho", "proj_pts_i", "line_sp", "vp_dir", "imu_covariance", "prior_J", "prior_x0"):
        assert np.array_equal(getattr(w, n), getattr(w2, n)), n
    assert w.cam_dim == 15 * w.n_frames and w.tangent_dim == w.cam_dim + w.n_points + 4 * w.n_lines
    jac, res = w.sweep_bytes()
    n = w.prior_n
    assert jac == 384 * w.n_proj + 232 * w.n_line_obs + 120 * w.n_vp_obs + 6024 * w.n_imu + 8 * (n * n + 3 * n) + 8 * (16 * w.n_frames + 8 + w.n_points + 4 * w.n_lines)
    s = w.as_struct()
    assert s.n_proj == w.n_proj and s.prior_n == n


def test_generator_contract():
    """eligibility rules of estimator.cpp:826,873 and the grouping the library requires"""
    w = gw.make_window("C1")
    F = w.n_frames
    assert np.all(np.diff(w.proj_point) >= 0) and np.all(np.diff(w.line_idx) >= 0)
    for k in range(w.n_points):
        idx = np.nonzero(w.proj_point == k)[0]
        assert len(idx) >= 1 and len(set(w.proj_frame_i[idx])) == 1      # used_num >= 2, one anchor
        assert w.proj_frame_i[idx[0]] < F - 3 + 1
    for k in range(w.n_lines):
        assert np.count_nonzero(w.line_idx == k) >= gw.LINE_WINDOW
    pairs = set(zip(w.line_frame.tolist(), w.line_idx.tolist()))
    assert all((f, l) in pairs for f, l in zip(w.vp_frame.tolist(), w.vp_line.tolist()))
    assert np.all(w.vp_dir[:, 2] == 1.0)                                  # vp(2) == 1 (estimator.cpp:920)

"""The C++ writer of the `uvs_window v1` format (uv-slam_b200/host/window_io.cpp, the dump hook of SURVEY.md §8f row 4)
against the Python container: byte-identical files, loss-free round trip.  CPU only - no compute call."""
import ctypes as C
import os

import numpy as np
import pytest

import uvs_b200
from uvs_b200.window import _F64, _I32, UvsWindowStruct, Window
from tools import gen_window as gw

HOST_LIB = os.path.join(os.path.dirname(os.path.abspath(uvs_b200.__file__)), "host", "libuvs_host.so")


@pytest.fixture(scope="module")
def host():
    if not os.path.exists(HOST_LIB):
        pytest.skip("libuvs_host.so not built (python -c 'import __graft_entry__ as g; g.build()')")
    lib = C.CDLL(HOST_LIB)
    lib.uvs_host_save_window.argtypes = [C.POINTER(UvsWindowStruct), C.c_char_p]
    lib.uvs_host_save_window.restype = C.c_int
    return lib


@pytest.mark.parametrize("cfg", ["tiny", "C1", "C2"])
def test_cpp_writer_is_byte_identical_to_the_python_container(host, tmp_path, cfg):
    w = gw.make_window(cfg)
    path = str(tmp_path / ("%s.uvsw" % cfg))
    s = w.as_struct()
    assert host.uvs_host_save_window(C.byref(s), path.encode()) == 0
    blob = open(path, "rb").read()
    assert blob == w.to_bytes()
    back = Window.load(path)
    for n in _F64 + _I32:
        a, b = getattr(w, n), getattr(back, n)
        assert a.shape == b.shape and np.array_equal(a, b), n
    assert (back.estimate_extrinsic, back.estimate_td) == (w.estimate_extrinsic, w.estimate_td)


def test_cpp_writer_handles_missing_prior_and_bad_arguments(host, tmp_path):
    w = gw.make_window("tiny")
    w.set_prior(np.zeros((0, 0)), np.zeros(0), np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0))
    path = str(tmp_path / "noprior.uvsw")
    s = w.as_struct()
    assert host.uvs_host_save_window(C.byref(s), path.encode()) == 0
    back = Window.load(path)
    assert back.prior_n == 0 and back.n_proj == w.n_proj
    assert host.uvs_host_save_window(None, path.encode()) == -1                      # UVS_ERR_INVALID_ARG
    assert host.uvs_host_save_window(C.byref(s), None) == -1
    assert host.uvs_host_save_window(C.byref(s), str(tmp_path / "no" / "such" / "dir.uvsw").encode()) != 0

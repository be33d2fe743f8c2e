"""ctypes binding of libuvs_b200.so — the same C ABI a cgo/JNI/C++ caller would bind
(include/uvs.h).  There is NO CPU fallback: if the CUDA library is missing or no GPU is present,
every entry point raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .window import (EVAL_CERES_LAYOUT, EVAL_DEVICE_OUT, EVAL_LOCAL_LAYOUT, UvsOptionsStruct, UvsPriorStruct,
                     UvsSummaryStruct, UvsWindowStruct, Window, c_double_p, c_int32_p, default_options,
                     window_array)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = "libuvs_b200.so"
_lib = None

EXPORTS = (
    "uvs_abi_version", "uvs_default_options", "uvs_status_string", "uvs_create", "uvs_destroy", "uvs_last_error",
    "uvs_upload_windows", "uvs_download_state", "uvs_eval_proj", "uvs_eval_line", "uvs_eval_vp", "uvs_eval_imu",
    "uvs_eval_prior", "uvs_eval_cost", "uvs_solve", "uvs_batch_solve", "uvs_marginalize", "uvs_sweep_bytes",
    "uvs_launch_count", "uvs_last_solve_ms", "uvs_last_sweep_ms", "uvs_comm_init", "uvs_reset_state",
    "uvs_set_profiling", "uvs_last_stage_ms", "uvs_preintegrate", "uvs_batch_solve_pipelined",
    "uvs_triangulate_points", "uvs_triangulate_lines", "uvs_validate_lines", "uvs_set_graph_replay", "uvs_upload_state", "uvs_jacobian_sweep", "uvs_comm_unique_id", "uvs_comm_init_nccl", "uvs_collective_count",
    "uvs_window_create", "uvs_window_push_frame", "uvs_window_counts", "uvs_window_upload", "uvs_window_marginalize", "uvs_window_slide",
    "uvs_window_remove_tracks", "uvs_download_factors", "uvs_h2d_bytes",
)

N_STAGES = 10
STAGE_NAMES = ("sweep_proj", "sweep_line", "sweep_vp", "sweep_imu", "sweep_prior", "build", "chol", "backsub",
               "resid_sweep", "step")
ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p)


class UvsImuRecordStruct(C.Structure):
    _fields_ = [(n, c_double_p) for n in ("delta_p", "delta_q", "delta_v", "sum_dt", "lin_ba", "lin_bg", "jacobian", "covariance")]


class UvsFrameInputStruct(C.Structure):
    _fields_ = [("imu", C.POINTER(UvsImuRecordStruct)), ("n_points", C.c_int32), ("n_lines", C.c_int32), ("point_id", c_int32_p),
                ("point_xyz", c_double_p), ("line_id", c_int32_p), ("line_sp", c_double_p), ("line_ep", c_double_p), ("line_vp", c_double_p)]


def _imu_record(rec):
    """dict with the keys of tools/gen_window.preintegrate (+ lin_ba / lin_bg) -> (struct, keep-alive arrays)"""
    keep = {k: np.ascontiguousarray(np.atleast_1d(rec[k]), dtype=np.float64) for k in
            ("delta_p", "delta_q", "delta_v", "sum_dt", "lin_ba", "lin_bg", "jacobian", "covariance")}
    st = UvsImuRecordStruct()
    for k, a in keep.items():
        setattr(st, k, a.ctypes.data_as(c_double_p))
    return st, keep


class UvsError(RuntimeError):
    def __init__(self, status, where, detail=""):
        self.status = status
        super().__init__("%s failed with status %d%s" % (where, status, (": " + detail) if detail else ""))


def library_path() -> str:
    # UVS_LIB: developer override (e.g. a -DUVS_CHOL_TIMING build); the product library lives next to this package
    return os.environ.get("UVS_LIB") or os.path.join(_HERE, "csrc", _LIB_NAME)


def load_library():
    """Load the CUDA library; raises (no fallback) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise ImportError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)" % path)
    lib = C.CDLL(path)
    H = C.c_void_p
    lib.uvs_abi_version.restype = C.c_int
    lib.uvs_default_options.argtypes = [C.POINTER(UvsOptionsStruct)]
    lib.uvs_status_string.restype = C.c_char_p
    lib.uvs_status_string.argtypes = [C.c_int]
    lib.uvs_create.argtypes = [C.c_int, C.POINTER(H)]
    lib.uvs_destroy.argtypes = [H]
    lib.uvs_last_error.restype = C.c_char_p
    lib.uvs_last_error.argtypes = [H]
    lib.uvs_upload_windows.argtypes = [H, C.c_int32, C.POINTER(UvsWindowStruct), C.POINTER(UvsOptionsStruct)]
    lib.uvs_download_state.argtypes = [H, C.c_int32, C.POINTER(UvsWindowStruct)]
    lib.uvs_upload_state.argtypes = [H, C.c_int32, C.POINTER(UvsWindowStruct)]
    lib.uvs_jacobian_sweep.argtypes = [H, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_float * 4)]
    for n in ("uvs_eval_proj", "uvs_eval_line", "uvs_eval_vp", "uvs_eval_imu", "uvs_eval_prior"):
        getattr(lib, n).argtypes = [H, C.c_void_p, C.c_void_p, C.c_int32]
    lib.uvs_eval_cost.argtypes = [H, c_double_p]
    lib.uvs_solve.argtypes = [H, C.POINTER(UvsSummaryStruct)]
    lib.uvs_batch_solve.argtypes = [H, C.c_int32, C.POINTER(UvsWindowStruct), C.POINTER(UvsOptionsStruct),
                                    C.POINTER(UvsSummaryStruct)]
    lib.uvs_batch_solve_pipelined.argtypes = [H, C.c_int32, C.POINTER(UvsWindowStruct), C.POINTER(UvsOptionsStruct),
                                              C.POINTER(UvsSummaryStruct), C.c_int32]
    lib.uvs_marginalize.argtypes = [H, C.c_int32, C.c_int32, C.POINTER(UvsPriorStruct)]
    lib.uvs_sweep_bytes.argtypes = [H, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.uvs_launch_count.restype = C.c_int64
    lib.uvs_launch_count.argtypes = [H]
    lib.uvs_last_solve_ms.argtypes = [H, C.POINTER(C.c_float)]
    lib.uvs_last_sweep_ms.argtypes = [H, C.POINTER(C.c_float), C.POINTER(C.c_int32)]
    lib.uvs_comm_init.argtypes = [H, C.c_int32, C.c_int32, ALLREDUCE_FN, C.c_void_p]
    lib.uvs_reset_state.argtypes = [H]
    lib.uvs_comm_unique_id.argtypes = [C.c_char_p]
    lib.uvs_comm_init_nccl.argtypes = [H, C.c_char_p, C.c_int32, C.c_int32]
    lib.uvs_collective_count.restype = C.c_int64
    lib.uvs_collective_count.argtypes = [H]
    lib.uvs_preintegrate.argtypes = [H, C.c_int32, c_int32_p] + [c_double_p] * 14
    lib.uvs_triangulate_points.argtypes = [H, C.c_int32] + [c_double_p] * 4 + [C.c_int32, c_int32_p, c_int32_p, c_double_p, C.c_double, c_double_p]
    lib.uvs_triangulate_lines.argtypes = [H, C.c_int32] + [c_double_p] * 4 + [C.c_int32, c_int32_p, c_int32_p] + [c_double_p] * 5
    lib.uvs_validate_lines.argtypes = [H, C.c_int32] + [c_double_p] * 4 + [C.c_int32, c_int32_p] + [c_double_p] * 3 + [c_int32_p, c_double_p]
    lib.uvs_set_profiling.argtypes = [H, C.c_int32]
    lib.uvs_set_graph_replay.argtypes = [H, C.c_int32]
    lib.uvs_last_stage_ms.argtypes = [H, C.POINTER(C.c_float * N_STAGES), C.POINTER(C.c_int32)]
    lib.uvs_window_create.argtypes = [H, C.c_int32, C.c_int32, C.c_int32, C.c_int32]
    lib.uvs_window_push_frame.argtypes = [H, C.POINTER(UvsFrameInputStruct)]
    lib.uvs_window_counts.argtypes = [H, C.POINTER(C.c_int32 * 8)]
    lib.uvs_window_upload.argtypes = [H, C.POINTER(UvsWindowStruct), C.POINTER(UvsOptionsStruct)]
    lib.uvs_window_marginalize.argtypes = [H, C.c_int32, C.POINTER(UvsPriorStruct)]
    lib.uvs_window_slide.argtypes = [H, C.c_int32, C.POINTER(UvsImuRecordStruct)]
    lib.uvs_window_remove_tracks.argtypes = [H, C.c_int32, c_int32_p, C.c_int32, c_int32_p]
    lib.uvs_download_factors.argtypes = [H, C.c_int32, C.POINTER(UvsWindowStruct)]
    lib.uvs_h2d_bytes.restype = C.c_int64
    lib.uvs_h2d_bytes.argtypes = [H]
    _lib = lib
    return lib


_JDIM = {  # doubles per factor: (nres, ceres-layout jac, local-layout jac)
    "proj": (2, 44, 38), "line": (2, 22, 20), "vp": (1, 11, 10), "imu": (15, 480, 450),
}


class Solver:
    """Thin object wrapper over one UvsHandle (one CUDA stream, not re-entrant)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        self.h = C.c_void_p()
        rc = self.lib.uvs_create(device, C.byref(self.h))
        if rc != 0:
            raise UvsError(rc, "uvs_create", self.lib.uvs_status_string(rc).decode())
        self.windows = []
        self._arr = None
        self._cb = None

    def close(self):
        if self.h:
            self.lib.uvs_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, where):
        if rc != 0:
            raise UvsError(rc, where, self.lib.uvs_last_error(self.h).decode())

    # -- data movement -------------------------------------------------------------------------
    def upload(self, windows, opts: UvsOptionsStruct | None = None):
        if isinstance(windows, Window):
            windows = [windows]
        self.windows = list(windows)
        self.opts = opts if opts is not None else default_options()
        self._arr = window_array(self.windows)
        self._check(self.lib.uvs_upload_windows(self.h, len(self.windows), self._arr, C.byref(self.opts)), "uvs_upload_windows")

    def download(self):
        """Writes the device state back into the Window objects given to upload()."""
        self._check(self.lib.uvs_download_state(self.h, len(self.windows), self._arr), "uvs_download_state")
        return self.windows

    def upload_state(self, windows=None):
        """Replaces the device state of the uploaded batch by the state arrays of `windows` (default: the Window objects
        given to upload(), e.g. after the caller changed them in place); factors stay as uploaded."""
        arr = self._arr if windows is None else window_array(list(windows))
        n = len(self.windows)
        self._check(self.lib.uvs_upload_state(self.h, n, arr), "uvs_upload_state")

    def jacobian_sweep(self, repeats=10, each=True):
        """-> (ms per materialised Jacobian sweep with the four kernels side by side, [ms of proj, line+vp, imu, prior alone])"""
        g, e = C.c_float(), (C.c_float * 4)()
        self._check(self.lib.uvs_jacobian_sweep(self.h, int(repeats), C.byref(g), C.byref(e) if each else None), "uvs_jacobian_sweep")
        return g.value, [e[k] for k in range(4)]

    # -- factor sweeps -------------------------------------------------------------------------
    def _count(self, kind):
        # relocalisation factors are point factors inside the library (each one follows its point's own factors)
        return sum({"proj": w.n_proj + w.n_relo, "line": w.n_line_obs, "vp": w.n_vp_obs, "imu": w.n_imu}[kind] for w in self.windows)

    def eval(self, kind: str, local: bool = False, want_jac: bool = True):
        """-> (residuals [n, nr], jacobians [n, jd] or None) on the host."""
        if kind == "prior":
            return self.eval_prior(local, want_jac)
        nr, jc, jl = _JDIM[kind]
        if kind == "proj" and self.windows and self.windows[0].estimate_td:
            jc, jl = jc + 2, jl + 2
        n = self._count(kind)
        r = np.zeros((n, nr))
        J = np.zeros((n, jl if local else jc)) if want_jac else None
        fn = getattr(self.lib, "uvs_eval_" + kind)
        flags = EVAL_LOCAL_LAYOUT if local else EVAL_CERES_LAYOUT
        self._check(fn(self.h, r.ctypes.data, J.ctypes.data if want_jac else None, flags), "uvs_eval_" + kind)
        return r, J

    def eval_prior(self, local=False, want_jac=True):
        """Per-window lists: residual [n], jacobian [n, cols]."""
        tot_r, tot_j, shapes = 0, 0, []
        for w in self.windows:
            n = w.prior_n
            cols = sum((6 if local else 7) if k in (0, 2) else (9 if k == 1 else 1) for k in w.prior_block_kind)
            shapes.append((n, cols))
            tot_r += n
            tot_j += n * cols
        r = np.zeros(max(tot_r, 1)); J = np.zeros(max(tot_j, 1))
        flags = EVAL_LOCAL_LAYOUT if local else EVAL_CERES_LAYOUT
        self._check(self.lib.uvs_eval_prior(self.h, r.ctypes.data, J.ctypes.data if want_jac else None, flags), "uvs_eval_prior")
        rs, Js, ro, jo = [], [], 0, 0
        for n, cols in shapes:
            rs.append(r[ro:ro + n].copy()); ro += n
            Js.append(J[jo:jo + n * cols].copy()); jo += n * cols
        return rs, (Js if want_jac else None)

    def cost(self):
        c = np.zeros(len(self.windows))
        self._check(self.lib.uvs_eval_cost(self.h, c.ctypes.data_as(c_double_p)), "uvs_eval_cost")
        return c

    # -- solve ---------------------------------------------------------------------------------
    def solve(self):
        sums = (UvsSummaryStruct * len(self.windows))()
        self._check(self.lib.uvs_solve(self.h, sums), "uvs_solve")
        return sums

    def batch_solve(self, windows, opts=None, prepared=None, groups=None):
        """upload + solve + download through the single reference-facing call (host buffers).
        `prepared` = window_array(windows) built beforehand (the ctypes view of the same host arrays)."""
        if isinstance(windows, Window):
            windows = [windows]
        self.windows = list(windows)
        self.opts = opts if opts is not None else default_options()
        self._arr = prepared if prepared is not None else window_array(self.windows)
        sums = (UvsSummaryStruct * len(self.windows))()
        if groups is None:
            self._check(self.lib.uvs_batch_solve(self.h, len(self.windows), self._arr, C.byref(self.opts), sums), "uvs_batch_solve")
        else:   # pipelined over sub-batches (one-shot: the handle keeps no batch afterwards)
            self._check(self.lib.uvs_batch_solve_pipelined(self.h, len(self.windows), self._arr, C.byref(self.opts), sums, int(groups)),
                        "uvs_batch_solve_pipelined")
        return sums

    # -- device-resident sliding window (include/uvs.h, uvs_window_*) --------------------------------
    def window_create(self, window_size=10, line_window=5, max_points=1024, max_lines=512):
        self._check(self.lib.uvs_window_create(self.h, window_size, line_window, max_points, max_lines), "uvs_window_create")

    def window_push_frame(self, point_id, point_xyz, line_id, line_sp, line_ep, line_vp, imu=None):
        f64 = lambda a, sh: np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(sh))
        pid, lid = np.ascontiguousarray(point_id, dtype=np.int32), np.ascontiguousarray(line_id, dtype=np.int32)
        xyz, sp, ep, vp = f64(point_xyz, (-1, 3)), f64(line_sp, (-1, 2)), f64(line_ep, (-1, 2)), f64(line_vp, (-1, 3))
        fr = UvsFrameInputStruct()
        keep = None
        if imu is not None:
            st, keep = _imu_record(imu)
            fr.imu = C.pointer(st)
        fr.n_points, fr.n_lines = len(pid), len(lid)
        fr.point_id, fr.point_xyz = pid.ctypes.data_as(c_int32_p), xyz.ctypes.data_as(c_double_p)
        fr.line_id, fr.line_sp, fr.line_ep, fr.line_vp = (lid.ctypes.data_as(c_int32_p), sp.ctypes.data_as(c_double_p),
                                                          ep.ctypes.data_as(c_double_p), vp.ctypes.data_as(c_double_p))
        self._check(self.lib.uvs_window_push_frame(self.h, C.byref(fr)), "uvs_window_push_frame")
        del keep

    def window_counts(self):
        c = (C.c_int32 * 8)()
        self._check(self.lib.uvs_window_counts(self.h, C.byref(c)), "uvs_window_counts")
        return dict(zip(("n_frames", "n_points", "n_lines", "n_proj", "n_line_obs", "n_vp_obs", "n_imu", "prior_n"), list(c)))

    def window_upload(self, state: Window, opts: UvsOptionsStruct | None = None):
        """assemble the resident window on the device around the caller's state; `state` then plays the role of the uploaded
        window for solve() / download() / upload_state()"""
        self.windows = [state]
        self.opts = opts if opts is not None else default_options()
        self._arr = window_array(self.windows)
        self._check(self.lib.uvs_window_upload(self.h, self._arr, C.byref(self.opts)), "uvs_window_upload")

    def window_marginalize(self, flag=0):
        cap_n, cap_b = 16 * 32 + 16, 2 * 32 + 8
        J = np.zeros((cap_n, cap_n)); r = np.zeros(cap_n); A = np.zeros((cap_n, cap_n)); b = np.zeros(cap_n)
        kind = np.zeros(cap_b, np.int32); bid = np.zeros(cap_b, np.int32); x0 = np.zeros(9 * cap_b)
        p = UvsPriorStruct()
        p.J, p.r, p.A, p.b, p.x0 = (a.ctypes.data_as(c_double_p) for a in (J, r, A, b, x0))
        p.block_kind, p.block_id = kind.ctypes.data_as(c_int32_p), bid.ctypes.data_as(c_int32_p)
        p.cap_n, p.cap_blocks = cap_n, cap_b
        self._check(self.lib.uvs_window_marginalize(self.h, flag, C.byref(p)), "uvs_window_marginalize")
        n, nb = p.n, p.n_blocks
        if n == 0:
            return None
        kinds = kind[:nb].copy(); ids = bid[:nb].copy()
        gs = np.array([7 if k in (0, 2) else (9 if k == 1 else 1) for k in kinds])
        return dict(n=n, m=p.m, J=J.ravel()[:n * n].reshape(n, n).copy(), r=r[:n].copy(), A=A.ravel()[:n * n].reshape(n, n).copy(),
                    b=b[:n].copy(), kinds=kinds, ids=ids, x0=x0[:gs.sum()].copy())

    def window_slide(self, flag=0, merged_imu=None):
        if merged_imu is None:
            self._check(self.lib.uvs_window_slide(self.h, flag, None), "uvs_window_slide")
        else:
            st, keep = _imu_record(merged_imu)
            self._check(self.lib.uvs_window_slide(self.h, flag, C.byref(st)), "uvs_window_slide")
            del keep

    def window_remove_tracks(self, point_ids=(), line_ids=()):
        p, l = np.ascontiguousarray(point_ids, dtype=np.int32), np.ascontiguousarray(line_ids, dtype=np.int32)
        self._check(self.lib.uvs_window_remove_tracks(self.h, len(p), p.ctypes.data_as(c_int32_p), len(l), l.ctypes.data_as(c_int32_p)),
                    "uvs_window_remove_tracks")

    def download_factors(self, counts, window_index=0) -> Window:
        """the factor arrays of an uploaded window as the device holds them -> a Window sized by `counts` (window_counts() or
        a host-packed Window's sizes); state arrays are zeros"""
        g = lambda k: int(counts[k]) if isinstance(counts, dict) else int(getattr(counts, k))
        F, npnt, nl, npj, nlo, nvo, nim, pn = (g(k) for k in ("n_frames", "n_points", "n_lines", "n_proj", "n_line_obs", "n_vp_obs", "n_imu", "prior_n"))
        z, zi = (lambda *sh: np.zeros(sh)), (lambda n: np.zeros(n, np.int32))
        w = Window(pose=z(F, 7), speed_bias=z(F, 9), ex_pose=z(7), inv_depth=z(npnt), ortho=z(nl, 4),
                   proj_frame_i=zi(npj), proj_frame_j=zi(npj), proj_point=zi(npj), proj_pts_i=z(npj, 3), proj_pts_j=z(npj, 3),
                   line_frame=zi(nlo), line_idx=zi(nlo), line_sp=z(nlo, 2), line_ep=z(nlo, 2), vp_frame=zi(nvo), vp_line=zi(nvo), vp_dir=z(nvo, 3),
                   imu_frame_i=zi(nim), imu_delta_p=z(nim, 3), imu_delta_q=z(nim, 4), imu_delta_v=z(nim, 3), imu_sum_dt=z(nim),
                   imu_lin_ba=z(nim, 3), imu_lin_bg=z(nim, 3), imu_jacobian=z(nim, 225), imu_covariance=z(nim, 225),
                   prior_J=z(pn, pn), prior_r=z(pn))
        arr = window_array([w])
        self._check(self.lib.uvs_download_factors(self.h, window_index, arr), "uvs_download_factors")
        return w

    def h2d_bytes(self):
        return int(self.lib.uvs_h2d_bytes(self.h))

    def marginalize(self, window_index=0, flag=0):
        w = self.windows[window_index]
        cap_n, cap_b = 16 * w.n_frames + 16, 2 * w.n_frames + 8
        J = np.zeros((cap_n, cap_n)); r = np.zeros(cap_n); A = np.zeros((cap_n, cap_n)); b = np.zeros(cap_n)
        kind = np.zeros(cap_b, np.int32); bid = np.zeros(cap_b, np.int32); x0 = np.zeros(9 * cap_b)
        p = UvsPriorStruct()
        p.J, p.r, p.A, p.b, p.x0 = (a.ctypes.data_as(c_double_p) for a in (J, r, A, b, x0))
        p.block_kind, p.block_id = kind.ctypes.data_as(c_int32_p), bid.ctypes.data_as(c_int32_p)
        p.cap_n, p.cap_blocks = cap_n, cap_b
        self._check(self.lib.uvs_marginalize(self.h, window_index, flag, C.byref(p)), "uvs_marginalize")
        n, nb = p.n, p.n_blocks
        if n == 0:
            return None
        kinds = kind[:nb].copy(); ids = bid[:nb].copy()
        gs = np.array([7 if k in (0, 2) else (9 if k == 1 else 1) for k in kinds])
        return dict(n=n, m=p.m, J=J.ravel()[:n * n].reshape(n, n).copy(), r=r[:n].copy(),
                    A=A.ravel()[:n * n].reshape(n, n).copy(), b=b[:n].copy(), kinds=kinds, ids=ids,
                    x0=x0[:gs.sum()].copy())

    # -- introspection -------------------------------------------------------------------------
    def sweep_bytes(self):
        a, b = C.c_int64(), C.c_int64()
        self._check(self.lib.uvs_sweep_bytes(self.h, C.byref(a), C.byref(b)), "uvs_sweep_bytes")
        return a.value, b.value

    def launch_count(self):
        return int(self.lib.uvs_launch_count(self.h))

    def last_solve_ms(self):
        ms = C.c_float()
        self._check(self.lib.uvs_last_solve_ms(self.h, C.byref(ms)), "uvs_last_solve_ms")
        return ms.value

    def last_sweep_ms(self):
        ms, n = C.c_float(), C.c_int32()
        self._check(self.lib.uvs_last_sweep_ms(self.h, C.byref(ms), C.byref(n)), "uvs_last_sweep_ms")
        return ms.value, n.value

    def reset_state(self):
        self._check(self.lib.uvs_reset_state(self.h), "uvs_reset_state")

    def set_profiling(self, level: int):
        self._check(self.lib.uvs_set_profiling(self.h, level), "uvs_set_profiling")

    def last_stage_ms(self):
        """-> ({stage: ms accumulated over the last solve}, iterations run)"""
        ms, n = (C.c_float * N_STAGES)(), C.c_int32()
        self._check(self.lib.uvs_last_stage_ms(self.h, C.byref(ms), C.byref(n)), "uvs_last_stage_ms")
        return {k: ms[i] for i, k in enumerate(STAGE_NAMES)}, n.value

    def preintegrate(self, sample_off, dt, acc, gyr, acc0, gyr0, lin_ba, lin_bg, noise):
        """IMU mid-point preintegration of n intervals on the device -> dict of the UvsWindow imu_* arrays"""
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        off = np.ascontiguousarray(sample_off, dtype=np.int32)
        n = len(off) - 1
        dt, acc, gyr, acc0, gyr0, lin_ba, lin_bg, noise = map(f, (dt, acc, gyr, acc0, gyr0, lin_ba, lin_bg, noise))
        out = dict(delta_p=np.zeros((n, 3)), delta_q=np.zeros((n, 4)), delta_v=np.zeros((n, 3)), sum_dt=np.zeros(n),
                   jacobian=np.zeros((n, 225)), covariance=np.zeros((n, 225)))
        p = lambda a: a.ctypes.data_as(c_double_p)
        self._check(self.lib.uvs_preintegrate(self.h, n, off.ctypes.data_as(c_int32_p), p(dt), p(acc), p(gyr), p(acc0), p(gyr0), p(lin_ba),
                                              p(lin_bg), p(noise), p(out["delta_p"]), p(out["delta_q"]), p(out["delta_v"]), p(out["sum_dt"]),
                                              p(out["jacobian"]), p(out["covariance"])), "uvs_preintegrate")
        return out

    def set_graph_replay(self, enable=True):
        self._check(self.lib.uvs_set_graph_replay(self.h, 1 if enable else 0), "uvs_set_graph_replay")

    def triangulate_points(self, Rs, Ps, ric, tic, start_frame, obs_off, obs_pts, init_depth=5.0):
        """FeatureManager::triangulate on the device: depth of every track (see include/uvs.h)."""
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        Rs, Ps, ric, tic, obs_pts = f64(Rs), f64(Ps), f64(ric), f64(tic), f64(obs_pts)
        start_frame, obs_off = i32(start_frame), i32(obs_off)
        n = len(start_frame)
        out = np.zeros(n)
        p = lambda a: a.ctypes.data_as(c_double_p)
        self._check(self.lib.uvs_triangulate_points(self.h, len(Ps), p(Rs), p(Ps), p(ric), p(tic), n, start_frame.ctypes.data_as(c_int32_p),
                                                    obs_off.ctypes.data_as(c_int32_p), p(obs_pts), float(init_depth), p(out)),
                    "uvs_triangulate_points")
        return out

    def validate_lines(self, Rs, Ps, ric, tic, start_frame, ortho, sp_first, ep_first, want_end_points=False):
        """validity test of FeatureManager::setLineOrtho on the device: solve_flag [n] (1 valid, 2 behind the camera)
        and, on request, the world end points [n][6] (see include/uvs.h)."""
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        Rs, Ps, ric, tic, ortho, sp_first, ep_first = (f64(a) for a in (Rs, Ps, ric, tic, ortho, sp_first, ep_first))
        start_frame = np.ascontiguousarray(start_frame, dtype=np.int32)
        n = len(start_frame)
        flag = np.zeros(n, np.int32)
        ends = np.zeros((n, 6)) if want_end_points else None
        p = lambda a: a.ctypes.data_as(c_double_p)
        self._check(self.lib.uvs_validate_lines(self.h, len(Ps), p(Rs), p(Ps), p(ric), p(tic), n, start_frame.ctypes.data_as(c_int32_p),
                                                p(ortho), p(sp_first), p(ep_first), flag.ctypes.data_as(c_int32_p),
                                                p(ends) if want_end_points else None), "uvs_validate_lines")
        return (flag, ends) if want_end_points else flag

    def triangulate_lines(self, Rs, Ps, ric, tic, frame_first, frame_last, sp_first, ep_first, sp_last, ep_last):
        """FeatureManager::triangulateLine on the device: orthonormal parameters [n][4] (see include/uvs.h)."""
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        Rs, Ps, ric, tic = f64(Rs), f64(Ps), f64(ric), f64(tic)
        sp_first, ep_first, sp_last, ep_last = f64(sp_first), f64(ep_first), f64(sp_last), f64(ep_last)
        frame_first, frame_last = i32(frame_first), i32(frame_last)
        n = len(frame_first)
        out = np.zeros((n, 4))
        p = lambda a: a.ctypes.data_as(c_double_p)
        self._check(self.lib.uvs_triangulate_lines(self.h, len(Ps), p(Rs), p(Ps), p(ric), p(tic), n, frame_first.ctypes.data_as(c_int32_p),
                                                   frame_last.ctypes.data_as(c_int32_p), p(sp_first), p(ep_first), p(sp_last), p(ep_last),
                                                   p(out)), "uvs_triangulate_lines")
        return out

    @staticmethod
    def comm_unique_id() -> bytes:
        """ncclGetUniqueId through the library (call on one rank, hand the 128 bytes to all)"""
        lib = load_library()
        buf = C.create_string_buffer(128)
        rc = lib.uvs_comm_unique_id(buf)
        if rc != 0:
            raise UvsError(rc, "uvs_comm_unique_id", lib.uvs_status_string(rc).decode())
        return buf.raw

    def comm_init_nccl(self, unique_id: bytes, rank: int, nranks: int):
        """factor-parallel mode over the library's own NCCL communicator (collective call)"""
        assert len(unique_id) == 128
        self._check(self.lib.uvs_comm_init_nccl(self.h, unique_id, rank, nranks), "uvs_comm_init_nccl")

    def collective_count(self):
        return int(self.lib.uvs_collective_count(self.h))

    def comm_init(self, rank, nranks, reduce_fn):
        """reduce_fn(device_ptr:int, count:int, stream:int) -> int, summing in place over ranks."""
        def _tramp(user, buf, count, stream):
            try:
                return int(reduce_fn(buf, count, stream) or 0)
            except Exception:  # never let an exception cross the C ABI
                return -7
        self._cb = ALLREDUCE_FN(_tramp)
        self._check(self.lib.uvs_comm_init(self.h, rank, nranks, self._cb, None), "uvs_comm_init")

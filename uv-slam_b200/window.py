"""Host-side description of one sliding window ("uvs_window v1") and its ctypes mirror of
`include/uvs.h`.

The reference never serialises its estimator window (SURVEY.md §5 "checkpoint / resume"); this is
the flat, pointer-free description of what `Estimator::optimization()` hands to Ceres
(vins_estimator/src/estimator.cpp:761-978): parameter blocks packed as `vector2double()` does
(:526-594) plus one record per residual block.
"""
from __future__ import annotations

import ctypes as C
import io
import struct
from dataclasses import dataclass, field

import numpy as np

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)

UVS_MAX_ITER_LOG = 64
BLOCK_POSE, BLOCK_SPEEDBIAS, BLOCK_EXPOSE, BLOCK_TD = 0, 1, 2, 3
MARGIN_OLD, MARGIN_SECOND_NEW = 0, 1
EVAL_CERES_LAYOUT, EVAL_LOCAL_LAYOUT, EVAL_DEVICE_OUT = 0, 1, 2


class UvsWindowStruct(C.Structure):
    _fields_ = (
        [(n, C.c_int32) for n in (
            "n_frames", "n_points", "n_lines", "n_proj", "n_line_obs", "n_vp_obs", "n_imu", "prior_n",
            "prior_n_blocks", "estimate_extrinsic", "estimate_td", "reserved0")]
        + [(n, c_double_p) for n in ("pose", "speed_bias", "ex_pose", "td", "inv_depth", "ortho")]
        + [(n, c_int32_p) for n in ("proj_frame_i", "proj_frame_j", "proj_point")]
        + [(n, c_double_p) for n in ("proj_pts_i", "proj_pts_j", "proj_vel_i", "proj_vel_j", "proj_td_i",
                                     "proj_td_j", "proj_row_i", "proj_row_j")]
        + [(n, c_int32_p) for n in ("line_frame", "line_idx")]
        + [(n, c_double_p) for n in ("line_sp", "line_ep")]
        + [(n, c_int32_p) for n in ("vp_frame", "vp_line")]
        + [(n, c_double_p) for n in ("vp_dir", "line_ric", "line_tic")]
        + [("imu_frame_i", c_int32_p)]
        + [(n, c_double_p) for n in ("imu_delta_p", "imu_delta_q", "imu_delta_v", "imu_sum_dt", "imu_lin_ba",
                                     "imu_lin_bg", "imu_jacobian", "imu_covariance")]
        + [("prior_J", c_double_p), ("prior_r", c_double_p), ("prior_block_kind", c_int32_p),
           ("prior_block_id", c_int32_p), ("prior_x0", c_double_p)]
        + [("n_relo", C.c_int32), ("reserved1", C.c_int32), ("relo_pose", c_double_p), ("relo_point", c_int32_p),
           ("relo_pts_j", c_double_p)]
    )


class UvsOptionsStruct(C.Structure):
    _fields_ = [
        ("focal_length", C.c_double), ("gravity", C.c_double * 3), ("line_factor", C.c_double),
        ("vp_factor", C.c_double), ("cauchy_point", C.c_double), ("cauchy_line", C.c_double),
        ("cauchy_vp", C.c_double), ("tr", C.c_double), ("row", C.c_double),
        ("max_num_iterations", C.c_int32), ("fixed_iterations", C.c_int32), ("max_solver_time", C.c_double),
        ("initial_radius", C.c_double), ("max_radius", C.c_double), ("min_radius", C.c_double),
        ("min_relative_decrease", C.c_double), ("min_lm_diagonal", C.c_double), ("max_lm_diagonal", C.c_double),
        ("function_tolerance", C.c_double), ("gradient_tolerance", C.c_double), ("parameter_tolerance", C.c_double),
    ]


class UvsSummaryStruct(C.Structure):
    _fields_ = [
        ("num_iterations", C.c_int32), ("num_successful_steps", C.c_int32), ("termination", C.c_int32),
        ("status", C.c_int32), ("initial_cost", C.c_double), ("final_cost", C.c_double),
        ("cost", C.c_double * UVS_MAX_ITER_LOG), ("radius", C.c_double * UVS_MAX_ITER_LOG),
        ("relative_decrease", C.c_double * UVS_MAX_ITER_LOG), ("step_norm", C.c_double * UVS_MAX_ITER_LOG),
        ("gradient_max_norm", C.c_double * UVS_MAX_ITER_LOG), ("step_accepted", C.c_int32 * UVS_MAX_ITER_LOG),
    ]


class UvsPriorStruct(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("n_blocks", C.c_int32), ("m", C.c_int32), ("reserved0", C.c_int32),
        ("J", c_double_p), ("r", c_double_p), ("block_kind", c_int32_p), ("block_id", c_int32_p),
        ("x0", c_double_p), ("A", c_double_p), ("b", c_double_p), ("cap_n", C.c_int32), ("cap_blocks", C.c_int32),
    ]


def default_options(**kw) -> UvsOptionsStruct:
    """EuRoC values of the reference (config/euroc/euroc_config.yaml:20,55-56,64,85-87) and Ceres'
    trust-region defaults."""
    o = UvsOptionsStruct()
    o.focal_length = 461.6
    o.gravity[0], o.gravity[1], o.gravity[2] = 0.0, 0.0, 9.81007
    o.line_factor, o.vp_factor = 300.0, 10.0
    o.cauchy_point, o.cauchy_line, o.cauchy_vp = 1.0, 0.1, 1.0
    o.tr, o.row = 0.0, 480.0
    o.max_num_iterations, o.fixed_iterations, o.max_solver_time = 10, 0, 0.0
    o.initial_radius, o.max_radius, o.min_radius = 1e4, 1e16, 1e-32
    o.min_relative_decrease, o.min_lm_diagonal, o.max_lm_diagonal = 1e-3, 1e-6, 1e32
    o.function_tolerance, o.gradient_tolerance, o.parameter_tolerance = 1e-6, 1e-10, 1e-8
    for k, v in kw.items():
        if k == "gravity":
            for i in range(3):
                o.gravity[i] = v[i]
        else:
            setattr(o, k, v)
    return o


_F64 = ("pose", "speed_bias", "ex_pose", "td", "inv_depth", "ortho", "proj_pts_i", "proj_pts_j", "proj_vel_i",
        "proj_vel_j", "proj_td_i", "proj_td_j", "proj_row_i", "proj_row_j", "line_sp", "line_ep", "vp_dir",
        "line_ric", "line_tic", "imu_delta_p", "imu_delta_q", "imu_delta_v", "imu_sum_dt", "imu_lin_ba",
        "imu_lin_bg", "imu_jacobian", "imu_covariance", "prior_J", "prior_r", "prior_x0")
_I32 = ("proj_frame_i", "proj_frame_j", "proj_point", "line_frame", "line_idx", "vp_frame", "vp_line",
        "imu_frame_i", "prior_block_kind", "prior_block_id")
# optional relocalisation section (estimator.cpp:944-978), written after the arrays above only when n_relo > 0
_RELO_F64 = ("relo_pose", "relo_pts_j")
_RELO_I32 = ("relo_point",)
_MAGIC = b"UVSWIN01"


def _z(shape, dtype=np.float64):
    return np.zeros(shape, dtype=dtype)


@dataclass
class Window:
    """numpy-backed window; every array is C-contiguous float64 / int32."""
    pose: np.ndarray                      # [F,7]
    speed_bias: np.ndarray                # [F,9]
    ex_pose: np.ndarray                   # [7]
    td: np.ndarray = field(default_factory=lambda: _z(1))
    inv_depth: np.ndarray = field(default_factory=lambda: _z(0))
    ortho: np.ndarray = field(default_factory=lambda: _z((0, 4)))
    proj_frame_i: np.ndarray = field(default_factory=lambda: _z(0, np.int32))
    proj_frame_j: np.ndarray = field(default_factory=lambda: _z(0, np.int32))
    proj_point: np.ndarray = field(default_factory=lambda: _z(0, np.int32))
    proj_pts_i: np.ndarray = field(default_factory=lambda: _z((0, 3)))
    proj_pts_j: np.ndarray = field(default_factory=lambda: _z((0, 3)))
    proj_vel_i: np.ndarray = field(default_factory=lambda: _z((0, 2)))
    proj_vel_j: np.ndarray = field(default_factory=lambda: _z((0, 2)))
    proj_td_i: np.ndarray = field(default_factory=lambda: _z(0))
    proj_td_j: np.ndarray = field(default_factory=lambda: _z(0))
    proj_row_i: np.ndarray = field(default_factory=lambda: _z(0))
    proj_row_j: np.ndarray = field(default_factory=lambda: _z(0))
    line_frame: np.ndarray = field(default_factory=lambda: _z(0, np.int32))
    line_idx: np.ndarray = field(default_factory=lambda: _z(0, np.int32))
    line_sp: np.ndarray = field(default_factory=lambda: _z((0, 2)))
    line_ep: np.ndarray = field(default_factory=lambda: _z((0, 2)))
    vp_frame: np.ndarray = field(default_factory=lambda: _z(0, np.int32))
    vp_line: np.ndarray = field(default_factory=lambda: _z(0, np.int32))
    vp_dir: np.ndarray = field(default_factory=lambda: _z((0, 3)))
    line_ric: np.ndarray = field(default_factory=lambda: np.eye(3))
    line_tic: np.ndarray = field(default_factory=lambda: _z(3))
    imu_frame_i: np.ndarray = field(default_factory=lambda: _z(0, np.int32))
    imu_delta_p: np.ndarray = field(default_factory=lambda: _z((0, 3)))
    imu_delta_q: np.ndarray = field(default_factory=lambda: _z((0, 4)))
    imu_delta_v: np.ndarray = field(default_factory=lambda: _z((0, 3)))
    imu_sum_dt: np.ndarray = field(default_factory=lambda: _z(0))
    imu_lin_ba: np.ndarray = field(default_factory=lambda: _z((0, 3)))
    imu_lin_bg: np.ndarray = field(default_factory=lambda: _z((0, 3)))
    imu_jacobian: np.ndarray = field(default_factory=lambda: _z((0, 225)))
    imu_covariance: np.ndarray = field(default_factory=lambda: _z((0, 225)))
    prior_J: np.ndarray = field(default_factory=lambda: _z((0, 0)))
    prior_r: np.ndarray = field(default_factory=lambda: _z(0))
    prior_block_kind: np.ndarray = field(default_factory=lambda: _z(0, np.int32))
    prior_block_id: np.ndarray = field(default_factory=lambda: _z(0, np.int32))
    prior_x0: np.ndarray = field(default_factory=lambda: _z(0))
    relo_pose: np.ndarray = field(default_factory=lambda: _z(0))          # [7] with relocalisation factors, else empty
    relo_pts_j: np.ndarray = field(default_factory=lambda: _z((0, 3)))
    relo_point: np.ndarray = field(default_factory=lambda: _z(0, np.int32))
    estimate_extrinsic: int = 0
    estimate_td: int = 0

    def __post_init__(self):
        self.normalize()

    def normalize(self):
        for n in _F64 + _RELO_F64:
            setattr(self, n, np.ascontiguousarray(getattr(self, n), dtype=np.float64))
        for n in _I32 + _RELO_I32:
            setattr(self, n, np.ascontiguousarray(getattr(self, n), dtype=np.int32))
        return self

    # ---- sizes -------------------------------------------------------------------------------
    @property
    def n_frames(self): return int(self.pose.shape[0])
    @property
    def n_points(self): return int(self.inv_depth.shape[0])
    @property
    def n_lines(self): return int(self.ortho.shape[0])
    @property
    def n_proj(self): return int(self.proj_frame_i.shape[0])
    @property
    def n_line_obs(self): return int(self.line_frame.shape[0])
    @property
    def n_vp_obs(self): return int(self.vp_frame.shape[0])
    @property
    def n_imu(self): return int(self.imu_frame_i.shape[0])
    @property
    def prior_n(self): return int(self.prior_r.shape[0])
    @property
    def prior_n_blocks(self): return int(self.prior_block_kind.shape[0])
    @property
    def n_relo(self): return int(self.relo_point.shape[0])
    @property
    def cam_dim(self): return 15 * self.n_frames + (6 if self.estimate_extrinsic else 0) + (1 if self.estimate_td else 0)
    @property
    def tangent_dim(self): return self.cam_dim + self.n_points + 4 * self.n_lines

    def copy(self) -> "Window":
        kw = {n: getattr(self, n).copy() for n in _F64 + _I32 + _RELO_F64 + _RELO_I32}
        return Window(estimate_extrinsic=self.estimate_extrinsic, estimate_td=self.estimate_td, **kw)

    def state_vector(self) -> np.ndarray:
        return np.concatenate([self.pose.ravel(), self.speed_bias.ravel(), self.ex_pose.ravel(), self.td.ravel(),
                               self.inv_depth.ravel(), self.ortho.ravel()])

    # ---- ctypes view -------------------------------------------------------------------------
    def as_struct(self) -> UvsWindowStruct:
        """Struct of pointers into this object's arrays (keep `self` alive while it is in use)."""
        self.normalize()
        s = UvsWindowStruct()
        for n in ("n_frames", "n_points", "n_lines", "n_proj", "n_line_obs", "n_vp_obs", "n_imu", "prior_n",
                  "prior_n_blocks", "estimate_extrinsic", "estimate_td", "n_relo"):
            setattr(s, n, int(getattr(self, n)))
        if self.n_relo and self.relo_pose.size != 7:
            raise ValueError("relocalisation factors need relo_pose[7]")
        for n in _F64 + _RELO_F64:
            a = getattr(self, n)
            setattr(s, n, a.ctypes.data_as(c_double_p) if a.size else C.cast(None, c_double_p))
        for n in _I32 + _RELO_I32:
            a = getattr(self, n)
            setattr(s, n, a.ctypes.data_as(c_int32_p) if a.size else C.cast(None, c_int32_p))
        if self.estimate_td and self.proj_vel_i.shape[0] != self.n_proj:
            raise ValueError("estimate_td needs the td extras for every projection factor")
        return s

    # ---- uvs_window v1 file format -----------------------------------------------------------
    def to_bytes(self) -> bytes:
        self.normalize()
        out = io.BytesIO()
        out.write(_MAGIC)
        out.write(struct.pack("<2i", self.estimate_extrinsic, self.estimate_td))
        for n in _F64 + _I32:
            a = getattr(self, n)
            out.write(struct.pack("<i", a.ndim))
            out.write(struct.pack("<%di" % a.ndim, *a.shape))
            out.write(a.astype("<f8" if n in _F64 else "<i4").tobytes())
        if self.n_relo:
            for n in _RELO_F64 + _RELO_I32:
                a = getattr(self, n)
                out.write(struct.pack("<i", a.ndim))
                out.write(struct.pack("<%di" % a.ndim, *a.shape))
                out.write(a.astype("<f8" if n in _RELO_F64 else "<i4").tobytes())
        return out.getvalue()

    @staticmethod
    def from_bytes(buf: bytes) -> "Window":
        if buf[:8] != _MAGIC:
            raise ValueError("not a uvs_window v1 blob")
        off = 8
        ee, et = struct.unpack_from("<2i", buf, off); off += 8
        kw = {}
        for n in _F64 + _I32 + _RELO_F64 + _RELO_I32:
            if off >= len(buf):
                break           # no relocalisation section
            (nd,) = struct.unpack_from("<i", buf, off); off += 4
            shape = struct.unpack_from("<%di" % nd, buf, off); off += 4 * nd
            cnt = int(np.prod(shape)) if nd else 1
            f64 = n in _F64 or n in _RELO_F64
            kw[n] = np.frombuffer(buf, dtype="<f8" if f64 else "<i4", count=cnt, offset=off).reshape(shape).copy()
            off += cnt * (8 if f64 else 4)
        return Window(estimate_extrinsic=ee, estimate_td=et, **kw)

    def save(self, path):
        with open(path, "wb") as f:
            f.write(self.to_bytes())

    @staticmethod
    def load(path) -> "Window":
        with open(path, "rb") as f:
            return Window.from_bytes(f.read())

    def set_prior(self, J, r, kinds, ids, x0):
        self.prior_J = np.ascontiguousarray(J, dtype=np.float64)
        self.prior_r = np.ascontiguousarray(r, dtype=np.float64)
        self.prior_block_kind = np.ascontiguousarray(kinds, dtype=np.int32)
        self.prior_block_id = np.ascontiguousarray(ids, dtype=np.int32)
        self.prior_x0 = np.ascontiguousarray(x0, dtype=np.float64)

    # ---- algorithmic bytes of one sweep ---------------------------------------
    def sweep_bytes(self):
        n = self.prior_n
        state = 8 * (16 * self.n_frames + 8 + self.n_points + 4 * self.n_lines)
        jac = 384 * self.n_proj + 232 * self.n_line_obs + 120 * self.n_vp_obs + 6024 * self.n_imu + 8 * (n * n + 3 * n) + state
        res = ((64 + 16) * self.n_proj + (56 + 16) * self.n_line_obs + (32 + 8) * self.n_vp_obs
               + (2304 + 120) * self.n_imu + 8 * (n * n + 3 * n) + state)
        return jac, res


def window_array(windows):
    """ctypes array of UvsWindow structs for a list of Window objects (keeps pointers valid while
    the Window objects are alive)."""
    arr = (UvsWindowStruct * len(windows))()
    for i, w in enumerate(windows):
        arr[i] = w.as_struct()
    return arr

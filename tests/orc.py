"""ctypes binding of the CPU oracle (oracle/liborc.so).  TEST INFRASTRUCTURE: imported only by
tests/, __graft_entry__.smoke() and bench.py's CPU legs - never by the product package."""
import ctypes as C
import os
import subprocess

import numpy as np

import uvs_b200
from uvs_b200.window import (UvsOptionsStruct, UvsPriorStruct, UvsSummaryStruct, UvsWindowStruct, Window,
                             c_double_p, c_int32_p, window_array)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.environ.get("UVS_ORC_LIB") or os.path.join(ROOT, "oracle", "liborc.so")   # bench.py points this at its -march=native build
F_PRIOR, F_IMU, F_PROJ, F_LINE, F_VP = 0, 1, 2, 3, 4
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.orc_eval.argtypes = [C.POINTER(UvsWindowStruct), C.POINTER(UvsOptionsStruct), C.c_int, C.c_int,
                                  c_double_p, c_double_p, c_double_p]
        _lib.orc_factor_dims.argtypes = [C.POINTER(UvsWindowStruct), C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        _lib.orc_solve.argtypes = [C.POINTER(UvsWindowStruct), C.POINTER(UvsOptionsStruct), C.POINTER(UvsSummaryStruct), C.c_int]
        _lib.orc_solve_batch.argtypes = [C.c_int, C.POINTER(UvsWindowStruct), C.POINTER(UvsOptionsStruct),
                                         C.POINTER(UvsSummaryStruct), C.c_int]
        _lib.orc_first_step.argtypes = [C.POINTER(UvsWindowStruct), C.POINTER(UvsOptionsStruct), C.c_double] + [c_double_p] * 6
        _lib.orc_marginalize.argtypes = [C.POINTER(UvsWindowStruct), C.POINTER(UvsOptionsStruct), C.c_int, C.POINTER(UvsPriorStruct)]
        _lib.orc_total_cost.argtypes = [C.POINTER(UvsWindowStruct), C.POINTER(UvsOptionsStruct), c_double_p]
        _lib.orc_preintegrate.argtypes = [C.c_int] + [c_double_p] * 14
        _lib.orc_imu_sqrt_info.argtypes = [c_double_p, c_double_p]
        _lib.orc_pose_plus.argtypes = [c_double_p] * 3
        _lib.orc_cauchy.argtypes = [C.c_double, C.c_double, c_double_p]
        _lib.orc_sym_eig.argtypes = [C.c_int, c_double_p, c_double_p, c_double_p]
    return _lib


def _p(a):
    return a.ctypes.data_as(c_double_p)


def n_factors(w: Window, ftype):
    return {F_PRIOR: 1 if w.prior_n > 0 else 0, F_IMU: w.n_imu, F_PROJ: w.n_proj, F_LINE: w.n_line_obs, F_VP: w.n_vp_obs}[ftype]


def factor_dims(w: Window, ftype, local):
    nr, jd = C.c_int(), C.c_int()
    s = w.as_struct()
    lib().orc_factor_dims(C.byref(s), ftype, int(local), C.byref(nr), C.byref(jd))
    return nr.value, jd.value


def eval_factors(w: Window, opts, ftype, local=False, want_jac=True):
    """-> (residuals [n, nr], jacobians [n, jd] or None, cost [n])"""
    n = n_factors(w, ftype)
    nr, jd = factor_dims(w, ftype, local)
    r = np.zeros((n, nr)); J = np.zeros((n, jd)) if want_jac else None; cost = np.zeros(max(n, 1))
    if n:
        s = w.as_struct()
        rc = lib().orc_eval(C.byref(s), C.byref(opts), ftype, int(local), _p(r), _p(J) if want_jac else None, _p(cost))
        assert rc == 0
    return r, J, cost[:n]


def total_cost(w: Window, opts):
    c = C.c_double()
    s = w.as_struct()
    lib().orc_total_cost(C.byref(s), C.byref(opts), C.cast(C.byref(c), c_double_p))
    return c.value


def solve(w: Window, opts, dense_check=False):
    """Solves in place (w's state arrays are updated); returns the summary struct."""
    s = w.as_struct()
    sm = UvsSummaryStruct()
    rc = lib().orc_solve(C.byref(s), C.byref(opts), C.byref(sm), int(dense_check))
    assert rc == 0
    return sm


def solve_batch(ws, opts, n_threads=1):
    arr = window_array(ws)
    sums = (UvsSummaryStruct * len(ws))()
    rc = lib().orc_solve_batch(len(ws), arr, C.byref(opts), sums, n_threads)
    assert rc == 0
    return sums


def first_step(w: Window, opts, radius=1e4):
    T, d = w.tangent_dim, w.cam_dim
    delta = np.zeros(T); S = np.zeros((d, d)); g = np.zeros(d); scale = np.zeros(T)
    mc = np.zeros(1); cost = np.zeros(1)
    s = w.as_struct()
    rc = lib().orc_first_step(C.byref(s), C.byref(opts), radius, _p(delta), _p(S), _p(g), _p(mc), _p(cost), _p(scale))
    assert rc == 0, rc
    return dict(delta=delta, S=S, g=g, model_change=mc[0], cost=cost[0], scale=scale)


def marginalize(w: Window, opts, flag=0):
    cap_n, cap_b = 16 * w.n_frames + 16, 2 * w.n_frames + 8
    J = np.zeros((cap_n, cap_n)); r = np.zeros(cap_n); A = np.zeros((cap_n, cap_n)); b = np.zeros(cap_n)
    kind = np.zeros(cap_b, np.int32); bid = np.zeros(cap_b, np.int32); x0 = np.zeros(9 * cap_b)
    p = UvsPriorStruct()
    p.J, p.r, p.A, p.b, p.x0 = _p(J), _p(r), _p(A), _p(b), _p(x0)
    p.block_kind, p.block_id = kind.ctypes.data_as(c_int32_p), bid.ctypes.data_as(c_int32_p)
    p.cap_n, p.cap_blocks = cap_n, cap_b
    s = w.as_struct()
    rc = lib().orc_marginalize(C.byref(s), C.byref(opts), flag, C.byref(p))
    assert rc == 0, rc
    n, nb = p.n, p.n_blocks
    if n == 0:
        return None
    kinds = kind[:nb].copy(); ids = bid[:nb].copy()
    gs = np.array([7 if k in (0, 2) else (9 if k == 1 else 1) for k in kinds])
    return dict(n=n, m=p.m, J=J.ravel()[:n * n].reshape(n, n).copy(), r=r[:n].copy(),
                A=A.ravel()[:n * n].reshape(n, n).copy(), b=b[:n].copy(), kinds=kinds, ids=ids,
                x0=x0[:gs.sum()].copy())


def marginalize_system(w: Window, opts, flag=0):
    """-> (A [(m+n) x (m+n)], b, m, n): the system MarginalizationInfo::marginalize builds before its Schur complement
    (dropped blocks first), or None"""
    s = w.as_struct()
    cap = 16 * w.n_frames + 16 + w.n_points + 4 * w.n_lines
    A = np.zeros(cap * cap); b = np.zeros(cap)
    m, n = C.c_int(), C.c_int()
    fn = lib().orc_marginalize_system
    fn.argtypes = [C.POINTER(UvsWindowStruct), C.POINTER(UvsOptionsStruct), C.c_int, c_double_p, c_double_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    rc = fn(C.byref(s), C.byref(opts), flag, _p(A), _p(b), cap, C.byref(m), C.byref(n))
    assert rc == 0, rc
    if m.value + n.value == 0:
        return None
    pos = m.value + n.value
    return A[:pos * pos].reshape(pos, pos).copy(), b[:pos].copy(), m.value, n.value


def preintegrate(dt, acc, gyr, acc0, gyr0, ba, bg, noise):
    n = len(dt)
    f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    dt, acc, gyr, acc0, gyr0, ba, bg, noise = map(f, (dt, acc, gyr, acc0, gyr0, ba, bg, noise))
    dp = np.zeros(3); dq = np.zeros(4); dv = np.zeros(3); sdt = np.zeros(1); jac = np.zeros(225); cov = np.zeros(225)
    lib().orc_preintegrate(n, _p(dt), _p(acc), _p(gyr), _p(acc0), _p(gyr0), _p(ba), _p(bg), _p(noise), _p(dp), _p(dq),
                           _p(dv), _p(sdt), _p(jac), _p(cov))
    return dict(delta_p=dp, delta_q=dq, delta_v=dv, sum_dt=sdt[0], jacobian=jac.reshape(15, 15), covariance=cov.reshape(15, 15))


def imu_sqrt_info(cov):
    cov = np.ascontiguousarray(cov, dtype=np.float64); out = np.zeros((15, 15))
    rc = lib().orc_imu_sqrt_info(_p(cov), _p(out))
    assert rc == 0
    return out


def pose_plus(x, delta):
    x = np.ascontiguousarray(x, dtype=np.float64); d = np.ascontiguousarray(delta, dtype=np.float64); o = np.zeros(7)
    lib().orc_pose_plus(_p(x), _p(d), _p(o))
    return o


def cauchy(a, s):
    rho = np.zeros(3)
    lib().orc_cauchy(a, s, _p(rho))
    return rho


def sym_eig(A):
    A = np.ascontiguousarray(A, dtype=np.float64); n = A.shape[0]
    ev = np.zeros(n); V = np.zeros((n, n))
    lib().orc_sym_eig(n, _p(A), _p(ev), _p(V))
    return ev, V

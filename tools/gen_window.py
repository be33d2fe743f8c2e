"""Seeded synthetic sliding windows in the `uvs_window v1` description.

TOOLING (tests / bench input generator), not product code.  The reference has no window
serialisation and ships no recorded windows, so inputs are synthesised:

  * trajectory: smooth Lissajous, 10 Hz keyframes, IMU at 200 Hz with EuRoC noise values
    (config/euroc/euroc_config.yaml:60-64), preintegrated with a numpy restatement of
    IntegrationBase::midPointIntegration (factor/integration_base.h:54-158) — an implementation
    independent of oracle/factors.h, cross-checked against it in tests/test_oracle.py;
  * points / Manhattan lines / vanishing points observed under the eligibility rules of
    estimator.cpp:826 (points) and :873 (lines);
  * prior: an (F+1)-frame window is built first, solved and marginalised (MARGIN_OLD) with the CPU
    oracle, giving the F-frame window a real prior (estimator.cpp:1003-1158).

Configs (BASELINE.json): C1 11 fr/50 pt/20 ln/1 VP, C2 11/200/80/3, C5 31/2000/500/3,
"10k" 11/1500/500/3.  Seeds 1001.. as in SURVEY.md.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from uvs_b200.window import Window, default_options  # noqa: E402

G_NORM = 9.81007
ACC_N, GYR_N, ACC_W, GYR_W = 0.08, 0.004, 0.00004, 2.0e-6
FOCAL = 461.6
LINE_WINDOW = 5
# imu^R_cam / imu^T_cam literals, config/euroc/euroc_config.yaml:35-37,43
RIC_RAW = np.array([[0.0148655429818, -0.999880929698, 0.00414029679422],
                    [0.999557249008, 0.0149672133247, 0.025715529948],
                    [-0.0257744366974, 0.00375618835797, 0.999660727178]])
TIC = np.array([-0.0216401454975, -0.064676986768, 0.00981073058949])

CONFIGS = {
    "C1": dict(n_frames=11, n_points=50, n_lines=20, n_vp=1, seed=1001),
    "C2": dict(n_frames=11, n_points=200, n_lines=80, n_vp=3, seed=1002),
    "C5": dict(n_frames=31, n_points=2000, n_lines=500, n_vp=3, seed=1005),
    "10k": dict(n_frames=11, n_points=1500, n_lines=500, n_vp=3, seed=1006),
    "tiny": dict(n_frames=5, n_points=12, n_lines=6, n_vp=2, seed=1000),
}


# ---- small quaternion / rotation helpers (x,y,z,w storage like the parameter blocks) -----------
def q_mul(a, b):
    ax, ay, az, aw = a; bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx,
                     aw * bw - ax * bx - ay * by - az * bz])


def q_to_R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def R_to_q(R):
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
        q = np.zeros(4)
        q[i] = 0.25 * s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
        q[3] = (R[k, j] - R[j, k]) / s
    q = q / np.linalg.norm(q)
    return q if q[3] >= 0 else -q


def skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0.0]])


def rot_xyz(a, b, c):
    ca, sa, cb, sb, cc, sc = np.cos(a), np.sin(a), np.cos(b), np.sin(b), np.cos(c), np.sin(c)
    Rx = np.array([[1, 0, 0], [0, ca, -sa], [0, sa, ca]])
    Ry = np.array([[cb, 0, sb], [0, 1, 0], [-sb, 0, cb]])
    Rz = np.array([[cc, -sc, 0], [sc, cc, 0], [0, 0, 1]])
    return Rx @ Ry @ Rz


def plucker_to_ortho(n, d):
    """(n, d) -> [psi1, psi2, psi3, phi] with U = Rx Ry Rz, n_w = cos(phi) U[:,0], d_w = sin(phi) U[:,1]
    (the parameterisation LineProjectionFactor decodes, line_projection_factor.h:23-35)."""
    u1 = n / np.linalg.norm(n)
    u2 = d / np.linalg.norm(d)
    u2 = u2 - u1 * (u1 @ u2)
    u2 /= np.linalg.norm(u2)
    u3 = np.cross(u1, u2)
    U = np.stack([u1, u2, u3], axis=1)
    b = np.arcsin(np.clip(U[0, 2], -1, 1))
    a = np.arctan2(-U[1, 2], U[2, 2])
    c = np.arctan2(-U[0, 1], U[0, 0])
    phi = np.arctan2(np.linalg.norm(d), np.linalg.norm(n))
    return np.array([a, b, c, phi])


# ---- trajectory ----------------------------------------------------------------------------------
class Trajectory:
    def __init__(self, rng):
        self.ph = rng.uniform(0, 2 * np.pi, size=6)
        self.w = 2 * np.pi / 12.0  # 12 s period, ~1 m/s

    def p(self, t):
        w, ph = self.w, self.ph
        return np.array([2.0 * np.sin(w * t + ph[0]), 2.0 * np.sin(2 * w * t + ph[1]), 0.5 * np.sin(1.5 * w * t + ph[2])])

    def v(self, t):
        w, ph = self.w, self.ph
        return np.array([2.0 * w * np.cos(w * t + ph[0]), 4.0 * w * np.cos(2 * w * t + ph[1]), 0.75 * w * np.cos(1.5 * w * t + ph[2])])

    def a(self, t):
        w, ph = self.w, self.ph
        return np.array([-2.0 * w * w * np.sin(w * t + ph[0]), -8.0 * w * w * np.sin(2 * w * t + ph[1]),
                         -1.125 * w * w * np.sin(1.5 * w * t + ph[2])])

    def R(self, t):
        v = self.v(t)
        yaw = np.arctan2(v[1], v[0])
        roll = np.deg2rad(10) * np.sin(0.7 * self.w * t + self.ph[3])
        pitch = np.deg2rad(10) * np.sin(0.9 * self.w * t + self.ph[4])
        cy, sy, cp, sp, cr, sr = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch), np.cos(roll), np.sin(roll)
        Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
        Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
        Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
        return Rz @ Ry @ Rx

    def omega_body(self, t, h=1e-5):
        Rm, Rp = self.R(t - h), self.R(t + h)
        W = self.R(t).T @ (Rp - Rm) / (2 * h)
        return np.array([W[2, 1] - W[1, 2], W[0, 2] - W[2, 0], W[1, 0] - W[0, 1]]) * 0.5


# ---- preintegration (numpy restatement of integration_base.h:54-158) ------------------------------
def preintegrate(dts, accs, gyrs, acc0, gyr0, ba, bg):
    dp = np.zeros(3); dv = np.zeros(3); dq = np.array([0, 0, 0, 1.0])
    jac = np.eye(15); cov = np.zeros((15, 15))
    noise = np.diag(np.repeat([ACC_N ** 2, GYR_N ** 2, ACC_N ** 2, GYR_N ** 2, ACC_W ** 2, GYR_W ** 2], 3))
    a0, g0 = np.array(acc0, float), np.array(gyr0, float)
    sum_dt = 0.0
    I = np.eye(3)
    for dt, a1, g1 in zip(dts, accs, gyrs):
        Rq = q_to_R(dq)
        un_acc_0 = Rq @ (a0 - ba)
        un_gyr = 0.5 * (g0 + g1) - bg
        rq = q_mul(dq, np.array([un_gyr[0] * dt / 2, un_gyr[1] * dt / 2, un_gyr[2] * dt / 2, 1.0]))
        Rr = q_to_R(rq)
        un_acc_1 = Rr @ (a1 - ba)
        un_acc = 0.5 * (un_acc_0 + un_acc_1)
        rp = dp + dv * dt + 0.5 * un_acc * dt * dt
        rv = dv + un_acc * dt
        Rw, Ra0, Ra1 = skew(un_gyr), skew(a0 - ba), skew(a1 - ba)
        F = np.zeros((15, 15))
        F[0:3, 0:3] = I
        F[0:3, 3:6] = -0.25 * Rq @ Ra0 * dt * dt + -0.25 * Rr @ Ra1 @ (I - Rw * dt) * dt * dt
        F[0:3, 6:9] = I * dt
        F[0:3, 9:12] = -0.25 * (Rq + Rr) * dt * dt
        F[0:3, 12:15] = -0.25 * Rr @ Ra1 * dt * dt * -dt
        F[3:6, 3:6] = I - Rw * dt
        F[3:6, 12:15] = -I * dt
        F[6:9, 3:6] = -0.5 * Rq @ Ra0 * dt + -0.5 * Rr @ Ra1 @ (I - Rw * dt) * dt
        F[6:9, 6:9] = I
        F[6:9, 9:12] = -0.5 * (Rq + Rr) * dt
        F[6:9, 12:15] = -0.5 * Rr @ Ra1 * dt * -dt
        F[9:12, 9:12] = I
        F[12:15, 12:15] = I
        V = np.zeros((15, 18))
        V[0:3, 0:3] = 0.25 * Rq * dt * dt
        V[0:3, 3:6] = 0.25 * -Rr @ Ra1 * dt * dt * 0.5 * dt
        V[0:3, 6:9] = 0.25 * Rr * dt * dt
        V[0:3, 9:12] = V[0:3, 3:6]
        V[3:6, 3:6] = 0.5 * I * dt
        V[3:6, 9:12] = 0.5 * I * dt
        V[6:9, 0:3] = 0.5 * Rq * dt
        V[6:9, 3:6] = 0.5 * -Rr @ Ra1 * dt * 0.5 * dt
        V[6:9, 6:9] = 0.5 * Rr * dt
        V[6:9, 9:12] = V[6:9, 3:6]
        V[9:12, 12:15] = I * dt
        V[12:15, 15:18] = I * dt
        jac = F @ jac
        cov = F @ cov @ F.T + V @ noise @ V.T
        dp, dv = rp, rv
        dq = rq / np.linalg.norm(rq)
        sum_dt += dt
        a0, g0 = np.array(a1, float), np.array(g1, float)
    return dict(delta_p=dp, delta_q=dq, delta_v=dv, sum_dt=sum_dt, jacobian=jac, covariance=cov)


def ric_normalized():
    """parameters.cpp:107-112: the YAML rotation goes through a normalised quaternion."""
    return q_to_R(R_to_q(RIC_RAW))


# ---- window synthesis ------------------------------------------------------------------------------
def _build(n_frames, n_points, n_lines, n_vp, rng, noise_rng, with_prior_inputs=None, estimate_extrinsic=0):
    """Build a window over `n_frames` frames (no prior).  Returns (Window, truth dict)."""
    F = n_frames
    traj = Trajectory(rng)
    t0 = rng.uniform(0, 12.0)
    kf_t = t0 + 0.1 * np.arange(F)
    ric = ric_normalized()
    Rs = [traj.R(t) for t in kf_t]
    Ps = [traj.p(t) for t in kf_t]
    Vs = [traj.v(t) for t in kf_t]
    ba_true = rng.normal(0, 0.02, 3)
    bg_true = rng.normal(0, 0.002, 3)
    g = np.array([0, 0, G_NORM])

    # IMU preintegration between consecutive keyframes (20 samples of 5 ms)
    imu = []
    dt = 0.005
    for f in range(F - 1):
        ts = kf_t[f] + dt * np.arange(21)
        acc = np.array([traj.R(t).T @ (traj.a(t) + g) + ba_true for t in ts]) + rng.normal(0, ACC_N, (21, 3))
        gyr = np.array([traj.omega_body(t) + bg_true for t in ts]) + rng.normal(0, GYR_N, (21, 3))
        lin_ba = ba_true + rng.normal(0, 0.005, 3)
        lin_bg = bg_true + rng.normal(0, 0.0005, 3)
        pre = preintegrate([dt] * 20, acc[1:], gyr[1:], acc[0], gyr[0], lin_ba, lin_bg)
        pre.update(lin_ba=lin_ba, lin_bg=lin_bg, frame_i=f)
        imu.append(pre)

    Rwc = [Rs[f] @ ric for f in range(F)]
    twc = [Rs[f] @ TIC + Ps[f] for f in range(F)]

    def to_cam(f, X):
        return (X - twc[f]) @ Rwc[f]  # R^T (X - t), row-vector form

    px = 1.0 / FOCAL
    # ---- points: eligible means used_num >= 2 and start_frame < WINDOW_SIZE - 2 = F - 3
    proj = dict(fi=[], fj=[], pt=[], pi=[], pj=[])
    inv_depth_true = []
    k = 0
    tries = 0
    while k < n_points and tries < 20 * n_points + 100:
        tries += 1
        start = int(rng.integers(0, max(1, F - 3)))
        remaining = F - start
        length = int(rng.integers(2, remaining + 1))
        depth = rng.uniform(2, 10)
        xy = rng.uniform(-0.6, 0.6, 2)
        Xc = np.array([xy[0] * depth, xy[1] * depth, depth])
        Xw = Rwc[start] @ Xc + twc[start]
        obs = []
        for f in range(start, start + length):
            pc = to_cam(f, Xw)
            if pc[2] < 0.5 or abs(pc[0] / pc[2]) > 1.5 or abs(pc[1] / pc[2]) > 1.5:
                break
            obs.append(np.array([pc[0] / pc[2], pc[1] / pc[2], 1.0]) + np.append(rng.normal(0, px, 2), 0.0))
        if len(obs) < 2:
            continue
        for j in range(1, len(obs)):
            proj["fi"].append(start); proj["fj"].append(start + j); proj["pt"].append(k)
            proj["pi"].append(obs[0]); proj["pj"].append(obs[j])
        inv_depth_true.append(1.0 / depth)
        k += 1
    n_points = k

    # ---- lines: used_num >= LINE_WINDOW, one factor per observation including the start frame
    axes = np.eye(3)[:max(1, n_vp)]
    lobs = dict(f=[], l=[], sp=[], ep=[])
    vobs = dict(f=[], l=[], vp=[])
    ortho_true = []
    k = 0
    tries = 0
    while k < n_lines and tries < 20 * n_lines + 100 and F >= LINE_WINDOW:
        tries += 1
        start = int(rng.integers(0, F - LINE_WINDOW + 1))
        remaining = F - start
        length = int(rng.integers(LINE_WINDOW, remaining + 1))
        depth = rng.uniform(2, 10)
        xy = rng.uniform(-0.5, 0.5, 2)
        Xc = np.array([xy[0] * depth, xy[1] * depth, depth])
        mid = Rwc[start] @ Xc + twc[start]
        axis_id = -1
        if rng.uniform() < 0.8:
            axis_id = int(rng.integers(0, len(axes)))
            dirw = axes[axis_id].copy()
        else:
            dirw = rng.normal(0, 1, 3)
            dirw /= np.linalg.norm(dirw)
        seg = rng.uniform(0.5, 3.0)
        A, B = mid - 0.5 * seg * dirw, mid + 0.5 * seg * dirw
        obs = []
        for f in range(start, start + length):
            a = A + dirw * seg * rng.uniform(-0.1, 0.1)
            b = B + dirw * seg * rng.uniform(-0.1, 0.1)
            ac, bc = to_cam(f, a), to_cam(f, b)
            if ac[2] < 0.5 or bc[2] < 0.5:
                break
            sp = ac[:2] / ac[2] + rng.normal(0, px, 2)
            ep = bc[:2] / bc[2] + rng.normal(0, px, 2)
            vp = None
            if axis_id >= 0 and rng.uniform() < 0.8:
                dc = Rwc[f].T @ dirw
                if abs(dc[2]) >= 0.05:
                    ang = np.deg2rad(0.5)
                    dcn = dc / np.linalg.norm(dc) + rng.normal(0, ang, 3)
                    cosang = abs(dcn @ dc) / (np.linalg.norm(dcn) * np.linalg.norm(dc))
                    if abs(dcn[2]) >= 0.05 and np.arccos(min(1.0, cosang)) > 1e-4:
                        vp = np.array([dcn[0] / dcn[2], dcn[1] / dcn[2], 1.0])
            obs.append((f, sp, ep, vp))
        if len(obs) < LINE_WINDOW:
            continue
        for (f, sp, ep, vp) in obs:
            lobs["f"].append(f); lobs["l"].append(k); lobs["sp"].append(sp); lobs["ep"].append(ep)
            if vp is not None:
                vobs["f"].append(f); vobs["l"].append(k); vobs["vp"].append(vp)
        ortho_true.append(plucker_to_ortho(np.cross(A, dirw), dirw))
        k += 1
    n_lines = k

    # ---- state = truth (+) noise
    nr = noise_rng
    pose = np.zeros((F, 7)); sb = np.zeros((F, 9))
    for f in range(F):
        dth = nr.normal(0, np.deg2rad(1.0), 3)
        q = q_mul(R_to_q(Rs[f]), np.array([dth[0] / 2, dth[1] / 2, dth[2] / 2, 1.0]))
        q /= np.linalg.norm(q)
        pose[f, :3] = Ps[f] + nr.normal(0, 0.05, 3)
        pose[f, 3:] = q
        sb[f, :3] = Vs[f] + nr.normal(0, 0.05, 3)
        sb[f, 3:6] = ba_true + nr.normal(0, 0.01, 3)
        sb[f, 6:9] = bg_true + nr.normal(0, 0.001, 3)
    ex = np.concatenate([TIC, R_to_q(ric)])
    inv_depth = np.array(inv_depth_true) * (1 + nr.normal(0, 0.1, n_points)) if n_points else np.zeros(0)
    ortho = (np.array(ortho_true) + nr.normal(0, 0.02, (n_lines, 4))) if n_lines else np.zeros((0, 4))

    def arr(x, shape, dt=np.float64):
        return np.array(x, dtype=dt).reshape(shape)

    w = Window(
        pose=pose, speed_bias=sb, ex_pose=ex, td=np.zeros(1), inv_depth=inv_depth, ortho=ortho,
        proj_frame_i=arr(proj["fi"], (-1,), np.int32), proj_frame_j=arr(proj["fj"], (-1,), np.int32),
        proj_point=arr(proj["pt"], (-1,), np.int32), proj_pts_i=arr(proj["pi"], (-1, 3)), proj_pts_j=arr(proj["pj"], (-1, 3)),
        line_frame=arr(lobs["f"], (-1,), np.int32), line_idx=arr(lobs["l"], (-1,), np.int32),
        line_sp=arr(lobs["sp"], (-1, 2)), line_ep=arr(lobs["ep"], (-1, 2)),
        vp_frame=arr(vobs["f"], (-1,), np.int32), vp_line=arr(vobs["l"], (-1,), np.int32), vp_dir=arr(vobs["vp"], (-1, 3)),
        line_ric=ric.copy(), line_tic=TIC.copy(),
        imu_frame_i=arr([p["frame_i"] for p in imu], (-1,), np.int32),
        imu_delta_p=arr([p["delta_p"] for p in imu], (-1, 3)), imu_delta_q=arr([p["delta_q"] for p in imu], (-1, 4)),
        imu_delta_v=arr([p["delta_v"] for p in imu], (-1, 3)), imu_sum_dt=arr([p["sum_dt"] for p in imu], (-1,)),
        imu_lin_ba=arr([p["lin_ba"] for p in imu], (-1, 3)), imu_lin_bg=arr([p["lin_bg"] for p in imu], (-1, 3)),
        imu_jacobian=arr([p["jacobian"] for p in imu], (-1, 225)), imu_covariance=arr([p["covariance"] for p in imu], (-1, 225)),
        estimate_extrinsic=estimate_extrinsic,
    )
    truth = dict(Rs=Rs, Ps=Ps, Vs=Vs, ba=ba_true, bg=bg_true, inv_depth=np.array(inv_depth_true),
                 ortho=np.array(ortho_true).reshape(-1, 4), Rwc=Rwc, twc=twc)
    return w, truth


def truth_window(w: Window, truth) -> Window:
    """Copy of `w` with the state set to the noise-free truth."""
    t = w.copy()
    F = w.n_frames
    for f in range(F):
        t.pose[f, :3] = truth["Ps"][f]
        t.pose[f, 3:] = R_to_q(truth["Rs"][f])
        t.speed_bias[f, :3] = truth["Vs"][f]
        t.speed_bias[f, 3:6] = truth["ba"]
        t.speed_bias[f, 6:9] = truth["bg"]
    t.inv_depth[:] = truth["inv_depth"]
    t.ortho[:] = truth["ortho"]
    return t


def drop_first_frame(w: Window, truth, rng) -> tuple:
    """slideWindow for MARGIN_OLD restated on the flat description (estimator.cpp:1235-1287,
    feature_manager removeBackShiftDepth): frame 0 and its observations go away; features that
    started there are re-anchored at their next observation."""
    F = w.n_frames
    keep_f = slice(1, F)
    # points
    new = dict(fi=[], fj=[], pt=[], pi=[], pj=[])
    inv_depth, inv_true, remap = [], [], {}
    for k in range(w.n_points):
        idx = np.nonzero(w.proj_point == k)[0]
        fi = int(w.proj_frame_i[idx[0]])
        frames = [fi] + [int(f) for f in w.proj_frame_j[idx]]
        pts = [w.proj_pts_i[idx[0]]] + [w.proj_pts_j[i] for i in idx]
        depth_inv, depth_inv_true = w.inv_depth[k], truth["inv_depth"][k]
        if fi == 0:
            frames, pts = frames[1:], pts[1:]
            if len(frames) < 2:
                continue
            # true depth in the new anchor frame
            Xc0 = pts and None
            Xw = truth["Rwc"][0] @ (np.array([w.proj_pts_i[idx[0]][0], w.proj_pts_i[idx[0]][1], 1.0]) / depth_inv_true) + truth["twc"][0]
            z = ((Xw - truth["twc"][frames[0]]) @ truth["Rwc"][frames[0]])[2]
            depth_inv_true = 1.0 / z
            depth_inv = depth_inv_true * (1 + rng.normal(0, 0.1))
        start = frames[0] - 1
        if not (len(frames) >= 2 and start < (F - 1) - 3):
            continue
        kk = len(inv_depth)
        remap[k] = kk
        inv_depth.append(depth_inv); inv_true.append(depth_inv_true)
        for j in range(1, len(frames)):
            new["fi"].append(start); new["fj"].append(frames[j] - 1); new["pt"].append(kk)
            new["pi"].append(pts[0]); new["pj"].append(pts[j])
    # lines
    lnew = dict(f=[], l=[], sp=[], ep=[]); vnew = dict(f=[], l=[], vp=[])
    ortho, ortho_true, lremap = [], [], {}
    for k in range(w.n_lines):
        idx = np.nonzero((w.line_idx == k) & (w.line_frame >= 1))[0]
        if len(idx) < LINE_WINDOW:
            continue
        kk = len(ortho)
        lremap[k] = kk
        ortho.append(w.ortho[k]); ortho_true.append(truth["ortho"][k])
        for i in idx:
            lnew["f"].append(int(w.line_frame[i]) - 1); lnew["l"].append(kk); lnew["sp"].append(w.line_sp[i]); lnew["ep"].append(w.line_ep[i])
        for i in np.nonzero((w.vp_line == k) & (w.vp_frame >= 1))[0]:
            vnew["f"].append(int(w.vp_frame[i]) - 1); vnew["l"].append(kk); vnew["vp"].append(w.vp_dir[i])

    def arr(x, shape, dt=np.float64):
        return np.array(x, dtype=dt).reshape(shape)
    imu_keep = w.imu_frame_i >= 1
    out = Window(
        pose=w.pose[keep_f].copy(), speed_bias=w.speed_bias[keep_f].copy(), ex_pose=w.ex_pose.copy(), td=w.td.copy(),
        inv_depth=arr(inv_depth, (-1,)), ortho=arr(ortho, (-1, 4)),
        proj_frame_i=arr(new["fi"], (-1,), np.int32), proj_frame_j=arr(new["fj"], (-1,), np.int32),
        proj_point=arr(new["pt"], (-1,), np.int32), proj_pts_i=arr(new["pi"], (-1, 3)), proj_pts_j=arr(new["pj"], (-1, 3)),
        line_frame=arr(lnew["f"], (-1,), np.int32), line_idx=arr(lnew["l"], (-1,), np.int32),
        line_sp=arr(lnew["sp"], (-1, 2)), line_ep=arr(lnew["ep"], (-1, 2)),
        vp_frame=arr(vnew["f"], (-1,), np.int32), vp_line=arr(vnew["l"], (-1,), np.int32), vp_dir=arr(vnew["vp"], (-1, 3)),
        line_ric=w.line_ric.copy(), line_tic=w.line_tic.copy(),
        imu_frame_i=(w.imu_frame_i[imu_keep] - 1).astype(np.int32),
        imu_delta_p=w.imu_delta_p[imu_keep], imu_delta_q=w.imu_delta_q[imu_keep], imu_delta_v=w.imu_delta_v[imu_keep],
        imu_sum_dt=w.imu_sum_dt[imu_keep], imu_lin_ba=w.imu_lin_ba[imu_keep], imu_lin_bg=w.imu_lin_bg[imu_keep],
        imu_jacobian=w.imu_jacobian[imu_keep], imu_covariance=w.imu_covariance[imu_keep],
        estimate_extrinsic=w.estimate_extrinsic, estimate_td=w.estimate_td,
    )
    t2 = dict(Rs=truth["Rs"][1:], Ps=truth["Ps"][1:], Vs=truth["Vs"][1:], ba=truth["ba"], bg=truth["bg"],
              inv_depth=np.array(inv_true), ortho=np.array(ortho_true).reshape(-1, 4), Rwc=truth["Rwc"][1:], twc=truth["twc"][1:])
    return out, t2


def truncate_landmarks(w: Window, n_points, n_lines, truth=None) -> Window:
    """Keep the first n_points points / n_lines lines (the generator over-produces because features
    anchored in the dropped frame can fall out of the window)."""
    if w.n_points <= n_points and w.n_lines <= n_lines:
        return w
    n_points, n_lines = min(n_points, w.n_points), min(n_lines, w.n_lines)
    pk, lk, vk = w.proj_point < n_points, w.line_idx < n_lines, w.vp_line < n_lines
    out = w.copy()
    out.inv_depth, out.ortho = w.inv_depth[:n_points].copy(), w.ortho[:n_lines].copy()
    for n in ("proj_frame_i", "proj_frame_j", "proj_point", "proj_pts_i", "proj_pts_j"):
        setattr(out, n, getattr(w, n)[pk].copy())
    for n in ("line_frame", "line_idx", "line_sp", "line_ep"):
        setattr(out, n, getattr(w, n)[lk].copy())
    for n in ("vp_frame", "vp_line", "vp_dir"):
        setattr(out, n, getattr(w, n)[vk].copy())
    if truth is not None:
        truth["inv_depth"] = truth["inv_depth"][:n_points]
        truth["ortho"] = truth["ortho"][:n_lines]
    return out.normalize()


def make_window(config="C2", seed=None, with_prior=True, estimate_extrinsic=0, return_truth=False, **over):
    """Generate one window.  `config` is a key of CONFIGS or a dict with the same keys."""
    cfg = dict(CONFIGS[config]) if isinstance(config, str) else dict(config)
    cfg.update(over)
    if seed is None:
        seed = cfg["seed"]
    rng = np.random.default_rng(seed)
    noise_rng = np.random.default_rng(seed + 7919)
    F = cfg["n_frames"]
    # over-generate: features anchored in the dropped frame can fall out of the window
    grow = (F + 1) / F * 1.12 if with_prior else 1.0
    w, truth = _build(F + (1 if with_prior else 0), int(np.ceil(cfg["n_points"] * grow)), int(np.ceil(cfg["n_lines"] * grow)),
                      cfg["n_vp"], rng, noise_rng, estimate_extrinsic=estimate_extrinsic)
    if with_prior:
        from tests import orc  # CPU oracle: tooling use only
        opts = default_options()
        orc.solve(w, opts)  # the reference marginalises after solving (estimator.cpp:994-1003)
        prior = orc.marginalize(w, opts, 0)
        w2, truth = drop_first_frame(w, truth, noise_rng)
        # fresh initial guess for the new window: truth (+) noise, as in SURVEY.md 8d
        F2 = w2.n_frames
        for f in range(F2):
            dth = noise_rng.normal(0, np.deg2rad(1.0), 3)
            q = q_mul(R_to_q(truth["Rs"][f]), np.array([dth[0] / 2, dth[1] / 2, dth[2] / 2, 1.0]))
            w2.pose[f, :3] = truth["Ps"][f] + noise_rng.normal(0, 0.05, 3)
            w2.pose[f, 3:] = q / np.linalg.norm(q)
            w2.speed_bias[f, :3] = truth["Vs"][f] + noise_rng.normal(0, 0.05, 3)
            w2.speed_bias[f, 3:6] = truth["ba"] + noise_rng.normal(0, 0.01, 3)
            w2.speed_bias[f, 6:9] = truth["bg"] + noise_rng.normal(0, 0.001, 3)
        if prior is not None:
            w2.set_prior(prior["J"], prior["r"], prior["kinds"], prior["ids"], prior["x0"])
        w = truncate_landmarks(w2, cfg["n_points"], cfg["n_lines"], truth)
    w.normalize()
    return (w, truth) if return_truth else w


def make_batch(config="C2", n=64, seed0=1004, **kw):
    """Config C4: n independent windows, seeds seed0 + i."""
    return [make_window(config, seed=seed0 + i, **kw) for i in range(n)]


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C2")
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    win = make_window(a.config, seed=a.seed)
    print("frames %d points %d lines %d proj %d line_obs %d vp_obs %d imu %d prior_n %d" % (
        win.n_frames, win.n_points, win.n_lines, win.n_proj, win.n_line_obs, win.n_vp_obs, win.n_imu, win.prior_n))
    if a.out:
        win.save(a.out)


def add_relocalisation(w: Window, truth, frame=3, n_match=15, seed=0) -> Window:
    """Relocalisation factors (estimator.cpp:944-978, :1366-1378) for window `w`: a loop-closure frame that saw the same place
    as window frame `frame` from a slightly different pose; `n_match` eligible features whose track starts at or before that
    frame are matched in it (match_points), relo_Pose starts at the window's estimate of that frame (estimator.cpp:1376).
    -> a copy of `w` with relo_pose / relo_point / relo_pts_j set; truth["relo_pose"] = the loop frame's true pose."""
    rng = np.random.default_rng(seed + 31)
    ric = ric_normalized()
    dth = rng.normal(0, np.deg2rad(2.0), 3)
    q_loop = q_mul(R_to_q(truth["Rs"][frame]), np.array([dth[0] / 2, dth[1] / 2, dth[2] / 2, 1.0]))
    q_loop /= np.linalg.norm(q_loop)
    R_loop, P_loop = q_to_R(q_loop), truth["Ps"][frame] + rng.normal(0, 0.1, 3)
    Rwc, twc = R_loop @ ric, R_loop @ TIC + P_loop
    first = {}
    for k in range(w.n_proj):
        first.setdefault(int(w.proj_point[k]), k)
    pts, pj = [], []
    for p in sorted(first):
        a = first[p]
        fi = int(w.proj_frame_i[a])
        if fi > frame:
            continue
        Xc = w.proj_pts_i[a] / truth["inv_depth"][p]
        Xw = truth["Rwc"][fi] @ Xc + truth["twc"][fi]
        pc = (Xw - twc) @ Rwc
        if pc[2] < 0.5 or abs(pc[0] / pc[2]) > 1.5 or abs(pc[1] / pc[2]) > 1.5:
            continue
        pts.append(p)
        pj.append(np.array([pc[0] / pc[2], pc[1] / pc[2], 1.0]) + np.append(rng.normal(0, 1.0 / FOCAL, 2), 0.0))
        if len(pts) == n_match:
            break
    out = w.copy()
    out.relo_pose = w.pose[frame].copy()
    out.relo_point = np.array(pts, np.int32)
    out.relo_pts_j = np.array(pj).reshape(-1, 3)
    out.normalize()
    truth["relo_pose"] = np.concatenate([P_loop, q_loop])
    return out


def expand_relocalisation(w: Window) -> Window:
    """The same problem written without the relocalisation fields: relo_Pose as one more frame at the end of the window
    (no IMU factor, a speed-bias block no factor touches), its factors as ordinary point factors observing from that frame -
    what the library does internally (uvs_api.cu) and what the oracle's F_RELO evaluation is checked against."""
    F = w.n_frames
    out = w.copy()
    out.pose = np.vstack([w.pose, w.relo_pose.reshape(1, 7)])
    out.speed_bias = np.vstack([w.speed_bias, np.zeros((1, 9))])
    fi, fj, pt, pi, pj = [], [], [], [], []
    relo_of = {int(p): r for r, p in enumerate(w.relo_point)}
    for k in range(w.n_proj):
        fi.append(w.proj_frame_i[k]); fj.append(w.proj_frame_j[k]); pt.append(w.proj_point[k]); pi.append(w.proj_pts_i[k]); pj.append(w.proj_pts_j[k])
        p = int(w.proj_point[k])
        if (k + 1 == w.n_proj or int(w.proj_point[k + 1]) != p) and p in relo_of:
            fi.append(w.proj_frame_i[k]); fj.append(F); pt.append(p); pi.append(w.proj_pts_i[k]); pj.append(w.relo_pts_j[relo_of[p]])
    out.proj_frame_i, out.proj_frame_j, out.proj_point = (np.array(a, np.int32) for a in (fi, fj, pt))
    out.proj_pts_i, out.proj_pts_j = np.array(pi).reshape(-1, 3), np.array(pj).reshape(-1, 3)
    out.relo_pose, out.relo_point, out.relo_pts_j = np.zeros(0), np.zeros(0, np.int32), np.zeros((0, 3))
    out.normalize()
    return out

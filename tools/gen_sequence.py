"""Seeded synthetic FRAME STREAM for the device-resident sliding window: what the feature tracker and
the IMU hand the estimator frame by frame - tracked point / line observations with persistent ids, one preintegration
record per keyframe interval - plus the ground truth needed to seed the state.  TOOLING (tests / bench input), not product
code.  Same trajectory / noise model as tools/gen_window.py."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tools import gen_window as gw  # noqa: E402


class Frame:
    """one image frame: truth pose, tracked observations, IMU samples since the previous frame"""
    def __init__(self):
        self.R = self.P = self.V = None
        self.point_id, self.point_xyz = [], []
        self.line_id, self.line_sp, self.line_ep, self.line_vp = [], [], [], []
        self.imu_samples = None      # (dts, accs, gyrs, acc0, gyr0, lin_ba, lin_bg) from the previous frame to this one
        self.imu = None              # preintegration record of those samples


def imu_record(samples):
    dts, accs, gyrs, acc0, gyr0, lin_ba, lin_bg = samples
    pre = gw.preintegrate(dts, accs, gyrs, acc0, gyr0, lin_ba, lin_bg)
    pre.update(lin_ba=lin_ba, lin_bg=lin_bg)
    return pre


def merge_samples(a, b):
    """the samples of two consecutive intervals as one (Estimator::slideWindow pushes the newest interval's samples into the
    one before, estimator.cpp:1301-1312): starts from a's first measurement with a's linearisation biases"""
    return (list(a[0]) + list(b[0]), np.concatenate([a[1], b[1]]), np.concatenate([a[2], b[2]]), a[3], a[4], a[5], a[6])


class Sequence:
    """n_frames keyframes at 10 Hz; about n_points point tracks and n_lines line tracks alive per frame"""

    def __init__(self, n_frames=32, n_points=120, n_lines=40, n_vp=3, seed=7):
        rng = np.random.default_rng(seed)
        self.rng = rng
        traj = gw.Trajectory(rng)
        t0 = rng.uniform(0, 12.0)
        self.t = t0 + 0.1 * np.arange(n_frames)
        self.ric = gw.ric_normalized()
        self.tic = gw.TIC.copy()
        self.ba = rng.normal(0, 0.02, 3)
        self.bg = rng.normal(0, 0.002, 3)
        g = np.array([0, 0, gw.G_NORM])
        px = 1.0 / gw.FOCAL
        self.frames = []
        self.point_truth, self.line_truth = {}, {}     # id -> world point / (A, dir)
        active_pts, active_lns = {}, {}                # id -> frames left
        next_pid, next_lid = 0, 0
        axes = np.eye(3)[:max(1, n_vp)]
        dt = 0.005
        for k in range(n_frames):
            fr = Frame()
            t = self.t[k]
            fr.R, fr.P, fr.V = traj.R(t), traj.p(t), traj.v(t)
            Rwc, twc = fr.R @ self.ric, fr.R @ self.tic + fr.P
            if k > 0:
                ts = self.t[k - 1] + dt * np.arange(21)
                acc = np.array([traj.R(s).T @ (traj.a(s) + g) + self.ba for s in ts]) + rng.normal(0, gw.ACC_N, (21, 3))
                gyr = np.array([traj.omega_body(s) + self.bg for s in ts]) + rng.normal(0, gw.GYR_N, (21, 3))
                lin_ba = self.ba + rng.normal(0, 0.005, 3)
                lin_bg = self.bg + rng.normal(0, 0.0005, 3)
                fr.imu_samples = ([dt] * 20, acc[1:], gyr[1:], acc[0], gyr[0], lin_ba, lin_bg)
                fr.imu = imu_record(fr.imu_samples)
            to_cam = lambda X: (X - twc) @ Rwc
            # ---- points: continue the tracks that stay in view, spawn new ones up to n_points
            for pid in list(active_pts):
                pc = to_cam(self.point_truth[pid])
                if active_pts[pid] <= 0 or pc[2] < 0.5 or abs(pc[0] / pc[2]) > 1.5 or abs(pc[1] / pc[2]) > 1.5:
                    del active_pts[pid]
                    continue
                active_pts[pid] -= 1
                fr.point_id.append(pid)
                fr.point_xyz.append(np.array([pc[0] / pc[2], pc[1] / pc[2], 1.0]) + np.append(rng.normal(0, px, 2), 0.0))
            while len(active_pts) < n_points:
                depth = rng.uniform(2, 10)
                xy = rng.uniform(-0.6, 0.6, 2)
                Xw = Rwc @ np.array([xy[0] * depth, xy[1] * depth, depth]) + twc
                pid = next_pid; next_pid += 1
                self.point_truth[pid] = Xw
                active_pts[pid] = int(rng.integers(1, 16))          # frames it lives on after this one
                fr.point_id.append(pid)
                fr.point_xyz.append(np.array([xy[0], xy[1], 1.0]) + np.append(rng.normal(0, px, 2), 0.0))
            # ---- lines
            def observe_line(lid):
                A, dirw, seg, axis_id = self.line_truth[lid]
                a = A + dirw * seg * rng.uniform(-0.1, 0.1)
                b = A + dirw * seg * (1.0 + rng.uniform(-0.1, 0.1))
                ac, bc = to_cam(a), to_cam(b)
                if ac[2] < 0.5 or bc[2] < 0.5:
                    return None
                sp = ac[:2] / ac[2] + rng.normal(0, px, 2)
                ep = bc[:2] / bc[2] + rng.normal(0, px, 2)
                vp = np.zeros(3)
                if axis_id >= 0 and rng.uniform() < 0.8:
                    dc = Rwc.T @ dirw
                    if abs(dc[2]) >= 0.05:
                        dcn = dc / np.linalg.norm(dc) + rng.normal(0, np.deg2rad(0.5), 3)
                        cosang = abs(dcn @ dc) / (np.linalg.norm(dcn) * np.linalg.norm(dc))
                        if abs(dcn[2]) >= 0.05 and np.arccos(min(1.0, cosang)) > 1e-4:
                            vp = np.array([dcn[0] / dcn[2], dcn[1] / dcn[2], 1.0])
                return sp, ep, vp
            for lid in list(active_lns):
                ob = observe_line(lid) if active_lns[lid] > 0 else None
                if ob is None:
                    del active_lns[lid]
                    continue
                active_lns[lid] -= 1
                fr.line_id.append(lid); fr.line_sp.append(ob[0]); fr.line_ep.append(ob[1]); fr.line_vp.append(ob[2])
            tries = 0
            while len(active_lns) < n_lines and tries < 10 * n_lines:
                tries += 1
                depth = rng.uniform(2, 10)
                xy = rng.uniform(-0.5, 0.5, 2)
                mid = Rwc @ np.array([xy[0] * depth, xy[1] * depth, depth]) + twc
                axis_id = -1
                if rng.uniform() < 0.8:
                    axis_id = int(rng.integers(0, len(axes)))
                    dirw = axes[axis_id].copy()
                else:
                    dirw = rng.normal(0, 1, 3)
                    dirw /= np.linalg.norm(dirw)
                seg = rng.uniform(0.5, 3.0)
                lid = next_lid
                self.line_truth[lid] = (mid - 0.5 * seg * dirw, dirw, seg, axis_id)
                ob = observe_line(lid)
                if ob is None:
                    del self.line_truth[lid]
                    continue
                next_lid += 1
                active_lns[lid] = int(rng.integers(2, 18))
                fr.line_id.append(lid); fr.line_sp.append(ob[0]); fr.line_ep.append(ob[1]); fr.line_vp.append(ob[2])
            fr.point_id = np.array(fr.point_id, np.int32); fr.point_xyz = np.array(fr.point_xyz).reshape(-1, 3)
            fr.line_id = np.array(fr.line_id, np.int32); fr.line_sp = np.array(fr.line_sp).reshape(-1, 2)
            fr.line_ep = np.array(fr.line_ep).reshape(-1, 2); fr.line_vp = np.array(fr.line_vp).reshape(-1, 3)
            self.frames.append(fr)

    # ---- state the estimator would hold for a window made of the absolute frames `frame_ids`
    def noisy_pose_sb(self, frame_ids, nrng):
        F = len(frame_ids)
        pose, sb = np.zeros((F, 7)), np.zeros((F, 9))
        for i, k in enumerate(frame_ids):
            fr = self.frames[k]
            dth = nrng.normal(0, np.deg2rad(0.5), 3)
            q = gw.q_mul(gw.R_to_q(fr.R), np.array([dth[0] / 2, dth[1] / 2, dth[2] / 2, 1.0]))
            pose[i, :3] = fr.P + nrng.normal(0, 0.03, 3)
            pose[i, 3:] = q / np.linalg.norm(q)
            sb[i, :3] = fr.V + nrng.normal(0, 0.03, 3)
            sb[i, 3:6] = self.ba + nrng.normal(0, 0.005, 3)
            sb[i, 6:9] = self.bg + nrng.normal(0, 0.0005, 3)
        return pose, sb

    def inv_depth_of(self, pid, start_abs):
        fr = self.frames[start_abs]
        Rwc, twc = fr.R @ self.ric, fr.R @ self.tic + fr.P
        return 1.0 / ((self.point_truth[pid] - twc) @ Rwc)[2]

    def ortho_of(self, lid):
        A, dirw, _, _ = self.line_truth[lid]
        return gw.plucker_to_ortho(np.cross(A, dirw), dirw)

    def ex_pose(self):
        return np.concatenate([self.tic, gw.R_to_q(self.ric)])

"""Host-side restatement of the reference's per-frame bookkeeping, as the device-resident window's checker (TEST
INFRASTRUCTURE): FeatureManager's track lists (feature_manager.cpp:73-133 addFeatureCheckParallax - bookkeeping part,
:607-723 removeBackShiftDepth / removeBack / removeFront / removeLineBack / removeLineFront), Estimator::slideWindow's IMU
shuffling (estimator.cpp:1235-1334) and the problem-assembly loops of optimization() (estimator.cpp:823-931).
pack() builds the window the reference would hand to Ceres, as a uvs_b200.Window packed entirely on the host."""
import numpy as np

from uvs_b200 import Window

MARGIN_OLD, MARGIN_SECOND_NEW = 0, 1


class Track:
    def __init__(self, fid, start):
        self.id, self.start, self.obs = fid, start, []


class HostWindow:
    def __init__(self, window_size=10, line_window=5):
        self.W, self.line_window = window_size, line_window
        self.points, self.lines = [], []          # std::list<FeaturePerId> / <LineFeaturePerId>, insertion order
        self.frame_ids = []                       # absolute ids of the frames in the window (test-side convenience)
        self.imu = []                             # one record per interval
        self.imu_samples = []

    def push(self, abs_id, fr):
        pos = len(self.frame_ids)
        by_id = {t.id: t for t in self.points}
        for pid, xyz in zip(fr.point_id, fr.point_xyz):
            t = by_id.get(int(pid))
            if t is None:
                t = Track(int(pid), pos)
                self.points.append(t)
            t.obs.append(np.array(xyz))
        by_id = {t.id: t for t in self.lines}
        for lid, sp, ep, vp in zip(fr.line_id, fr.line_sp, fr.line_ep, fr.line_vp):
            t = by_id.get(int(lid))
            if t is None:
                t = Track(int(lid), pos)
                self.lines.append(t)
            t.obs.append((np.array(sp), np.array(ep), np.array(vp)))
        if pos > 0:
            self.imu.append(fr.imu)
            self.imu_samples.append(fr.imu_samples)
        self.frame_ids.append(abs_id)

    def eligible_points(self):
        return [t for t in self.points if len(t.obs) >= 2 and t.start < self.W - 2]      # estimator.cpp:826

    def eligible_lines(self):
        return [t for t in self.lines if len(t.obs) >= self.line_window]                 # estimator.cpp:873

    def slide(self, flag, merged=None, merged_samples=None):
        fc = len(self.frame_ids) - 1
        if flag == MARGIN_OLD:
            for lst, min_left in ((self.points, 2), (self.lines, 1)):
                keep = []
                for t in lst:
                    if t.start != 0:
                        t.start -= 1
                    else:
                        t.obs.pop(0)
                        if len(t.obs) < min_left:
                            continue
                    keep.append(t)
                lst[:] = keep
            self.frame_ids.pop(0)
            if self.imu:
                self.imu.pop(0); self.imu_samples.pop(0)
        else:
            for lst in (self.points, self.lines):
                keep = []
                for t in lst:
                    if t.start == fc:
                        t.start -= 1
                    else:
                        j = fc - 1 - t.start
                        if t.start + len(t.obs) - 1 >= fc - 1:
                            t.obs.pop(j)
                            if len(t.obs) == 0:
                                continue
                    keep.append(t)
                lst[:] = keep
            self.frame_ids.pop(fc - 1)
            if len(self.imu) >= 2:
                self.imu[-2:] = [merged]
                self.imu_samples[-2:] = [merged_samples]
            elif self.imu:
                self.imu.clear(); self.imu_samples.clear()

    def pack(self, pose, speed_bias, ex_pose, inv_depth, ortho, ric, tic, prior=None) -> Window:
        fi, fj, pt, pi, pj = [], [], [], [], []
        for k, t in enumerate(self.eligible_points()):
            for j in range(1, len(t.obs)):
                fi.append(t.start); fj.append(t.start + j); pt.append(k); pi.append(t.obs[0]); pj.append(t.obs[j])
        lf, li, sp, ep, vf, vl, vd = [], [], [], [], [], [], []
        for k, t in enumerate(self.eligible_lines()):
            for j, (s, e, v) in enumerate(t.obs):
                lf.append(t.start + j); li.append(k); sp.append(s); ep.append(e)
                if v[2] == 1.0:
                    vf.append(t.start + j); vl.append(k); vd.append(v)
        arr = lambda x, shape, dt=np.float64: np.array(x, dtype=dt).reshape(shape)
        n = len(self.imu)
        w = Window(pose=pose, speed_bias=speed_bias, ex_pose=ex_pose, td=np.zeros(1), inv_depth=inv_depth, ortho=ortho,
                   proj_frame_i=arr(fi, (-1,), np.int32), proj_frame_j=arr(fj, (-1,), np.int32), proj_point=arr(pt, (-1,), np.int32),
                   proj_pts_i=arr(pi, (-1, 3)), proj_pts_j=arr(pj, (-1, 3)),
                   line_frame=arr(lf, (-1,), np.int32), line_idx=arr(li, (-1,), np.int32), line_sp=arr(sp, (-1, 2)), line_ep=arr(ep, (-1, 2)),
                   vp_frame=arr(vf, (-1,), np.int32), vp_line=arr(vl, (-1,), np.int32), vp_dir=arr(vd, (-1, 3)),
                   line_ric=np.array(ric), line_tic=np.array(tic),
                   imu_frame_i=np.arange(n, dtype=np.int32),
                   imu_delta_p=arr([p["delta_p"] for p in self.imu], (-1, 3)), imu_delta_q=arr([p["delta_q"] for p in self.imu], (-1, 4)),
                   imu_delta_v=arr([p["delta_v"] for p in self.imu], (-1, 3)), imu_sum_dt=arr([p["sum_dt"] for p in self.imu], (-1,)),
                   imu_lin_ba=arr([p["lin_ba"] for p in self.imu], (-1, 3)), imu_lin_bg=arr([p["lin_bg"] for p in self.imu], (-1, 3)),
                   imu_jacobian=arr([p["jacobian"] for p in self.imu], (-1, 225)), imu_covariance=arr([p["covariance"] for p in self.imu], (-1, 225)))
        if prior is not None:
            w.set_prior(prior["J"], prior["r"], prior["kinds"], prior["ids"], prior["x0"])
        return w

// ORACLE — TEST INFRASTRUCTURE ONLY (see smallmat.h).  PARITY UNPINNED at the Ceres/Eigen boundary.
//
// problem.h: the residual-block list Estimator::optimization() hands to Ceres
// (vins_estimator/src/estimator.cpp:761-978), restated over the flat `UvsWindow` description, and the
// per-block evaluation Ceres performs: Evaluate -> local-parameterisation Jacobian (first 6 columns
// of each 7-wide pose block, pose_local_parameterization.cpp:20-27) -> loss correction
// (restated in-tree at factor/marginalization_factor.cpp:37-68).
#pragma once
#include <cstdint>
#include <limits>
#include <vector>

#include "../include/uvs.h"
#include "factors.h"

namespace orc {

enum FactorType { F_PRIOR = 0, F_IMU = 1, F_PROJ = 2, F_LINE = 3, F_VP = 4, F_RELO = 5 };   // F_RELO: estimator.cpp:944-978
constexpr int F_LAST = F_RELO;

// Tangent-space layout used by the oracle's linear algebra:
//   camera set:   pose f -> 15f, speed-bias f -> 15f+6, [ex-pose -> 15F], [td -> 15F(+6)], [relo_Pose -> after those]
//   landmark set: point k -> d + k, line k -> d + Np + 4k
struct Layout {
  int F = 0, Np = 0, Nl = 0;
  int ex_off = -1, td_off = -1, relo_off = -1;
  int d = 0;      // camera dims
  int total = 0;  // camera + landmark dims
  int pose(int f) const { return 15 * f; }
  int sb(int f) const { return 15 * f + 6; }
  int point(int k) const { return d + k; }
  int line(int k) const { return d + Np + 4 * k; }
};

struct State {
  std::vector<double> pose, sb, ex, td, inv_depth, ortho, relo;
};

// One evaluated (non-prior) residual block.  Jacobians are row-major nr x gs[b] ("Ceres layout")
// when raw, nr x ls[b] (tangent columns, loss-corrected) after localize_and_correct().
struct BlockEval {
  static const int MAXB = 5, MAXJ = 15 * 9;
  int nr = 0, nb = 0;
  int off[MAXB];  // tangent offset of each parameter block, -1 = constant block
  int gs[MAXB];   // global size
  int ls[MAXB];   // local size
  double r[15];
  double J[MAXB][MAXJ];
  double cost = 0.0;
};

class Problem {
 public:
  UvsWindow w;  // shallow view; caller keeps the buffers alive
  UvsOptions o;
  Layout lay;
  std::vector<double> imu_sqrt_info_;  // [n_imu][225]
  std::vector<int> prior_gs_, prior_off_, prior_col_;  // global size, tangent offset, column in J0
  std::vector<int> relo_anchor_;                       // first projection factor of every matched point (its start frame / pts_i)

  Problem(const UvsWindow &win, const UvsOptions &opt) : w(win), o(opt) {
    lay.F = w.n_frames; lay.Np = w.n_points; lay.Nl = w.n_lines;
    lay.d = 15 * lay.F;
    if (w.estimate_extrinsic) { lay.ex_off = lay.d; lay.d += 6; }
    if (w.estimate_td) { lay.td_off = lay.d; lay.d += 1; }
    if (w.n_relo > 0) {   // problem.AddParameterBlock(relo_Pose, SIZE_POSE, local_parameterization), estimator.cpp:947-948
      lay.relo_off = lay.d; lay.d += 6;
      for (int r = 0; r < w.n_relo; r++) {
        int a = -1;
        for (int k = 0; k < w.n_proj && a < 0; k++) if (w.proj_point[k] == w.relo_point[r]) a = k;
        relo_anchor_.push_back(a);
      }
    }
    lay.total = lay.d + lay.Np + 4 * lay.Nl;
    imu_sqrt_info_.resize((size_t)w.n_imu * 225);
    for (int k = 0; k < w.n_imu; k++) imu_sqrt_info(w.imu_covariance + (size_t)k * 225, &imu_sqrt_info_[(size_t)k * 225]);
    int col = 0;
    for (int b = 0; b < (w.prior_n > 0 ? w.prior_n_blocks : 0); b++) {
      const int kind = w.prior_block_kind[b], id = w.prior_block_id[b];
      int gs, off;
      if (kind == UVS_BLOCK_POSE) { gs = 7; off = lay.pose(id); }
      else if (kind == UVS_BLOCK_SPEEDBIAS) { gs = 9; off = lay.sb(id); }
      else if (kind == UVS_BLOCK_EXPOSE) { gs = 7; off = lay.ex_off; }
      else { gs = 1; off = lay.td_off; }
      prior_gs_.push_back(gs); prior_off_.push_back(off); prior_col_.push_back(col);
      col += gs == 7 ? 6 : gs;
    }
  }

  State initial_state() const {
    State s;
    s.pose.assign(w.pose, w.pose + 7 * w.n_frames);
    s.sb.assign(w.speed_bias, w.speed_bias + 9 * w.n_frames);
    s.ex.assign(w.ex_pose, w.ex_pose + 7);
    s.td.assign(1, w.td ? w.td[0] : 0.0);
    s.inv_depth.assign(w.inv_depth, w.inv_depth + w.n_points);
    s.ortho.assign(w.ortho, w.ortho + 4 * w.n_lines);
    if (w.n_relo > 0) s.relo.assign(w.relo_pose, w.relo_pose + 7);
    return s;
  }
  void store_state(const State &s, UvsWindow &dst) const {
    std::copy(s.pose.begin(), s.pose.end(), dst.pose);
    std::copy(s.sb.begin(), s.sb.end(), dst.speed_bias);
    std::copy(s.ex.begin(), s.ex.end(), dst.ex_pose);
    if (dst.td) dst.td[0] = s.td[0];
    std::copy(s.inv_depth.begin(), s.inv_depth.end(), dst.inv_depth);
    std::copy(s.ortho.begin(), s.ortho.end(), dst.ortho);
    if (lay.relo_off >= 0 && dst.relo_pose) std::copy(s.relo.begin(), s.relo.end(), dst.relo_pose);
  }

  int num_factors(int type) const {
    switch (type) {
      case F_PRIOR: return w.prior_n > 0 ? 1 : 0;
      case F_IMU: return w.n_imu;
      case F_PROJ: return w.n_proj;
      case F_LINE: return w.n_line_obs;
      case F_RELO: return w.n_relo;
      default: return w.n_vp_obs;
    }
  }

  // x+ = Plus(x, delta): PoseLocalParameterization for 7-blocks, plain addition otherwise.
  void plus(const State &s, const double *delta, State &t) const {
    t = s;
    for (int f = 0; f < lay.F; f++) {
      pose_plus(&s.pose[7 * f], &delta[lay.pose(f)], &t.pose[7 * f]);
      for (int k = 0; k < 9; k++) t.sb[9 * f + k] = s.sb[9 * f + k] + delta[lay.sb(f) + k];
    }
    if (lay.ex_off >= 0) pose_plus(s.ex.data(), &delta[lay.ex_off], t.ex.data());
    if (lay.td_off >= 0) t.td[0] = s.td[0] + delta[lay.td_off];
    if (lay.relo_off >= 0) pose_plus(s.relo.data(), &delta[lay.relo_off], t.relo.data());
    for (int k = 0; k < lay.Np; k++) t.inv_depth[k] = s.inv_depth[k] + delta[lay.point(k)];
    for (int k = 0; k < 4 * lay.Nl; k++) t.ortho[k] = s.ortho[k] + delta[lay.d + lay.Np + k];
  }

  // squared norms over the ambient (global-size) non-constant blocks, as Ceres' reduced program sees x
  double ambient_sqnorm(const State &s) const {
    double n = 0.0;
    for (double v : s.pose) n += v * v;
    for (double v : s.sb) n += v * v;
    if (lay.ex_off >= 0) for (double v : s.ex) n += v * v;
    if (lay.td_off >= 0) n += s.td[0] * s.td[0];
    for (double v : s.relo) n += v * v;
    for (double v : s.inv_depth) n += v * v;
    for (double v : s.ortho) n += v * v;
    return n;
  }
  double ambient_sqdist(const State &a, const State &b) const {
    double n = 0.0;
    auto acc = [&](const std::vector<double> &x, const std::vector<double> &y) { for (size_t i = 0; i < x.size(); i++) n += (x[i] - y[i]) * (x[i] - y[i]); };
    acc(a.pose, b.pose); acc(a.sb, b.sb);
    if (lay.ex_off >= 0) acc(a.ex, b.ex);
    if (lay.td_off >= 0) acc(a.td, b.td);
    acc(a.relo, b.relo);
    acc(a.inv_depth, b.inv_depth); acc(a.ortho, b.ortho);
    return n;
  }

  // Raw Evaluate() of non-prior factor `idx` of `type`, Ceres layout.
  void evaluate_raw(int type, int idx, const State &s, BlockEval &e, bool want_jac) const {
    auto pose_ptr = [&](int f) { return &s.pose[7 * f]; };
    auto jp = [&](int b) -> double * { return want_jac ? e.J[b] : nullptr; };
    if (type == F_PROJ) {
      const int fi = w.proj_frame_i[idx], fj = w.proj_frame_j[idx], pk = w.proj_point[idx];
      e.nr = 2; e.nb = 4;
      e.gs[0] = 7; e.off[0] = lay.pose(fi);
      e.gs[1] = 7; e.off[1] = lay.pose(fj);
      e.gs[2] = 7; e.off[2] = lay.ex_off;
      e.gs[3] = 1; e.off[3] = lay.point(pk);
      TdTerms tdt; const TdTerms *tp = nullptr;
      if (w.estimate_td) {
        e.gs[4] = 1; e.off[4] = lay.td_off; e.nb = 5;
        tdt.td = s.td[0]; tdt.td_i = w.proj_td_i[idx]; tdt.td_j = w.proj_td_j[idx];
        tdt.row_i = w.proj_row_i[idx]; tdt.row_j = w.proj_row_j[idx];
        tdt.vel_i = Vec3(w.proj_vel_i[2 * idx], w.proj_vel_i[2 * idx + 1], 0.0);
        tdt.vel_j = Vec3(w.proj_vel_j[2 * idx], w.proj_vel_j[2 * idx + 1], 0.0);
        tdt.tr = o.tr; tdt.row = o.row;
        tp = &tdt;
      }
      eval_projection(pose_ptr(fi), pose_ptr(fj), s.ex.data(), s.inv_depth[pk], vec_from(w.proj_pts_i + 3 * idx),
                      vec_from(w.proj_pts_j + 3 * idx), o.focal_length / 1.6, tp, e.r, jp(0), jp(1), jp(2), jp(3),
                      e.nb == 5 ? jp(4) : nullptr);
    } else if (type == F_RELO) {
      // ProjectionFactor(pts_i, pts_j) on {para_Pose[start], relo_Pose, para_Ex_Pose[0], para_Feature[k]} (estimator.cpp:964-970):
      // pts_i = the feature's first observation, pts_j = (match.x, match.y, 1)
      const int a = relo_anchor_[idx], fi = w.proj_frame_i[a], pk = w.relo_point[idx];
      e.nr = 2; e.nb = 4;
      e.gs[0] = 7; e.off[0] = lay.pose(fi);
      e.gs[1] = 7; e.off[1] = lay.relo_off;
      e.gs[2] = 7; e.off[2] = lay.ex_off;
      e.gs[3] = 1; e.off[3] = lay.point(pk);
      eval_projection(pose_ptr(fi), s.relo.data(), s.ex.data(), s.inv_depth[pk], vec_from(w.proj_pts_i + 3 * a),
                      vec_from(w.relo_pts_j + 3 * idx), o.focal_length / 1.6, nullptr, e.r, jp(0), jp(1), jp(2), jp(3), nullptr);
    } else if (type == F_LINE || type == F_VP) {
      const bool is_line = type == F_LINE;
      const int fj = is_line ? w.line_frame[idx] : w.vp_frame[idx];
      const int lk = is_line ? w.line_idx[idx] : w.vp_line[idx];
      e.nr = is_line ? 2 : 1; e.nb = 2;
      e.gs[0] = 7; e.off[0] = lay.pose(fj);
      e.gs[1] = 4; e.off[1] = lay.line(lk);
      Mat3 ric; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) ric(i, j) = w.line_ric[3 * i + j];
      Vec3 tic = vec_from(w.line_tic);
      if (is_line) {
        Vec3 sp(w.line_sp[2 * idx], w.line_sp[2 * idx + 1], 1.0), ep(w.line_ep[2 * idx], w.line_ep[2 * idx + 1], 1.0);
        eval_line(pose_ptr(fj), &s.ortho[4 * lk], ric, tic, sp, ep, o.line_factor, e.r, jp(0), jp(1));
      } else {
        eval_vp(pose_ptr(fj), &s.ortho[4 * lk], ric, tic, vec_from(w.vp_dir + 3 * idx), o.vp_factor, e.r, jp(0), jp(1));
      }
    } else {  // F_IMU
      const int fi = w.imu_frame_i[idx], fj = fi + 1;
      e.nr = 15; e.nb = 4;
      e.gs[0] = 7; e.off[0] = lay.pose(fi);
      e.gs[1] = 9; e.off[1] = lay.sb(fi);
      e.gs[2] = 7; e.off[2] = lay.pose(fj);
      e.gs[3] = 9; e.off[3] = lay.sb(fj);
      ImuConst c;
      c.delta_p = vec_from(w.imu_delta_p + 3 * idx);
      c.delta_v = vec_from(w.imu_delta_v + 3 * idx);
      c.lin_ba = vec_from(w.imu_lin_ba + 3 * idx);
      c.lin_bg = vec_from(w.imu_lin_bg + 3 * idx);
      const double *dq = w.imu_delta_q + 4 * idx;
      c.delta_q = Quatd(dq[3], dq[0], dq[1], dq[2]);
      c.sum_dt = w.imu_sum_dt[idx];
      c.jacobian = w.imu_jacobian + (size_t)225 * idx;
      c.sqrt_info = &imu_sqrt_info_[(size_t)225 * idx];
      eval_imu(pose_ptr(fi), &s.sb[9 * fi], pose_ptr(fj), &s.sb[9 * fj], c, Vec3(o.gravity[0], o.gravity[1], o.gravity[2]),
               e.r, jp(0), jp(1), jp(2), jp(3));
    }
    for (int b = 0; b < e.nb; b++) e.ls[b] = e.gs[b] == 7 ? 6 : e.gs[b];
  }

  double loss_scale(int type) const {
    return (type == F_PROJ || type == F_RELO) ? o.cauchy_point : (type == F_LINE ? o.cauchy_line : (type == F_VP ? o.cauchy_vp : -1.0));
  }

  // Ceres' ResidualBlock::Evaluate: raw -> tangent columns -> corrector.  J[b] becomes nr x ls[b].
  void evaluate_block(int type, int idx, const State &s, BlockEval &e, bool want_jac) const {
    evaluate_raw(type, idx, s, e, want_jac);
    double *jp[BlockEval::MAXB];
    for (int b = 0; b < e.nb; b++) {
      jp[b] = nullptr;
      if (!want_jac || e.off[b] < 0) continue;
      if (e.gs[b] != e.ls[b])  // compact in place: drop the 7th column (ascending order keeps it safe)
        for (int i = 0; i < e.nr; i++) for (int c = 0; c < e.ls[b]; c++) e.J[b][i * e.ls[b] + c] = e.J[b][i * e.gs[b] + c];
      jp[b] = e.J[b];
    }
    e.cost = apply_corrector(loss_scale(type), e.nr, e.r, e.nb, want_jac ? jp : nullptr, e.ls);
  }

  // Prior residual r = r0 + J0 dx (its Jacobian is J0 itself, column block prior_col_[b]).
  void evaluate_prior(const State &s, double *r) const {
    const double *params[64];
    for (int b = 0; b < w.prior_n_blocks; b++) {
      const int kind = w.prior_block_kind[b], id = w.prior_block_id[b];
      params[b] = kind == UVS_BLOCK_POSE ? &s.pose[7 * id] : (kind == UVS_BLOCK_SPEEDBIAS ? &s.sb[9 * id] : (kind == UVS_BLOCK_EXPOSE ? s.ex.data() : s.td.data()));
    }
    eval_prior(w.prior_n, w.prior_n_blocks, prior_gs_.data(), params, w.prior_x0, w.prior_J, w.prior_r, r, nullptr);
  }

  double total_cost(const State &s) const {
    double c = 0.0;
    BlockEval e;
    for (int t = F_IMU; t <= F_LAST; t++) for (int i = 0; i < num_factors(t); i++) { evaluate_block(t, i, s, e, false); c += e.cost; }
    if (w.prior_n > 0) {
      std::vector<double> r(w.prior_n);
      evaluate_prior(s, r.data());
      double sq = 0; for (double v : r) sq += v * v;
      c += 0.5 * sq;
    }
    return c;
  }
};

}  // namespace orc

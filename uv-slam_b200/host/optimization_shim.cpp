// optimization_shim.cpp — see optimization_shim.h.  Host glue over the C ABI; no arithmetic of the
// solve happens here.
#include "optimization_shim.h"
#include "window_io.h"

#include <algorithm>
#include <cstring>

namespace uvs_host {

GpuWindowProblem::GpuWindowProblem(int n_frames, double (*para_Pose)[7], double (*para_SpeedBias)[9], double (*para_Ex_Pose)[7],
                                   double *para_Td, double (*para_Feature)[1], double (*para_Ortho_plucker)[4])
    : n_frames_(n_frames), pose_(para_Pose), sb_(para_SpeedBias), ex_(para_Ex_Pose), td_(para_Td), feat_(para_Feature),
      ortho_(para_Ortho_plucker) {
  opts_ = shared_options();
  const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, z3[3] = {0, 0, 0};
  std::memcpy(ric_, I3, sizeof(I3)); std::memcpy(tic_, z3, sizeof(z3));
}

GpuWindowProblem::~GpuWindowProblem() {
  if (h_) uvs_destroy(h_);
}

UvsHandle *GpuWindowProblem::handle() {
  if (!h_ && uvs_create(shared_device(), &h_) != UVS_OK) h_ = nullptr;
  return h_;
}

// the td arrays are filled by addProjectionTd() only: a mix of addProjection() and addProjectionTd() calls (or
// setEstimateTd(true) with plain addProjection()) would leave them shorter than n_proj
int GpuWindowProblem::check_sizes() {
  const size_t n = p_fi_.size();
  if (estimate_td_ && (p_vi_.size() != 2 * n || p_vj_.size() != 2 * n || p_tdi_.size() != n || p_tdj_.size() != n || p_rwi_.size() != n ||
                       p_rwj_.size() != n)) {
    err_ = "estimate_td is set but not every projection factor was added with addProjectionTd()";
    return UVS_ERR_INVALID_ARG;
  }
  return UVS_OK;
}

void GpuWindowProblem::setLineExtrinsic(const double ric[9], const double tic[3]) {
  std::memcpy(ric_, ric, 72); std::memcpy(tic_, tic, 24);
}

void GpuWindowProblem::addIMU(int frame_i, const PreintegrationView &pre) {
  i_fr_.push_back(frame_i);
  i_dp_.insert(i_dp_.end(), pre.delta_p, pre.delta_p + 3); i_dq_.insert(i_dq_.end(), pre.delta_q_xyzw, pre.delta_q_xyzw + 4);
  i_dv_.insert(i_dv_.end(), pre.delta_v, pre.delta_v + 3); i_dt_.push_back(pre.sum_dt);
  i_ba_.insert(i_ba_.end(), pre.linearized_ba, pre.linearized_ba + 3); i_bg_.insert(i_bg_.end(), pre.linearized_bg, pre.linearized_bg + 3);
  i_jac_.insert(i_jac_.end(), pre.jacobian, pre.jacobian + 225); i_cov_.insert(i_cov_.end(), pre.covariance, pre.covariance + 225);
}

void GpuWindowProblem::addProjection(int fi, int fj, int k, const double pts_i[3], const double pts_j[3]) {
  p_fi_.push_back(fi); p_fj_.push_back(fj); p_pt_.push_back(k);
  p_pi_.insert(p_pi_.end(), pts_i, pts_i + 3); p_pj_.insert(p_pj_.end(), pts_j, pts_j + 3);
  n_points_ = std::max(n_points_, k + 1);
}

void GpuWindowProblem::addProjectionTd(int fi, int fj, int k, const double pts_i[3], const double pts_j[3], const double vel_i[2],
                                       const double vel_j[2], double td_i, double td_j, double row_i, double row_j) {
  addProjection(fi, fj, k, pts_i, pts_j);
  p_vi_.insert(p_vi_.end(), vel_i, vel_i + 2); p_vj_.insert(p_vj_.end(), vel_j, vel_j + 2);
  p_tdi_.push_back(td_i); p_tdj_.push_back(td_j); p_rwi_.push_back(row_i); p_rwj_.push_back(row_j);
}

void GpuWindowProblem::addLine(int fj, int k, const double sp[2], const double ep[2]) {
  l_fr_.push_back(fj); l_idx_.push_back(k);
  l_sp_.insert(l_sp_.end(), sp, sp + 2); l_ep_.insert(l_ep_.end(), ep, ep + 2);
  n_lines_ = std::max(n_lines_, k + 1);
}

void GpuWindowProblem::addVP(int fj, int k, const double vp[3]) {
  v_fr_.push_back(fj); v_idx_.push_back(k);
  v_dir_.insert(v_dir_.end(), vp, vp + 3);
  n_lines_ = std::max(n_lines_, k + 1);
}

void GpuWindowProblem::addRelocalization(int k, const double pts_j[3]) {
  r_pt_.push_back(k);
  r_pj_.insert(r_pj_.end(), pts_j, pts_j + 3);
}

UvsWindow GpuWindowProblem::view() {
  UvsWindow w{};
  w.n_frames = n_frames_; w.n_points = n_points_; w.n_lines = n_lines_;
  w.n_proj = (int32_t)p_fi_.size(); w.n_line_obs = (int32_t)l_fr_.size(); w.n_vp_obs = (int32_t)v_fr_.size(); w.n_imu = (int32_t)i_fr_.size();
  w.estimate_extrinsic = estimate_extrinsic_ ? 1 : 0; w.estimate_td = estimate_td_ ? 1 : 0;
  w.pose = &pose_[0][0]; w.speed_bias = &sb_[0][0]; w.ex_pose = &ex_[0][0]; w.td = td_;
  w.inv_depth = feat_ ? &feat_[0][0] : nullptr; w.ortho = ortho_ ? &ortho_[0][0] : nullptr;
  w.proj_frame_i = p_fi_.data(); w.proj_frame_j = p_fj_.data(); w.proj_point = p_pt_.data(); w.proj_pts_i = p_pi_.data(); w.proj_pts_j = p_pj_.data();
  w.proj_vel_i = p_vi_.data(); w.proj_vel_j = p_vj_.data(); w.proj_td_i = p_tdi_.data(); w.proj_td_j = p_tdj_.data();
  w.proj_row_i = p_rwi_.data(); w.proj_row_j = p_rwj_.data();
  w.line_frame = l_fr_.data(); w.line_idx = l_idx_.data(); w.line_sp = l_sp_.data(); w.line_ep = l_ep_.data();
  w.vp_frame = v_fr_.data(); w.vp_line = v_idx_.data(); w.vp_dir = v_dir_.data(); w.line_ric = ric_; w.line_tic = tic_;
  w.imu_frame_i = i_fr_.data(); w.imu_delta_p = i_dp_.data(); w.imu_delta_q = i_dq_.data(); w.imu_delta_v = i_dv_.data();
  w.imu_sum_dt = i_dt_.data(); w.imu_lin_ba = i_ba_.data(); w.imu_lin_bg = i_bg_.data(); w.imu_jacobian = i_jac_.data(); w.imu_covariance = i_cov_.data();
  if (prior_.valid()) {
    w.prior_n = prior_.n; w.prior_n_blocks = (int32_t)prior_.block_kind.size();
    w.prior_J = prior_.J.data(); w.prior_r = prior_.r.data(); w.prior_block_kind = prior_.block_kind.data();
    w.prior_block_id = prior_.block_id.data(); w.prior_x0 = prior_.x0.data();
  }
  if (relo_pose_ && !r_pt_.empty()) {
    w.n_relo = (int32_t)r_pt_.size(); w.relo_pose = relo_pose_; w.relo_point = r_pt_.data(); w.relo_pts_j = r_pj_.data();
  }
  return w;
}

int GpuWindowProblem::solve(UvsSummary *summary) {
  UvsHandle *h = handle();
  if (!h) { err_ = "no CUDA device (there is no CPU fallback)"; return UVS_ERR_CUDA; }
  if (int rc = check_sizes()) return rc;
  UvsWindow w = view();
  const int rc = uvs_batch_solve(h, 1, &w, &opts_, summary);
  if (rc != UVS_OK) err_ = uvs_last_error(h);
  uploaded_ = rc == UVS_OK;
  return rc;
}

int GpuWindowProblem::save(const char *path) {
  if (int rc = check_sizes()) return rc;
  const UvsWindow w = view();
  const int rc = save_window(w, path);
  if (rc) err_ = "GpuWindowProblem::save: cannot write the window";
  return rc;
}

int GpuWindowProblem::marginalize(int flag, PriorData &out) {
  UvsHandle *h = handle();
  if (!h) { err_ = "no CUDA device (there is no CPU fallback)"; return UVS_ERR_CUDA; }
  if (!uploaded_) { err_ = "marginalize() needs a solved window"; return UVS_ERR_NO_WINDOW; }
  {
    // linearise at the state the caller holds NOW (gauge-fixed and re-packed, see the header), with the current ric / tic
    UvsWindow w = view();
    const int rc = uvs_upload_state(h, 1, &w);
    if (rc != UVS_OK) { err_ = uvs_last_error(h); out.n = 0; return rc; }
  }
  const int cap_n = 16 * n_frames_ + 16, cap_b = 2 * n_frames_ + 8;
  out.J.assign((size_t)cap_n * cap_n, 0.0); out.r.assign(cap_n, 0.0); out.x0.assign((size_t)9 * cap_b, 0.0);
  out.block_kind.assign(cap_b, 0); out.block_id.assign(cap_b, 0);
  UvsPrior p{};
  p.J = out.J.data(); p.r = out.r.data(); p.x0 = out.x0.data(); p.block_kind = out.block_kind.data(); p.block_id = out.block_id.data();
  p.cap_n = cap_n; p.cap_blocks = cap_b;
  const int rc = uvs_marginalize(h, 0, flag, &p);
  if (rc != UVS_OK) { err_ = uvs_last_error(h); out.n = 0; return rc; }
  out.n = p.n;
  out.J.resize((size_t)p.n * p.n); out.r.resize(p.n);
  out.block_kind.resize(p.n_blocks); out.block_id.resize(p.n_blocks);
  size_t gs = 0;
  for (int k : out.block_kind) gs += (k == UVS_BLOCK_POSE || k == UVS_BLOCK_EXPOSE) ? 7 : (k == UVS_BLOCK_SPEEDBIAS ? 9 : 1);
  out.x0.resize(gs);
  return UVS_OK;
}

}  // namespace uvs_host

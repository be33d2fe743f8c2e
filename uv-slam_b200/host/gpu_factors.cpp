// gpu_factors.cpp — see gpu_factors.h.  Host glue only: packs one factor into a UvsWindow, calls the
// C ABI (uvs_upload_windows + uvs_eval_*, Ceres layout = raw Evaluate() output) and scatters the result
// into the caller's residual / Jacobian pointers, honouring NULL jacobians / NULL jacobians[i].
#include "gpu_factors.h"

#include <cmath>
#include <cstring>
#include <mutex>

namespace uvs_host {

namespace {
UvsHandle *g_handle = nullptr;
UvsOptions g_opts;
std::once_flag g_once, g_handle_once;
int g_device = 0;
const double kIdentityPose[7] = {0, 0, 0, 0, 0, 0, 1};

void copy_block(double **jacobians, int b, const double *src, int rows, int cols) {
  if (jacobians && jacobians[b]) std::memcpy(jacobians[b], src, sizeof(double) * rows * cols);
}
}  // namespace

UvsOptions &shared_options() {
  std::call_once(g_once, [] { uvs_default_options(&g_opts); });
  return g_opts;
}

int shared_device() { return g_device; }
void set_shared_device(int device) { g_device = device; }

// the handle of the single-factor Evaluate() drop-ins only; every GpuWindowProblem owns its own handle
UvsHandle *shared_handle() {
  shared_options();
  std::call_once(g_handle_once, [] { if (uvs_create(g_device, &g_handle) != UVS_OK) g_handle = nullptr; });
  return g_handle;
}

// ---- IMU ---------------------------------------------------------------------------------------------
bool GpuIMUFactor::Evaluate(double const *const *p, double *residuals, double **jacobians) const {
  UvsHandle *h = shared_handle();
  if (!h) return false;
  double pose[14], sb[18];
  std::memcpy(pose, p[0], 56); std::memcpy(pose + 7, p[2], 56);
  std::memcpy(sb, p[1], 72); std::memcpy(sb + 9, p[3], 72);
  const int32_t frame_i = 0;
  UvsWindow w{};
  w.n_frames = 2; w.n_imu = 1;
  w.pose = pose; w.speed_bias = sb; w.ex_pose = const_cast<double *>(kIdentityPose);
  w.imu_frame_i = &frame_i;
  w.imu_delta_p = pre_.delta_p; w.imu_delta_q = pre_.delta_q_xyzw; w.imu_delta_v = pre_.delta_v; w.imu_sum_dt = &pre_.sum_dt;
  w.imu_lin_ba = pre_.linearized_ba; w.imu_lin_bg = pre_.linearized_bg; w.imu_jacobian = pre_.jacobian; w.imu_covariance = pre_.covariance;
  if (uvs_upload_windows(h, 1, &w, &shared_options()) != UVS_OK) return false;
  double J[15 * 32];
  if (uvs_eval_imu(h, residuals, jacobians ? J : nullptr, UVS_EVAL_CERES_LAYOUT) != UVS_OK) return false;
  copy_block(jacobians, 0, J, 15, 7); copy_block(jacobians, 1, J + 105, 15, 9);
  copy_block(jacobians, 2, J + 240, 15, 7); copy_block(jacobians, 3, J + 345, 15, 9);
  return true;
}

// ---- point reprojection ------------------------------------------------------------------------------
GpuProjectionFactor::GpuProjectionFactor(const double pts_i[3], const double pts_j[3]) {
  std::memcpy(pts_i_, pts_i, 24); std::memcpy(pts_j_, pts_j, 24);
}

static bool eval_projection(const double *const *p, double *residuals, double **jacobians, const double *pts_i, const double *pts_j,
                            bool td, const double *vel_i, const double *vel_j, double td_i, double td_j, double row_i, double row_j) {
  UvsHandle *h = shared_handle();
  if (!h) return false;
  double pose[14], sb[18] = {0}, ex[7], lam = p[3][0], tdv = td ? p[4][0] : 0.0;
  std::memcpy(pose, p[0], 56); std::memcpy(pose + 7, p[1], 56); std::memcpy(ex, p[2], 56);
  const int32_t fi = 0, fj = 1, pt = 0;
  UvsWindow w{};
  w.n_frames = 2; w.n_points = 1; w.n_proj = 1; w.estimate_td = td ? 1 : 0;
  w.pose = pose; w.speed_bias = sb; w.ex_pose = ex; w.td = &tdv; w.inv_depth = &lam;
  w.proj_frame_i = &fi; w.proj_frame_j = &fj; w.proj_point = &pt; w.proj_pts_i = pts_i; w.proj_pts_j = pts_j;
  if (td) { w.proj_vel_i = vel_i; w.proj_vel_j = vel_j; w.proj_td_i = &td_i; w.proj_td_j = &td_j; w.proj_row_i = &row_i; w.proj_row_j = &row_j; }
  if (uvs_upload_windows(h, 1, &w, &shared_options()) != UVS_OK) return false;
  double J[46];
  if (uvs_eval_proj(h, residuals, jacobians ? J : nullptr, UVS_EVAL_CERES_LAYOUT) != UVS_OK) return false;
  copy_block(jacobians, 0, J, 2, 7); copy_block(jacobians, 1, J + 14, 2, 7); copy_block(jacobians, 2, J + 28, 2, 7);
  copy_block(jacobians, 3, J + 42, 2, 1);
  if (td) copy_block(jacobians, 4, J + 44, 2, 1);
  return true;
}

bool GpuProjectionFactor::Evaluate(double const *const *p, double *residuals, double **jacobians) const {
  return eval_projection(p, residuals, jacobians, pts_i_, pts_j_, false, nullptr, nullptr, 0, 0, 0, 0);
}

GpuProjectionTdFactor::GpuProjectionTdFactor(const double pts_i[3], const double pts_j[3], const double vel_i[2], const double vel_j[2],
                                             double td_i, double td_j, double row_i, double row_j)
    : td_i_(td_i), td_j_(td_j), row_i_(row_i), row_j_(row_j) {
  std::memcpy(pts_i_, pts_i, 24); std::memcpy(pts_j_, pts_j, 24); std::memcpy(vel_i_, vel_i, 16); std::memcpy(vel_j_, vel_j, 16);
}

bool GpuProjectionTdFactor::Evaluate(double const *const *p, double *residuals, double **jacobians) const {
  return eval_projection(p, residuals, jacobians, pts_i_, pts_j_, true, vel_i_, vel_j_, td_i_, td_j_, row_i_, row_j_);
}

// ---- line / vanishing point ----------------------------------------------------------------------------
GpuLineProjectionFactor::GpuLineProjectionFactor(const double ric[9], const double tic[3], const double sp[2], const double ep[2]) {
  std::memcpy(ric_, ric, 72); std::memcpy(tic_, tic, 24); std::memcpy(sp_, sp, 16); std::memcpy(ep_, ep, 16);
}

bool GpuLineProjectionFactor::Evaluate(double const *const *p, double *residuals, double **jacobians) const {
  UvsHandle *h = shared_handle();
  if (!h) return false;
  double pose[7], sb[9] = {0}, line[4];
  std::memcpy(pose, p[0], 56); std::memcpy(line, p[1], 32);
  const int32_t fr = 0, li = 0;
  UvsWindow w{};
  w.n_frames = 1; w.n_lines = 1; w.n_line_obs = 1;
  w.pose = pose; w.speed_bias = sb; w.ex_pose = const_cast<double *>(kIdentityPose); w.ortho = line;
  w.line_frame = &fr; w.line_idx = &li; w.line_sp = sp_; w.line_ep = ep_; w.line_ric = ric_; w.line_tic = tic_;
  if (uvs_upload_windows(h, 1, &w, &shared_options()) != UVS_OK) return false;
  double J[22];
  if (uvs_eval_line(h, residuals, jacobians ? J : nullptr, UVS_EVAL_CERES_LAYOUT) != UVS_OK) return false;
  copy_block(jacobians, 0, J, 2, 7); copy_block(jacobians, 1, J + 14, 2, 4);
  return true;
}

GpuVPProjectionFactor::GpuVPProjectionFactor(const double ric[9], const double tic[3], const double vp[3]) {
  std::memcpy(ric_, ric, 72); std::memcpy(tic_, tic, 24); std::memcpy(vp_, vp, 24);
}

bool GpuVPProjectionFactor::Evaluate(double const *const *p, double *residuals, double **jacobians) const {
  UvsHandle *h = shared_handle();
  if (!h) return false;
  double pose[7], sb[9] = {0}, line[4];
  std::memcpy(pose, p[0], 56); std::memcpy(line, p[1], 32);
  const int32_t fr = 0, li = 0;
  // the library pairs every VP factor with the line factor of the same observation (estimator.cpp:916-925
  // always adds both); a dummy line observation satisfies that here
  const double sp[2] = {0.0, 0.0}, ep[2] = {0.1, 0.0};
  UvsWindow w{};
  w.n_frames = 1; w.n_lines = 1; w.n_line_obs = 1; w.n_vp_obs = 1;
  w.pose = pose; w.speed_bias = sb; w.ex_pose = const_cast<double *>(kIdentityPose); w.ortho = line;
  w.line_frame = &fr; w.line_idx = &li; w.line_sp = sp; w.line_ep = ep; w.line_ric = ric_; w.line_tic = tic_;
  w.vp_frame = &fr; w.vp_line = &li; w.vp_dir = vp_;
  if (uvs_upload_windows(h, 1, &w, &shared_options()) != UVS_OK) return false;
  double J[11];
  if (uvs_eval_vp(h, residuals, jacobians ? J : nullptr, UVS_EVAL_CERES_LAYOUT) != UVS_OK) return false;
  copy_block(jacobians, 0, J, 1, 7); copy_block(jacobians, 1, J + 7, 1, 4);
  return true;
}

// ---- prior ---------------------------------------------------------------------------------------------
GpuMarginalizationFactor::GpuMarginalizationFactor(int n, const double *J0, const double *r0, const std::vector<int> &block_kind,
                                                   const double *x0)
    : n_(n), J0_(J0, J0 + (size_t)n * n), r0_(r0, r0 + n), kind_(block_kind) {
  size_t gs = 0;
  for (int k : kind_) {
    const int g = (k == UVS_BLOCK_POSE || k == UVS_BLOCK_EXPOSE) ? 7 : (k == UVS_BLOCK_SPEEDBIAS ? 9 : 1);
    mutable_parameter_block_sizes()->push_back(g);
    gs += g;
  }
  x0_.assign(x0, x0 + gs);
  set_num_residuals(n);
}

bool GpuMarginalizationFactor::Evaluate(double const *const *p, double *residuals, double **jacobians) const {
  UvsHandle *h = shared_handle();
  if (!h) return false;
  // lay the kept blocks out as a window: pose / speed-bias block k -> frame index = its rank among its kind
  const int nb = (int)kind_.size();
  std::vector<double> pose, sb;
  double ex[7] = {0, 0, 0, 0, 0, 0, 1}, tdv = 0.0;
  std::vector<int32_t> ids(nb, 0), kinds(kind_.begin(), kind_.end());
  bool has_ex = false, has_td = false;
  for (int b = 0; b < nb; b++) {
    if (kind_[b] == UVS_BLOCK_POSE) { ids[b] = (int32_t)(pose.size() / 7); pose.insert(pose.end(), p[b], p[b] + 7); }
    else if (kind_[b] == UVS_BLOCK_SPEEDBIAS) { ids[b] = (int32_t)(sb.size() / 9); sb.insert(sb.end(), p[b], p[b] + 9); }
    else if (kind_[b] == UVS_BLOCK_EXPOSE) { std::memcpy(ex, p[b], 56); has_ex = true; }
    else { tdv = p[b][0]; has_td = true; }
  }
  const int F = (int)std::max(pose.size() / 7, sb.size() / 9);
  pose.resize((size_t)std::max(F, 1) * 7, 0.0); sb.resize((size_t)std::max(F, 1) * 9, 0.0);
  for (int f = 0; f < std::max(F, 1); f++) if (pose[7 * f + 6] == 0.0 && pose[7 * f + 3] == 0.0 && pose[7 * f + 4] == 0.0 && pose[7 * f + 5] == 0.0) pose[7 * f + 6] = 1.0;
  UvsWindow w{};
  w.n_frames = std::max(F, 1); w.prior_n = n_; w.prior_n_blocks = nb;
  w.estimate_extrinsic = has_ex ? 1 : 0; w.estimate_td = 0; (void)has_td;
  w.pose = pose.data(); w.speed_bias = sb.data(); w.ex_pose = ex; w.td = &tdv;
  w.prior_J = J0_.data(); w.prior_r = r0_.data(); w.prior_block_kind = kinds.data(); w.prior_block_id = ids.data(); w.prior_x0 = x0_.data();
  if (uvs_upload_windows(h, 1, &w, &shared_options()) != UVS_OK) return false;
  size_t cols = 0;
  for (int g : parameter_block_sizes()) cols += g;
  std::vector<double> J(jacobians ? (size_t)n_ * cols : 0);
  if (uvs_eval_prior(h, residuals, jacobians ? J.data() : nullptr, UVS_EVAL_CERES_LAYOUT) != UVS_OK) return false;
  size_t off = 0;
  for (int b = 0; b < nb; b++) {
    const int g = parameter_block_sizes()[b];
    copy_block(jacobians, b, J.data() + off, n_, g);
    off += (size_t)n_ * g;
  }
  return true;
}

// ---- pose manifold (pose_local_parameterization.cpp:3-27) -------------------------------------------------
bool PoseLocalParameterization::Plus(const double *x, const double *d, double *out) const {
  for (int k = 0; k < 3; k++) out[k] = x[k] + d[k];
  const double ax = x[3], ay = x[4], az = x[5], aw = x[6];
  const double bx = d[3] / 2.0, by = d[4] / 2.0, bz = d[5] / 2.0, bw = 1.0;   // Utility::deltaQ (utility.h:11-24)
  double q[4] = {aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                 aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz};
  const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int k = 0; k < 4; k++) out[3 + k] = q[k] / n;
  return true;
}

bool PoseLocalParameterization::ComputeJacobian(const double *, double *J) const {
  for (int i = 0; i < 7; i++) for (int j = 0; j < 6; j++) J[i * 6 + j] = (i == j) ? 1.0 : 0.0;
  return true;
}

}  // namespace uvs_host

// ORACLE — TEST INFRASTRUCTURE ONLY (see smallmat.h).  PARITY UNPINNED at the Ceres/Eigen boundary.
//
// factors.h: CPU restatement of the cost functions on UV-SLAM's sliding-window hot path.  Every
// function reproduces `Evaluate()` of the cited reference class: raw residuals and row-major
// num_residuals x global_size Jacobians, before any loss correction.
#pragma once
#include <cmath>
#include <limits>
#include <algorithm>
#include <cstring>
#include <vector>

#include "jet.h"
#include "smallmat.h"

namespace orc {

typedef V3<double> Vec3;
typedef M3<double> Mat3;
typedef Quat<double> Quatd;

inline Quatd quat_from_block(const double *b) { return Quatd(b[6], b[3], b[4], b[5]); }  // [p, qx qy qz qw]
inline Vec3 vec_from(const double *b) { return Vec3(b[0], b[1], b[2]); }

// ------------------------------------------------------------------------------------------------
// a15 / a11: ceres::CauchyLoss and the corrector restated in ResidualBlockInfo::Evaluate
// (factor/marginalization_factor.cpp:37-68).  rho = {rho, rho', rho''}.
inline void cauchy_loss(double a, double s, double rho[3]) {
  const double b = a * a, c = 1.0 / b;
  const double sum = 1.0 + s * c;
  const double inv = 1.0 / sum;
  rho[0] = b * std::log(sum);
  rho[1] = inv > std::numeric_limits<double>::min() ? inv : std::numeric_limits<double>::min();
  rho[2] = -c * (inv * inv);
}

// Applies the correction in place to residuals r[nr] and jacobian blocks (row-major nr x cols[i]).
// Returns the cost contribution 1/2 rho(s).  loss_a <= 0 means "no loss function" (cost = s/2).
inline double apply_corrector(double loss_a, int nr, double *r, int nblocks, double **J, const int *cols) {
  double sq_norm = 0.0;
  for (int i = 0; i < nr; i++) sq_norm += r[i] * r[i];
  if (loss_a <= 0.0) return 0.5 * sq_norm;
  double rho[3];
  cauchy_loss(loss_a, sq_norm, rho);
  const double sqrt_rho1 = std::sqrt(rho[1]);
  double residual_scaling, alpha_sq_norm;
  if (sq_norm == 0.0 || rho[2] <= 0.0) {
    residual_scaling = sqrt_rho1;
    alpha_sq_norm = 0.0;
  } else {
    const double D = 1.0 + 2.0 * sq_norm * rho[2] / rho[1];
    const double alpha = 1.0 - std::sqrt(D);
    residual_scaling = sqrt_rho1 / (1 - alpha);
    alpha_sq_norm = alpha / sq_norm;
  }
  for (int b = 0; b < nblocks; b++) {
    if (!J || !J[b]) continue;
    const int nc = cols[b];
    if (alpha_sq_norm == 0.0) {
      for (int k = 0; k < nr * nc; k++) J[b][k] *= sqrt_rho1;
    } else {
      // J = sqrt_rho1 * (J - alpha_sq_norm * r * (r^T J))
      for (int c = 0; c < nc; c++) {
        double rtj = 0.0;
        for (int i = 0; i < nr; i++) rtj += r[i] * J[b][i * nc + c];
        for (int i = 0; i < nr; i++) J[b][i * nc + c] = sqrt_rho1 * (J[b][i * nc + c] - alpha_sq_norm * r[i] * rtj);
      }
    }
  }
  for (int i = 0; i < nr; i++) r[i] *= residual_scaling;
  return 0.5 * rho[0];
}

// ------------------------------------------------------------------------------------------------
// a5: ProjectionFactor::Evaluate, factor/projection_factor.cpp:22-175 (UNIT_SPHERE_ERROR undefined,
// parameters.h:17).  sqrt_info = S * I2 with S = FOCAL_LENGTH / 1.6 (estimator.cpp:17).
// a6: ProjectionTdFactor::Evaluate, factor/projection_td_factor.cpp:34-145, when `td` != nullptr.
struct TdTerms {
  double td;             // parameters[4][0]
  double td_i, td_j;     // cur_td at observation time
  double row_i, row_j;   // raw uv.y
  Vec3 vel_i, vel_j;     // z = 0
  double tr, row;        // TR, ROW
};

inline void eval_projection(const double *pose_i, const double *pose_j, const double *ex, double inv_dep_i,
                            const Vec3 &pts_i_in, const Vec3 &pts_j_in, double S, const TdTerms *tdt,
                            double *residuals, double *J_pi, double *J_pj, double *J_ex, double *J_feat,
                            double *J_td) {
  Vec3 Pi = vec_from(pose_i), Pj = vec_from(pose_j), tic = vec_from(ex);
  Quatd Qi = quat_from_block(pose_i), Qj = quat_from_block(pose_j), qic = quat_from_block(ex);

  Vec3 pts_i = pts_i_in, pts_j = pts_j_in;
  if (tdt) {  // projection_td_factor.cpp:18-19, 50-52
    const double row_i = tdt->row_i - tdt->row / 2, row_j = tdt->row_j - tdt->row / 2;
    pts_i = pts_i_in - (tdt->td - tdt->td_i + tdt->tr / tdt->row * row_i) * tdt->vel_i;
    pts_j = pts_j_in - (tdt->td - tdt->td_j + tdt->tr / tdt->row * row_j) * tdt->vel_j;
  }
  Vec3 pts_camera_i = pts_i / inv_dep_i;
  Vec3 pts_imu_i = rotate(qic, pts_camera_i) + tic;
  Vec3 pts_w = rotate(Qi, pts_imu_i) + Pi;
  Vec3 pts_imu_j = rotate(inverse(Qj), pts_w - Pj);
  Vec3 pts_camera_j = rotate(inverse(qic), pts_imu_j - tic);

  const double dep_j = pts_camera_j.z;
  residuals[0] = S * ((pts_camera_j.x / dep_j) - pts_j.x);
  residuals[1] = S * ((pts_camera_j.y / dep_j) - pts_j.y);

  if (!(J_pi || J_pj || J_ex || J_feat || J_td)) return;
  Mat3 Ri = toRotationMatrix(Qi), Rj = toRotationMatrix(Qj), ric = toRotationMatrix(qic);
  double reduce[2][3] = {{S * (1. / dep_j), 0.0, S * (-pts_camera_j.x / (dep_j * dep_j))},
                         {0.0, S * (1. / dep_j), S * (-pts_camera_j.y / (dep_j * dep_j))}};
  auto reduce_mul = [&](const Mat3 &M, double out[2][3]) {
    for (int r = 0; r < 2; r++)
      for (int c = 0; c < 3; c++) out[r][c] = reduce[r][0] * M(0, c) + reduce[r][1] * M(1, c) + reduce[r][2] * M(2, c);
  };
  auto store_pose = [&](double *J, const Mat3 &left, const Mat3 &right) {
    double a[2][3], b[2][3];
    reduce_mul(left, a);
    reduce_mul(right, b);
    for (int r = 0; r < 2; r++) {
      for (int c = 0; c < 3; c++) { J[r * 7 + c] = a[r][c]; J[r * 7 + 3 + c] = b[r][c]; }
      J[r * 7 + 6] = 0.0;
    }
  };
  Mat3 ricT = transpose(ric), RjT = transpose(Rj);
  if (J_pi) store_pose(J_pi, ricT * RjT, ricT * RjT * Ri * (-skew(pts_imu_i)));
  if (J_pj) store_pose(J_pj, ricT * (-RjT), ricT * skew(pts_imu_j));
  Mat3 tmp_r = ricT * RjT * Ri * ric;
  if (J_ex) {
    Mat3 left = ricT * (RjT * Ri - Mat3::Identity());
    Mat3 right = (-tmp_r) * skew(pts_camera_i) + skew(tmp_r * pts_camera_i) +
                 skew(ricT * (RjT * (Ri * tic + Pi - Pj) - tic));
    store_pose(J_ex, left, right);
  }
  if (J_feat) {
    Vec3 v = (tmp_r * pts_i) * (-1.0 / (inv_dep_i * inv_dep_i));
    for (int r = 0; r < 2; r++) J_feat[r] = reduce[r][0] * v.x + reduce[r][1] * v.y + reduce[r][2] * v.z;
  }
  if (J_td && tdt) {  // projection_td_factor.cpp:135-140
    Vec3 v = (tmp_r * tdt->vel_i) * (-1.0 / inv_dep_i);
    J_td[0] = reduce[0][0] * v.x + reduce[0][1] * v.y + reduce[0][2] * v.z + S * tdt->vel_j.x;
    J_td[1] = reduce[1][0] * v.x + reduce[1][1] * v.y + reduce[1][2] * v.z + S * tdt->vel_j.y;
  }
}

// ------------------------------------------------------------------------------------------------
// a7/a8: LineProjectionFactor::operator()<T> (factor/line_projection_factor.h:16-60) and
// VPProjectionFactor::operator()<T> (factor/vp_projection_factor.h:19-66): the shared transform.
template <class T>
inline void line_to_camera(const T *pose, const T *line, const Mat3 &ric, const Vec3 &tic, V3<T> &n_c, V3<T> &d_c) {
  using std::cos; using std::sin;
  const V3<T> t_wb(pose[0], pose[1], pose[2]);
  const Quat<T> q_wb(pose[6], pose[3], pose[4], pose[5]);
  const Quat<T> roll = fromAngleAxis<T>(line[0], 0), pitch = fromAngleAxis<T>(line[1], 1), yaw = fromAngleAxis<T>(line[2], 2);
  const T pi = line[3];

  M3<T> ricT_;  // ric.cast<T>()
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) ricT_(i, j) = T(ric(i, j));
  V3<T> ticT(T(tic.x), T(tic.y), T(tic.z));

  M3<T> R_wc = toRotationMatrix(q_wb) * ricT_;   // Quaternion * Matrix3 -> toRotationMatrix() * m
  V3<T> t_wc = rotate(q_wb, ticT) + t_wb;        // Quaternion * Vector3 -> _transformVector
  M3<T> Rotation_psi = toRotationMatrix((roll * pitch) * yaw);

  V3<T> n_w = cos(pi) * Rotation_psi.col(0);
  V3<T> d_w = sin(pi) * Rotation_psi.col(1);

  M3<T> R_cw = transpose(R_wc);
  V3<T> t_cw = -(R_cw * t_wc);
  M3<T> t_cw_ss = skew(t_cw);
  M3<T> upper_right = t_cw_ss * R_cw;
  n_c = R_cw * n_w + upper_right * d_w;
  d_c = R_cw * d_w;
}

template <class T>
inline void line_functor(const T *pose, const T *line, const Mat3 &ric, const Vec3 &tic, const Vec3 &sp, const Vec3 &ep,
                         double line_factor, T *residuals) {
  using std::sqrt; using std::pow;
  V3<T> n_c, d_c;
  line_to_camera(pose, line, ric, tic, n_c, d_c);
  V3<T> spT(T(sp.x), T(sp.y), T(sp.z)), epT(T(ep.x), T(ep.y), T(ep.z));
  residuals[0] = T(line_factor) * dot(spT, n_c) / sqrt(pow(n_c[0], 2) + pow(n_c[1], 2));
  residuals[1] = T(line_factor) * dot(epT, n_c) / sqrt(pow(n_c[0], 2) + pow(n_c[1], 2));
}

template <class T>
inline void vp_functor(const T *pose, const T *line, const Mat3 &ric, const Vec3 &tic, const Vec3 &vp,
                       double vp_factor, T *residuals) {
  using std::sqrt; using std::acos; using std::abs;
  V3<T> n_c, d_c;
  line_to_camera(pose, line, ric, tic, n_c, d_c);
  V3<T> vp3(T(vp.x), T(vp.y), T(vp.z));
  T d_norm = sqrt(dot(d_c, d_c)), vp_norm = sqrt(dot(vp3, vp3));
  residuals[0] = T(vp_factor) * acos(abs(dot(d_c, vp3) / (d_norm * vp_norm)));
}

// ceres::AutoDiffCostFunction<F, NR, 7, 4>::Evaluate: Jets over the 11 raw inputs.
inline void eval_line(const double *pose, const double *line, const Mat3 &ric, const Vec3 &tic, const Vec3 &sp,
                      const Vec3 &ep, double line_factor, double *residuals, double *J_pose /*2x7*/, double *J_line /*2x4*/) {
  if (!J_pose && !J_line) { line_functor<double>(pose, line, ric, tic, sp, ep, line_factor, residuals); return; }
  typedef Jet<11> J11;
  J11 p[7], l[4], r[2];
  for (int i = 0; i < 7; i++) p[i] = J11(pose[i], i);
  for (int i = 0; i < 4; i++) l[i] = J11(line[i], 7 + i);
  line_functor<J11>(p, l, ric, tic, sp, ep, line_factor, r);
  for (int k = 0; k < 2; k++) {
    residuals[k] = r[k].a;
    if (J_pose) for (int i = 0; i < 7; i++) J_pose[k * 7 + i] = r[k].v[i];
    if (J_line) for (int i = 0; i < 4; i++) J_line[k * 4 + i] = r[k].v[7 + i];
  }
}
inline void eval_vp(const double *pose, const double *line, const Mat3 &ric, const Vec3 &tic, const Vec3 &vp,
                    double vp_factor, double *residuals, double *J_pose /*1x7*/, double *J_line /*1x4*/) {
  if (!J_pose && !J_line) { vp_functor<double>(pose, line, ric, tic, vp, vp_factor, residuals); return; }
  typedef Jet<11> J11;
  J11 p[7], l[4], r[1];
  for (int i = 0; i < 7; i++) p[i] = J11(pose[i], i);
  for (int i = 0; i < 4; i++) l[i] = J11(line[i], 7 + i);
  vp_functor<J11>(p, l, ric, tic, vp, vp_factor, r);
  residuals[0] = r[0].a;
  if (J_pose) for (int i = 0; i < 7; i++) J_pose[i] = r[0].v[i];
  if (J_line) for (int i = 0; i < 4; i++) J_line[i] = r[0].v[7 + i];
}

// ------------------------------------------------------------------------------------------------
// Dense helpers for the 15x15 IMU algebra (row-major).
// General inverse by LU with partial pivoting, then Cholesky: restates
//   Eigen::LLT<Matrix15d>(covariance.inverse()).matrixL().transpose()      imu_factor.h:64
inline bool invert_general(int n, const double *A, double *Ainv) {
  std::vector<double> lu(A, A + n * n);
  std::vector<int> piv(n);
  for (int k = 0; k < n; k++) {
    int p = k; double best = std::fabs(lu[k * n + k]);
    for (int i = k + 1; i < n; i++) if (std::fabs(lu[i * n + k]) > best) { best = std::fabs(lu[i * n + k]); p = i; }
    piv[k] = p;
    if (best == 0.0) return false;
    if (p != k) for (int j = 0; j < n; j++) std::swap(lu[k * n + j], lu[p * n + j]);
    for (int i = k + 1; i < n; i++) {
      lu[i * n + k] /= lu[k * n + k];
      const double f = lu[i * n + k];
      for (int j = k + 1; j < n; j++) lu[i * n + j] -= f * lu[k * n + j];
    }
  }
  for (int c = 0; c < n; c++) {
    std::vector<double> x(n, 0.0);
    x[c] = 1.0;
    for (int k = 0; k < n; k++) if (piv[k] != k) std::swap(x[k], x[piv[k]]);
    for (int i = 0; i < n; i++) { double s = x[i]; for (int j = 0; j < i; j++) s -= lu[i * n + j] * x[j]; x[i] = s; }
    for (int i = n - 1; i >= 0; i--) { double s = x[i]; for (int j = i + 1; j < n; j++) s -= lu[i * n + j] * x[j]; x[i] = s / lu[i * n + i]; }
    for (int i = 0; i < n; i++) Ainv[i * n + c] = x[i];
  }
  return true;
}
// Lower Cholesky factor L (row-major, upper part zeroed) reading the lower triangle of A, as Eigen LLT<Lower>.
inline bool cholesky_lower(int n, const double *A, double *L) {
  std::fill(L, L + n * n, 0.0);
  for (int j = 0; j < n; j++) {
    double d = A[j * n + j];
    for (int k = 0; k < j; k++) d -= L[j * n + k] * L[j * n + k];
    if (!(d > 0.0)) return false;
    d = std::sqrt(d);
    L[j * n + j] = d;
    for (int i = j + 1; i < n; i++) {
      double s = A[i * n + j];
      for (int k = 0; k < j; k++) s -= L[i * n + k] * L[j * n + k];
      L[i * n + j] = s / d;
    }
  }
  return true;
}
inline bool imu_sqrt_info(const double *cov /*15x15*/, double *sqrt_info /*15x15 upper*/) {
  double inv[225], L[225];
  if (!invert_general(15, cov, inv)) return false;
  if (!cholesky_lower(15, inv, L)) return false;
  for (int i = 0; i < 15; i++) for (int j = 0; j < 15; j++) sqrt_info[i * 15 + j] = L[j * 15 + i];
  return true;
}

struct ImuConst {     // the members of IntegrationBase the factor reads (integration_base.h:188-207)
  Vec3 delta_p, delta_v, lin_ba, lin_bg;
  Quatd delta_q;
  double sum_dt;
  const double *jacobian;    // 15x15 row-major
  const double *sqrt_info;   // 15x15, precomputed by imu_sqrt_info (constant during a solve, imu_factor.h:52-58)
};
enum { O_P = 0, O_R = 3, O_V = 6, O_BA = 9, O_BG = 12 };  // parameters.h:59-66

inline Mat3 block33(const double *M15, int r, int c) {
  Mat3 B; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) B(i, j) = M15[(r + i) * 15 + c + j]; return B;
}

// a3: IntegrationBase::evaluate (integration_base.h:160-186) + IMUFactor::Evaluate (imu_factor.h:19-182)
inline void eval_imu(const double *pose_i, const double *sb_i, const double *pose_j, const double *sb_j,
                     const ImuConst &c, const Vec3 &G, double *residuals /*15*/,
                     double *J_pi /*15x7*/, double *J_sbi /*15x9*/, double *J_pj /*15x7*/, double *J_sbj /*15x9*/) {
  Vec3 Pi = vec_from(pose_i), Pj = vec_from(pose_j);
  Quatd Qi = quat_from_block(pose_i), Qj = quat_from_block(pose_j);
  Vec3 Vi(sb_i[0], sb_i[1], sb_i[2]), Bai(sb_i[3], sb_i[4], sb_i[5]), Bgi(sb_i[6], sb_i[7], sb_i[8]);
  Vec3 Vj(sb_j[0], sb_j[1], sb_j[2]), Baj(sb_j[3], sb_j[4], sb_j[5]), Bgj(sb_j[6], sb_j[7], sb_j[8]);

  Mat3 dp_dba = block33(c.jacobian, O_P, O_BA), dp_dbg = block33(c.jacobian, O_P, O_BG);
  Mat3 dq_dbg = block33(c.jacobian, O_R, O_BG);
  Mat3 dv_dba = block33(c.jacobian, O_V, O_BA), dv_dbg = block33(c.jacobian, O_V, O_BG);
  const double sum_dt = c.sum_dt;

  Vec3 dba = Bai - c.lin_ba, dbg = Bgi - c.lin_bg;
  Quatd corrected_delta_q = c.delta_q * deltaQ(dq_dbg * dbg);
  Vec3 corrected_delta_v = c.delta_v + dv_dba * dba + dv_dbg * dbg;
  Vec3 corrected_delta_p = c.delta_p + dp_dba * dba + dp_dbg * dbg;

  Quatd Qi_inv = inverse(Qi);
  double raw[15];
  Vec3 rp = rotate(Qi_inv, 0.5 * G * sum_dt * sum_dt + Pj - Pi - Vi * sum_dt) - corrected_delta_p;
  Vec3 rq = 2.0 * (inverse(corrected_delta_q) * (Qi_inv * Qj)).vec();
  Vec3 rv = rotate(Qi_inv, G * sum_dt + Vj - Vi) - corrected_delta_v;
  Vec3 rba = Baj - Bai, rbg = Bgj - Bgi;
  for (int k = 0; k < 3; k++) { raw[O_P + k] = rp[k]; raw[O_R + k] = rq[k]; raw[O_V + k] = rv[k]; raw[O_BA + k] = rba[k]; raw[O_BG + k] = rbg[k]; }
  const double *SI = c.sqrt_info;
  for (int i = 0; i < 15; i++) { double s = 0.0; for (int k = 0; k < 15; k++) s += SI[i * 15 + k] * raw[k]; residuals[i] = s; }

  auto put = [](double *J, int ncols, int r0, int c0, const Mat3 &B) {
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) J[(r0 + i) * ncols + c0 + j] = B(i, j);
  };
  auto premul = [&](double *J, int ncols) {  // J = sqrt_info * J
    double t[15 * 9];
    for (int i = 0; i < 15; i++)
      for (int j = 0; j < ncols; j++) { double s = 0.0; for (int k = i; k < 15; k++) s += SI[i * 15 + k] * J[k * ncols + j]; t[i * ncols + j] = s; }
    std::memcpy(J, t, sizeof(double) * 15 * ncols);
  };
  Mat3 RiT = toRotationMatrix(Qi_inv);
  if (J_pi) {
    std::fill(J_pi, J_pi + 15 * 7, 0.0);
    put(J_pi, 7, O_P, O_P, -RiT);
    put(J_pi, 7, O_P, O_R, skew(rotate(Qi_inv, 0.5 * G * sum_dt * sum_dt + Pj - Pi - Vi * sum_dt)));
    // -(Qleft(Qj^-1 * Qi) * Qright(corrected_delta_q)).bottomRightCorner<3,3>()     imu_factor.h:100-101
    double L4[4][4], R4[4][4];
    Qleft4(inverse(Qj) * Qi, L4);
    Qright4(corrected_delta_q, R4);
    Mat3 br;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double s = 0.0; for (int k = 0; k < 4; k++) s += L4[i + 1][k] * R4[k][j + 1]; br(i, j) = -s; }
    put(J_pi, 7, O_R, O_R, br);
    put(J_pi, 7, O_V, O_R, skew(rotate(Qi_inv, G * sum_dt + Vj - Vi)));
    premul(J_pi, 7);
  }
  if (J_sbi) {
    std::fill(J_sbi, J_sbi + 15 * 9, 0.0);
    put(J_sbi, 9, O_P, O_V - O_V, -(RiT * sum_dt));
    put(J_sbi, 9, O_P, O_BA - O_V, -dp_dba);
    put(J_sbi, 9, O_P, O_BG - O_V, -dp_dbg);
    // -Qleft(Qj^-1 * Qi * delta_q).bottomRightCorner<3,3>() * dq_dbg   (delta_q, not corrected)   imu_factor.h:128
    put(J_sbi, 9, O_R, O_BG - O_V, -(QleftBR((inverse(Qj) * Qi) * c.delta_q) * dq_dbg));
    put(J_sbi, 9, O_V, O_V - O_V, -RiT);
    put(J_sbi, 9, O_V, O_BA - O_V, -dv_dba);
    put(J_sbi, 9, O_V, O_BG - O_V, -dv_dbg);
    put(J_sbi, 9, O_BA, O_BA - O_V, -Mat3::Identity());
    put(J_sbi, 9, O_BG, O_BG - O_V, -Mat3::Identity());
    premul(J_sbi, 9);
  }
  if (J_pj) {
    std::fill(J_pj, J_pj + 15 * 7, 0.0);
    put(J_pj, 7, O_P, O_P, RiT);
    put(J_pj, 7, O_R, O_R, QleftBR((inverse(corrected_delta_q) * Qi_inv) * Qj));
    premul(J_pj, 7);
  }
  if (J_sbj) {
    std::fill(J_sbj, J_sbj + 15 * 9, 0.0);
    put(J_sbj, 9, O_V, O_V - O_V, RiT);
    put(J_sbj, 9, O_BA, O_BA - O_V, Mat3::Identity());
    put(J_sbj, 9, O_BG, O_BG - O_V, Mat3::Identity());
    premul(J_sbj, 9);
  }
}

// a4: IntegrationBase::{propagate,midPointIntegration} (integration_base.h:54-158): one IMU sample.
struct Preintegration {
  Vec3 acc_0, gyr_0, lin_ba, lin_bg, delta_p, delta_v;
  Quatd delta_q;
  double sum_dt;
  double jacobian[225], covariance[225];
  double noise_diag[18];
  Preintegration(const Vec3 &a0, const Vec3 &g0, const Vec3 &ba, const Vec3 &bg, double acc_n, double gyr_n, double acc_w, double gyr_w)
      : acc_0(a0), gyr_0(g0), lin_ba(ba), lin_bg(bg), sum_dt(0.0) {
    std::fill(jacobian, jacobian + 225, 0.0);
    std::fill(covariance, covariance + 225, 0.0);
    for (int i = 0; i < 15; i++) jacobian[i * 15 + i] = 1.0;
    const double nd[6] = {acc_n * acc_n, gyr_n * gyr_n, acc_n * acc_n, gyr_n * gyr_n, acc_w * acc_w, gyr_w * gyr_w};
    for (int b = 0; b < 6; b++) for (int k = 0; k < 3; k++) noise_diag[3 * b + k] = nd[b];
  }
  void push_back(double dt, const Vec3 &acc_1, const Vec3 &gyr_1) {
    Vec3 un_acc_0 = rotate(delta_q, acc_0 - lin_ba);
    Vec3 un_gyr = 0.5 * (gyr_0 + gyr_1) - lin_bg;
    Quatd result_delta_q = delta_q * Quatd(1, un_gyr.x * dt / 2, un_gyr.y * dt / 2, un_gyr.z * dt / 2);
    Vec3 un_acc_1 = rotate(result_delta_q, acc_1 - lin_ba);
    Vec3 un_acc = 0.5 * (un_acc_0 + un_acc_1);
    Vec3 result_delta_p = delta_p + delta_v * dt + 0.5 * un_acc * dt * dt;
    Vec3 result_delta_v = delta_v + un_acc * dt;

    Vec3 w_x = un_gyr, a_0_x = acc_0 - lin_ba, a_1_x = acc_1 - lin_ba;
    Mat3 R_w_x = skew(w_x), R_a_0_x = skew(a_0_x), R_a_1_x = skew(a_1_x);
    Mat3 Rq = toRotationMatrix(delta_q), Rr = toRotationMatrix(result_delta_q), I = Mat3::Identity();
    double F[225] = {0}, V[15 * 18] = {0};
    auto putF = [&](int r, int c, const Mat3 &B) { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) F[(r + i) * 15 + c + j] = B(i, j); };
    auto putV = [&](int r, int c, const Mat3 &B) { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) V[(r + i) * 18 + c + j] = B(i, j); };
    Mat3 ImW = I - R_w_x * dt;
    putF(0, 0, I);
    putF(0, 3, (Rq * R_a_0_x) * (-0.25 * dt * dt) + ((Rr * R_a_1_x) * ImW) * (-0.25 * dt * dt));
    putF(0, 6, I * dt);
    putF(0, 9, (Rq + Rr) * (-0.25 * dt * dt));
    putF(0, 12, (Rr * R_a_1_x) * (-0.25 * dt * dt * -dt));
    putF(3, 3, ImW);
    putF(3, 12, I * (-1.0 * dt));
    putF(6, 3, (Rq * R_a_0_x) * (-0.5 * dt) + ((Rr * R_a_1_x) * ImW) * (-0.5 * dt));
    putF(6, 6, I);
    putF(6, 9, (Rq + Rr) * (-0.5 * dt));
    putF(6, 12, (Rr * R_a_1_x) * (-0.5 * dt * -dt));
    putF(9, 9, I);
    putF(12, 12, I);
    Mat3 V03 = ((-Rr) * R_a_1_x) * (0.25 * dt * dt * 0.5 * dt);
    Mat3 V63 = ((-Rr) * R_a_1_x) * (0.5 * dt * 0.5 * dt);
    putV(0, 0, Rq * (0.25 * dt * dt));
    putV(0, 3, V03);
    putV(0, 6, Rr * (0.25 * dt * dt));
    putV(0, 9, V03);
    putV(3, 3, I * (0.5 * dt));
    putV(3, 9, I * (0.5 * dt));
    putV(6, 0, Rq * (0.5 * dt));
    putV(6, 3, V63);
    putV(6, 6, Rr * (0.5 * dt));
    putV(6, 9, V63);
    putV(9, 12, I * dt);
    putV(12, 15, I * dt);
    // jacobian = F * jacobian; covariance = F cov F^T + V noise V^T
    double t[225], t2[225];
    for (int i = 0; i < 15; i++) for (int j = 0; j < 15; j++) { double s = 0; for (int k = 0; k < 15; k++) s += F[i * 15 + k] * jacobian[k * 15 + j]; t[i * 15 + j] = s; }
    std::memcpy(jacobian, t, sizeof(t));
    for (int i = 0; i < 15; i++) for (int j = 0; j < 15; j++) { double s = 0; for (int k = 0; k < 15; k++) s += F[i * 15 + k] * covariance[k * 15 + j]; t[i * 15 + j] = s; }
    for (int i = 0; i < 15; i++) for (int j = 0; j < 15; j++) { double s = 0; for (int k = 0; k < 15; k++) s += t[i * 15 + k] * F[j * 15 + k]; t2[i * 15 + j] = s; }
    for (int i = 0; i < 15; i++) for (int j = 0; j < 15; j++) { double s = 0; for (int k = 0; k < 18; k++) s += V[i * 18 + k] * noise_diag[k] * V[j * 18 + k]; covariance[i * 15 + j] = t2[i * 15 + j] + s; }

    delta_p = result_delta_p;
    delta_q = normalized(result_delta_q);
    delta_v = result_delta_v;
    sum_dt += dt;
    acc_0 = acc_1;
    gyr_0 = gyr_1;
  }
};

// ------------------------------------------------------------------------------------------------
// a10: MarginalizationFactor::Evaluate (factor/marginalization_factor.cpp:333-381).
// Blocks are the kept blocks in column order; global sizes gs[b] in {7,9,1}.
inline void eval_prior(int n, int nblocks, const int *gs, const double *const *params, const double *x0,
                       const double *J0 /*n x n*/, const double *r0, double *residuals, double **jacobians) {
  std::vector<double> dx(n, 0.0);
  int idx = 0, x0off = 0;
  for (int b = 0; b < nblocks; b++) {
    const int size = gs[b];
    const double *x = params[b], *xz = x0 + x0off;
    if (size != 7) {
      for (int k = 0; k < size; k++) dx[idx + k] = x[k] - xz[k];
      idx += size;
    } else {
      for (int k = 0; k < 3; k++) dx[idx + k] = x[k] - xz[k];
      Quatd dq = inverse(quat_from_block(xz)) * quat_from_block(x);
      Vec3 v = 2.0 * dq.vec();
      if (!(dq.w >= 0)) v = -v;
      for (int k = 0; k < 3; k++) dx[idx + 3 + k] = v[k];
      idx += 6;
    }
    x0off += size;
  }
  for (int i = 0; i < n; i++) { double s = r0[i]; for (int k = 0; k < n; k++) s += J0[i * n + k] * dx[k]; residuals[i] = s; }
  if (jacobians) {
    idx = 0;
    for (int b = 0; b < nblocks; b++) {
      const int size = gs[b], local = size == 7 ? 6 : size;
      if (jacobians[b]) {
        for (int i = 0; i < n; i++) {
          for (int k = 0; k < size; k++) jacobians[b][i * size + k] = 0.0;
          for (int k = 0; k < local; k++) jacobians[b][i * size + k] = J0[i * n + idx + k];
        }
      }
      idx += local;
    }
  }
}

// a9: PoseLocalParameterization::Plus (factor/pose_local_parameterization.cpp:3-19)
inline void pose_plus(const double *x, const double *delta, double *out) {
  for (int k = 0; k < 3; k++) out[k] = x[k] + delta[k];
  Quatd q = normalized(quat_from_block(x) * deltaQ(Vec3(delta[3], delta[4], delta[5])));
  out[3] = q.x; out[4] = q.y; out[5] = q.z; out[6] = q.w;
}

}  // namespace orc

// gpu_factors.h — drop-in cost functions with the reference's exact Evaluate() signatures, forwarding
// to libuvs_b200.so (C ABI, include/uvs.h) with a batch of one factor.
//
//   reference class (vins_estimator/src/factor/)              replacement
//   IMUFactor : SizedCostFunction<15,7,9,7,9>  imu_factor.h:12          GpuIMUFactor
//   ProjectionFactor : SizedCostFunction<2,7,7,7,1>  projection_factor.h:12   GpuProjectionFactor
//   ProjectionTdFactor : SizedCostFunction<2,7,7,7,1,1>  projection_td_factor.h:10   GpuProjectionTdFactor
//   AutoDiffCostFunction<LineProjectionFactor,2,7,4>  estimator.cpp:916      GpuLineProjectionFactor
//   AutoDiffCostFunction<VPProjectionFactor,1,7,4>    estimator.cpp:923      GpuVPProjectionFactor
//   MarginalizationFactor  marginalization_factor.h:74-81                    GpuMarginalizationFactor
//
// A batch of one is the literal drop-in (and what the signature-level parity tests use); the fast path
// is the batched solve behind gpu_optimization() (optimization_shim.h).  Evaluate() returns false when
// the device call fails - there is no CPU fallback.
#pragma once
#include <vector>

#include "../../include/uvs.h"
#include "../../include/uvs/ceres_compat.h"

namespace uvs_host {

// One process-wide handle for the single-factor Evaluate() calls (the reference calls Evaluate() from one thread,
// ceres num_threads = 1, estimator.cpp:985 commented out); created once (std::call_once), not re-entrant.  A
// GpuWindowProblem never uses it: each problem owns a handle, so an Evaluate() call cannot replace its device batch.
UvsHandle *shared_handle();
int shared_device();                 // CUDA device of the handles created by this library (default 0)
void set_shared_device(int device);  // call before the first factor / problem is used
UvsOptions &shared_options();   // FOCAL_LENGTH, G, LINE_FACTOR, VP_FACTOR, TR, ROW of parameters.h:11-47

struct PreintegrationView {   // the members of IntegrationBase the factor reads (integration_base.h:188-207)
  const double *delta_p, *delta_q_xyzw, *delta_v;
  double sum_dt;
  const double *linearized_ba, *linearized_bg;
  const double *jacobian, *covariance;   // 15x15 row-major
};

class GpuIMUFactor : public ceres::SizedCostFunction<15, 7, 9, 7, 9> {
 public:
  explicit GpuIMUFactor(const PreintegrationView &pre) : pre_(pre) {}
  bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const override;
 private:
  PreintegrationView pre_;
};

class GpuProjectionFactor : public ceres::SizedCostFunction<2, 7, 7, 7, 1> {
 public:
  GpuProjectionFactor(const double pts_i[3], const double pts_j[3]);
  bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const override;
 private:
  double pts_i_[3], pts_j_[3];
};

class GpuProjectionTdFactor : public ceres::SizedCostFunction<2, 7, 7, 7, 1, 1> {
 public:
  GpuProjectionTdFactor(const double pts_i[3], const double pts_j[3], const double vel_i[2], const double vel_j[2], double td_i,
                        double td_j, double row_i, double row_j);
  bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const override;
 private:
  double pts_i_[3], pts_j_[3], vel_i_[2], vel_j_[2], td_i_, td_j_, row_i_, row_j_;
};

class GpuLineProjectionFactor : public ceres::SizedCostFunction<2, 7, 4> {
 public:
  GpuLineProjectionFactor(const double ric_rowmajor[9], const double tic[3], const double sp[2], const double ep[2]);
  bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const override;
 private:
  double ric_[9], tic_[3], sp_[2], ep_[2];
};

class GpuVPProjectionFactor : public ceres::SizedCostFunction<1, 7, 4> {
 public:
  GpuVPProjectionFactor(const double ric_rowmajor[9], const double tic[3], const double vp[3]);
  bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const override;
 private:
  double ric_[9], tic_[3], vp_[3];
};

// Prior described as the C ABI does: J0 (n x n), r0, kept blocks (kind, global size implied) and x0.
class GpuMarginalizationFactor : public ceres::CostFunction {
 public:
  GpuMarginalizationFactor(int n, const double *J0, const double *r0, const std::vector<int> &block_kind, const double *x0);
  bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const override;
 private:
  int n_;
  std::vector<double> J0_, r0_, x0_;
  std::vector<int> kind_;
};

// PoseLocalParameterization (pose_local_parameterization.h): Plus = p + dp, q (x) deltaQ(dtheta) normalised;
// ComputeJacobian = [I6; 0].  Pure host arithmetic (7 numbers), kept for interface completeness.
class PoseLocalParameterization : public ceres::LocalParameterization {
 public:
  bool Plus(const double *x, const double *delta, double *x_plus_delta) const override;
  bool ComputeJacobian(const double *x, double *jacobian) const override;
  int GlobalSize() const override { return 7; }
  int LocalSize() const override { return 6; }
};

}  // namespace uvs_host

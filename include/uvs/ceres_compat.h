/*
 * ceres_compat.h — the slice of the Ceres interface the hot path is written against, for builds
 * where Ceres itself is not installed (this image has no Ceres; the reference links it through
 * find_package(Ceres REQUIRED), vins_estimator/CMakeLists.txt:22).
 *
 * Signatures are the ones the reference's factors override:
 *   bool CostFunction::Evaluate(double const* const* parameters, double* residuals,
 *                               double** jacobians) const        (imu_factor.h:19, projection_factor.h:17,
 *                                                                  marginalization_factor.h:79)
 * jacobians may be NULL, any jacobians[i] may be NULL, each is row-major num_residuals x block size.
 * Against a real Ceres, define UVS_USE_REAL_CERES and include <ceres/ceres.h> first: the Gpu*Factor
 * classes of uv-slam_b200/host/gpu_factors.h then derive from ceres::CostFunction unchanged.
 */
#ifndef UVS_CERES_COMPAT_H_
#define UVS_CERES_COMPAT_H_

#ifdef UVS_USE_REAL_CERES
#include <ceres/ceres.h>
#else
#include <cstdint>
#include <vector>

namespace ceres {

class CostFunction {
 public:
  virtual ~CostFunction() {}
  virtual bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const = 0;
  const std::vector<int32_t> &parameter_block_sizes() const { return parameter_block_sizes_; }
  int num_residuals() const { return num_residuals_; }

 protected:
  std::vector<int32_t> *mutable_parameter_block_sizes() { return &parameter_block_sizes_; }
  void set_num_residuals(int n) { num_residuals_ = n; }

 private:
  std::vector<int32_t> parameter_block_sizes_;
  int num_residuals_ = 0;
};

template <int kNumResiduals, int... Ns>
class SizedCostFunction : public CostFunction {
 public:
  SizedCostFunction() {
    set_num_residuals(kNumResiduals);
    *mutable_parameter_block_sizes() = std::vector<int32_t>{Ns...};
  }
};

class LossFunction {
 public:
  virtual ~LossFunction() {}
  virtual void Evaluate(double sq_norm, double out[3]) const = 0;
};

class LocalParameterization {
 public:
  virtual ~LocalParameterization() {}
  virtual bool Plus(const double *x, const double *delta, double *x_plus_delta) const = 0;
  virtual bool ComputeJacobian(const double *x, double *jacobian) const = 0;
  virtual int GlobalSize() const = 0;
  virtual int LocalSize() const = 0;
};

}  // namespace ceres
#endif /* UVS_USE_REAL_CERES */
#endif /* UVS_CERES_COMPAT_H_ */

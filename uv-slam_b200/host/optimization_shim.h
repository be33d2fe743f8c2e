// optimization_shim.h — the call surface that replaces the body of Estimator::optimization()
// (vins_estimator/src/estimator.cpp:761-1233) between vector2double() and double2vector().
//
// The reference builds a ceres::Problem by calling problem.AddResidualBlock(...) inside its loops over
// pre_integrations / f_manager.feature / f_manager.line_feature (:811-934) and then ceres::Solve(:994).
// GpuWindowProblem keeps those loops: every AddResidualBlock call becomes one add*() call with the same
// data, solve() replaces ceres::Solve (results land in the same para_* arrays), marginalize() replaces the
// MarginalizationInfo block (:1003-1228).  See INTEGRATION.md for the patched optimization().
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/uvs.h"
#include "gpu_factors.h"

namespace uvs_host {

struct PriorData {            // what Estimator keeps in last_marginalization_info / _parameter_blocks
  int n = 0;
  std::vector<double> J, r, x0;
  std::vector<int32_t> block_kind, block_id;
  bool valid() const { return n > 0; }
};

class GpuWindowProblem {
 public:
  // para_* are the Estimator's own arrays (estimator.h:114-121); n_frames = WINDOW_SIZE + 1
  GpuWindowProblem(int n_frames, double (*para_Pose)[7], double (*para_SpeedBias)[9], double (*para_Ex_Pose)[7], double *para_Td,
                   double (*para_Feature)[1], double (*para_Ortho_plucker)[4]);
  void setOptions(const UvsOptions &o) { opts_ = o; }
  UvsOptions &options() { return opts_; }
  void setEstimateExtrinsic(bool e) { estimate_extrinsic_ = e; }      // estimator.cpp:781-790
  void setEstimateTd(bool e) { estimate_td_ = e; }                    // :846
  void setLineExtrinsic(const double ric_rowmajor[9], const double tic[3]);   // ric[0], tic[0] frozen into the functors (:917)

  void setPrior(const PriorData &p) { prior_ = p; }                   // :803-809
  void addIMU(int frame_i, const PreintegrationView &pre);            // :811-818 (caller skips sum_dt > 10)
  void addProjection(int frame_i, int frame_j, int feature_index, const double pts_i[3], const double pts_j[3]);   // :857-862
  void addProjectionTd(int frame_i, int frame_j, int feature_index, const double pts_i[3], const double pts_j[3], const double vel_i[2],
                       const double vel_j[2], double td_i, double td_j, double row_i, double row_j);              // :846-855
  void addLine(int frame_j, int line_index, const double sp[2], const double ep[2]);                              // :916-918
  void addVP(int frame_j, int line_index, const double vp[3]);                                                    // :920-925
  // relocalisation (estimator.cpp:944-978): relo_Pose becomes a parameter block of this solve (updated in place), one factor
  // per matched feature on {para_Pose[start], relo_Pose, para_Ex_Pose[0], para_Feature[feature_index]}; pts_i is taken from
  // the feature's own factors.  Add the matches in ascending feature_index (the order of the reference's loop).
  void setRelocalization(double relo_Pose[7]) { relo_pose_ = relo_Pose; }
  void addRelocalization(int feature_index, const double pts_j[3]);

  ~GpuWindowProblem();
  GpuWindowProblem(const GpuWindowProblem &) = delete;
  GpuWindowProblem &operator=(const GpuWindowProblem &) = delete;

  // ceres::Solve replacement.  Returns the UvsStatus; parameters are updated in place.
  int solve(UvsSummary *summary = nullptr);
  // Next prior (flag = marginalization_flag); returns UvsStatus, out.n == 0 when the reference would build nothing.
  // The prior is linearised at the CURRENT contents of the para_* arrays, not at the device's raw solver state: the
  // reference runs double2vector() (yaw / position gauge fix, estimator.cpp:596-711) and vector2double() again
  // (:1006 / :1168) between ceres::Solve and the marginalization, so call this after those two, exactly where the
  // reference builds its MarginalizationInfo.  The state (and ric / tic of the line functors) is re-uploaded first.
  int marginalize(int flag, PriorData &out);
  // dump hook: the assembled window in the `uvs_window v1` format of tests/ and bench.py - call it
  // before solve() to record what the reference would hand to ceres::Solve.  Returns the UvsStatus.
  int save(const char *path);
  const std::string &lastError() const { return err_; }

 private:
  int n_frames_;
  double (*pose_)[7]; double (*sb_)[9]; double (*ex_)[7]; double *td_; double (*feat_)[1]; double (*ortho_)[4];
  bool estimate_extrinsic_ = false, estimate_td_ = false, uploaded_ = false;
  UvsOptions opts_;
  double ric_[9], tic_[3];
  PriorData prior_;
  int n_points_ = 0, n_lines_ = 0;
  std::vector<int32_t> p_fi_, p_fj_, p_pt_, l_fr_, l_idx_, v_fr_, v_idx_, i_fr_;
  std::vector<double> p_pi_, p_pj_, p_vi_, p_vj_, p_tdi_, p_tdj_, p_rwi_, p_rwj_, l_sp_, l_ep_, v_dir_;
  std::vector<double> i_dp_, i_dq_, i_dv_, i_dt_, i_ba_, i_bg_, i_jac_, i_cov_;
  double *relo_pose_ = nullptr;
  std::vector<int32_t> r_pt_;
  std::vector<double> r_pj_;
  std::string err_;
  UvsHandle *h_ = nullptr;   // this problem's own device batch: nothing another object uploads can get between solve() and marginalize()
  UvsHandle *handle();
  int check_sizes();
  UvsWindow view();
};

}  // namespace uvs_host

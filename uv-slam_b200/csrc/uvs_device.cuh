// uvs_device.cuh — HBM layout of a batch of sliding windows and kernel launch prototypes.
//
// All windows of a batch are concatenated ("flat over the batch") so that every sweep kernel is one
// launch over thousands of factors, whatever the window boundaries.  The caller's arrays are copied
// verbatim (one pinned staging buffer, one H2D copy); k_prep then derives the global index records.
//
//   state      pose[nF][7] sb[nF][9] ex[B][7] td[B] inv_depth[nP] ortho[nL][4], double-buffered:
//              cur[w] selects the buffer holding window w's current iterate, the other one holds the
//              LM candidate, so accepting a step is a one-int flip (no copy).
//   records    one AoS record per factor in LOCAL layout (tangent columns, loss-corrected):
//                proj  [r(2) | Ji 2x6 | Jj 2x6 | Jex 2x6 | Jl 2 (| Jtd 2)]  = 40 (42) doubles
//                line  [r(2) | Jpose 2x6 | Jline 2x4]                        = 22
//                vp    [r(1) | Jpose 6 | Jline 4]                            = 11
//                imu   [r(15) | J 15x30]  columns pose_i(6) sb_i(9) pose_j(6) sb_j(9) = 465
//              factors of one landmark are contiguous, so a landmark's records are one contiguous
//              chunk for the Schur stage.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/uvs.h"

namespace uvs {

constexpr int REC_PROJ = 40, REC_PROJ_TD = 42, REC_LINE = 22, REC_VP = 11, REC_IMU = 465;
// raw Evaluate() records in Ceres layout (7-wide pose blocks): [r | J...]
constexpr int CREC_PROJ = 46, CREC_PROJ_TD = 48, CREC_LINE = 24, CREC_VP = 12;

constexpr int WF_EXTRINSIC = 1, WF_TD = 2;
constexpr int RANGE_UNSET = 0x7f7f7f7f;   // cudaMemset(0x7f) pattern of pt_begin / ln_begin before k_prep
// per-window solver status bits (Dev::win_state)
constexpr int WS_ACTIVE = 1, WS_NEED_JAC = 2, WS_STEP_OK = 4;

struct Params {
  double S;                 // FOCAL_LENGTH / 1.6
  double g[3];
  double line_factor, vp_factor;
  double cauchy_point, cauchy_line, cauchy_vp;
  double tr_over_row, half_row;
  double min_lm_diag, max_lm_diag, min_relative_decrease, max_radius, min_radius, initial_radius;
  double function_tolerance, gradient_tolerance, parameter_tolerance;
  int fixed_iterations;
  int max_num_iterations;
};

// Per-window LM bookkeeping, one struct per window in device memory.
struct WinCtl {
  double cost, radius, decrease_factor, x_norm;
  int state;        // WS_* bits
  int iter;         // iterations logged so far (including iteration 0)
  int n_success, n_invalid, termination, status;
  int have_scale;   // Jacobi scaling fixed (from the first Jacobian)
  int pad;
};

// Per-window accumulators of one LM iteration: ACC_STRIDE doubles per window, summed over ranks in
// the factor-parallel multi-GPU mode (the per-rank gradient maxima go to separate slots so that a
// SUM all-reduce transports them).
constexpr int ACC_STRIDE = 16;
constexpr int ACC_CAND_COST = 0, ACC_MODEL = 1, ACC_STEP2 = 2, ACC_XNORM2 = 3, ACC_FAIL = 4, ACC_COST0 = 5, ACC_GMAX = 8;
constexpr int MAX_RANKS = 8;

struct Dev {
  int B;
  int nF, nP, nL, nProj, nLobs, nVobs, nImu, nCam, nPriorR, nPriorBlk;
  int estimate_td;          // uniform over the batch
  int rank, nranks;         // factor-parallel multi-GPU: landmark k is owned by rank k % nranks
  int max_frames, max_lines, max_lobs;   // largest per-window counts of the batch (grids of the window-chunk kernels)
  // per-window tables [B+1]
  const int *frame_off, *point_off, *line_off, *proj_off, *lobs_off, *vobs_off, *imu_off, *cam_off, *prior_off,
      *pblk_off;
  const long long *S_off, *priorJ_off;
  const int *win_flags;     // [B] WF_*
  // state, double buffered
  double *pose[2], *sb[2], *ex[2], *td[2], *inv_depth[2], *ortho[2];
  int *cur;                 // [B]
  WinCtl *ctl;              // [B]
  double *acc;              // [B][ACC_STRIDE]
  UvsSummary *summary;      // [B]
  // caller arrays, concatenated over the batch
  const int *proj_fi, *proj_fj, *proj_pt;
  const double *proj_pts_i, *proj_pts_j;   // [nProj][3]
  const double *proj_vel_i, *proj_vel_j;   // [nProj][2]
  const double *proj_td_i, *proj_td_j, *proj_row_i, *proj_row_j;
  const int *line_frame, *line_idx;
  const double *line_sp, *line_ep;         // [nLobs][2]
  const int *vp_frame, *vp_line;
  const double *vp_dir;                    // [nVobs][3]
  const double *ric, *tic;                 // [B][9], [B][3]
  const int *imu_frame;
  const double *imu_dp, *imu_dq, *imu_dv, *imu_sum_dt, *imu_lin_ba, *imu_lin_bg, *imu_jac, *imu_cov;
  const double *prior_J, *prior_r0, *prior_x0;  // prior_x0: [nPriorBlk][9] padded
  const int *pblk_kind, *pblk_id;               // [nPriorBlk]
  // derived by k_prep
  int4 *proj_idx;           // {pose row i, pose row j, global point, window}
  int4 *line_idx4;          // {pose row, global line, window, vp obs of the same (frame,line) or -1}
  int4 *vp_idx4;            // {pose row, global line, window, line obs of the same (frame,line)}
  int2 *imu_idx;            // {pose row i, window}
  int *pt_begin, *pt_end;   // [nP] proj factor range of a point
  int *ln_begin, *ln_end;   // [nL] line obs range of a line
  int *pt_win, *ln_win;     // [nP], [nL]
  const int *fr_win;        // [nF] window of a frame (host-built)
  double *ftab[2];          // [nF][48] per-frame tables of the line / VP factors for state buffer 0 / 1 (uvs_linefast.cuh)
  double *lsc[2];           // [nL][8] sin, cos of the four orthonormal line parameters, per state buffer
  int *pt_order;            // [32 nPW] points in the processing order of the fused point kernel (k_prep_point_order), -1 = empty lane
  const int *pw_off;        // [B+1] warp slots of every window in pt_order
  int nPW;
  int *pblk_col, *pblk_cam, *pblk_row;   // column in J0, tangent offset in the window (-1 const), state row
  double *imu_sqrt_info;    // [nImu][225]
  double *chain_lw;         // [B][chain_lw_stride] L_w blocks of the chain-mode reduced solve (k_chol_chain)
  long long chain_lw_stride;
  double *chol_frag;        // [B][chol_frag_stride] fragment-order factor of the large-window reduced solve (d > packed limit, k_chol mode 0)
  long long chol_frag_stride;
  double *imu_comp;         // [IMU_COMP + 15 = 123][nImu] compact blocks of the unweighted IMU Jacobians + unweighted residual (k_imu_geom -> k_imu_weight)
  double *prior_H;          // [sum n^2]  J0^T J0 (constant during a solve)
  int *err;                 // [1] validation flag
  // records
  double *rec_proj, *rec_line, *rec_vp, *rec_imu, *rec_prior;  // rec_prior: [nPriorR] residuals
  // solver buffers
  double *scale_cam, *scale_pt, *scale_ln;   // Jacobi scaling [nCam], [nP], [nL][4]
  double *colsq_cam, *colsq_pt, *colsq_ln;   // squared column norms of the current Jacobian (unscaled)
  double *Smat, *gS, *gfull, *delta_cam;     // [sum d^2], [nCam] x3
  double *delta_pt, *delta_ln;               // [nP], [nL][4]
};

}  // namespace uvs

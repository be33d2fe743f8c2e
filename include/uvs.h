/*
 * uvs.h — C ABI of libuvs_b200.so: the B200-native replacement for the sliding-window
 * backend solve of UV-SLAM (vins_estimator).
 *
 * Every entry point is `extern "C"`, takes plain pointers / sizes / POD structs and returns an
 * int status (0 = ok, negative = UvsStatus error).  No C++ or torch types cross this boundary.
 * All host buffers are caller-owned; all device memory lives behind the opaque UvsHandle.
 *
 * Reference interfaces replaced (paths relative to the UV-SLAM checkout):
 *   void Estimator::optimization()                           vins_estimator/src/estimator.h:52,
 *                                                            estimator.cpp:761-1233
 *   bool ceres::CostFunction::Evaluate(double const* const*, double*, double**) const
 *        IMUFactor                                           factor/imu_factor.h:12,19
 *        ProjectionFactor                                    factor/projection_factor.h:12,17
 *        ProjectionTdFactor                                  factor/projection_td_factor.h:10
 *        AutoDiffCostFunction<LineProjectionFactor,2,7,4>    estimator.cpp:916-918
 *        AutoDiffCostFunction<VPProjectionFactor,1,7,4>      estimator.cpp:923-925
 *        MarginalizationFactor                               factor/marginalization_factor.h:74-81
 *   MarginalizationInfo::{preMarginalize,marginalize}        factor/marginalization_factor.cpp:110-297
 *   ceres::Solve (SPARSE_SCHUR + LEVENBERG_MARQUARDT)        estimator.cpp:982-994
 *
 * Layout conventions (identical to the reference's parameter blocks, estimator.cpp:526-594):
 *   pose block        [px,py,pz,qx,qy,qz,qw]            7 doubles, tangent size 6
 *   speed-bias block  [v(3), ba(3), bg(3)]              9 doubles
 *   inverse depth     [lambda]                          1 double
 *   line block        [psi_x, psi_y, psi_z, phi]        4 doubles (orthonormal Pluecker)
 * Jacobians in "Ceres layout" are row-major num_residuals x global_block_size, blocks of one
 * factor concatenated in parameter-block order; "local layout" keeps only the tangent columns
 * (first 6 of every 7-wide pose block, pose_local_parameterization.cpp:20-27).
 */
#ifndef UVS_H_
#define UVS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UVS_ABI_VERSION 2

typedef enum UvsStatus {
  UVS_OK = 0,
  UVS_ERR_INVALID_ARG = -1,
  UVS_ERR_CUDA = -2,
  UVS_ERR_CAPACITY = -3,
  UVS_ERR_NOT_FINITE = -4,      /* non-finite cost / step */
  UVS_ERR_NOT_PD = -5,          /* reduced camera system not positive definite */
  UVS_ERR_NO_WINDOW = -6,       /* call needs an uploaded window */
  UVS_ERR_COMM = -7,            /* multi-GPU exchange failed */
  UVS_ERR_UNSUPPORTED = -8
} UvsStatus;

/* kinds of parameter block a marginalization prior can refer to */
typedef enum UvsBlockKind {
  UVS_BLOCK_POSE = 0,       /* id = frame index, global 7 / local 6 */
  UVS_BLOCK_SPEEDBIAS = 1,  /* id = frame index, 9 */
  UVS_BLOCK_EXPOSE = 2,     /* id = 0, global 7 / local 6 */
  UVS_BLOCK_TD = 3          /* id = 0, 1 */
} UvsBlockKind;

/* marginalization_flag of the reference (estimator.h:26-30) */
typedef enum UvsMarginFlag { UVS_MARGIN_OLD = 0, UVS_MARGIN_SECOND_NEW = 1 } UvsMarginFlag;

/*
 * One sliding window ("uvs_window v1").  Pointers address caller-owned host memory.
 * State arrays are read by uvs_upload_windows and written by uvs_download_state.
 */
typedef struct UvsWindow {
  int32_t n_frames;           /* WINDOW_SIZE + 1 (11 in the reference, parameters.h:12) */
  int32_t n_points;           /* eligible point features  (estimator.cpp:826-829) */
  int32_t n_lines;            /* eligible line features   (estimator.cpp:873-878) */
  int32_t n_proj;             /* point reprojection factors */
  int32_t n_line_obs;         /* line reprojection factors */
  int32_t n_vp_obs;           /* vanishing-point factors */
  int32_t n_imu;              /* IMU factors */
  int32_t prior_n;            /* residual dimension of the marginalization prior, 0 = none */
  int32_t prior_n_blocks;     /* number of kept parameter blocks of the prior */
  int32_t estimate_extrinsic; /* 0: ex_pose constant (estimator.cpp:786-790) */
  int32_t estimate_td;        /* 0: ProjectionFactor, 1: ProjectionTdFactor (estimator.cpp:846-863) */
  int32_t reserved0;

  /* ---- state (parameter blocks) ---- */
  double *pose;        /* [n_frames][7] */
  double *speed_bias;  /* [n_frames][9] */
  double *ex_pose;     /* [7] */
  double *td;          /* [1] */
  double *inv_depth;   /* [n_points] */
  double *ortho;       /* [n_lines][4] */

  /* ---- point reprojection factors (projection_factor.h:12) ---- */
  const int32_t *proj_frame_i; /* [n_proj] start frame of the feature */
  const int32_t *proj_frame_j; /* [n_proj] observing frame, != frame_i */
  const int32_t *proj_point;   /* [n_proj] index into inv_depth */
  const double *proj_pts_i;    /* [n_proj][3] normalised image point in frame i (z = 1) */
  const double *proj_pts_j;    /* [n_proj][3] */
  /* time-offset variant only (projection_td_factor.h:10); may be NULL when !estimate_td */
  const double *proj_vel_i;    /* [n_proj][2] */
  const double *proj_vel_j;    /* [n_proj][2] */
  const double *proj_td_i;     /* [n_proj] td at which pts_i was taken */
  const double *proj_td_j;     /* [n_proj] */
  const double *proj_row_i;    /* [n_proj] raw image row uv.y (ROW/2 is subtracted inside) */
  const double *proj_row_j;    /* [n_proj] */

  /* ---- line reprojection factors (line_projection_factor.h:11-68) ---- */
  const int32_t *line_frame;   /* [n_line_obs] observing frame */
  const int32_t *line_idx;     /* [n_line_obs] index into ortho */
  const double *line_sp;       /* [n_line_obs][2] start point, z = 1 implied */
  const double *line_ep;       /* [n_line_obs][2] end point */
  /* ---- vanishing-point factors (vp_projection_factor.h:14-73) ---- */
  const int32_t *vp_frame;     /* [n_vp_obs] */
  const int32_t *vp_line;      /* [n_vp_obs] */
  const double *vp_dir;        /* [n_vp_obs][3] vp = (x/z, y/z, 1) */
  /* extrinsics frozen into the line / VP functors at construction (estimator.cpp:917,924) */
  const double *line_ric;      /* [9] row-major 3x3 */
  const double *line_tic;      /* [3] */

  /* ---- IMU factors: constants of IntegrationBase (integration_base.h:188-207) ---- */
  const int32_t *imu_frame_i;  /* [n_imu] factor links frame_i and frame_i + 1 */
  const double *imu_delta_p;   /* [n_imu][3] */
  const double *imu_delta_q;   /* [n_imu][4]  x,y,z,w */
  const double *imu_delta_v;   /* [n_imu][3] */
  const double *imu_sum_dt;    /* [n_imu] */
  const double *imu_lin_ba;    /* [n_imu][3] linearized_ba */
  const double *imu_lin_bg;    /* [n_imu][3] linearized_bg */
  const double *imu_jacobian;  /* [n_imu][15*15] row-major */
  const double *imu_covariance;/* [n_imu][15*15] row-major */

  /* ---- marginalization prior (marginalization_factor.h:52-71) ---- */
  const double *prior_J;           /* [prior_n][prior_n] row-major linearized_jacobians */
  const double *prior_r;           /* [prior_n] linearized_residuals */
  const int32_t *prior_block_kind; /* [prior_n_blocks] UvsBlockKind, in column order */
  const int32_t *prior_block_id;   /* [prior_n_blocks] */
  const double *prior_x0;          /* global-size linearisation points, concatenated */

  /* ---- relocalisation factors (estimator.cpp:944-978; ABI version 2) ----
   * With relocalization_info set, optimization() adds the parameter block relo_Pose (7, PoseLocalParameterization) and,
   * for every eligible feature matched in the loop-closure frame, one ProjectionFactor(pts_i, pts_j) on the blocks
   * {para_Pose[start_frame], relo_Pose, para_Ex_Pose[0], para_Feature[k]} with the point loss; pts_i is the feature's
   * first observation (as in its other factors), pts_j = (match.x, match.y, 1).  The marginalization does not see them
   * (estimator.cpp:1003-1228 builds its own list).  n_relo = 0: none.  Not combinable with estimate_td. */
  int32_t n_relo;
  int32_t reserved1;
  double *relo_pose;               /* [7] in: initial value (the loop frame's pose), out: uvs_download_state */
  const int32_t *relo_point;       /* [n_relo] index into inv_depth; the point must own at least one proj factor; ascending */
  const double *relo_pts_j;        /* [n_relo][3] */
} UvsWindow;

/* Globals of parameters.h:11-47 that the hot path reads, plus Ceres' trust-region defaults. */
typedef struct UvsOptions {
  double focal_length;       /* FOCAL_LENGTH; sqrt_info = focal/1.6 * I2 (estimator.cpp:17) */
  double gravity[3];         /* G */
  double line_factor;        /* LINE_FACTOR */
  double vp_factor;          /* VP_FACTOR */
  double cauchy_point;       /* CauchyLoss scale a for points (estimator.cpp:765) */
  double cauchy_line;        /* (estimator.cpp:768) */
  double cauchy_vp;          /* (estimator.cpp:772) */
  double tr;                 /* TR rolling-shutter read-out time */
  double row;                /* ROW image height */
  /* ceres::Solver::Options, defaults unless the reference sets them (estimator.cpp:982-991) */
  int32_t max_num_iterations;    /* NUM_ITERATIONS */
  int32_t fixed_iterations;      /* 1: disable convergence exits (benchmark mode) */
  double max_solver_time;        /* seconds; <= 0 disables the wall-clock cap */
  double initial_radius;         /* 1e4 */
  double max_radius;             /* 1e16 */
  double min_radius;             /* 1e-32 */
  double min_relative_decrease;  /* 1e-3 */
  double min_lm_diagonal;        /* 1e-6 */
  double max_lm_diagonal;        /* 1e32 */
  double function_tolerance;     /* 1e-6 */
  double gradient_tolerance;     /* 1e-10 */
  double parameter_tolerance;    /* 1e-8 */
} UvsOptions;

#define UVS_MAX_ITER_LOG 64

typedef enum UvsTermination {
  UVS_TERM_NO_CONVERGENCE = 0,   /* iteration cap */
  UVS_TERM_FUNCTION_TOL = 1,
  UVS_TERM_PARAMETER_TOL = 2,
  UVS_TERM_GRADIENT_TOL = 3,
  UVS_TERM_MIN_RADIUS = 4,
  UVS_TERM_FAILURE = 5,
  UVS_TERM_TIME = 6
} UvsTermination;

/* Counterpart of ceres::Solver::Summary for one window. */
typedef struct UvsSummary {
  int32_t num_iterations;          /* summary.iterations.size(): includes iteration 0 */
  int32_t num_successful_steps;
  int32_t termination;             /* UvsTermination */
  int32_t status;                  /* UvsStatus of this window */
  double initial_cost;
  double final_cost;
  double cost[UVS_MAX_ITER_LOG];             /* cost after each iteration */
  double radius[UVS_MAX_ITER_LOG];           /* trust-region radius after each iteration */
  double relative_decrease[UVS_MAX_ITER_LOG];
  double step_norm[UVS_MAX_ITER_LOG];
  double gradient_max_norm[UVS_MAX_ITER_LOG];
  int32_t step_accepted[UVS_MAX_ITER_LOG];
} UvsSummary;

/* Next marginalization prior, caller-allocated (see uvs_marginalize_size). */
typedef struct UvsPrior {
  int32_t n;                 /* out: residual dimension */
  int32_t n_blocks;          /* out */
  int32_t m;                 /* out: marginalised dimension */
  int32_t reserved0;
  double *J;                 /* [cap_n][cap_n] row-major, first n*n used */
  double *r;                 /* [cap_n] */
  int32_t *block_kind;       /* [cap_blocks] */
  int32_t *block_id;         /* [cap_blocks] ids AFTER the window shift (estimator.cpp:1139-1153) */
  double *x0;                /* [7*cap_blocks] */
  double *A;                 /* optional [cap_n][cap_n]: Schur complement A' (nullable) */
  double *b;                 /* optional [cap_n]: b' (nullable) */
  int32_t cap_n;
  int32_t cap_blocks;
} UvsPrior;

/* uvs_eval_* flags */
#define UVS_EVAL_CERES_LAYOUT 0x0  /* raw Evaluate() output, global-size Jacobian blocks */
#define UVS_EVAL_LOCAL_LAYOUT 0x1  /* tangent columns only, loss-corrected (what the solver eats) */
#define UVS_EVAL_DEVICE_OUT   0x2  /* output pointers are device pointers */

typedef struct UvsHandle UvsHandle;

int uvs_abi_version(void);
void uvs_default_options(UvsOptions *opts);
const char *uvs_status_string(int status);

/* Create a solver bound to CUDA device `device` with its own stream.  Fails (UVS_ERR_CUDA) when no
 * GPU is present: there is no CPU fallback.
 *
 * Threading and lifetime.  A handle is NOT thread-safe: all calls on one handle must come from one thread at a time (the
 * reference calls optimization() under its m_estimator lock, estimator_node.cpp:352).  Different handles are independent
 * (own stream, own device and pinned arenas) and may be used from different threads concurrently; one process may hold
 * handles on several devices.  A handle owns ONE device batch: every upload replaces it, and uvs_marginalize / uvs_eval_* /
 * uvs_download_* always refer to the batch uploaded last through THAT handle - give every independent problem its own
 * handle (uvs_host::GpuWindowProblem does).  uvs_batch_solve_pipelined runs its sub-batches on child handles that the
 * parent creates once and keeps; it starts one short-lived host thread per sub-batch and joins them before it returns, so
 * the call itself is synchronous.  Every entry point returns only after its results are in the caller's buffers
 * (the library synchronises its stream); caller-owned arrays are never referenced after the call returns. */
int uvs_create(int device, UvsHandle **out);
int uvs_destroy(UvsHandle *h);
const char *uvs_last_error(const UvsHandle *h);

/* Copy `n_windows` windows (H2D) into the handle's device batch; replaces any previous batch. */
int uvs_upload_windows(UvsHandle *h, int32_t n_windows, const UvsWindow *windows,
                       const UvsOptions *opts);
/* Copy the current state of every window back into the caller's state arrays (D2H). */
int uvs_download_state(UvsHandle *h, int32_t n_windows, UvsWindow *windows);

/* Replace the STATE of the uploaded batch (pose, speed_bias, ex_pose, td, inv_depth, ortho and the line_ric / line_tic
 * frozen into the line functors) by the caller's arrays; the factors stay as uploaded.  The reference re-packs its
 * state between the solve and the marginalization - double2vector() applies the yaw / position gauge fix and
 * vector2double() packs the fixed state again (estimator.cpp:1006, :596-711, :1168) - so the prior must be built at
 * THAT state: call this between uvs_solve / uvs_batch_solve and uvs_marginalize.  Sizes must match the upload. */
int uvs_upload_state(UvsHandle *h, int32_t n_windows, const UvsWindow *windows);

/* Batched factor sweeps over the uploaded batch, factors concatenated in window order.
 * residuals: [n_total][nres]; jacobians (nullable): per factor, layout per flags. */
int uvs_eval_proj(UvsHandle *h, double *residuals, double *jacobians, int32_t flags);
int uvs_eval_line(UvsHandle *h, double *residuals, double *jacobians, int32_t flags);
int uvs_eval_vp(UvsHandle *h, double *residuals, double *jacobians, int32_t flags);
int uvs_eval_imu(UvsHandle *h, double *residuals, double *jacobians, int32_t flags);
int uvs_eval_prior(UvsHandle *h, double *residuals, double *jacobians, int32_t flags);
/* cost = 1/2 sum rho(|r|^2) per window, [n_windows] */
int uvs_eval_cost(UvsHandle *h, double *cost);

/* Levenberg-Marquardt + Schur solve of every uploaded window (ceres::Solve replacement). */
int uvs_solve(UvsHandle *h, UvsSummary *summaries /* [n_windows], nullable */);

/* Convenience for the optimization() shim: upload + solve + download for one batch. */
int uvs_batch_solve(UvsHandle *h, int32_t n_windows, UvsWindow *windows, const UvsOptions *opts,
                    UvsSummary *summaries);

/* The same one-call service for LARGE batches, pipelined: the batch is cut into n_groups sub-batches (0 = choose),
 * each with its own stream and arenas; sub-batch k+1 is packed and copied to the device while sub-batch k already
 * iterates.  Results are identical to uvs_batch_solve (windows are independent).  One-shot: the handle keeps no
 * batch afterwards (uvs_marginalize / uvs_eval_* need uvs_upload_windows or uvs_batch_solve). */
int uvs_batch_solve_pipelined(UvsHandle *h, int32_t n_windows, UvsWindow *windows, const UvsOptions *opts,
                              UvsSummary *summaries, int32_t n_groups);

/* Build the next prior of window `window_index` from its current state
 * (MarginalizationInfo::preMarginalize + marginalize, estimator.cpp:1003-1228). */
int uvs_marginalize(UvsHandle *h, int32_t window_index, int32_t flag, UvsPrior *out);

/* Sum of the sweep's algorithmic bytes for the uploaded batch */
int uvs_sweep_bytes(UvsHandle *h, int64_t *jacobian_sweep_bytes, int64_t *residual_sweep_bytes);
/* Number of kernel launches issued through this handle since creation. */
int64_t uvs_launch_count(const UvsHandle *h);
/* Device time of the last uvs_solve (CUDA events on the handle's stream), milliseconds. */
int uvs_last_solve_ms(const UvsHandle *h, float *ms);
/* Device time [ms] of the Jacobian-sweep kernels accumulated over the last uvs_solve. */
int uvs_last_sweep_ms(const UvsHandle *h, float *ms, int32_t *n_sweeps);

/* Materialised Jacobian sweep of the uploaded batch, for measurement: the four factor-type kernels (point, line + VP,
 * IMU, prior) evaluate every residual and tangent-space Jacobian block of the batch into the device record arrays
 * (the layout of UVS_EVAL_LOCAL_LAYOUT), `repeats` times.  ms_group = CUDA-event time per repetition with the four
 * kernels side by side on the handle's streams; ms_each[4] (nullable) = proj, line + VP, IMU, prior each on its own.
 * (uvs_solve itself evaluates point / line / VP factors inside the landmark elimination and never writes these records
 * when the batch takes the fused path; bytes per repetition: uvs_sweep_bytes.) */
int uvs_jacobian_sweep(UvsHandle *h, int32_t repeats, float *ms_group, float *ms_each);

/* Put every window back to the state it was uploaded with (device-to-device; no host traffic), so
 * that the same batch can be solved again - used to time solves with the inputs resident in HBM. */
int uvs_reset_state(UvsHandle *h);

/* Stage timing with CUDA events on the handle's stream.  level 0: whole solve only; level 1: one
 * event per pipeline stage and LM iteration, read back (no extra synchronisation) when the solve ends. */
#define UVS_N_STAGES 10
/* stage order: sweep_proj, sweep_line, sweep_vp, sweep_imu, sweep_prior, build, chol, backsub,
 * resid_sweep, step */
int uvs_set_profiling(UvsHandle *h, int32_t level);
/* Replay the LM iteration of the uploaded batch from a CUDA graph (captured during the first solve after an upload)
 * instead of launching its ~20 kernels one by one.  Off by default: building the graph costs more than one solve
 * saves; enable it when the same upload is solved many times.  Ignored while stage profiling is on. */
int uvs_set_graph_replay(UvsHandle *h, int32_t enable);
/* accumulated device time [ms] of every stage over the last uvs_solve and the number of iterations run */
int uvs_last_stage_ms(const UvsHandle *h, float ms[UVS_N_STAGES], int32_t *n_iterations);

/* IMU mid-point preintegration of `n_intervals` keyframe intervals in one launch
 * (IntegrationBase::{push_back,propagate,midPointIntegration,repropagate}, factor/integration_base.h:30-158; the step
 * that produces the IMU-factor constants of UvsWindow).  Interval k owns samples [sample_off[k], sample_off[k+1]):
 * dt[S], acc[S][3], gyr[S][3]; acc0/gyr0[n][3] = the measurement the interval starts from; lin_ba/lin_bg[n][3] =
 * linearisation biases; noise = {ACC_N, GYR_N, ACC_W, GYR_W}.  Outputs are the arrays imu_delta_p ... imu_covariance
 * of UvsWindow ([n][3], [n][4] xyzw, [n][3], [n], [n][225], [n][225]). */
int uvs_preintegrate(UvsHandle *h, int32_t n_intervals, const int32_t *sample_off, const double *dt, const double *acc,
                     const double *gyr, const double *acc0, const double *gyr0, const double *lin_ba, const double *lin_bg,
                     const double *noise, double *delta_p, double *delta_q, double *delta_v, double *sum_dt, double *jacobian,
                     double *covariance);

/* Triangulation of the features that have no estimate yet - the step that seeds para_Feature / para_Ortho_plucker
 * before optimization().  Rs[n_frames][9] (row-major) and Ps[n_frames][3] are Estimator::Rs / Ps,
 * ric[9] (row-major) / tic[3] are ric[0] / tic[0].
 *
 * uvs_triangulate_points = FeatureManager::triangulate (feature_manager.cpp:427-481): track t starts at frame
 * start_frame[t] and owns the observations [obs_off[t], obs_off[t+1]) of consecutive frames, obs_pts[.][3] =
 * FeaturePerFrame::point; depth_out[t] = V(2)/V(3) of the smallest right singular vector of the 2n x 4 DLT system, or
 * init_depth (INIT_DEPTH, parameters.h) when that is < 0.1.  The caller applies the reference's eligibility tests
 * (used_num >= 2, start_frame < WINDOW_SIZE - 2, estimated_depth <= 0) when it builds the list.
 *
 * uvs_triangulate_lines = FeatureManager::triangulateLine (feature_manager.cpp:504-589, calcPluckerLine :827-902):
 * line t is seen first in frame frame_first[t] with end points sp_first / ep_first and last in frame_last[t] with
 * sp_last / ep_last ([n_lines][3], z = 1); ortho_out[t][4] = (eulerAngles(0,1,2) of [n_w d_w n_w x d_w] normalised,
 * atan2(|d_w|, |n_w|)) - the orthonormal_vec of the feature. */
int uvs_triangulate_points(UvsHandle *h, int32_t n_frames, const double *Rs, const double *Ps, const double *ric, const double *tic,
                           int32_t n_tracks, const int32_t *start_frame, const int32_t *obs_off, const double *obs_pts,
                           double init_depth, double *depth_out);
int uvs_triangulate_lines(UvsHandle *h, int32_t n_frames, const double *Rs, const double *Ps, const double *ric, const double *tic,
                          int32_t n_lines, const int32_t *frame_first, const int32_t *frame_last, const double *sp_first,
                          const double *ep_first, const double *sp_last, const double *ep_last, double *ortho_out);
/* uvs_validate_lines = the validity test of FeatureManager::setLineOrtho (feature_manager.cpp:333-423), run after the solve
 * (double2vector, estimator.cpp:706): line t, given by the orthonormal parameters the FEATURE holds (ortho[t][4]: the test
 * uses the stored ones, the solved ones are written back only for valid lines), is taken into the camera of its first frame
 * start_frame[t] and intersected with the planes through the end points sp_first / ep_first ([n_lines][3], z = 1) of its first
 * observation; solve_flag[t] = 2 when either 3-D end point lies behind that camera, else 1.  end_points (nullable,
 * [n_lines][6]) receives the world end points D_s_w, D_e_w the reference computes on the way.  Rs / Ps: the solved poses. */
int uvs_validate_lines(UvsHandle *h, int32_t n_frames, const double *Rs, const double *Ps, const double *ric, const double *tic,
                       int32_t n_lines, const int32_t *start_frame, const double *ortho, const double *sp_first,
                       const double *ep_first, int32_t *solve_flag, double *end_points);

/* ---- Device-resident sliding window --------------------------------------------------------------
 * The reference rebuilds its Ceres problem from FeatureManager every frame (estimator.cpp:823-934) although only one frame
 * of observations, one IMU interval and the state are new.  With a resident window the observation tracks, the IMU records
 * and the marginalization prior stay on the device; per frame the caller sends the new frame (uvs_window_push_frame), the
 * packed state (the para_* arrays of vector2double, through uvs_window_upload) and nothing else:
 *
 *   uvs_window_create(h, WINDOW_SIZE, LINE_WINDOW, max tracks)           once
 *   per frame:  uvs_window_push_frame                                    FeatureManager::addFeatureCheckParallax (bookkeeping part)
 *               uvs_window_counts -> sizes of para_Feature / para_Ortho_plucker
 *               uvs_window_upload(state)                                 problem assembly of estimator.cpp:823-934, on the device
 *               uvs_solve, uvs_download_state, [double2vector / vector2double], uvs_upload_state
 *               uvs_window_marginalize(flag)                             estimator.cpp:1003-1228, prior kept on the device
 *               uvs_window_slide(flag)                                   Estimator::slideWindow, estimator.cpp:1235-1359
 *               uvs_window_remove_tracks                                 FeatureManager::removeFailures / removeOutlier / removeLineFailures
 * After uvs_window_upload the handle holds the window exactly as after uvs_upload_windows of the same window packed on the
 * host (same factor order: list order of the tracks, eligibility filters of estimator.cpp:826 and :873, running feature
 * indices), so every other entry point works on it.  One window per handle; estimate_td is not supported in this mode. */
typedef struct UvsImuRecord {      /* constants of one IntegrationBase (integration_base.h:188-207) */
  const double *delta_p;     /* [3] */
  const double *delta_q;     /* [4] x,y,z,w */
  const double *delta_v;     /* [3] */
  const double *sum_dt;      /* [1] */
  const double *lin_ba;      /* [3] */
  const double *lin_bg;      /* [3] */
  const double *jacobian;    /* [225] row-major */
  const double *covariance;  /* [225] row-major */
} UvsImuRecord;

typedef struct UvsFrameInput {     /* what one image frame adds (feature_manager.cpp:73-133; the tracker's ids) */
  const UvsImuRecord *imu;   /* preintegration from the previous frame to this one; NULL for the first frame only */
  int32_t n_points;
  int32_t n_lines;
  const int32_t *point_id;   /* [n_points] feature_id */
  const double *point_xyz;   /* [n_points][3] FeaturePerFrame::point */
  const int32_t *line_id;    /* [n_lines] */
  const double *line_sp;     /* [n_lines][2] LineFeaturePerFrame::start_point */
  const double *line_ep;     /* [n_lines][2] end_point */
  const double *line_vp;     /* [n_lines][3] vp; a VP factor exists where vp[2] == 1 (estimator.cpp:920) */
} UvsFrameInput;

int uvs_window_create(UvsHandle *h, int32_t window_size, int32_t line_window, int32_t max_points, int32_t max_lines);
int uvs_window_push_frame(UvsHandle *h, const UvsFrameInput *frame);
/* counts[8] = n_frames, eligible points, eligible lines, point factors, line factors, VP factors, IMU factors, prior_n */
int uvs_window_counts(UvsHandle *h, int32_t counts[8]);
/* `state`: n_frames / n_points / n_lines as uvs_window_counts reports them, the state arrays, line_ric / line_tic,
 * estimate_extrinsic; every factor / IMU / prior field is ignored (they come from the device-resident store). */
int uvs_window_upload(UvsHandle *h, const UvsWindow *state, const UvsOptions *opts);
/* uvs_marginalize of the resident window; the result also becomes the resident prior of the next uvs_window_upload
 * (block ids already shifted).  `out` may be NULL. */
int uvs_window_marginalize(UvsHandle *h, int32_t flag, UvsPrior *out);
/* UVS_MARGIN_OLD: frame 0 leaves (removeBackShiftDepth / removeBack / removeLineBack; the re-anchored depth is caller
 * state).  UVS_MARGIN_SECOND_NEW: the second newest frame leaves (removeFront / removeLineFront); `merged_imu` = the
 * preintegration over the last two intervals (estimator.cpp:1301-1312), required when the window has >= 2 IMU factors. */
int uvs_window_slide(UvsHandle *h, int32_t flag, const UvsImuRecord *merged_imu);
int uvs_window_remove_tracks(UvsHandle *h, int32_t n_points, const int32_t *point_id, int32_t n_lines, const int32_t *line_id);

/* Copy the factor arrays of an uploaded window back into the caller's arrays (w's counts must match; NULL pointers are
 * skipped; IMU / prior arrays only when n_imu / prior_n match): the dump hook of the device-resident window, and its test. */
int uvs_download_factors(UvsHandle *h, int32_t window_index, UvsWindow *w);
/* host-to-device bytes copied by the upload paths of this handle since creation */
int64_t uvs_h2d_bytes(const UvsHandle *h);

/* Factor-parallel multi-GPU mode: this rank owns the landmarks with (index % nranks == rank);
 * IMU factors and the prior belong to rank 0.  `reduce` is called once per LM iteration with the
 * device buffer holding the rank's partial reduced camera system (count doubles) and must sum it
 * over ranks in place (e.g. ncclAllReduce on `stream`). */
typedef int (*UvsAllReduceFn)(void *user, void *device_buf, int64_t count, void *cuda_stream);
int uvs_comm_init(UvsHandle *h, int32_t rank, int32_t nranks, UvsAllReduceFn reduce, void *user);

/* The same mode with the library's own NCCL communicator (uvs_comm_init(h, ncclUniqueId, rank, nranks)):
 * one rank calls uvs_comm_unique_id (ncclGetUniqueId; the 128 bytes of a ncclUniqueId), the caller hands the bytes to
 * every rank by whatever means it has (MPI, a file, torch.distributed), every rank calls uvs_comm_init_nccl
 * (ncclCommInitRank, collective).  Per LM iteration the library then issues ONE ncclAllReduce (sum, double) of
 * [S | gS | g | column norms | per-window accumulators] on the handle's stream, plus one of 16 doubles per window for
 * the candidate cost.  libnccl.so.2 is bound at run time (dlopen); nranks == 1 leaves the mode. */
#define UVS_NCCL_UNIQUE_ID_BYTES 128
int uvs_comm_unique_id(unsigned char *id /* [UVS_NCCL_UNIQUE_ID_BYTES] */);
int uvs_comm_init_nccl(UvsHandle *h, const unsigned char *id, int32_t rank, int32_t nranks);
/* number of all-reduces issued through this handle since creation */
int64_t uvs_collective_count(const UvsHandle *h);

#ifdef __cplusplus
}
#endif
#endif /* UVS_H_ */

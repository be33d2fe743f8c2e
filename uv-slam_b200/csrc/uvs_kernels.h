// uvs_kernels.h — host-callable launch wrappers of the sm_100a kernels (internal to libuvs_b200).
// Every wrapper returns the number of kernel launches it issued.
#pragma once
#include <cuda_runtime.h>

#include "uvs_device.cuh"

namespace uvs {

// uvs_prep.cu
int launch_prep(const Dev &D, cudaStream_t st);
int launch_split(const double *rec, long long n, int REC, int NR, double *r_out, double *J_out, cudaStream_t st);
int launch_export_imu(const double *rec, int n, int PW, double *r_out, double *J_out, cudaStream_t st);
int launch_export_prior(const Dev &D, int PW, double *J_out, const long long *out_off, cudaStream_t st);

// uvs_sweep.cu — mode 0: API evaluation of every factor; mode 1: solver (per-window state gates)
int launch_proj(const Dev &D, const Params &P, bool jac, bool ceres, int mode, int cand, double *out, double *res_out,
                double *cost, int cost_stride, cudaStream_t st);
int launch_line(const Dev &D, const Params &P, bool jac, bool ceres, int mode, int cand, double *out, double *res_out,
                double *cost, int cost_stride, cudaStream_t st);
int launch_vp(const Dev &D, const Params &P, bool jac, bool ceres, int mode, int cand, double *out, double *res_out,
              double *cost, int cost_stride, cudaStream_t st);
// solver path: line + VP factors of one observation in one pass (every VP factor is paired with a line factor at upload)
int launch_line_vp(const Dev &D, const Params &P, bool jac, int mode, int cand, double *out_line, double *out_vp, double *cost,
                   int cost_stride, cudaStream_t st);
// tables of the line / VP factors for state buffer cur ^ cand (must precede launch_line_vp / the fused line kernel at that state)
int launch_line_tables(const Dev &D, int cand, cudaStream_t st);
int launch_imu(const Dev &D, const Params &P, bool jac, int mode, int cand, double *out, double *res_out, double *cost,
               int cost_stride, cudaStream_t st);
int launch_prior(const Dev &D, int max_prior_n, bool jac_phase, int mode, int cand, double *res_out, double *cost,
                 int cost_stride, cudaStream_t st);

// uvs_build.cu
int launch_build(const Dev &D, const Params &P, int max_prior_n, cudaStream_t st);
int launch_backsub(const Dev &D, const Params &P, cudaStream_t st);

// uvs_build3.cu — atomics-free landmark path for windows of <= 12 six-wide camera blocks without td
struct Build3Layout {
  int mp;                                        // doubles per dense landmark column (6 blocks + z + pad)
  size_t o_Y, o_ph, o_lh, o_items, o_off;
};
// Auxiliary streams of a handle: independent kernels of one solver stage run side by side (most of them are latency-
// bound, so they overlap almost perfectly); every stage forks from and joins the handle's main stream with events.
struct Fork {
  cudaStream_t aux[3];
  cudaEvent_t fork, join[3];
};
inline void fork_from(const Fork *fk, cudaStream_t main, int n) {
  cudaEventRecord(fk->fork, main);
  for (int k = 0; k < n; k++) cudaStreamWaitEvent(fk->aux[k], fk->fork, 0);
}
inline void join_to(const Fork *fk, cudaStream_t main, int k) {
  cudaEventRecord(fk->join[k], fk->aux[k]);
  cudaStreamWaitEvent(main, fk->join[k], 0);
}

size_t build3_bytes(const Dev &D, int max_frames, bool any_ex, Build3Layout *lay);
size_t build3_smem(int max_frames, bool any_ex, int max_prior_n);
int launch_stash_init(const Dev &D, char *base, const Build3Layout &lay, cudaStream_t st);
int launch_build3_prep(const Dev &D, char *base, const Build3Layout &lay, bool any_ex, cudaStream_t st);
int launch_build3(const Dev &D, const Params &P, char *base, const Build3Layout &lay, int max_frames, bool any_ex, int max_prior_n,
                  cudaStream_t st, const Fork *fk);
int launch_back3(const Dev &D, char *base, const Build3Layout &lay, cudaStream_t st, const Fork *fk, bool unscaled_pts);
// record path: Jacobian sweep + build stage as one dependency graph over the handle's streams (small batches, constant extrinsic)
int launch_record_linearisation(const Dev &D, const Params &P, char *base, const Build3Layout &lay, int max_frames, int max_prior_n,
                                cudaStream_t st, const Fork *fk);
int launch_build_cam(const Dev &D, int max_prior_n, cudaStream_t st);
// fused linearisation (uvs_lin.cu): factors evaluated in registers, no Jacobian records in HBM; launch_build3_fused is the
// build stage of that path (point + line linearisation, IMU / prior tail, rank update)
int launch_build3_fused(const Dev &D, const Params &P, char *base, const Build3Layout &lay, int max_frames, int max_lines, int max_prior_n,
                        cudaStream_t st, const Fork *fk);

// uvs_lin.cu
int launch_prep_point_order(const Dev &D, int *key_scratch, cudaStream_t st);
size_t lin_lines_smem(int max_frames);
int lin_max_line_obs();
int launch_lin_points(const Dev &D, const Params &P, char *base, const Build3Layout &lay, cudaStream_t st);
int launch_lin_lines(const Dev &D, const Params &P, char *base, const Build3Layout &lay, int max_frames, int max_lines, cudaStream_t st);

// uvs_solve.cu
int chol_packed_limit(size_t max_smem);
size_t chol_max_dynamic_smem(size_t optin_bytes);
int launch_solve_init(const Dev &D, const Params &P, cudaStream_t st);
int launch_chol(const Dev &D, const Params &P, int max_d, int packed_limit, bool mc_identity, cudaStream_t st);
// chain mode (uvs_solve.cu): speed-bias blocks eliminated one by one, dense tensor-core Cholesky of the rest
size_t chol_chain_smem(int max_frames, bool any_ex);
int chol_chain_lw_doubles(int max_frames);
long long chol_frag_doubles(int d);   // doubles of global scratch per window of the large-window reduced solve
int launch_chol_chain(const Dev &D, const Params &P, int max_frames, bool any_ex, bool mc_identity, cudaStream_t st);
int launch_step(const Dev &D, const Params &P, bool clear_system, cudaStream_t st);
int launch_finish(const Dev &D, cudaStream_t st);
int launch_count_active(const Dev &D, int *out, cudaStream_t st);
int launch_copy_acc(const Dev &D, int slot, double *out, int zero, cudaStream_t st);

int launch_gather_state(const Dev &D, double *out_pose, double *out_sb, double *out_ex, double *out_td, double *out_inv,
                        double *out_ortho, cudaStream_t st);

}  // namespace uvs

// uvs_marg.cu
struct UvsHandle;
int uvs_marginalize_impl(UvsHandle *h, int window_index, int flag, UvsPrior *out);

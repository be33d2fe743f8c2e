// uvs_window.cu — device-resident sliding window.
//
// The reference rebuilds its whole Ceres problem every frame from FeatureManager (estimator.cpp:823-934) although only
// one frame of observations, one IMU interval and the state are new.  Here the observation tracks, the IMU records and
// the marginalization prior STAY on the device between frames:
//   uvs_window_push_frame   appends the new frame's observations / IMU record to the resident track store
//                           (FeatureManager::addFeatureCheckParallax bookkeeping, feature_manager.cpp:73-133)
//   uvs_window_upload       assembles the flat factor arrays of uvs_upload_windows ON THE DEVICE in the order of the
//                           reference's assembly loops (eligibility filters of estimator.cpp:826 / :873, running feature
//                           indices), from a few-KB plan; only the state (para_* arrays) comes from the caller
//   uvs_window_marginalize  uvs_marginalize whose result also stays on the device as the next window's prior
//   uvs_window_slide        Estimator::slideWindow (estimator.cpp:1235-1359) on the store: removeBack[ShiftDepth] /
//                           removeLineBack or removeFront / removeLineFront (feature_manager.cpp:607-723), in place
// The host keeps a mirror of the integer bookkeeping only (track id -> slot, start frame, length, VP flags); every
// observation payload is uploaded once, when its frame arrives.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/uvs.h"
#include "uvs_device.cuh"
#include "uvs_handle.h"
#include "uvs_kernels.h"

namespace uvs {

constexpr int IMU_REC = 467;   // dp3 dq4 dv3 dt1 ba3 bg3 jac225 cov225

struct RTrack {
  int id = 0, start = 0, n = 0;
  bool live = false;
  std::vector<unsigned char> vp;   // lines: VP factor present per observation (it_per_frame.vp(2) == 1, estimator.cpp:920)
};

struct ResidentWindow {
  int W = 10, line_window = 5, nobs = 11;   // WINDOW_SIZE, LINE_WINDOW, observations a track can hold (W + 1)
  int max_pt = 0, max_ln = 0;
  int n_frames = 0, n_imu = 0, imu_base = 0;
  std::vector<RTrack> pt, ln;               // per slot
  std::vector<int> pt_order, ln_order;      // live slots in list (insertion) order = FeatureManager::feature / line_feature
  std::vector<int> pt_free, ln_free;
  std::unordered_map<int, int> pt_slot, ln_slot;
  std::vector<double> imu_dt;               // sum_dt per resident interval (marginalization plan, estimator.cpp:1036)
  // prior kept from uvs_window_marginalize
  int prior_n = 0;
  std::vector<int> prior_kind, prior_id;
  std::vector<double> prior_x0;
  // device store
  char *dev = nullptr;
  double *d_pt_obs = nullptr, *d_ln_obs = nullptr, *d_imu = nullptr, *d_prior_J = nullptr, *d_prior_r = nullptr;
  int *d_pt_start = nullptr, *d_pt_n = nullptr, *d_ln_start = nullptr, *d_ln_n = nullptr;
  char *d_cmd = nullptr, *h_cmd = nullptr;  // per-call command / plan buffer (device, pinned host)
  size_t cmd_bytes = 0;
  int prior_cap = 0;
  // plan of the last assembly (uvs_window_upload)
  std::vector<int> plan, plan_off;
  int e_np = 0, e_nl = 0, e_nproj = 0, e_nlobs = 0, e_nvobs = 0, line_run_max = 0;
};

// ---- kernels ------------------------------------------------------------------------------------------------------
// append one observation per command: cmd = {slot | NEW_BIT}; a new track starts at window position `pos`
constexpr int NEW_BIT = 1 << 30;
__global__ void k_rw_push(int n, const int *slots, const double *payload, int width, int pos, int nobs, double *obs, int *start, int *cnt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int s = slots[i];
  if (s & NEW_BIT) { s &= ~NEW_BIT; start[s] = pos; cnt[s] = 0; }
  const int k = cnt[s];
  double *dst = obs + ((size_t)s * nobs + k) * width;
  for (int c = 0; c < width; c++) dst[c] = payload[(size_t)i * width + c];
  cnt[s] = k + 1;
}

// MARGIN_OLD (removeBackShiftDepth / removeBack / removeLineBack): tracks that start at frame 0 lose their first
// observation, the others move one frame down; a point track left with fewer than min_left observations dies
__global__ void k_rw_slide_old(int nslots, int width, int nobs, int min_left, double *obs, int *start, int *cnt) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nslots) return;
  const int n = cnt[s];
  if (n <= 0) return;
  if (start[s] != 0) { start[s]--; return; }
  double *o = obs + (size_t)s * nobs * width;
  for (int e = 0; e < (n - 1) * width; e++) o[e] = o[e + width];
  cnt[s] = (n - 1 < min_left) ? 0 : n - 1;
}

// MARGIN_SECOND_NEW (removeFront / removeLineFront with frame_count = fc): the observation of frame fc - 1 goes, what was
// seen in frame fc moves to fc - 1
__global__ void k_rw_slide_new(int nslots, int width, int nobs, int fc, double *obs, int *start, int *cnt) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nslots) return;
  const int n = cnt[s];
  if (n <= 0) return;
  if (start[s] == fc) { start[s]--; return; }
  const int j = fc - 1 - start[s];
  if (start[s] + n - 1 < fc - 1) return;
  double *o = obs + (size_t)s * nobs * width;
  for (int e = j * width; e < (n - 1) * width; e++) o[e] = o[e + width];
  cnt[s] = n - 1;
}

__global__ void k_rw_kill(int n, const int *slots, int *cnt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) cnt[slots[i]] = 0;
}

// exclusive prefix sum over the block (blockDim.x = 1024 >= n values, one per thread)
__device__ __forceinline__ int block_excl_scan(int v, int *sh) {
  const int t = threadIdx.x;
  sh[t] = v;
  __syncthreads();
  for (int o = 1; o < (int)blockDim.x; o <<= 1) {
    const int add = t >= o ? sh[t - o] : 0;
    __syncthreads();
    sh[t] += add;
    __syncthreads();
  }
  return sh[t] - v;
}

// point factors of the k-th eligible track (plan: its slot): one ProjectionFactor per observation after the first
// (estimator.cpp:838-865), feature index = k; the first factor of a track = what the tracks before it hold (block scan)
__global__ void __launch_bounds__(1024) k_rw_assemble_points(int np, const int *plan, int nobs, const double *obs, const int *start, const int *cnt,
                                                             int *fi, int *fj, int *pt, double *pi, double *pj) {
  __shared__ int sh[1024];
  const int k = threadIdx.x;
  const int s = k < np ? plan[k] : 0;
  const int n = k < np ? cnt[s] : 1;
  const int off = block_excl_scan(n - 1, sh);
  if (k >= np) return;
  const int st = start[s];
  const double *o = obs + (size_t)s * nobs * 3;
  for (int j = 1; j < n; j++) {
    const int f = off + j - 1;
    fi[f] = st; fj[f] = st + j; pt[f] = k;
    for (int c = 0; c < 3; c++) { pi[3 * (size_t)f + c] = o[c]; pj[3 * (size_t)f + c] = o[3 * j + c]; }
  }
}

// line + VP factors of the k-th eligible line: one LineProjectionFactor per observation including the start frame, one
// VPProjectionFactor where vp(2) == 1 (estimator.cpp:881-931)
__global__ void __launch_bounds__(1024) k_rw_assemble_lines(int nl, const int *plan, int nobs, const double *obs, const int *start, const int *cnt,
                                                            int *lf, int *li, double *sp, double *ep, int *vf, int *vl, double *vd) {
  __shared__ int sh[1024];
  const int k = threadIdx.x;
  const int s = k < nl ? plan[k] : 0;
  const int n = k < nl ? cnt[s] : 0;
  const double *o = obs + (size_t)s * nobs * 7;
  int nv = 0;
  for (int j = 0; j < n; j++) nv += o[7 * j + 6] == 1.0 ? 1 : 0;
  const int off = block_excl_scan(n, sh);
  int voff = block_excl_scan(nv, sh);
  if (k >= nl) return;
  const int st = start[s];
  for (int j = 0; j < n; j++) {
    const int f = off + j;
    lf[f] = st + j; li[f] = k;
    sp[2 * (size_t)f] = o[7 * j]; sp[2 * (size_t)f + 1] = o[7 * j + 1];
    ep[2 * (size_t)f] = o[7 * j + 2]; ep[2 * (size_t)f + 1] = o[7 * j + 3];
    if (o[7 * j + 6] == 1.0) {
      vf[voff] = st + j; vl[voff] = k;
      for (int c = 0; c < 3; c++) vd[3 * (size_t)voff + c] = o[7 * j + 4 + c];
      voff++;
    }
  }
}

// line observations arrive as [sp(2) ep(2) vp.xy(2)] + a flag bit in the slot word (vp(2) == 1, the only value the assembly
// tests): the store keeps the reference's 7 numbers
constexpr int VP_BIT = 1 << 29;
__global__ void k_rw_push_lines(int n, const int *slots, const double *payload, int pos, int nobs, double *obs, int *start, int *cnt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int s = slots[i];
  const bool has_vp = (s & VP_BIT) != 0, fresh = (s & NEW_BIT) != 0;
  s &= ~(NEW_BIT | VP_BIT);
  if (fresh) { start[s] = pos; cnt[s] = 0; }
  const int k = cnt[s];
  double *dst = obs + ((size_t)s * nobs + k) * 7;
  for (int c = 0; c < 6; c++) dst[c] = payload[(size_t)i * 6 + c];
  dst[6] = has_vp ? 1.0 : 0.0;
  if (!has_vp) { dst[4] = 0.0; dst[5] = 0.0; }
  cnt[s] = k + 1;
}

// IMU records of the ring -> the SoA sections of the input region; factor k links frames k and k + 1
__global__ void k_rw_assemble_imu(int n, int base, int nobs, const double *ring, int *fr, double *dp, double *dq, double *dv, double *dt,
                                  double *ba, double *bg, double *jac, double *cov) {
  const int k = blockIdx.x;
  if (k >= n) return;
  const double *r = ring + (size_t)((base + k) % nobs) * IMU_REC;
  for (int e = threadIdx.x; e < IMU_REC; e += blockDim.x) {
    const double v = r[e];
    if (e < 3) dp[3 * k + e] = v;
    else if (e < 7) dq[4 * k + e - 3] = v;
    else if (e < 10) dv[3 * k + e - 7] = v;
    else if (e < 11) dt[k] = v;
    else if (e < 14) ba[3 * k + e - 11] = v;
    else if (e < 17) bg[3 * k + e - 14] = v;
    else if (e < 242) jac[225 * (size_t)k + e - 17] = v;
    else cov[225 * (size_t)k + e - 242] = v;
  }
  if (threadIdx.x == 0) fr[k] = k;
}

void resident_destroy(UvsHandle *h) {
  if (!h || !h->resident) return;
  ResidentWindow *R = (ResidentWindow *)h->resident;
  if (R->dev) cudaFree(R->dev);
  if (R->d_cmd) cudaFree(R->d_cmd);
  if (R->h_cmd) cudaFreeHost(R->h_cmd);
  delete R;
  h->resident = nullptr;
}

}  // namespace uvs

using namespace uvs;

#define CKW(call)                                                                                   \
  do {                                                                                              \
    cudaError_t e_ = (call);                                                                        \
    if (e_ != cudaSuccess) return handle_fail(h, UVS_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

namespace {

ResidentWindow *resident(UvsHandle *h) { return h ? (ResidentWindow *)h->resident : nullptr; }

int ensure_cmd(UvsHandle *h, ResidentWindow *R, size_t bytes) {
  if (bytes <= R->cmd_bytes) return UVS_OK;
  if (R->d_cmd) cudaFree(R->d_cmd);
  if (R->h_cmd) cudaFreeHost(R->h_cmd);
  R->d_cmd = nullptr; R->h_cmd = nullptr; R->cmd_bytes = 0;
  const size_t want = bytes * 2 + 4096;
  CKW(cudaMalloc((void **)&R->d_cmd, want));
  CKW(cudaMallocHost((void **)&R->h_cmd, want));
  R->cmd_bytes = want;
  return UVS_OK;
}

void erase_track(std::vector<RTrack> &tr, std::vector<int> &order, std::vector<int> &freel, std::unordered_map<int, int> &map, int slot) {
  map.erase(tr[slot].id);
  tr[slot].live = false; tr[slot].n = 0; tr[slot].vp.clear();
  order.erase(std::find(order.begin(), order.end(), slot));
  freel.push_back(slot);
}

// eligibility filters of the assembly loops (estimator.cpp:822-829, :871-878): plan = eligible slots in list order (what goes
// to the device); plan_off = first point / line / VP factor of each (host mirror only, the device scans)
void make_plan(ResidentWindow *R) {
  R->plan.clear(); R->plan_off.clear();
  int k = 0, off = 0;
  for (int s : R->pt_order) {
    const RTrack &t = R->pt[s];
    if (!(t.n >= 2 && t.start < R->W - 2)) continue;
    R->plan.push_back(s); R->plan_off.push_back(off);
    off += t.n - 1; k++;
  }
  R->e_np = k; R->e_nproj = off;
  k = 0; off = 0;
  int voff = 0, run = 0;
  for (int s : R->ln_order) {
    const RTrack &t = R->ln[s];
    if (t.n < R->line_window) continue;
    R->plan.push_back(s); R->plan_off.push_back(off); R->plan_off.push_back(voff);
    off += t.n; k++;
    run = std::max(run, t.n);
    for (int j = 0; j < t.n; j++) voff += t.vp[j] ? 1 : 0;
  }
  R->e_nl = k; R->e_nlobs = off; R->e_nvobs = voff; R->line_run_max = run;
}

// ResidentHook::fill: plan -> device, assembly kernels on the upload's stream, integer index arrays into the pinned
// staging mirror (uvs_marginalize plans on the host from there)
int fill_from_store(void *user, UvsHandle *h, char *Dv, char *S, const InputOffsets &o) {
  ResidentWindow *R = (ResidentWindow *)user;
  cudaStream_t st = h->stream;
  const size_t pbytes = R->plan.size() * sizeof(int);
  int rc = ensure_cmd(h, R, pbytes + 64); if (rc) return rc;
  if (pbytes) {
    std::memcpy(R->h_cmd, R->plan.data(), pbytes);
    CKW(cudaMemcpyAsync(R->d_cmd, R->h_cmd, pbytes, cudaMemcpyHostToDevice, st));
    h->h2d_bytes += (int64_t)pbytes;
  }
  const int *plan_p = (const int *)R->d_cmd, *plan_l = plan_p + R->e_np;
  if (R->e_np) {
    k_rw_assemble_points<<<1, 1024, 0, st>>>(R->e_np, plan_p, R->nobs, R->d_pt_obs, R->d_pt_start, R->d_pt_n, (int *)(Dv + o.pfi),
                                                               (int *)(Dv + o.pfj), (int *)(Dv + o.ppt), (double *)(Dv + o.ppi), (double *)(Dv + o.ppj));
    h->launches++;
  }
  if (R->e_nl) {
    k_rw_assemble_lines<<<1, 1024, 0, st>>>(R->e_nl, plan_l, R->nobs, R->d_ln_obs, R->d_ln_start, R->d_ln_n, (int *)(Dv + o.lf),
                                                              (int *)(Dv + o.li), (double *)(Dv + o.lsp), (double *)(Dv + o.lep), (int *)(Dv + o.vf),
                                                              (int *)(Dv + o.vl), (double *)(Dv + o.vd));
    h->launches++;
  }
  if (R->n_imu) {
    k_rw_assemble_imu<<<R->n_imu, 128, 0, st>>>(R->n_imu, R->imu_base, R->nobs, R->d_imu, (int *)(Dv + o.imu_f), (double *)(Dv + o.idp), (double *)(Dv + o.idq),
                                               (double *)(Dv + o.idv), (double *)(Dv + o.idt), (double *)(Dv + o.iba), (double *)(Dv + o.ibg),
                                               (double *)(Dv + o.ijac), (double *)(Dv + o.icov));
    h->launches++;
  }
  if (R->prior_n) {
    CKW(cudaMemcpyAsync(Dv + o.prJ, R->d_prior_J, (size_t)R->prior_n * R->prior_n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    CKW(cudaMemcpyAsync(Dv + o.prr, R->d_prior_r, (size_t)R->prior_n * sizeof(double), cudaMemcpyDeviceToDevice, st));
  }
  // host mirror of the integer arrays (and sum_dt) the marginalization plan reads
  int *fi = (int *)(S + o.pfi), *fj = (int *)(S + o.pfj), *pt = (int *)(S + o.ppt);
  for (int k = 0; k < R->e_np; k++) {
    const RTrack &t = R->pt[R->plan[k]];
    const int off = R->plan_off[k];
    for (int j = 1; j < t.n; j++) { fi[off + j - 1] = t.start; fj[off + j - 1] = t.start + j; pt[off + j - 1] = k; }
  }
  int *lf = (int *)(S + o.lf), *li = (int *)(S + o.li), *vf = (int *)(S + o.vf), *vl = (int *)(S + o.vl);
  for (int k = 0; k < R->e_nl; k++) {
    const RTrack &t = R->ln[R->plan[R->e_np + k]];
    const int off = R->plan_off[R->e_np + 2 * k];
    int voff = R->plan_off[R->e_np + 2 * k + 1];
    for (int j = 0; j < t.n; j++) {
      lf[off + j] = t.start + j; li[off + j] = k;
      if (t.vp[j]) { vf[voff] = t.start + j; vl[voff] = k; voff++; }
    }
  }
  int *imf = (int *)(S + o.imu_f);
  double *idt = (double *)(S + o.idt);
  for (int k = 0; k < R->n_imu; k++) { imf[k] = k; idt[k] = R->imu_dt[k]; }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return handle_fail(h, UVS_ERR_CUDA, std::string("device-resident window assembly: ") + cudaGetErrorString(e));
  return UVS_OK;
}

}  // namespace

extern "C" {

int uvs_window_create(UvsHandle *h, int32_t window_size, int32_t line_window, int32_t max_points, int32_t max_lines) {
  if (!h || window_size < 2 || window_size > 31 || line_window < 1 || max_points < 1 || max_lines < 0)
    return handle_fail(h, UVS_ERR_INVALID_ARG, "uvs_window_create: bad arguments");
  CKW(cudaSetDevice(h->device));
  resident_destroy(h);
  ResidentWindow *R = new ResidentWindow();
  R->W = window_size; R->line_window = line_window; R->nobs = window_size + 1;
  R->max_pt = max_points; R->max_ln = std::max(1, (int)max_lines);
  R->pt.resize(R->max_pt); R->ln.resize(R->max_ln);
  for (int s = R->max_pt - 1; s >= 0; s--) R->pt_free.push_back(s);
  for (int s = R->max_ln - 1; s >= 0; s--) R->ln_free.push_back(s);
  R->prior_cap = 15 * (window_size + 1) + 7;
  auto al = [](size_t v) { return (v + 255) / 256 * 256; };
  size_t o = 0;
  const size_t o_po = o; o += al((size_t)R->max_pt * R->nobs * 3 * sizeof(double));
  const size_t o_lo = o; o += al((size_t)R->max_ln * R->nobs * 7 * sizeof(double));
  const size_t o_im = o; o += al((size_t)R->nobs * IMU_REC * sizeof(double));
  const size_t o_pj = o; o += al((size_t)R->prior_cap * R->prior_cap * sizeof(double));
  const size_t o_pr = o; o += al((size_t)R->prior_cap * sizeof(double));
  const size_t o_ps = o; o += al((size_t)R->max_pt * sizeof(int));
  const size_t o_pn = o; o += al((size_t)R->max_pt * sizeof(int));
  const size_t o_ls = o; o += al((size_t)R->max_ln * sizeof(int));
  const size_t o_ln = o; o += al((size_t)R->max_ln * sizeof(int));
  if (cudaMalloc((void **)&R->dev, o) != cudaSuccess) { delete R; cudaGetLastError(); return handle_fail(h, UVS_ERR_CUDA, "uvs_window_create: cudaMalloc"); }
  cudaMemsetAsync(R->dev, 0, o, h->stream);
  R->d_pt_obs = (double *)(R->dev + o_po); R->d_ln_obs = (double *)(R->dev + o_lo); R->d_imu = (double *)(R->dev + o_im);
  R->d_prior_J = (double *)(R->dev + o_pj); R->d_prior_r = (double *)(R->dev + o_pr);
  R->d_pt_start = (int *)(R->dev + o_ps); R->d_pt_n = (int *)(R->dev + o_pn); R->d_ln_start = (int *)(R->dev + o_ls); R->d_ln_n = (int *)(R->dev + o_ln);
  h->resident = R;
  h->have_window = false;
  return UVS_OK;
}

int uvs_window_push_frame(UvsHandle *h, const UvsFrameInput *f) {
  ResidentWindow *R = resident(h);
  if (!R || !f) return handle_fail(h, UVS_ERR_INVALID_ARG, "uvs_window_push_frame: no resident window (uvs_window_create) / null frame");
  if (R->n_frames > R->W) return handle_fail(h, UVS_ERR_CAPACITY, "uvs_window_push_frame: the window is full - slide first");
  if (f->n_points < 0 || f->n_lines < 0 || (f->n_points && (!f->point_id || !f->point_xyz)) ||
      (f->n_lines && (!f->line_id || !f->line_sp || !f->line_ep || !f->line_vp)))
    return handle_fail(h, UVS_ERR_INVALID_ARG, "uvs_window_push_frame: null observation array");
  if (R->n_frames > 0 && !f->imu) return handle_fail(h, UVS_ERR_INVALID_ARG, "uvs_window_push_frame: every frame after the first needs its IMU record");
  if (f->imu && (!f->imu->delta_p || !f->imu->delta_q || !f->imu->delta_v || !f->imu->sum_dt || !f->imu->lin_ba || !f->imu->lin_bg || !f->imu->jacobian ||
                 !f->imu->covariance))
    return handle_fail(h, UVS_ERR_INVALID_ARG, "uvs_window_push_frame: null IMU array");
  CKW(cudaSetDevice(h->device));
  CKW(cudaStreamSynchronize(h->stream));   // the pinned command buffer may still feed an earlier copy
  const int pos = R->n_frames, np = f->n_points, nl = f->n_lines;
  const size_t b_ps = (size_t)np * sizeof(int), b_ls = (size_t)nl * sizeof(int);
  auto al = [](size_t v) { return (v + 15) / 16 * 16; };
  const size_t o_ps = 0, o_ls = al(o_ps + b_ps), o_pp = al(o_ls + b_ls), o_lp = al(o_pp + (size_t)np * 3 * sizeof(double)),
               o_im = al(o_lp + (size_t)nl * 6 * sizeof(double)), total = o_im + (f->imu ? IMU_REC * sizeof(double) : 0);
  int rc = ensure_cmd(h, R, total); if (rc) return rc;
  // ---- host bookkeeping first (nothing is committed on failure): id -> slot, new tracks start here
  std::vector<int> ps(np), ls(nl);
  std::vector<int> new_pt, new_ln;
  auto rollback = [&]() {
    for (int s : new_pt) { R->pt_slot.erase(R->pt[s].id); R->pt[s].live = false; R->pt_order.pop_back(); R->pt_free.push_back(s); }
    for (int s : new_ln) { R->ln_slot.erase(R->ln[s].id); R->ln[s].live = false; R->ln_order.pop_back(); R->ln_free.push_back(s); }
  };
  for (int i = 0; i < np; i++) {
    auto it = R->pt_slot.find(f->point_id[i]);
    if (it != R->pt_slot.end()) {
      const RTrack &t = R->pt[it->second];
      if (t.start + t.n != pos) { rollback(); return handle_fail(h, UVS_ERR_INVALID_ARG, "uvs_window_push_frame: point track is not continuous"); }
      ps[i] = it->second;
    } else {
      if (R->pt_free.empty()) { rollback(); return handle_fail(h, UVS_ERR_CAPACITY, "uvs_window_push_frame: out of point track slots"); }
      const int s = R->pt_free.back(); R->pt_free.pop_back();
      RTrack &t = R->pt[s]; t.id = f->point_id[i]; t.start = pos; t.n = 0; t.live = true;
      R->pt_slot[t.id] = s; R->pt_order.push_back(s); new_pt.push_back(s);
      ps[i] = s | NEW_BIT;
    }
  }
  for (int i = 0; i < nl; i++) {
    auto it = R->ln_slot.find(f->line_id[i]);
    if (it != R->ln_slot.end()) {
      const RTrack &t = R->ln[it->second];
      if (t.start + t.n != pos) { rollback(); return handle_fail(h, UVS_ERR_INVALID_ARG, "uvs_window_push_frame: line track is not continuous"); }
      ls[i] = it->second;
    } else {
      if (R->ln_free.empty()) { rollback(); return handle_fail(h, UVS_ERR_CAPACITY, "uvs_window_push_frame: out of line track slots"); }
      const int s = R->ln_free.back(); R->ln_free.pop_back();
      RTrack &t = R->ln[s]; t.id = f->line_id[i]; t.start = pos; t.n = 0; t.live = true; t.vp.clear();
      R->ln_slot[t.id] = s; R->ln_order.push_back(s); new_ln.push_back(s);
      ls[i] = s | NEW_BIT;
    }
  }
  for (int i = 0; i < np; i++) R->pt[ps[i] & ~NEW_BIT].n++;
  for (int i = 0; i < nl; i++) {
    RTrack &t = R->ln[ls[i] & ~NEW_BIT];
    const bool has_vp = f->line_vp[3 * i + 2] == 1.0;
    t.n++; t.vp.push_back(has_vp ? 1 : 0);
    if (has_vp) ls[i] |= VP_BIT;
  }
  // ---- payload: one copy, two append kernels
  char *H = R->h_cmd;
  if (np) { std::memcpy(H + o_ps, ps.data(), b_ps); std::memcpy(H + o_pp, f->point_xyz, (size_t)np * 3 * sizeof(double)); }
  double *lp = (double *)(H + o_lp);
  for (int i = 0; i < nl; i++) {
    lp[6 * i] = f->line_sp[2 * i]; lp[6 * i + 1] = f->line_sp[2 * i + 1]; lp[6 * i + 2] = f->line_ep[2 * i]; lp[6 * i + 3] = f->line_ep[2 * i + 1];
    lp[6 * i + 4] = f->line_vp[3 * i]; lp[6 * i + 5] = f->line_vp[3 * i + 1];
  }
  if (nl) std::memcpy(H + o_ls, ls.data(), b_ls);
  if (f->imu) {
    const UvsImuRecord &m = *f->imu;
    double *r = (double *)(H + o_im);
    std::memcpy(r, m.delta_p, 24); std::memcpy(r + 3, m.delta_q, 32); std::memcpy(r + 7, m.delta_v, 24); r[10] = m.sum_dt[0];
    std::memcpy(r + 11, m.lin_ba, 24); std::memcpy(r + 14, m.lin_bg, 24); std::memcpy(r + 17, m.jacobian, 225 * 8); std::memcpy(r + 242, m.covariance, 225 * 8);
  }
  if (total) { CKW(cudaMemcpyAsync(R->d_cmd, H, total, cudaMemcpyHostToDevice, h->stream)); h->h2d_bytes += (int64_t)total; }
  if (np) { k_rw_push<<<(np + 127) / 128, 128, 0, h->stream>>>(np, (const int *)(R->d_cmd + o_ps), (const double *)(R->d_cmd + o_pp), 3, pos, R->nobs, R->d_pt_obs, R->d_pt_start, R->d_pt_n); h->launches++; }
  if (nl) { k_rw_push_lines<<<(nl + 127) / 128, 128, 0, h->stream>>>(nl, (const int *)(R->d_cmd + o_ls), (const double *)(R->d_cmd + o_lp), pos, R->nobs, R->d_ln_obs, R->d_ln_start, R->d_ln_n); h->launches++; }
  if (f->imu && pos > 0) {
    CKW(cudaMemcpyAsync(R->d_imu + (size_t)((R->imu_base + R->n_imu) % R->nobs) * IMU_REC, R->d_cmd + o_im, IMU_REC * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    R->imu_dt.push_back(f->imu->sum_dt[0]);
    R->n_imu++;
  }
  R->n_frames++;
  h->have_window = false;   // the uploaded batch no longer describes the store
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return handle_fail(h, UVS_ERR_CUDA, std::string("uvs_window_push_frame: ") + cudaGetErrorString(e));
  return UVS_OK;
}

int uvs_window_counts(UvsHandle *h, int32_t counts[8]) {
  ResidentWindow *R = resident(h);
  if (!R || !counts) return handle_fail(h, UVS_ERR_INVALID_ARG, "uvs_window_counts: no resident window");
  make_plan(R);
  counts[0] = R->n_frames; counts[1] = R->e_np; counts[2] = R->e_nl; counts[3] = R->e_nproj; counts[4] = R->e_nlobs; counts[5] = R->e_nvobs;
  counts[6] = R->n_imu; counts[7] = R->prior_n;
  return UVS_OK;
}

int uvs_window_upload(UvsHandle *h, const UvsWindow *state, const UvsOptions *opts) {
  ResidentWindow *R = resident(h);
  if (!R || !state) return handle_fail(h, UVS_ERR_INVALID_ARG, "uvs_window_upload: no resident window / null state");
  if (R->n_frames < 2) return handle_fail(h, UVS_ERR_NO_WINDOW, "uvs_window_upload: fewer than two frames");
  make_plan(R);
  if (R->e_np > 1024 || R->e_nl > 1024) return handle_fail(h, UVS_ERR_CAPACITY, "uvs_window_upload: more than 1024 eligible points or lines");
  if (state->n_frames != R->n_frames || state->n_points != R->e_np || state->n_lines != R->e_nl)
    return handle_fail(h, UVS_ERR_INVALID_ARG, "uvs_window_upload: state sizes differ from the resident window (uvs_window_counts)");
  UvsWindow w = *state;
  w.n_proj = R->e_nproj; w.n_line_obs = R->e_nlobs; w.n_vp_obs = R->e_nvobs; w.n_imu = R->n_imu;
  w.proj_frame_i = w.proj_frame_j = w.proj_point = nullptr; w.proj_pts_i = w.proj_pts_j = nullptr;
  w.line_frame = w.line_idx = nullptr; w.line_sp = w.line_ep = nullptr; w.vp_frame = w.vp_line = nullptr; w.vp_dir = nullptr;
  w.imu_frame_i = nullptr; w.imu_delta_p = w.imu_delta_q = w.imu_delta_v = w.imu_sum_dt = w.imu_lin_ba = w.imu_lin_bg = w.imu_jacobian = w.imu_covariance = nullptr;
  w.estimate_td = 0;
  w.prior_n = R->prior_n; w.prior_n_blocks = (int)R->prior_kind.size();
  w.prior_J = nullptr; w.prior_r = nullptr;
  w.prior_block_kind = R->prior_kind.data(); w.prior_block_id = R->prior_id.data(); w.prior_x0 = R->prior_x0.data();
  ResidentHook hook{R, R->line_run_max, fill_from_store};
  return handle_upload(h, 1, &w, opts, &hook);
}

int uvs_window_marginalize(UvsHandle *h, int32_t flag, UvsPrior *out) {
  ResidentWindow *R = resident(h);
  if (!R) return handle_fail(h, UVS_ERR_INVALID_ARG, "uvs_window_marginalize: no resident window");
  if (!h->have_window) return handle_fail(h, UVS_ERR_NO_WINDOW, "uvs_window_marginalize: call uvs_window_upload first");
  const int cap = R->prior_cap, capb = 2 * (R->W + 1) + 2;
  std::vector<double> J, r, x0;
  std::vector<int> kind, id;
  UvsPrior tmp{};
  if (!out) {
    J.resize((size_t)cap * cap); r.resize(cap); x0.resize((size_t)9 * capb); kind.resize(capb); id.resize(capb);
    tmp.J = J.data(); tmp.r = r.data(); tmp.x0 = x0.data(); tmp.block_kind = kind.data(); tmp.block_id = id.data(); tmp.cap_n = cap; tmp.cap_blocks = capb;
    out = &tmp;
  }
  const int rc = uvs_marginalize(h, 0, flag, out);
  if (rc) return rc;
  // MARGIN_SECOND_NEW without the second newest pose in the prior: the reference builds nothing and KEEPS its prior
  // (estimator.cpp:1162-1164); frames 0 .. F-3 keep their indices through this slide, so the resident prior stays valid
  if (flag == UVS_MARGIN_SECOND_NEW && out->n == 0) return UVS_OK;
  if (out->n > cap) return handle_fail(h, UVS_ERR_CAPACITY, "uvs_window_marginalize: prior larger than the resident buffer");
  // the prior stays on the device: J0 / r0 straight from the marginalization's device result
  R->prior_n = out->n;
  R->prior_kind.assign(out->block_kind, out->block_kind + out->n_blocks);
  R->prior_id.assign(out->block_id, out->block_id + out->n_blocks);
  size_t xs = 0;
  for (int b = 0; b < out->n_blocks; b++) xs += (out->block_kind[b] == UVS_BLOCK_POSE || out->block_kind[b] == UVS_BLOCK_EXPOSE) ? 7 : (out->block_kind[b] == UVS_BLOCK_SPEEDBIAS ? 9 : 1);
  R->prior_x0.assign(out->x0, out->x0 + xs);
  if (out->n > 0) {
    if (!h->last_marg_J || h->last_marg_n != out->n) return handle_fail(h, UVS_ERR_CUDA, "uvs_window_marginalize: device result missing");
    CKW(cudaMemcpyAsync(R->d_prior_J, h->last_marg_J, (size_t)out->n * out->n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    CKW(cudaMemcpyAsync(R->d_prior_r, h->last_marg_r, (size_t)out->n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    CKW(cudaStreamSynchronize(h->stream));
  }
  return UVS_OK;
}

int uvs_window_slide(UvsHandle *h, int32_t flag, const UvsImuRecord *merged) {
  ResidentWindow *R = resident(h);
  if (!R) return handle_fail(h, UVS_ERR_INVALID_ARG, "uvs_window_slide: no resident window");
  if (flag != UVS_MARGIN_OLD && flag != UVS_MARGIN_SECOND_NEW) return handle_fail(h, UVS_ERR_INVALID_ARG, "uvs_window_slide: bad flag");
  if (R->n_frames < 2) return handle_fail(h, UVS_ERR_NO_WINDOW, "uvs_window_slide: fewer than two frames");
  CKW(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  const int fc = R->n_frames - 1;
  if (flag == UVS_MARGIN_OLD) {
    for (size_t k = 0; k < R->pt_order.size();) {   // removeBackShiftDepth / removeBack (feature_manager.cpp:607-663); the depth itself is caller state
      RTrack &t = R->pt[R->pt_order[k]];
      if (t.start != 0) { t.start--; k++; continue; }
      t.n--;
      if (t.n < 2) erase_track(R->pt, R->pt_order, R->pt_free, R->pt_slot, R->pt_order[k]); else k++;
    }
    for (size_t k = 0; k < R->ln_order.size();) {   // removeLineBack (feature_manager.cpp:687-703)
      RTrack &t = R->ln[R->ln_order[k]];
      if (t.start != 0) { t.start--; k++; continue; }
      t.n--; t.vp.erase(t.vp.begin());
      if (t.n == 0) erase_track(R->ln, R->ln_order, R->ln_free, R->ln_slot, R->ln_order[k]); else k++;
    }
    k_rw_slide_old<<<(R->max_pt + 127) / 128, 128, 0, st>>>(R->max_pt, 3, R->nobs, 2, R->d_pt_obs, R->d_pt_start, R->d_pt_n);
    k_rw_slide_old<<<(R->max_ln + 127) / 128, 128, 0, st>>>(R->max_ln, 7, R->nobs, 1, R->d_ln_obs, R->d_ln_start, R->d_ln_n);
    h->launches += 2;
    if (R->n_imu > 0) { R->imu_base = (R->imu_base + 1) % R->nobs; R->n_imu--; R->imu_dt.erase(R->imu_dt.begin()); }
  } else {
    if (R->n_imu >= 2 && !merged) return handle_fail(h, UVS_ERR_INVALID_ARG, "uvs_window_slide: MARGIN_SECOND_NEW needs the merged IMU record of the last two intervals");
    for (size_t k = 0; k < R->pt_order.size();) {   // removeFront (feature_manager.cpp:665-685)
      RTrack &t = R->pt[R->pt_order[k]];
      if (t.start == fc) { t.start--; k++; continue; }
      if (t.start + t.n - 1 < fc - 1) { k++; continue; }
      t.n--;
      if (t.n == 0) erase_track(R->pt, R->pt_order, R->pt_free, R->pt_slot, R->pt_order[k]); else k++;
    }
    for (size_t k = 0; k < R->ln_order.size();) {   // removeLineFront (feature_manager.cpp:705-723)
      RTrack &t = R->ln[R->ln_order[k]];
      if (t.start == fc) { t.start--; k++; continue; }
      if (t.start + t.n - 1 < fc - 1) { k++; continue; }
      t.vp.erase(t.vp.begin() + (fc - 1 - t.start)); t.n--;
      if (t.n == 0) erase_track(R->ln, R->ln_order, R->ln_free, R->ln_slot, R->ln_order[k]); else k++;
    }
    k_rw_slide_new<<<(R->max_pt + 127) / 128, 128, 0, st>>>(R->max_pt, 3, R->nobs, fc, R->d_pt_obs, R->d_pt_start, R->d_pt_n);
    k_rw_slide_new<<<(R->max_ln + 127) / 128, 128, 0, st>>>(R->max_ln, 7, R->nobs, fc, R->d_ln_obs, R->d_ln_start, R->d_ln_n);
    h->launches += 2;
    if (R->n_imu >= 2) {   // the last two intervals become one: pre_integrations[fc - 1]->push_back(...) (estimator.cpp:1301-1312)
      const UvsImuRecord &m = *merged;
      if (!m.delta_p || !m.delta_q || !m.delta_v || !m.sum_dt || !m.lin_ba || !m.lin_bg || !m.jacobian || !m.covariance)
        return handle_fail(h, UVS_ERR_INVALID_ARG, "uvs_window_slide: null IMU array");
      CKW(cudaStreamSynchronize(st));
      int rc = ensure_cmd(h, R, IMU_REC * sizeof(double)); if (rc) return rc;
      double *r = (double *)R->h_cmd;
      std::memcpy(r, m.delta_p, 24); std::memcpy(r + 3, m.delta_q, 32); std::memcpy(r + 7, m.delta_v, 24); r[10] = m.sum_dt[0];
      std::memcpy(r + 11, m.lin_ba, 24); std::memcpy(r + 14, m.lin_bg, 24); std::memcpy(r + 17, m.jacobian, 225 * 8); std::memcpy(r + 242, m.covariance, 225 * 8);
      CKW(cudaMemcpyAsync(R->d_imu + (size_t)((R->imu_base + R->n_imu - 2) % R->nobs) * IMU_REC, r, IMU_REC * sizeof(double), cudaMemcpyHostToDevice, st));
      h->h2d_bytes += (int64_t)(IMU_REC * sizeof(double));
      R->imu_dt[R->n_imu - 2] = m.sum_dt[0];
      R->imu_dt.pop_back();
      R->n_imu--;
    } else if (R->n_imu == 1) { R->n_imu = 0; R->imu_dt.clear(); }
  }
  R->n_frames--;
  h->have_window = false;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return handle_fail(h, UVS_ERR_CUDA, std::string("uvs_window_slide: ") + cudaGetErrorString(e));
  return UVS_OK;
}

int uvs_window_remove_tracks(UvsHandle *h, int32_t n_points, const int32_t *point_id, int32_t n_lines, const int32_t *line_id) {
  ResidentWindow *R = resident(h);
  if (!R || n_points < 0 || n_lines < 0 || (n_points && !point_id) || (n_lines && !line_id))
    return handle_fail(h, UVS_ERR_INVALID_ARG, "uvs_window_remove_tracks: bad arguments");
  CKW(cudaSetDevice(h->device));
  CKW(cudaStreamSynchronize(h->stream));
  std::vector<int> kp, kl;
  for (int i = 0; i < n_points; i++) { auto it = R->pt_slot.find(point_id[i]); if (it != R->pt_slot.end()) { kp.push_back(it->second); erase_track(R->pt, R->pt_order, R->pt_free, R->pt_slot, it->second); } }
  for (int i = 0; i < n_lines; i++) { auto it = R->ln_slot.find(line_id[i]); if (it != R->ln_slot.end()) { kl.push_back(it->second); erase_track(R->ln, R->ln_order, R->ln_free, R->ln_slot, it->second); } }
  const size_t bytes = (kp.size() + kl.size()) * sizeof(int);
  if (!bytes) return UVS_OK;
  int rc = ensure_cmd(h, R, bytes); if (rc) return rc;
  int *H = (int *)R->h_cmd;
  std::copy(kp.begin(), kp.end(), H); std::copy(kl.begin(), kl.end(), H + kp.size());
  CKW(cudaMemcpyAsync(R->d_cmd, H, bytes, cudaMemcpyHostToDevice, h->stream));
  h->h2d_bytes += (int64_t)bytes;
  if (!kp.empty()) { k_rw_kill<<<((int)kp.size() + 127) / 128, 128, 0, h->stream>>>((int)kp.size(), (const int *)R->d_cmd, R->d_pt_n); h->launches++; }
  if (!kl.empty()) { k_rw_kill<<<((int)kl.size() + 127) / 128, 128, 0, h->stream>>>((int)kl.size(), (const int *)R->d_cmd + kp.size(), R->d_ln_n); h->launches++; }
  h->have_window = false;
  return UVS_OK;
}

int uvs_download_factors(UvsHandle *h, int32_t window_index, UvsWindow *w) {
  if (!h || !w) return handle_fail(h, UVS_ERR_INVALID_ARG, "uvs_download_factors: bad arguments");
  if (!h->have_window) return handle_fail(h, UVS_ERR_NO_WINDOW, "uvs_download_factors: no window uploaded");
  if (window_index < 0 || window_index >= h->B) return handle_fail(h, UVS_ERR_INVALID_ARG, "uvs_download_factors: window index out of range");
  CKW(cudaSetDevice(h->device));
  const Dev &D = h->D;
  const int i = window_index;
  const int j0 = h->proj_off[i], np = h->proj_off[i + 1] - j0, a0 = h->lobs_off[i], na = h->lobs_off[i + 1] - a0, v0 = h->vobs_off[i], nv = h->vobs_off[i + 1] - v0;
  if (w->n_proj != np || w->n_line_obs != na || w->n_vp_obs != nv) return handle_fail(h, UVS_ERR_INVALID_ARG, "uvs_download_factors: factor counts differ from the upload");
  cudaStream_t st = h->stream;
  auto get = [&](const void *dst, const void *src, size_t bytes) -> cudaError_t {
    if (!bytes || !dst) return cudaSuccess;
    return cudaMemcpyAsync(const_cast<void *>(dst), src, bytes, cudaMemcpyDeviceToHost, st);
  };
  CKW(get(w->proj_frame_i, D.proj_fi + j0, np * sizeof(int))); CKW(get(w->proj_frame_j, D.proj_fj + j0, np * sizeof(int)));
  CKW(get(w->proj_point, D.proj_pt + j0, np * sizeof(int)));
  CKW(get(w->proj_pts_i, D.proj_pts_i + 3 * (size_t)j0, (size_t)np * 24)); CKW(get(w->proj_pts_j, D.proj_pts_j + 3 * (size_t)j0, (size_t)np * 24));
  CKW(get(w->line_frame, D.line_frame + a0, na * sizeof(int))); CKW(get(w->line_idx, D.line_idx + a0, na * sizeof(int)));
  CKW(get(w->line_sp, D.line_sp + 2 * (size_t)a0, (size_t)na * 16)); CKW(get(w->line_ep, D.line_ep + 2 * (size_t)a0, (size_t)na * 16));
  CKW(get(w->vp_frame, D.vp_frame + v0, nv * sizeof(int))); CKW(get(w->vp_line, D.vp_line + v0, nv * sizeof(int)));
  CKW(get(w->vp_dir, D.vp_dir + 3 * (size_t)v0, (size_t)nv * 24));
  const int m0 = h->imu_off[i], nm = h->imu_off[i + 1] - m0;
  if (w->n_imu == nm) {
    CKW(get(w->imu_frame_i, D.imu_frame + m0, nm * sizeof(int))); CKW(get(w->imu_delta_p, D.imu_dp + 3 * (size_t)m0, (size_t)nm * 24));
    CKW(get(w->imu_delta_q, D.imu_dq + 4 * (size_t)m0, (size_t)nm * 32)); CKW(get(w->imu_delta_v, D.imu_dv + 3 * (size_t)m0, (size_t)nm * 24));
    CKW(get(w->imu_sum_dt, D.imu_sum_dt + m0, (size_t)nm * 8)); CKW(get(w->imu_lin_ba, D.imu_lin_ba + 3 * (size_t)m0, (size_t)nm * 24));
    CKW(get(w->imu_lin_bg, D.imu_lin_bg + 3 * (size_t)m0, (size_t)nm * 24)); CKW(get(w->imu_jacobian, D.imu_jac + 225 * (size_t)m0, (size_t)nm * 1800));
    CKW(get(w->imu_covariance, D.imu_cov + 225 * (size_t)m0, (size_t)nm * 1800));
  }
  const int pn = h->prior_off[i + 1] - h->prior_off[i];
  if (w->prior_n == pn && pn > 0) {
    CKW(get(w->prior_J, D.prior_J + h->priorJ_off[i], (size_t)pn * pn * 8)); CKW(get(w->prior_r, D.prior_r0 + h->prior_off[i], (size_t)pn * 8));
  }
  CKW(cudaStreamSynchronize(st));
  return UVS_OK;
}

int64_t uvs_h2d_bytes(const UvsHandle *h) { return h ? h->h2d_bytes : 0; }

}  // extern "C"

// uvs_marg.cu — construction of the next marginalization prior on the device.
//
// Replaces the tail of Estimator::optimization() (vins_estimator/src/estimator.cpp:1003-1228) and
// MarginalizationInfo::{addResidualBlockInfo, preMarginalize, marginalize, getParameterBlocks}
// (factor/marginalization_factor.cpp:89-319):
//   1. every factor that touches a dropped block is evaluated at the current iterate (the factor
//      sweep kernels, tangent columns + loss correction = ResidualBlockInfo::Evaluate, :3-69);
//   2. A = sum J^T J, b = sum J^T r over [dropped | kept] columns (ThreadsConstructA, :141-172);
//   3. A_mm^+ by symmetric eigendecomposition with eigenvalues <= 1e-8 zeroed (:266-272),
//      A' = A_rr - A_rm A_mm^+ A_mr, b' likewise (:274-281);
//   4. second eigendecomposition: J0 = sqrt(S) V^T, r0 = S^-1/2 V^T b' (:283-291).
// Deliberate difference (the CPU checker makes the same choice): blocks are identified by (kind, id), not by host
// addresses, and their order is fixed (dropped: pose, speed-bias, points, lines; kept: pose_f,
// speedbias_f by frame, extrinsic, td), so that results are reproducible.
#include <cstring>
#include <vector>

#include "uvs_handle.h"
#include "uvs_kernels.h"

namespace uvs {

struct MargFactor {
  int type;          // 0 proj, 1 line, 2 vp, 3 imu
  int idx;           // global factor index (record row)
  int nr, nblk;
  int base[5], stride[5], width[5], aoff[5];   // J(row, b, c) = rec[base[b] + row * stride[b] + c] -> column aoff[b] + c
};

// A += J^T J, b += J^T r for one factor per CTA
__global__ void __launch_bounds__(128) k_marg_accum(Dev D, const MargFactor *__restrict__ facs, int pos, double *A, double *b) {
  const MargFactor f = facs[blockIdx.x];
  const double *rec;
  if (f.type == 0) rec = D.rec_proj + (size_t)f.idx * (D.estimate_td ? REC_PROJ_TD : REC_PROJ);
  else if (f.type == 1) rec = D.rec_line + (size_t)f.idx * REC_LINE;
  else if (f.type == 2) rec = D.rec_vp + (size_t)f.idx * REC_VP;
  else rec = D.rec_imu + (size_t)f.idx * REC_IMU;
  __shared__ int col_rec[40], col_stride[40], col_a[40];
  __shared__ int ncol;
  if (threadIdx.x == 0) {
    int n = 0;
    for (int k = 0; k < f.nblk; k++)
      for (int c = 0; c < f.width[k]; c++) { col_rec[n] = f.base[k] + c; col_stride[n] = f.stride[k]; col_a[n] = f.aoff[k] + c; n++; }
    ncol = n;
  }
  __syncthreads();
  const int n = ncol;
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    const int p = e / n, q = e - p * n;
    double h = 0.0;
    for (int row = 0; row < f.nr; row++) h += rec[col_rec[p] + row * col_stride[p]] * rec[col_rec[q] + row * col_stride[q]];
    atomicAdd(A + (size_t)col_a[p] * pos + col_a[q], h);
  }
  for (int p = threadIdx.x; p < n; p += blockDim.x) {
    double g = 0.0;
    for (int row = 0; row < f.nr; row++) g += rec[col_rec[p] + row * col_stride[p]] * rec[row];
    atomicAdd(b + col_a[p], g);
  }
}

// prior factor: A += J0^T J0 (precomputed), b += J0^T r(x)
__global__ void __launch_bounds__(256) k_marg_prior(Dev D, int w, const int *__restrict__ colmap, int pos, double *A, double *b) {
  const int n = D.prior_off[w + 1] - D.prior_off[w];
  const double *H = D.prior_H + D.priorJ_off[w], *J0 = D.prior_J + D.priorJ_off[w], *r = D.rec_prior + D.prior_off[w];
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n * n; e += gridDim.x * blockDim.x) {
    const int p = e / n, q = e - p * n;
    atomicAdd(A + (size_t)colmap[p] * pos + colmap[q], H[e]);
  }
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
    double g = 0.0;
    for (int i = 0; i < n; i++) g += J0[(size_t)i * n + p] * r[i];
    atomicAdd(b + colmap[p], g);
  }
}

// ---- one-CTA cyclic Jacobi eigensolver (round-robin ordering), A symmetric n x n in global memory.
// On exit diag(A) holds the eigenvalues and the columns of V the eigenvectors.
constexpr int ET = 1024;

__device__ void jacobi_eig(double *A, double *V, int n, double *cs /*[2*(n/2+1)]*/, int *pq /*[2*(n/2+1)]*/, double *red) {
  const int tid = threadIdx.x;
  for (int e = tid; e < n * n; e += ET) V[e] = (e / n == e % n) ? 1.0 : 0.0;
  __syncthreads();
  if (n < 2) return;
  const int N = (n & 1) ? n + 1 : n;   // players; index n is a bye when n is odd
  const int half = N / 2;
  for (int sweep = 0; sweep < 30; sweep++) {
    // convergence: largest off-diagonal against the largest diagonal magnitude
    double off = 0.0, dg = 0.0;
    for (int e = tid; e < n * n; e += ET) { const int i = e / n, j = e - i * n; const double v = fabs(A[e]); if (i == j) dg = fmax(dg, v); else off = fmax(off, v); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { off = fmax(off, __shfl_xor_sync(0xffffffffu, off, o)); dg = fmax(dg, __shfl_xor_sync(0xffffffffu, dg, o)); }
    if ((tid & 31) == 0) { red[tid >> 5] = off; red[32 + (tid >> 5)] = dg; }
    __syncthreads();
    off = 0.0; dg = 0.0;
    for (int k = 0; k < ET / 32; k++) { off = fmax(off, red[k]); dg = fmax(dg, red[32 + k]); }
    __syncthreads();
    if (off <= 2e-15 * dg || off == 0.0) break;
    for (int r = 0; r < N - 1; r++) {
      for (int k = tid; k < half; k += ET) {
        int p = k == 0 ? N - 1 : (r + k) % (N - 1);
        int q = k == 0 ? r : (r - k + (N - 1)) % (N - 1);
        if (p > q) { const int t = p; p = q; q = t; }
        double c = 1.0, s = 0.0;
        if (q < n) {
          const double apq = A[(size_t)p * n + q];
          if (apq != 0.0) {
            const double theta = (A[(size_t)q * n + q] - A[(size_t)p * n + p]) / (2.0 * apq);
            const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            c = rsqrt(t * t + 1.0); s = t * c;
          }
        } else { p = -1; }
        cs[2 * k] = c; cs[2 * k + 1] = s; pq[2 * k] = p; pq[2 * k + 1] = q;
      }
      __syncthreads();
      // columns p, q of A and V
      for (int e = tid; e < half * n; e += ET) {
        const int i = e / half, k = e - i * half;
        const int p = pq[2 * k], q = pq[2 * k + 1];
        if (p < 0) continue;
        const double c = cs[2 * k], s = cs[2 * k + 1];
        double *a = A + (size_t)i * n, *v = V + (size_t)i * n;
        const double aip = a[p], aiq = a[q];
        a[p] = c * aip - s * aiq; a[q] = s * aip + c * aiq;
        const double vip = v[p], viq = v[q];
        v[p] = c * vip - s * viq; v[q] = s * vip + c * viq;
      }
      __syncthreads();
      // rows p, q of A
      for (int e = tid; e < half * n; e += ET) {
        const int k = e / n, j = e - k * n;
        const int p = pq[2 * k], q = pq[2 * k + 1];
        if (p < 0) continue;
        const double c = cs[2 * k], s = cs[2 * k + 1];
        const double apj = A[(size_t)p * n + j], aqj = A[(size_t)q * n + j];
        A[(size_t)p * n + j] = c * apj - s * aqj; A[(size_t)q * n + j] = s * apj + c * aqj;
      }
      __syncthreads();
    }
  }
}

// Schur complement with the eigen-thresholded pseudo-inverse and the square-root factorisation.
// work: Amm[m*m] Vm[m*m] Ainv[m*m] T[n*m] Ap[n*n] Vn[n*n]
__global__ void __launch_bounds__(ET) k_marg_solve(const double *__restrict__ A, const double *__restrict__ b, int m, int n, double eps,
                                                   double *work, double *Aout, double *bout, double *Jout, double *rout) {
  extern __shared__ double sm[];
  const int pos = m + n, tid = threadIdx.x;
  double *Amm = work, *Vm = Amm + (size_t)m * m, *Ainv = Vm + (size_t)m * m, *T = Ainv + (size_t)m * m;
  double *Ap = T + (size_t)n * m, *Vn = Ap + (size_t)n * n;
  const int mx = (m > n ? m : n) / 2 + 2;
  double *cs = sm; int *pq = reinterpret_cast<int *>(cs + 2 * mx); double *red = reinterpret_cast<double *>(pq + 2 * mx + 2);
  // Amm = 0.5 (Amm + Amm^T)                                                   marginalization_factor.cpp:266
  for (int e = tid; e < m * m; e += ET) { const int i = e / m, j = e - i * m; Amm[e] = 0.5 * (A[(size_t)i * pos + j] + A[(size_t)j * pos + i]); }
  __syncthreads();
  if (m > 0) jacobi_eig(Amm, Vm, m, cs, pq, red);
  __syncthreads();
  // Ainv = V diag(1/lambda if lambda > eps else 0) V^T                          :269-272
  for (int e = tid; e < m * m; e += ET) {
    const int i = e / m, j = e - i * m;
    double s = 0.0;
    for (int k = 0; k < m; k++) { const double lam = Amm[(size_t)k * m + k]; if (lam > eps) s += Vm[(size_t)i * m + k] * Vm[(size_t)j * m + k] / lam; }
    Ainv[e] = s;
  }
  __syncthreads();
  // T = Arm Ainv
  for (int e = tid; e < n * m; e += ET) {
    const int i = e / m, j = e - i * m;
    double s = 0.0;
    for (int k = 0; k < m; k++) s += A[(size_t)(m + i) * pos + k] * Ainv[(size_t)k * m + j];
    T[e] = s;
  }
  __syncthreads();
  // A' = Arr - T Amr, b' = br - T bm                                            :274-281
  for (int e = tid; e < n * n; e += ET) {
    const int i = e / n, j = e - i * n;
    double s = 0.0;
    for (int k = 0; k < m; k++) s += T[(size_t)i * m + k] * A[(size_t)k * pos + m + j];
    Ap[e] = A[(size_t)(m + i) * pos + m + j] - s;
  }
  for (int i = tid; i < n; i += ET) {
    double s = 0.0;
    for (int k = 0; k < m; k++) s += T[(size_t)i * m + k] * b[k];
    bout[i] = b[m + i] - s;
  }
  __syncthreads();
  for (int e = tid; e < n * n; e += ET) Aout[e] = Ap[e];
  __syncthreads();
  // SelfAdjointEigenSolver reads the lower triangle                              :283
  for (int e = tid; e < n * n; e += ET) { const int i = e / n, j = e - i * n; if (j > i) Ap[e] = Aout[(size_t)j * n + i]; }
  __syncthreads();
  jacobi_eig(Ap, Vn, n, cs, pq, red);
  __syncthreads();
  // J0 = sqrt(S) V^T, r0 = S^-1/2 V^T b'                                          :285-291
  for (int e = tid; e < n * n; e += ET) {
    const int k = e / n, j = e - k * n;
    const double lam = Ap[(size_t)k * n + k];
    Jout[e] = lam > eps ? sqrt(lam) * Vn[(size_t)j * n + k] : 0.0;
  }
  for (int k = tid; k < n; k += ET) {
    const double lam = Ap[(size_t)k * n + k];
    double s = 0.0;
    for (int j = 0; j < n; j++) s += Vn[(size_t)j * n + k] * bout[j];
    rout[k] = lam > eps ? s / sqrt(lam) : 0.0;
  }
}

// camera state of one window at its current iterate: [pose F x 7 | sb F x 9 | ex 7 | td 1]
__global__ void k_gather_window_cam(Dev D, int w, double *out) {
  const int F = D.frame_off[w + 1] - D.frame_off[w], fo = D.frame_off[w], c = D.cur[w];
  for (int e = threadIdx.x; e < 7 * F; e += blockDim.x) out[e] = D.pose[c][7 * (size_t)fo + e];
  for (int e = threadIdx.x; e < 9 * F; e += blockDim.x) out[7 * F + e] = D.sb[c][9 * (size_t)fo + e];
  if (threadIdx.x < 7) out[16 * F + threadIdx.x] = D.ex[c][7 * (size_t)w + threadIdx.x];
  if (threadIdx.x == 7) out[16 * F + 7] = D.td[c][w];
}

}  // namespace uvs

using namespace uvs;

#define CKM(call)                                                                                   \
  do {                                                                                              \
    cudaError_t e_ = (call);                                                                        \
    if (e_ != cudaSuccess) return handle_fail(h, UVS_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

int uvs_marginalize_impl(UvsHandle *h, int wi, int flag, UvsPrior *out) {
  h->last_marg_J = h->last_marg_r = nullptr; h->last_marg_n = 0;
  if (wi < 0 || wi >= h->B) return handle_fail(h, UVS_ERR_INVALID_ARG, "uvs_marginalize: window index out of range");
  if (flag != UVS_MARGIN_OLD && flag != UVS_MARGIN_SECOND_NEW) return handle_fail(h, UVS_ERR_INVALID_ARG, "uvs_marginalize: bad flag");
  if (h->nranks > 1) return handle_fail(h, UVS_ERR_UNSUPPORTED, "uvs_marginalize: not available in the factor-parallel multi-GPU mode");
  CKM(cudaSetDevice(h->device));
  const Dev &D = h->D;
  cudaStream_t st = h->stream;
  // host views of the caller's index arrays (the pinned staging buffer mirrors the device input region)
  auto host = [&](const void *devp) -> const char * { return h->stage.base + ((const char *)devp - h->dev.base); };
  // a relocalisation pose (uvs.h, n_relo > 0) is the last internal frame: the reference's marginalization knows neither the
  // block nor its factors (estimator.cpp:1003-1228), so the plan runs over the window's own frames and skips those factors
  const int Fint = h->frame_off[wi + 1] - h->frame_off[wi];
  const int F = Fint - ((int)h->relo.size() > wi ? h->relo[wi] : 0);
  const int j0 = h->proj_off[wi], nproj = h->proj_off[wi + 1] - j0;
  const int a0 = h->lobs_off[wi], nlobs = h->lobs_off[wi + 1] - a0;
  const int v0 = h->vobs_off[wi], nvobs = h->vobs_off[wi + 1] - v0;
  const int m0 = h->imu_off[wi], nimu = h->imu_off[wi + 1] - m0;
  const int np = h->point_off[wi + 1] - h->point_off[wi], nl = h->line_off[wi + 1] - h->line_off[wi];
  const int b0 = h->pblk_off[wi], nblk = h->pblk_off[wi + 1] - b0;
  const int prior_n = h->prior_off[wi + 1] - h->prior_off[wi];
  const bool ex_est = (h->win_flags[wi] & WF_EXTRINSIC) != 0;
  (void)ex_est;
  const bool td = D.estimate_td != 0;
  const int *pfi = (const int *)host(D.proj_fi) + j0, *pfj = (const int *)host(D.proj_fj) + j0, *ppt = (const int *)host(D.proj_pt) + j0;
  const int *lf = (const int *)host(D.line_frame) + a0, *li = (const int *)host(D.line_idx) + a0;
  const int *vf = (const int *)host(D.vp_frame) + v0, *vl = (const int *)host(D.vp_line) + v0;
  const int *imf = (const int *)host(D.imu_frame) + m0;
  const double *imdt = (const double *)host(D.imu_sum_dt) + m0;
  const int *bk = (const int *)host(D.pblk_kind) + b0, *bi = (const int *)host(D.pblk_id) + b0;

  // ---- which blocks take part, which are dropped
  std::vector<char> has_pose(F, 0), has_sb(F, 0), has_pt(np, 0), has_ln(nl, 0);
  bool has_ex = false, has_td = false, use_prior = false;
  std::vector<int> sel_imu, sel_proj, sel_line, sel_vp;
  int drop_pose = -1, drop_sb = -1;
  if (flag == UVS_MARGIN_OLD) {
    drop_pose = 0; drop_sb = 0;
    if (prior_n > 0) use_prior = true;
    for (int k = 0; k < nimu; k++) if (imf[k] == 0 && imdt[k] < 10.0) { sel_imu.push_back(k); has_pose[0] = has_sb[0] = 1; if (F > 1) has_pose[1] = has_sb[1] = 1; }
    for (int k = 0; k < nproj; k++) if (pfi[k] == 0 && pfj[k] < F) { sel_proj.push_back(k); has_pose[0] = 1; has_pose[pfj[k]] = 1; has_ex = true; has_pt[ppt[k]] = 1; if (td) has_td = true; }
    std::vector<int> start(nl, 1 << 30);
    for (int k = 0; k < nlobs; k++) start[li[k]] = std::min(start[li[k]], lf[k]);
    for (int k = 0; k < nlobs; k++) if (start[li[k]] == 0 && lf[k] != 0) { sel_line.push_back(k); has_pose[lf[k]] = 1; has_ln[li[k]] = 1; }
    for (int k = 0; k < nvobs; k++) if (start[vl[k]] == 0 && vf[k] != 0) { sel_vp.push_back(k); has_pose[vf[k]] = 1; has_ln[vl[k]] = 1; }
  } else {
    bool has = false;   // estimator.cpp:1162-1164
    for (int b = 0; b < (prior_n > 0 ? nblk : 0); b++) if (bk[b] == UVS_BLOCK_POSE && bi[b] == F - 2) has = true;
    if (!has) { out->n = 0; out->n_blocks = 0; out->m = 0; return UVS_OK; }
    drop_pose = F - 2;
    use_prior = true;
  }
  if (use_prior)
    for (int b = 0; b < nblk; b++) {
      if (bk[b] == UVS_BLOCK_POSE) has_pose[bi[b]] = 1;
      else if (bk[b] == UVS_BLOCK_SPEEDBIAS) has_sb[bi[b]] = 1;
      else if (bk[b] == UVS_BLOCK_EXPOSE) has_ex = true;
      else has_td = true;
    }
  if (!use_prior && sel_imu.empty() && sel_proj.empty() && sel_line.empty() && sel_vp.empty()) { out->n = 0; out->n_blocks = 0; out->m = 0; return UVS_OK; }

  // ---- column assignment: dropped first
  std::vector<int> col_pose(F, -1), col_sb(F, -1), col_pt(np, -1), col_ln(nl, -1);
  int col_ex = -1, col_td = -1, pos = 0;
  if (drop_pose >= 0 && has_pose[drop_pose]) { col_pose[drop_pose] = pos; pos += 6; }
  if (drop_sb >= 0 && has_sb[drop_sb]) { col_sb[drop_sb] = pos; pos += 9; }
  for (int k = 0; k < np; k++) if (has_pt[k]) { col_pt[k] = pos; pos += 1; }
  for (int k = 0; k < nl; k++) if (has_ln[k]) { col_ln[k] = pos; pos += 4; }
  const int m = pos;
  std::vector<int> kept_kind, kept_id;
  for (int f = 0; f < F; f++) {
    if (has_pose[f] && col_pose[f] < 0) { col_pose[f] = pos; pos += 6; kept_kind.push_back(UVS_BLOCK_POSE); kept_id.push_back(f); }
    if (has_sb[f] && col_sb[f] < 0) { col_sb[f] = pos; pos += 9; kept_kind.push_back(UVS_BLOCK_SPEEDBIAS); kept_id.push_back(f); }
  }
  if (has_ex) { col_ex = pos; pos += 6; kept_kind.push_back(UVS_BLOCK_EXPOSE); kept_id.push_back(0); }
  if (has_td) { col_td = pos; pos += 1; kept_kind.push_back(UVS_BLOCK_TD); kept_id.push_back(0); }
  const int n = pos - m, nkept = (int)kept_kind.size();
  if (n > out->cap_n || nkept > out->cap_blocks) return handle_fail(h, UVS_ERR_CAPACITY, "uvs_marginalize: output capacity too small");

  // ---- factor plan
  std::vector<MargFactor> facs;
  auto blockset = [](MargFactor &f, int k, int base, int stride, int width, int aoff) { f.base[k] = base; f.stride[k] = stride; f.width[k] = width; f.aoff[k] = aoff; };
  for (int k : sel_imu) {
    MargFactor f{}; f.type = 3; f.idx = m0 + k; f.nr = 15; f.nblk = 4;
    const int fi = imf[k];
    blockset(f, 0, 15, 30, 6, col_pose[fi]); blockset(f, 1, 21, 30, 9, col_sb[fi]);
    blockset(f, 2, 30, 30, 6, col_pose[fi + 1]); blockset(f, 3, 36, 30, 9, col_sb[fi + 1]);
    facs.push_back(f);
  }
  for (int k : sel_proj) {
    MargFactor f{}; f.type = 0; f.idx = j0 + k; f.nr = 2; f.nblk = td ? 5 : 4;
    blockset(f, 0, 2, 6, 6, col_pose[pfi[k]]); blockset(f, 1, 14, 6, 6, col_pose[pfj[k]]); blockset(f, 2, 26, 6, 6, col_ex);
    blockset(f, 3, 38, 1, 1, col_pt[ppt[k]]);
    if (td) blockset(f, 4, 40, 1, 1, col_td);
    facs.push_back(f);
  }
  for (int k : sel_line) {
    MargFactor f{}; f.type = 1; f.idx = a0 + k; f.nr = 2; f.nblk = 2;
    blockset(f, 0, 2, 6, 6, col_pose[lf[k]]); blockset(f, 1, 14, 4, 4, col_ln[li[k]]);
    facs.push_back(f);
  }
  for (int k : sel_vp) {
    MargFactor f{}; f.type = 2; f.idx = v0 + k; f.nr = 1; f.nblk = 2;
    blockset(f, 0, 1, 0, 6, col_pose[vf[k]]); blockset(f, 1, 7, 0, 4, col_ln[vl[k]]);
    facs.push_back(f);
  }
  std::vector<int> prior_colmap(std::max(prior_n, 1), 0);
  if (use_prior) {
    int c = 0;
    for (int b = 0; b < nblk; b++) {
      const int ls = bk[b] == UVS_BLOCK_POSE || bk[b] == UVS_BLOCK_EXPOSE ? 6 : (bk[b] == UVS_BLOCK_SPEEDBIAS ? 9 : 1);
      const int base = bk[b] == UVS_BLOCK_POSE ? col_pose[bi[b]] : (bk[b] == UVS_BLOCK_SPEEDBIAS ? col_sb[bi[b]] : (bk[b] == UVS_BLOCK_EXPOSE ? col_ex : col_td));
      for (int k = 0; k < ls; k++) prior_colmap[c + k] = base + k;
      c += ls;
    }
  }

  // ---- device scratch: [A pos^2 | b pos | work | Aout n^2 | bout n | J n^2 | r n | cam state | plan]
  auto al = [](size_t v) { return (v + 255) / 256 * 256; };
  const size_t Dd = sizeof(double);
  size_t o = 0;
  const size_t oA = o; o += al((size_t)pos * pos * Dd);
  const size_t ob = o; o += al((size_t)pos * Dd);
  const size_t oW = o; o += al(((size_t)3 * m * m + (size_t)n * m + (size_t)2 * n * n + 8) * Dd);
  const size_t oAo = o; o += al((size_t)n * n * Dd);
  const size_t obo = o; o += al((size_t)n * Dd);
  const size_t oJ = o; o += al((size_t)n * n * Dd);
  const size_t orr = o; o += al((size_t)n * Dd);
  const size_t ocam = o; o += al((size_t)(16 * Fint + 8) * Dd);
  const size_t ofac = o; o += al(std::max<size_t>(facs.size(), 1) * sizeof(MargFactor));
  const size_t ocm = o; o += al(prior_colmap.size() * sizeof(int));
  int rc = handle_ensure_scratch(h, o); if (rc) return rc;
  rc = handle_ensure_hscratch(h, o); if (rc) return rc;
  char *ds = h->scratch.base, *hs = h->hscratch.base;
  CKM(cudaMemsetAsync(ds + oA, 0, ob + al((size_t)pos * Dd) - oA, st));
  if (!facs.empty()) { std::memcpy(hs + ofac, facs.data(), facs.size() * sizeof(MargFactor)); CKM(cudaMemcpyAsync(ds + ofac, hs + ofac, facs.size() * sizeof(MargFactor), cudaMemcpyHostToDevice, st)); }
  std::memcpy(hs + ocm, prior_colmap.data(), prior_colmap.size() * sizeof(int));
  CKM(cudaMemcpyAsync(ds + ocm, hs + ocm, prior_colmap.size() * sizeof(int), cudaMemcpyHostToDevice, st));

  // ---- 1. evaluate every factor at the current iterate (tangent columns, loss-corrected).  The sweeps cover the whole
  //         batch, so the records stay valid for the other windows of the batch: marginalizing B windows costs one sweep,
  //         not B (the epoch moves with every upload, solve, state change or evaluation call)
  if (h->marg_epoch != h->records_epoch) {
    h->launches += launch_proj(D, h->P, true, false, 0, 0, D.rec_proj, nullptr, nullptr, 0, st);
    h->launches += launch_line_tables(D, 0, st);
    h->launches += launch_line_vp(D, h->P, true, 0, 0, D.rec_line, D.rec_vp, nullptr, 0, st);
    h->launches += launch_imu(D, h->P, true, 0, 0, D.rec_imu, nullptr, nullptr, 0, st);
    h->launches += launch_prior(D, h->max_prior_n, false, 0, 0, D.rec_prior, nullptr, 0, st);
    h->marg_epoch = h->records_epoch;
  }
  // ---- 2. A, b
  double *A = (double *)(ds + oA), *b = (double *)(ds + ob);
  if (!facs.empty()) { k_marg_accum<<<(int)facs.size(), 128, 0, st>>>(D, (const MargFactor *)(ds + ofac), pos, A, b); h->launches++; }
  if (use_prior) { k_marg_prior<<<8, 256, 0, st>>>(D, wi, (const int *)(ds + ocm), pos, A, b); h->launches++; }
  // ---- 3./4. Schur complement + square-root factor
  const int mx = std::max(m, n) / 2 + 2;
  const size_t smem = (size_t)2 * mx * Dd + (size_t)(2 * mx + 2) * sizeof(int) + 64 * Dd + 16;
  k_marg_solve<<<1, ET, smem, st>>>(A, b, m, n, 1e-8, (double *)(ds + oW), (double *)(ds + oAo), (double *)(ds + obo), (double *)(ds + oJ), (double *)(ds + orr));
  h->launches++;
  k_gather_window_cam<<<1, 256, 0, st>>>(D, wi, (double *)(ds + ocam));
  h->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return handle_fail(h, UVS_ERR_CUDA, std::string("uvs_marginalize kernels: ") + cudaGetErrorString(e));
  CKM(cudaMemcpyAsync(hs + oAo, ds + oAo, ocam + al((size_t)(16 * Fint + 8) * Dd) - oAo, cudaMemcpyDeviceToHost, st));
  CKM(cudaStreamSynchronize(st));

  h->last_marg_J = (const double *)(ds + oJ); h->last_marg_r = (const double *)(ds + orr); h->last_marg_n = n;   // stays valid until the scratch arena is used again
  // ---- outputs (getParameterBlocks with the window shift, estimator.cpp:1139-1153 / 1199-1222)
  out->n = n; out->m = m; out->n_blocks = nkept;
  std::memcpy(out->J, hs + oJ, (size_t)n * n * Dd);
  std::memcpy(out->r, hs + orr, (size_t)n * Dd);
  if (out->A) std::memcpy(out->A, hs + oAo, (size_t)n * n * Dd);
  if (out->b) std::memcpy(out->b, hs + obo, (size_t)n * Dd);
  const double *cam = (const double *)(hs + ocam);
  size_t xo = 0;
  for (int k = 0; k < nkept; k++) {
    const int kind = kept_kind[k];
    int id = kept_id[k];
    const double *src; int gs;
    if (kind == UVS_BLOCK_POSE) { src = cam + 7 * id; gs = 7; }
    else if (kind == UVS_BLOCK_SPEEDBIAS) { src = cam + 7 * Fint + 9 * id; gs = 9; }
    else if (kind == UVS_BLOCK_EXPOSE) { src = cam + 16 * Fint; gs = 7; }
    else { src = cam + 16 * Fint + 7; gs = 1; }
    if (kind <= UVS_BLOCK_SPEEDBIAS) {
      if (flag == UVS_MARGIN_OLD) id -= 1;
      else if (id == F - 1) id -= 1;
    }
    out->block_kind[k] = kind; out->block_id[k] = id;
    std::memcpy(out->x0 + xo, src, gs * Dd);
    xo += gs;
  }
  return UVS_OK;
}

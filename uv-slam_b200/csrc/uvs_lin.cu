// uvs_lin.cu — fused linearisation of the landmark factors: Jacobian evaluation + landmark elimination + direct
// terms in ONE pass per landmark type, so that no point / line / VP Jacobian record is ever written to HBM
// ("fully fused (never write J to HBM) for the iteration-rate config").
//
// What Ceres does per residual block behind ceres::Solve (vins_estimator/src/estimator.cpp:982-994): Evaluate
// (ProjectionFactor projection_factor.cpp:22-175, LineProjectionFactor / VPProjectionFactor via AutoDiff,
// line_projection_factor.h:16-60, vp_projection_factor.h:19-66), loss correction (marginalization_factor.cpp:37-68
// restates it), J^T J accumulation and the SPARSE_SCHUR elimination of the landmark blocks.
//
//   k_lin_points   a LANE owns a point, step k of the loop evaluates the k-th factor of every lane's point.  Points are
//                  processed in an order sorted by (anchor frame, track length) (k_prep_point_order, once per upload), so
//                  the 32 factors of a step mostly share their camera-block pair (i, j): their [r | Ji | Jj] rows sit in a
//                  shared-memory stage and G^T G (G = [J_i | J_j | r]) of the whole group is ONE chain of FP64 tensor-core
//                  MMAs, flushed with ~3 reductions per factor into the window's system.  The lane keeps the point's
//                  sums (E, g, W_i), parks the unscaled W_j blocks in the Y stash and rescales them at the end.
//   k_lin_lines    a lane GROUP (8 lanes) owns a line, one observation (+ its VP factor) per lane, evaluated through the
//                  per-frame tables of uvs_linefast.cuh into a shared-memory stage; the elimination of the 4x4 block
//                  then reads the stage instead of HBM records, and the diagonal direct terms are grouped by frame on
//                  the tensor cores.
// Both write the same stash (uvs_stash.cuh) as k_core_points / k_core_lines of uvs_build3.cu; the rank update
// (k_window_system), the IMU / prior tail and the back-substitution are shared with that path.
#include <algorithm>
#include <cstdlib>

#include "uvs_device.cuh"
#include "uvs_factors.cuh"
#include "uvs_kernels.h"
#include "uvs_linefast.cuh"
#include "uvs_stash.cuh"

namespace uvs {

// ------------------------------------------------------------------------------------------------
// Processing order of the points: every WARP of k_lin_points gets points of ONE window with ONE anchor frame, sorted by
// track length (descending), so that step k of a warp sees a single camera-block pair (i, i + 1 + k) for the usual
// consecutive tracks.  Window w owns the warp slots [pw_off[w], pw_off[w + 1]) (host bound: ceil(np / 32) + frames);
// pt_order[32 slot + lane] = global point or -1 (preset by a memset).  One CTA per window; `key` is scratch [nP].
__global__ void __launch_bounds__(256) k_prep_point_order(Dev D, int *__restrict__ key, int dense) {
  __shared__ int cnt[33], wstart[34];
  const int w = blockIdx.x;
  const int p0 = D.point_off[w], np = D.point_off[w + 1] - p0, fo = D.frame_off[w];
  if (threadIdx.x < 33) cnt[threadIdx.x] = 0;
  __syncthreads();
  for (int p = threadIdx.x; p < np; p += blockDim.x) {
    const int gp = p0 + p, n = D.pt_end[gp] - D.pt_begin[gp];
    int k = -1;
    if (n > 0) {
      const int a = min(max(D.proj_idx[D.pt_begin[gp]].x - fo, 0), 31);
      k = a << 8 | (255 - min(n, 255));
      atomicAdd(&cnt[a], 1);
    }
    key[gp] = k;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    // dense: the anchors share warps (fewer, fuller warps; a warp that straddles two anchors sees two block pairs per step)
    for (int a = 0; a < 32; a++) { wstart[a] = run; run += dense ? cnt[a] : ((cnt[a] + 31) >> 5); }
    wstart[32] = run;
  }
  __syncthreads();
  const int wbase = D.pw_off[w], wcap = D.pw_off[w + 1] - wbase;
  for (int p = threadIdx.x; p < np; p += blockDim.x) {
    const int kp = key[p0 + p];
    if (kp < 0) continue;
    int rank = 0;   // among the points of the same anchor
    for (int q = 0; q < np; q++) { const int kq = key[p0 + q]; rank += (kq >= 0 && (kq >> 8) == (kp >> 8) && (kq < kp || (kq == kp && q < p))) ? 1 : 0; }
    if (dense) {
      const int pos = wstart[kp >> 8] + rank;   // position in the window's sorted list
      D.pt_order[32 * (size_t)wbase + pos] = p0 + p;
    } else {
      const int slot = wstart[kp >> 8] + (rank >> 5);
      if (slot < wcap) D.pt_order[32 * (size_t)(wbase + slot) + (rank & 31)] = p0 + p;
    }
  }
}

// ------------------------------------------------------------------------------------------------
constexpr int PST = 27;      // stage row: [r(2) | Ji 2x6 | Jj 2x6] + 1 (odd stride: conflict-free stores)

// LP_NT threads per CTA (32: one warp slot per CTA - a finished warp frees its registers at once, whatever the track
// lengths of its neighbours; 128: four), kPrefetch: the inputs of step k + 1 are loaded before step k is evaluated
template <int LP_NT, bool kPrefetch, int kThreadsPerSM = 512>
__global__ void __launch_bounds__(LP_NT, kThreadsPerSM / LP_NT) k_lin_points(Dev D, Params P, Stash S) {
  __shared__ double stage_all[LP_NT * PST];
  __shared__ unsigned char mlist_all[LP_NT / 32][32];
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double *stage = stage_all + warp * 32 * PST;
  double *row = stage + lane * PST;
  unsigned char *mlist = mlist_all[warp];
  const int t = blockIdx.x * LP_NT + threadIdx.x;
  int gp = (t >> 5) < D.nPW ? D.pt_order[t] : -1;
  bool act = gp >= 0;
  int w = 0, f0 = 0, n = 0;
  if (act) { w = D.pt_win[gp]; act = (D.ctl[w].state & WS_ACTIVE) != 0; }
  const int mp = S.mp;
  double *Y = nullptr, *ph = nullptr;
  if (act) {
    Y = S.Y + (colbase(D, w) + (gp - D.point_off[w])) * mp;
    ph = S.ph + 4 * (size_t)gp;
    f0 = D.pt_begin[gp]; n = D.pt_end[gp] - f0;
    const bool mine = D.nranks <= 1 || (gp % D.nranks) == D.rank;
    // the column was zero-filled at upload and its sparsity pattern never changes: only the blocks are rewritten
    if (n <= 0 || !mine) { ph[1] = 0.0; act = false; }
  }
  if (!act) n = 0;
  const int nmax = __reduce_max_sync(full, n);
  if (nmax == 0) return;   // uniform over the warp
  int fo = 0, co = 0, d = 0, cur = 0;
  long long s_off = 0;
  bool need_cost = false;   // the cost of the linearisation point is only used at iteration 0 (uvs_math.cuh cauchy_rho1)
  if (act) { fo = D.frame_off[w]; co = D.cam_off[w]; d = D.cam_off[w + 1] - co; s_off = D.S_off[w]; cur = D.cur[w]; need_cost = D.ctl[w].iter == 0; }
  const double *ex = D.ex[cur] + 7 * (size_t)w;
  const double lam = act ? D.inv_depth[cur][gp] : 1.0;
  double colsq = 0.0, gk = 0.0, half = 0.0;
  double wa[6] = {0, 0, 0, 0, 0, 0};
  int row_i = 0;
  // Direct terms: G = [J_i (6) | r | 0 || J_j (6) | 0 0], two rows per factor.  G^T G on the FP64 tensor cores, four rows
  // (two factors) per k-step: c00 = (i,i) block and, in column 6, the gradient of frame i; c01 = (i,j) block and, in row 6,
  // the gradient of frame j; c11 = (j,j) block.  The A fragment of a column tile and its B fragment are the same value
  // G[k = lane % 4][8 tile + lane / 4].  c00 belongs to the anchor frame, which a warp keeps over all steps: it is flushed
  // only when the anchor changes (never, for the sorted order above).
  const int fcol = lane >> 2, frow = lane & 1, fsub = (lane >> 1) & 1;
  const int o0 = fcol < 6 ? 2 + 6 * frow + fcol : (fcol == 6 ? frow : -1);
  const int o1 = fcol < 6 ? 14 + 6 * frow + fcol : -1;
  const int pr = lane >> 2, pc = 2 * (lane & 3);   // accumulator entries of this lane: (pr, pc), (pr, pc + 1)
  double c00[2] = {0.0, 0.0};
  int pk_i = -1, pk_fo = 0, pk_co = 0, pk_d = 0;   // anchor (pose row) c00 belongs to, and its window
  long long pk_so = 0;
  auto flush00 = [&]() {
    if (pk_i < 0) return;
    const int ra = 15 * (pk_i - pk_fo);
    double *Sg = D.Smat + pk_so;
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int q = pc + e;
      if (pr < 6 && q >= pr && q < 6) {
        atomicAdd(Sg + (size_t)(ra + pr) * pk_d + ra + q, c00[e]);
        if (q == pr) atomicAdd(D.colsq_cam + pk_co + ra + pr, c00[e]);
      } else if (pr < 6 && q == 6) {
        atomicAdd(D.gS + pk_co + ra + pr, c00[e]); atomicAdd(D.gfull + pk_co + ra + pr, c00[e]);
      }
    }
    c00[0] = c00[1] = 0.0;
  };
  // inputs of step 0 (the loads of step k + 1 are issued before step k is evaluated)
  int4 ix = make_int4(-1, -1, 0, 0);
  d3 pts_i = mk3(0, 0, 1), pts_j = mk3(0, 0, 1);
  auto load_step = [&](int k) {
    const int f = f0 + k;
    ix = D.proj_idx[f];
    const double *oi = D.proj_pts_i + 3 * (size_t)f, *oj = D.proj_pts_j + 3 * (size_t)f;
    pts_i = mk3(__ldg(oi), __ldg(oi + 1), __ldg(oi + 2)); pts_j = mk3(__ldg(oj), __ldg(oj + 1), __ldg(oj + 2));
  };
  if (kPrefetch && n > 0) load_step(0);
  for (int k = 0; k < nmax; k++) {
    const bool has = k < n;
    int ri = -1, rj = -1;
    if (has) {
      if (!kPrefetch) load_step(k);
      ri = ix.x; rj = ix.y; row_i = ri;
      const d3 pi = pts_i, pj = pts_j;
      if (kPrefetch && k + 1 < n) load_step(k + 1);
      double jl[2], hr = 0.0;
      proj_eval<true, false>(D.pose[cur] + 7 * (size_t)ri, D.pose[cur] + 7 * (size_t)rj, ex, lam, pi, pj, P.S, nullptr, false,
                             P.cauchy_point, true, 6, row, row + 2, row + 14, nullptr, jl, nullptr, need_cost ? &hr : nullptr);
      half += hr;
      const double j0 = jl[0], j1 = jl[1];
      colsq += j0 * j0 + j1 * j1;
      gk += j0 * row[0] + j1 * row[1];
      double u[6];
#pragma unroll
      for (int c = 0; c < 6; c++) {
        wa[c] += row[2 + c] * j0 + row[8 + c] * j1;
        u[c] = row[14 + c] * j0 + row[20 + c] * j1;
      }
      // W_j = Jj^T Jl goes to the stash unscaled: the column carries its scale (uvs_stash.cuh), nothing is rewritten later
      double2 *y2 = reinterpret_cast<double2 *>(Y + 6 * (rj - fo));
      y2[0] = make_double2(u[0], u[1]); y2[1] = make_double2(u[2], u[3]); y2[2] = make_double2(u[4], u[5]);
    }
    __syncwarp();
    // ---- direct terms of this step's factors, grouped by camera-block pair (one group per step for sorted consecutive tracks)
    unsigned todo = __ballot_sync(full, has);
    while (todo) {
      const int leader = __ffs(todo) - 1;
      const int li = __shfl_sync(full, ri, leader), lj = __shfl_sync(full, rj, leader);
      const unsigned grp = __ballot_sync(full, has && ri == li && rj == lj);
      todo &= ~grp;
      const int m = __popc(grp);
      const bool contiguous = grp == ((m == 32 ? full : ((1u << m) - 1u)) << leader);   // members = lanes leader .. leader + m - 1
      if (!contiguous) {
        if (grp >> lane & 1u) mlist[__popc(grp & ((1u << lane) - 1u))] = (unsigned char)lane;
        __syncwarp();
      }
      if (li != pk_i) {   // uniform
        flush00();
        pk_i = li; pk_fo = __shfl_sync(full, fo, leader); pk_co = __shfl_sync(full, co, leader); pk_d = __shfl_sync(full, d, leader);
        pk_so = __shfl_sync(full, s_off, leader);
      }
      double c01[2] = {0.0, 0.0}, c11[2] = {0.0, 0.0};
      for (int s = 0; 2 * s < m; s++) {
        const int it = 2 * s + fsub;
        double g0 = 0.0, g1 = 0.0;
        if (it < m) {
          const double *rec = stage + (contiguous ? leader + it : (int)mlist[it]) * PST;
          if (o0 >= 0) g0 = rec[o0];
          if (o1 >= 0) g1 = rec[o1];
        }
        dmma884(c00, g0, g0);
        dmma884(c01, g0, g1);
        dmma884(c11, g1, g1);
      }
      // flush of the (i,j) block + gradient of j (c01) and the (j,j) block (c11)
      {
        const int ra = 15 * (li - pk_fo), rb = 15 * (lj - pk_fo);
        double *Sg = D.Smat + pk_so;
        const bool up = li < lj;   // the system keeps its upper triangle: block (i,j) as is, or transposed when j comes first
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int q = pc + e;
          if (q < 6) {
            if (pr < 6) {
              atomicAdd(up ? Sg + (size_t)(ra + pr) * pk_d + rb + q : Sg + (size_t)(rb + q) * pk_d + ra + pr, c01[e]);
              if (q >= pr) {
                atomicAdd(Sg + (size_t)(rb + pr) * pk_d + rb + q, c11[e]);
                if (q == pr) atomicAdd(D.colsq_cam + pk_co + rb + pr, c11[e]);
              }
            } else if (pr == 6) {
              atomicAdd(D.gS + pk_co + rb + q, c01[e]); atomicAdd(D.gfull + pk_co + rb + q, c01[e]);
            }
          }
        }
      }
      __syncwarp();
    }
  }
  flush00();
  add_window_scalar(D.acc + ACC_COST0, ACC_STRIDE, w, half, act);
  if (!act) return;
  // ---- elimination of the 1x1 landmark block (same arithmetic as k_core_points)
  double sk;
  if (!D.ctl[w].have_scale) { sk = 1.0 / (1.0 + sqrt(colsq)); D.scale_pt[gp] = sk; }
  else sk = D.scale_pt[gp];
  const double Et = sk * sk * colsq;
  const double D2 = clamp4(Et, P.min_lm_diag, P.max_lm_diag) / D.ctl[w].radius;
  const double sh = rsqrt(Et + D2);
  const double ysc = sk * sh;
  {
    double2 *y2 = reinterpret_cast<double2 *>(Y + 6 * (row_i - fo));
    y2[0] = make_double2(wa[0], wa[1]); y2[1] = make_double2(wa[2], wa[3]); y2[2] = make_double2(wa[4], wa[5]);
  }
  *reinterpret_cast<double2 *>(Y + mp - 2) = make_double2(gk, ysc * ysc);   // z row and column scale (unscaled column: uvs_stash.cuh)
  ph[0] = sk; ph[1] = sh; ph[2] = D2; ph[3] = gk;
  atomic_max_nn3(D.acc + (size_t)w * ACC_STRIDE + ACC_GMAX + D.rank, fabs(gk));
}

// ------------------------------------------------------------------------------------------------
// lines
constexpr int LL_NT = 128;                 // threads per CTA: 16 lines x 8 lanes
constexpr int LSLOT = 11;                  // observations staged per line (the reference's window has 11 frames; longer tracks -> record path)
constexpr int WSLOTS = (32 / LPL) * LSLOT; // staged observations per warp
// doubles of shared memory per warp: line + VP stage, frame / VP flag of every slot (ints)
constexpr int LL_WARP_DOUBLES = WSLOTS * (REC_LINE + REC_VP) + WSLOTS;

template <bool kJac>
struct LineVpSinkF {
  LineSink<kJac, false> ln;
  VpSink<kJac, false> vp;
  bool has_vp;
  __device__ __forceinline__ void base(d3 n, d3 dd) { ln.base(n, dd); if (has_vp) vp.base(n, dd); }
  __device__ __forceinline__ void partial(int k, d3 dn, d3 du) { ln.partial(k, dn, du); if (has_vp) vp.partial(k, dn, du); }
};

// grid (ceil(max lines per window / 16), B): a CTA works on 16 lines of ONE window, so the frame tables are per CTA
template <int kOcc>
__global__ void __launch_bounds__(LL_NT, kOcc) k_lin_lines(Dev D, Params P, Stash S, int max_frames) {
  extern __shared__ __align__(16) double lsm[];
  const unsigned full = 0xffffffffu;
  const int w = blockIdx.y;
  const int nl = D.line_off[w + 1] - D.line_off[w];
  const int l0 = blockIdx.x * (LL_NT / LPL);
  if (l0 >= nl) return;
  if (!(D.ctl[w].state & WS_ACTIVE)) return;
  const bool need_cost = D.ctl[w].iter == 0;   // the cost of the linearisation point is only used at iteration 0
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // shared: frame tables [max_frames][FT_STRIDE] | per warp: line stage [WSLOTS][REC_LINE], VP stage [WSLOTS][REC_VP],
  //         frame of a slot (-1 = empty), VP flag, row list of the direct-term grouping
  double *ftab = lsm;
  const int ftab_doubles = (max_frames * FT_STRIDE + 1) & ~1;
  double *wbase = lsm + ftab_doubles + warp * LL_WARP_DOUBLES;
  signed char *slotof = reinterpret_cast<signed char *>(lsm + ftab_doubles + (LL_NT / 32) * LL_WARP_DOUBLES);   // [16 lines][16 frames] observation (| 0x40: VP) or -1
  double *lstage = wbase, *vstage = wbase + WSLOTS * REC_LINE;
  int *sframe = reinterpret_cast<int *>(vstage + WSLOTS * REC_VP);        // [WSLOTS]
  int *svp = sframe + WSLOTS;                                              // [WSLOTS]
  const int fo = D.frame_off[w], F = D.frame_off[w + 1] - fo;
  const int cur = D.cur[w];
  load_frame_tables(D.ftab[cur] + (size_t)fo * FT_DOUBLES, F, ftab, LL_NT);
  for (int e = lane; e < WSLOTS; e += 32) { sframe[e] = -1; svp[e] = 0; }
  for (int e = threadIdx.x; e < (LL_NT / LPL) * 16; e += LL_NT) slotof[e] = -1;
  __syncthreads();

  const int grp = lane / LPL, sub = lane - grp * LPL;
  const int li = l0 + threadIdx.x / LPL;
  const unsigned gmask = ((1u << LPL) - 1u) << (lane & ~(LPL - 1));
  bool act = li < nl;   // uniform over the lane group
  const int gl = D.line_off[w] + li;
  const int mp = S.mp;
  int f0 = 0, n = 0;
  double *Y = nullptr, *hd = nullptr;
  if (act) {
    f0 = D.ln_begin[gl]; n = D.ln_end[gl] - f0;
    Y = S.Y + (colbase(D, w) + (D.point_off[w + 1] - D.point_off[w]) + 4LL * li) * mp;   // 4 columns
    hd = S.lh + 24 * (size_t)gl;
    const bool mine = D.nranks <= 1 || (gl % D.nranks) == D.rank;
    // Linv[0][0] = 0 marks "no step" for the back-substitution
    if (n <= 0 || !mine) { if (sub == 0) hd[8] = 0.0; act = false; }
  }
  if (!act) n = 0;
  if (n > LSLOT) n = LSLOT;   // excluded at upload (the batch takes the record path instead)
  // ---- per-line table from the sines / cosines k_line_tables left for this state buffer
  LineTab LT;
  {
    const double *q = D.lsc[cur] + 8 * (size_t)(act ? gl : D.line_off[w]);
    line_table(__ldg(q), __ldg(q + 1), __ldg(q + 2), __ldg(q + 3), __ldg(q + 4), __ldg(q + 5), __ldg(q + 6), __ldg(q + 7), LT);
  }
  // ---- evaluation: observation f of the line -> slot grp * LSLOT + f of the warp's stage
  double half = 0.0;
  for (int f = sub; f < n; f += LPL) {
    const int slot = grp * LSLOT + f;
    const int4 ix = D.line_idx4[f0 + f];
    const double *sp = D.line_sp + 2 * (size_t)(f0 + f), *ep = D.line_ep + 2 * (size_t)(f0 + f);
    LineVpSinkF<true> sink;
    sink.ln.spx = __ldg(sp); sink.ln.spy = __ldg(sp + 1); sink.ln.epx = __ldg(ep); sink.ln.epy = __ldg(ep + 1);
    sink.ln.lf = P.line_factor; sink.ln.loss_a = P.cauchy_line; sink.ln.correct = true; sink.ln.PW = 6;
    sink.ln.want_cost = sink.vp.want_cost = need_cost;
    sink.has_vp = ix.w >= 0;
    sink.vp.half_rho = 0.0;
    if (ix.w >= 0) {
      const double *vp = D.vp_dir + 3 * (size_t)ix.w;
      sink.vp.vp = mk3(__ldg(vp), __ldg(vp + 1), __ldg(vp + 2));
      sink.vp.vf = P.vp_factor; sink.vp.loss_a = P.cauchy_vp; sink.vp.correct = true;
    }
    double *tl = lstage + slot * REC_LINE, *tv = vstage + slot * REC_VP;
    sink.ln.out_r = tl; sink.ln.out_jp = tl + 2; sink.ln.out_jl = tl + 14;
    sink.vp.out_r = tv; sink.vp.out_jp = tv + 1; sink.vp.out_jl = tv + 7;
    line_obs_eval<true, true>(ftab + (ix.x - fo) * FT_STRIDE, LT, sink);
    half += sink.ln.half_rho + sink.vp.half_rho;
    sframe[slot] = ix.x - fo;
    svp[slot] = ix.w >= 0 ? 1 : 0;
    slotof[(threadIdx.x / LPL) * 16 + (ix.x - fo)] = (signed char)(f | (ix.w >= 0 ? 0x40 : 0));
  }
  add_window_scalar(D.acc + ACC_COST0, ACC_STRIDE, w, half, act);

  // ---- diagonal direct terms of the CTA's observations, per frame: G = [J_pose (6) | r | 0], one row per residual row
  //      (two per line observation, one per VP factor), G^T G on the FP64 tensor cores.  Warp q takes the frames q, q + 4,
  //      q + 8, ...; row r of a frame's G is fixed to (line r / 3 of the CTA, row type r % 3) and read through the
  //      (line, frame) -> observation table, absent rows are zero: no lists, one flush per frame and CTA.
  __syncthreads();
  {
    const int co = D.cam_off[w], d = D.cam_off[w + 1] - co;
    double *Sg = D.Smat + D.S_off[w];
    const int fcol = lane >> 2, krow = lane & 3;
    const int pr = lane >> 2, pc = 2 * (lane & 3);
    for (int jj = warp; jj < F; jj += LL_NT / 32) {
      double cc[2] = {0.0, 0.0};
#pragma unroll 4
      for (int ks = 0; ks < 3 * (LL_NT / LPL) / 4; ks++) {
        const int r = 4 * ks + krow, ln = r / 3, ty = r - 3 * ln;
        const int sl = slotof[ln * 16 + jj];
        double g = 0.0;
        if (sl >= 0 && fcol < 7 && (ty < 2 || (sl & 0x40))) {
          const double *wb = lsm + ftab_doubles + (ln >> 2) * LL_WARP_DOUBLES;   // stage of the warp that owns line ln
          const int slot = (ln & 3) * LSLOT + (sl & 0x3f);
          const double *rec = ty < 2 ? wb + slot * REC_LINE : wb + WSLOTS * REC_LINE + slot * REC_VP;
          g = ty == 0 ? (fcol < 6 ? rec[2 + fcol] : rec[0]) : (ty == 1 ? (fcol < 6 ? rec[8 + fcol] : rec[1]) : (fcol < 6 ? rec[1 + fcol] : rec[0]));
        }
        dmma884(cc, g, g);
      }
      const int ra = 15 * jj;
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int q = pc + e;
        if (pr < 6 && q >= pr && q < 6) {
          atomicAdd(Sg + (size_t)(ra + pr) * d + ra + q, cc[e]);
          if (q == pr) atomicAdd(D.colsq_cam + co + ra + pr, cc[e]);
        } else if (pr < 6 && q == 6) {
          atomicAdd(D.gS + co + ra + pr, cc[e]); atomicAdd(D.gfull + co + ra + pr, cc[e]);
        }
      }
    }
  }
  if (!act) return;   // uniform over the lane group

  // ---- elimination of the 4x4 landmark block (same arithmetic as k_core_lines; records come from the stage)
  auto line_rec = [&](int f) { return lstage + (grp * LSLOT + f) * REC_LINE; };
  auto vp_rec = [&](int f) { return vstage + (grp * LSLOT + f) * REC_VP; };
  auto has_vp = [&](int f) { return svp[grp * LSLOT + f] != 0; };
  double E[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, g[4] = {0, 0, 0, 0};
  for (int f = sub; f < n; f += LPL) {
    const double2 *r2 = reinterpret_cast<const double2 *>(line_rec(f));
    const double2 rr = r2[0];
#pragma unroll
    for (int rw_ = 0; rw_ < 2; rw_++) {
      const double2 a01 = r2[7 + 2 * rw_], a23 = r2[8 + 2 * rw_];
      const double a0 = a01.x, a1 = a01.y, a2 = a23.x, a3 = a23.y, rw = rw_ ? rr.y : rr.x;
      E[0] += a0 * a0; E[1] += a0 * a1; E[2] += a0 * a2; E[3] += a0 * a3; E[4] += a1 * a1; E[5] += a1 * a2; E[6] += a1 * a3;
      E[7] += a2 * a2; E[8] += a2 * a3; E[9] += a3 * a3;
      g[0] += a0 * rw; g[1] += a1 * rw; g[2] += a2 * rw; g[3] += a3 * rw;
    }
    if (has_vp(f)) {
      const double *q = vp_rec(f);
      const double a0 = q[7], a1 = q[8], a2 = q[9], a3 = q[10], rw = q[0];
      E[0] += a0 * a0; E[1] += a0 * a1; E[2] += a0 * a2; E[3] += a0 * a3; E[4] += a1 * a1; E[5] += a1 * a2; E[6] += a1 * a3;
      E[7] += a2 * a2; E[8] += a2 * a3; E[9] += a3 * a3;
      g[0] += a0 * rw; g[1] += a1 * rw; g[2] += a2 * rw; g[3] += a3 * rw;
    }
  }
#pragma unroll
  for (int k = 0; k < 10; k++) E[k] = group_sum<LPL>(gmask, E[k]);
#pragma unroll
  for (int k = 0; k < 4; k++) g[k] = group_sum<LPL>(gmask, g[k]);
  double s[4];
  const double Ed[4] = {E[0], E[4], E[7], E[9]};
  if (!D.ctl[w].have_scale) {
#pragma unroll
    for (int c = 0; c < 4; c++) { s[c] = 1.0 / (1.0 + sqrt(Ed[c])); if (sub == 0) D.scale_ln[4 * (size_t)gl + c] = s[c]; }
  } else {
#pragma unroll
    for (int c = 0; c < 4; c++) s[c] = D.scale_ln[4 * (size_t)gl + c];
  }
  // M = D_s E D_s + D^2 (lower), Cholesky, inverse of the factor
  const double radius = D.ctl[w].radius;
  double M[4][4], D2[4];
  M[0][0] = s[0] * s[0] * E[0]; M[1][0] = s[1] * s[0] * E[1]; M[2][0] = s[2] * s[0] * E[2]; M[3][0] = s[3] * s[0] * E[3];
  M[1][1] = s[1] * s[1] * E[4]; M[2][1] = s[2] * s[1] * E[5]; M[3][1] = s[3] * s[1] * E[6];
  M[2][2] = s[2] * s[2] * E[7]; M[3][2] = s[3] * s[2] * E[8]; M[3][3] = s[3] * s[3] * E[9];
#pragma unroll
  for (int c = 0; c < 4; c++) { D2[c] = clamp4(M[c][c], P.min_lm_diag, P.max_lm_diag) / radius; M[c][c] += D2[c]; }
  double L[4][4], Li[4][4];
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    double dj = M[j][j];
#pragma unroll
    for (int k = 0; k < j; k++) dj -= L[j][k] * L[j][k];
    if (!(dj > 0.0)) ok = false;
    const double id = rsqrt(dj);
    L[j][j] = dj * id;
#pragma unroll
    for (int i = j + 1; i < 4; i++) {
      double tt = M[i][j];
#pragma unroll
      for (int k = 0; k < j; k++) tt -= L[i][k] * L[j][k];
      L[i][j] = tt * id;
    }
  }
  if (!ok) {
    if (sub == 0) { hd[8] = 0.0; atomicAdd(D.acc + (size_t)w * ACC_STRIDE + ACC_FAIL, 1.0); }
    return;
  }
#pragma unroll
  for (int col = 0; col < 4; col++)
#pragma unroll
    for (int i = col; i < 4; i++) {
      double tt = (i == col) ? 1.0 : 0.0;
#pragma unroll
      for (int k = col; k < i; k++) tt -= L[i][k] * Li[k][col];
      Li[i][col] = tt / L[i][i];
    }
  if (sub == 0) {
    // z = L^-1 (D_s g)
#pragma unroll
    for (int c = 0; c < 4; c++) {
      double tt = 0.0;
#pragma unroll
      for (int k = 0; k <= c; k++) tt += Li[c][k] * s[k] * g[k];
      Y[c * mp + mp - 2] = tt; hd[c] = s[c]; hd[4 + c] = D2[c];
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int k = 0; k < 4; k++) hd[8 + 4 * i + k] = k <= i ? Li[i][k] : 0.0;
    atomic_max_nn3(D.acc + (size_t)w * ACC_STRIDE + ACC_GMAX + D.rank, fmax(fmax(fabs(g[0]), fabs(g[1])), fmax(fabs(g[2]), fabs(g[3]))));
  }
  // Y blocks of this lane's observations
  for (int f = sub; f < n; f += LPL) {
    const double2 *r2 = reinterpret_cast<const double2 *>(line_rec(f));
    const double *q = has_vp(f) ? vp_rec(f) : nullptr;
    double *Yf = Y + 6 * sframe[grp * LSLOT + f];
    double jl[3][4], jp[2][6];
#pragma unroll
    for (int c = 0; c < 2; c++) {
      const double2 a = r2[7 + c], b2 = r2[9 + c];
      jl[0][2 * c] = a.x * s[2 * c]; jl[0][2 * c + 1] = a.y * s[2 * c + 1];
      jl[1][2 * c] = b2.x * s[2 * c]; jl[1][2 * c + 1] = b2.y * s[2 * c + 1];
    }
#pragma unroll
    for (int c = 0; c < 4; c++) jl[2][c] = q ? q[7 + c] * s[c] : 0.0;
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const double2 a = r2[1 + c], b2 = r2[4 + c];
      jp[0][2 * c] = a.x; jp[0][2 * c + 1] = a.y; jp[1][2 * c] = b2.x; jp[1][2 * c + 1] = b2.y;
    }
#pragma unroll
    for (int p = 0; p < 6; p += 2) {   // two rows at a time: the 48-byte block of a column is written as three 128-bit words
      double tt[2][4];
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const double a0 = jp[0][p + h], a1 = jp[1][p + h], a2 = q ? q[1 + p + h] : 0.0;
        double W4[4];
#pragma unroll
        for (int c = 0; c < 4; c++) W4[c] = a0 * jl[0][c] + a1 * jl[1][c] + a2 * jl[2][c];
#pragma unroll
        for (int c = 0; c < 4; c++) {
          double t1 = 0.0;
#pragma unroll
          for (int k = 0; k <= c; k++) t1 += W4[k] * Li[c][k];
          tt[h][c] = t1;
        }
      }
#pragma unroll
      for (int c = 0; c < 4; c++) *reinterpret_cast<double2 *>(Yf + c * mp + p) = make_double2(tt[0][c], tt[1][c]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
static inline int cdivl(int a, int b) { return (a + b - 1) / b; }

static int env_int(const char *name, int dflt) {
  const char *e = std::getenv(name);
  return e ? std::atoi(e) : dflt;
}

int launch_prep_point_order(const Dev &D, int *key_scratch, cudaStream_t st) {
  if (D.nP == 0) return 0;
  k_prep_point_order<<<D.B, 256, 0, st>>>(D, key_scratch, env_int("UVS_PT_DENSE", 1));
  return 1;
}

size_t lin_lines_smem(int max_frames) {
  return ((size_t)((max_frames * FT_STRIDE + 1) & ~1) + (size_t)(LL_NT / 32) * LL_WARP_DOUBLES) * sizeof(double) + (LL_NT / LPL) * 16;
}
int lin_max_line_obs() { return LSLOT; }

int launch_lin_points(const Dev &D, const Params &P, char *base, const Build3Layout &lay, cudaStream_t st) {
  if (D.nP == 0 || D.nPW == 0) return 0;
  Stash S; S.Y = (double *)(base + lay.o_Y); S.ph = (double *)(base + lay.o_ph); S.lh = (double *)(base + lay.o_lh); S.mp = lay.mp; S.unscaled_pts = 1;
  // developer switches (A/B runs): UVS_PT_CTA = threads per CTA (32 | 128), UVS_PT_PREFETCH = 0 | 1
  const int cta = env_int("UVS_PT_CTA", 32), pre = env_int("UVS_PT_PREFETCH", 0), occ = env_int("UVS_PT_OCC", 512);
  if (cta == 32 && occ == 640) k_lin_points<32, false, 640><<<D.nPW, 32, 0, st>>>(D, P, S);
  else if (cta == 32 && occ == 768) k_lin_points<32, false, 768><<<D.nPW, 32, 0, st>>>(D, P, S);
  else if (cta == 128) {
    if (pre) k_lin_points<128, true><<<cdivl(32 * D.nPW, 128), 128, 0, st>>>(D, P, S);
    else k_lin_points<128, false><<<cdivl(32 * D.nPW, 128), 128, 0, st>>>(D, P, S);
  } else {
    if (pre) k_lin_points<32, true><<<D.nPW, 32, 0, st>>>(D, P, S);
    else k_lin_points<32, false><<<D.nPW, 32, 0, st>>>(D, P, S);
  }
  return 1;
}

int launch_lin_lines(const Dev &D, const Params &P, char *base, const Build3Layout &lay, int max_frames, int max_lines, cudaStream_t st) {
  if (D.nL == 0 || max_lines == 0) return 0;
  Stash S; S.Y = (double *)(base + lay.o_Y); S.ph = (double *)(base + lay.o_ph); S.lh = (double *)(base + lay.o_lh); S.mp = lay.mp; S.unscaled_pts = 1;
  const size_t smem = lin_lines_smem(max_frames);
  static size_t raised = 0;
  if (smem > raised) {
    cudaFuncSetAttribute(k_lin_lines<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_lin_lines<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    raised = smem;
  }
  const dim3 grid(cdivl(max_lines, LL_NT / LPL), D.B);
  if (env_int("UVS_LL_OCC", 4) == 3) k_lin_lines<3><<<grid, LL_NT, smem, st>>>(D, P, S, max_frames);
  else k_lin_lines<4><<<grid, LL_NT, smem, st>>>(D, P, S, max_frames);
  return 1;
}

}  // namespace uvs

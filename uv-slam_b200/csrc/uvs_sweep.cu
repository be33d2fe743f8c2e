// uvs_sweep.cu — the factor sweep: one fused kernel per factor type evaluates every residual of
// the batch and (Jacobian mode) its tangent-space Jacobian blocks, loss-corrected, into the
// per-factor record arrays.  sm_100a, FP64.
//
// Reference functions replaced (per call, on the CPU, inside ceres::Solve):
//   ProjectionFactor::Evaluate / ProjectionTdFactor::Evaluate   factor/projection_factor.cpp:22-175,
//                                                                factor/projection_td_factor.cpp:34-145
//   AutoDiffCostFunction<LineProjectionFactor,2,7,4>::Evaluate   factor/line_projection_factor.h:16-60
//   AutoDiffCostFunction<VPProjectionFactor,1,7,4>::Evaluate     factor/vp_projection_factor.h:19-66
//   IMUFactor::Evaluate                                          factor/imu_factor.h:19-182
//   MarginalizationFactor::Evaluate                              factor/marginalization_factor.cpp:333-381
//   Ceres' loss corrector (restated at marginalization_factor.cpp:37-68)
//
// Memory behaviour: a thread owns one factor (a warp for IMU factors); index records and
// observations are read once; state blocks (a few KB per window) come through L1/L2; every CTA
// stages its tile of records in shared memory and writes it back as one contiguous, fully
// coalesced chunk.
#include <algorithm>
#include <cstdlib>

#include "uvs_device.cuh"
#include "uvs_factors.cuh"
#include "uvs_imu.cuh"
#include "uvs_kernels.h"
#include "uvs_linefast.cuh"

namespace uvs {

constexpr int NT = 128;  // threads per CTA of the per-factor kernels

// FP64 tensor-core MMA  D(8x8) += A(8x4) B(4x8)  (mma.sync m8n8k4: lane l holds A[l/4][l%4], B[l%4][l/4], D[l/4][2(l%4) + {0,1}])
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// does window state `st` want this factor evaluated?
//   mode 0: API evaluation (everything);  mode 1: solver
template <bool kJac>
__device__ __forceinline__ bool wants(int st, int mode) {
  if (mode == 0) return true;
  if (!(st & WS_ACTIVE)) return false;
  return kJac ? (st & WS_NEED_JAC) != 0 : (st & WS_STEP_OK) != 0;
}

// warp-aggregated accumulation of a per-window scalar
__device__ __forceinline__ void add_window_scalar(double *arr, int stride_doubles, int win, double v, bool valid) {
  const unsigned full = 0xffffffffu;
  const int w0 = __shfl_sync(full, win, 0);
  const bool uniform = __all_sync(full, !valid || win == w0) && __shfl_sync(full, (int)valid, 0);
  if (uniform) {
    double s = valid ? v : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(full, s, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(arr + (size_t)w0 * stride_doubles, s);
  } else if (valid) {
    atomicAdd(arr + (size_t)win * stride_doubles, v);
  }
}

// loss: returns sqrt(rho') and the cost 1/2 rho(s); a <= 0 means no loss function
__device__ __forceinline__ double corrector(double a, double s, double &half_rho) {
  if (a <= 0.0) { half_rho = 0.5 * s; return 1.0; }
  double rho0, rho1;
  cauchy(a, s, rho0, rho1);
  half_rho = 0.5 * rho0;
  return sqrt(rho1);
}

// contiguous, coalesced write-back of a CTA's tile of records staged in shared memory.  Call after a __syncthreads()
// that follows the writes of `tile` and `ok`; `all_ok` (uniform over the CTA) skips the per-record test.
template <int REC>
__device__ __forceinline__ void flush_tile(const double *tile, const unsigned char *ok, double *__restrict__ out,
                                           int first, int count, bool all_ok) {
  double *dst = out + (size_t)first * REC;
  const int total = count * REC;
  constexpr int DF = NT / REC, DC = NT % REC;   // element e + NT lies DF records and DC columns further
  int e = threadIdx.x;
  int f = e / REC, c = e - f * REC;
  if (all_ok && count == NT) {
    // full tile: element e = (f, c) sits at tile[f (REC + 1) + c] = tile[e + f]; the trip count and the divisor are compile-time
    // constants, so an element costs a multiply-high, an add, the load and the store (the carried (f, c) update cost twice that:
    // the write-back loops were a third of the instructions of k_line_vp<1>)
#pragma unroll 4
    for (int i = 0; i < REC; i++) {
      const int x = threadIdx.x + NT * i;
      dst[x] = tile[x + x / REC];
    }
  } else if (all_ok) {
#pragma unroll 4
    for (; e < total; e += NT) {
      dst[e] = tile[e + f];
      f += DF; c += DC;
      if (c >= REC) { c -= REC; f++; }
    }
  } else {
    for (; e < total; e += NT) {
      if (ok[f]) dst[e] = tile[e + f];
      f += DF; c += DC;
      if (c >= REC) { c -= REC; f++; }
    }
  }
}

// ------------------------------------------------------------------------------------------------
template <bool kJac, bool kTd, bool kCeres, int kOcc = 4>
__global__ void __launch_bounds__(NT, kOcc) k_proj(Dev D, Params P, int mode, int cand, double *__restrict__ out,
                                             double *__restrict__ res_out, double *cost, int cost_stride) {
  constexpr int REC = kCeres ? (kTd ? CREC_PROJ_TD : CREC_PROJ) : (kTd ? REC_PROJ_TD : REC_PROJ);
  constexpr int PW = kCeres ? 7 : 6;
  extern __shared__ double smem[];
  double *tile = smem;
  unsigned char *ok = reinterpret_cast<unsigned char *>(smem + NT * (REC + 1));
  const int first = blockIdx.x * NT;
  const int f = first + threadIdx.x;
  bool valid = f < D.nProj;
  int4 ix = make_int4(0, 0, 0, 0);
  int wflags = 0;
  if (valid) {
    ix = D.proj_idx[f];
    valid = wants<kJac>(D.ctl[ix.w].state, mode) && (D.nranks <= 1 || mode == 0 || (ix.z % D.nranks) == D.rank);
    wflags = D.win_flags[ix.w];
  }
  double half_rho = 0.0;
  if (valid) {
    const int buf = D.cur[ix.w] ^ cand;
    const double *pi = D.pose[buf] + 7 * (size_t)ix.x, *pj = D.pose[buf] + 7 * (size_t)ix.y;
    const double *ex = D.ex[buf] + 7 * (size_t)ix.w;
    const double lam = __ldg(D.inv_depth[buf] + ix.z);
    const double *oi = D.proj_pts_i + 3 * (size_t)f, *oj = D.proj_pts_j + 3 * (size_t)f;
    const d3 pts_i = mk3(__ldg(oi), __ldg(oi + 1), __ldg(oi + 2)), pts_j = mk3(__ldg(oj), __ldg(oj + 1), __ldg(oj + 2));
    ProjTd tdp;
    if (kTd) {
      tdp.td = __ldg(D.td[buf] + ix.w);
      tdp.td_i = __ldg(D.proj_td_i + f); tdp.td_j = __ldg(D.proj_td_j + f);
      tdp.row_i = __ldg(D.proj_row_i + f); tdp.row_j = __ldg(D.proj_row_j + f);
      tdp.vix = __ldg(D.proj_vel_i + 2 * (size_t)f); tdp.viy = __ldg(D.proj_vel_i + 2 * (size_t)f + 1);
      tdp.vjx = __ldg(D.proj_vel_j + 2 * (size_t)f); tdp.vjy = __ldg(D.proj_vel_j + 2 * (size_t)f + 1);
      tdp.tr_over_row = P.tr_over_row; tdp.half_row = P.half_row;
    }
    const bool want_ex = mode == 0 || (wflags & WF_EXTRINSIC);
    // the cost of a Jacobian-mode sweep is only used at iteration 0 of a solve (and by nobody when no accumulator is given)
    const bool want_cost = cost != nullptr && (!kJac || mode == 0 || D.ctl[ix.w].iter == 0);
    if (kJac) {
      double *t = tile + threadIdx.x * (REC + 1);
      proj_eval<true, kTd>(pi, pj, ex, lam, pts_i, pts_j, P.S, &tdp, want_ex, P.cauchy_point, !kCeres, PW, t, t + 2, t + 2 + 2 * PW,
                           t + 2 + 4 * PW, t + 2 + 6 * PW, t + 2 + 6 * PW + 2, want_cost ? &half_rho : nullptr);
      if (kCeres) {
#pragma unroll
        for (int row = 0; row < 2; row++) { t[2 + row * 7 + 6] = 0.0; t[2 + 14 + row * 7 + 6] = 0.0; t[2 + 28 + row * 7 + 6] = 0.0; }
      }
    } else {
      double r[2];
      proj_eval<false, kTd>(pi, pj, ex, lam, pts_i, pts_j, P.S, &tdp, want_ex, P.cauchy_point, true, 6, r, nullptr, nullptr, nullptr,
                            nullptr, nullptr, want_cost ? &half_rho : nullptr);
      if (res_out) { res_out[2 * (size_t)f] = r[0]; res_out[2 * (size_t)f + 1] = r[1]; }
    }
  }
  if (cost) add_window_scalar(cost, cost_stride, ix.w, half_rho, valid);
  if (kJac) {
    ok[threadIdx.x] = valid;
    const bool all_ok = __syncthreads_and(valid || f >= D.nProj) != 0;
    flush_tile<REC>(tile, ok, out, first, min(NT, D.nProj - first), all_ok);
  }
}

// ------------------------------------------------------------------------------------------------
template <bool kJac, bool kCeres>
__global__ void __launch_bounds__(NT, 3) k_line(Dev D, Params P, int mode, int cand, double *__restrict__ out,
                                             double *__restrict__ res_out, double *cost, int cost_stride) {
  constexpr int REC = kCeres ? CREC_LINE : REC_LINE;
  constexpr int NP = kCeres ? 11 : 10;
  constexpr int PW = kCeres ? 7 : 6;
  extern __shared__ double smem[];
  double *tile = smem;
  unsigned char *ok = reinterpret_cast<unsigned char *>(smem + NT * (REC + 1));
  const int first = blockIdx.x * NT;
  const int f = first + threadIdx.x;
  bool valid = f < D.nLobs;
  int4 ix = make_int4(0, 0, 0, 0);
  if (valid) {
    ix = D.line_idx4[f];
    valid = wants<kJac>(D.ctl[ix.z].state, mode) && (D.nranks <= 1 || mode == 0 || (ix.y % D.nranks) == D.rank);
  }
  double half_rho = 0.0;
  if (valid) {
    const int buf = D.cur[ix.z] ^ cand;
    const double *sp = D.line_sp + 2 * (size_t)f, *ep = D.line_ep + 2 * (size_t)f;
    LineSink<kJac, kCeres> sink;
    sink.spx = __ldg(sp); sink.spy = __ldg(sp + 1); sink.epx = __ldg(ep); sink.epy = __ldg(ep + 1);
    sink.lf = P.line_factor; sink.loss_a = P.cauchy_line; sink.correct = !kCeres; sink.PW = PW;
    double rloc[2];
    if (kJac) { double *t = tile + threadIdx.x * (REC + 1); sink.out_r = t; sink.out_jp = t + 2; sink.out_jl = t + 2 + 2 * PW; }
    else { sink.out_r = rloc; sink.out_jp = nullptr; sink.out_jl = nullptr; }
    line_to_camera<kJac, true, kCeres>(D.pose[buf] + 7 * (size_t)ix.x, D.ortho[buf] + 4 * (size_t)ix.y, D.ric + 9 * (size_t)ix.z,
                                       D.tic + 3 * (size_t)ix.z, sink);
    half_rho = sink.half_rho;
    if (!kJac && res_out) { res_out[2 * (size_t)f] = rloc[0]; res_out[2 * (size_t)f + 1] = rloc[1]; }
  }
  if (cost) add_window_scalar(cost, cost_stride, ix.z, half_rho, valid);
  if (kJac) {
    ok[threadIdx.x] = valid;
    const bool all_ok = __syncthreads_and(valid || f >= D.nLobs) != 0;
    flush_tile<REC>(tile, ok, out, first, min(NT, D.nLobs - first), all_ok);
  }
}

// ------------------------------------------------------------------------------------------------
template <bool kJac, bool kCeres>
__global__ void __launch_bounds__(NT, 4) k_vp(Dev D, Params P, int mode, int cand, double *__restrict__ out,
                                           double *__restrict__ res_out, double *cost, int cost_stride) {
  constexpr int REC = kCeres ? CREC_VP : REC_VP;
  constexpr int NP = kCeres ? 11 : 10;
  constexpr int PW = kCeres ? 7 : 6;
  extern __shared__ double smem[];
  double *tile = smem;
  unsigned char *ok = reinterpret_cast<unsigned char *>(smem + NT * (REC + 1));
  const int first = blockIdx.x * NT;
  const int f = first + threadIdx.x;
  bool valid = f < D.nVobs;
  int4 ix = make_int4(0, 0, 0, 0);
  if (valid) {
    ix = D.vp_idx4[f];
    valid = wants<kJac>(D.ctl[ix.z].state, mode) && (D.nranks <= 1 || mode == 0 || (ix.y % D.nranks) == D.rank);
  }
  double half_rho = 0.0;
  if (valid) {
    const int buf = D.cur[ix.z] ^ cand;
    const double *vp = D.vp_dir + 3 * (size_t)f;
    VpSink<kJac, kCeres> sink;
    sink.vp = mk3(__ldg(vp), __ldg(vp + 1), __ldg(vp + 2));
    sink.vf = P.vp_factor; sink.loss_a = P.cauchy_vp; sink.correct = !kCeres;
    double rloc[1];
    if (kJac) { double *t = tile + threadIdx.x * (REC + 1); sink.out_r = t; sink.out_jp = t + 1; sink.out_jl = t + 1 + PW; }
    else { sink.out_r = rloc; sink.out_jp = nullptr; sink.out_jl = nullptr; }
    line_to_camera<kJac, false, kCeres>(D.pose[buf] + 7 * (size_t)ix.x, D.ortho[buf] + 4 * (size_t)ix.y, D.ric + 9 * (size_t)ix.z,
                                        D.tic + 3 * (size_t)ix.z, sink);
    half_rho = sink.half_rho;
    if (!kJac && res_out) res_out[f] = rloc[0];
  }
  if (cost) add_window_scalar(cost, cost_stride, ix.z, half_rho, valid);
  if (kJac) {
    ok[threadIdx.x] = valid;
    const bool all_ok = __syncthreads_and(valid || f >= D.nVobs) != 0;
    flush_tile<REC>(tile, ok, out, first, min(NT, D.nVobs - first), all_ok);
  }
}

// ------------------------------------------------------------------------------------------------
// Line factor and the VP factor of the same (frame, line) observation in ONE pass, through per-frame and per-line
// tables (uvs_linefast.cuh, written by k_line_tables for the state buffer in question): a CTA works on up to NT
// consecutive line observations of ONE window (grid: chunks x windows), copies the window's frame tables into shared
// memory, then a thread owns an observation: ~500 FP64 instructions instead of the ~1500 of line_to_camera.  line_idx4[f].w names the paired VP observation (k_prep_vp) or
// -1.  Records of a tile are staged in shared memory and written back as contiguous chunks (the VP observations follow
// the order of the line observations, so neighbouring VP records are adjacent in memory).
template <bool kJac>
struct LineVpSink {
  LineSink<kJac, false> ln;
  VpSink<kJac, false> vp;
  bool has_vp;
  __device__ __forceinline__ void base(d3 n, d3 d) { ln.base(n, d); if (has_vp) vp.base(n, d); }
  __device__ __forceinline__ void partial(int k, d3 dn, d3 du) { ln.partial(k, dn, du); if (has_vp) vp.partial(k, dn, du); }
};

template <bool kJac, int kOcc = 3>
__global__ void __launch_bounds__(NT, kOcc) k_line_vp(Dev D, Params P, int mode, int cand, double *__restrict__ out_line,
                                                double *__restrict__ out_vp, double *cost, int cost_stride) {
  extern __shared__ __align__(16) double smem[];
  const int w = blockIdx.y;
  const int a0 = D.lobs_off[w], a1 = D.lobs_off[w + 1];
  const int first = a0 + blockIdx.x * NT;
  if (first >= a1) return;
  if (!wants<kJac>(D.ctl[w].state, mode)) return;
  const int fo = D.frame_off[w], F = D.frame_off[w + 1] - fo;
  double *ftab = smem;                                        // [F][FT_STRIDE]
  double *tile = ftab + ((D.max_frames * FT_STRIDE + 1) & ~1); // [NT][REC_LINE + 1]   (Jacobian mode only)
  double *vtile = tile + NT * (REC_LINE + 1);                 // [NT][REC_VP + 1]
  int *vslot = reinterpret_cast<int *>(vtile + NT * (REC_VP + 1));   // [NT] VP observation of the slot or -1
  unsigned char *ok = reinterpret_cast<unsigned char *>(vslot + NT);
  const int buf = D.cur[w] ^ cand;
  load_frame_tables(D.ftab[buf] + (size_t)fo * FT_DOUBLES, F, ftab, NT);
  const int f = first + threadIdx.x;
  bool valid = f < a1;
  int4 ix = make_int4(0, 0, w, -1);
  if (valid) {
    ix = D.line_idx4[f];
    valid = D.nranks <= 1 || mode == 0 || (ix.y % D.nranks) == D.rank;
  }
  __syncthreads();
  double half_rho = 0.0;
  int vi = -1;
  if (valid) {
    LineTab LT;
    { const double *q = D.lsc[buf] + 8 * (size_t)ix.y; line_table(__ldg(q), __ldg(q + 1), __ldg(q + 2), __ldg(q + 3), __ldg(q + 4), __ldg(q + 5), __ldg(q + 6), __ldg(q + 7), LT); }
    const double *sp = D.line_sp + 2 * (size_t)f, *ep = D.line_ep + 2 * (size_t)f;
    LineVpSink<kJac> sink;
    sink.ln.spx = __ldg(sp); sink.ln.spy = __ldg(sp + 1); sink.ln.epx = __ldg(ep); sink.ln.epy = __ldg(ep + 1);
    sink.ln.lf = P.line_factor; sink.ln.loss_a = P.cauchy_line; sink.ln.correct = true; sink.ln.PW = 6;
    sink.ln.want_cost = sink.vp.want_cost = cost != nullptr && (!kJac || mode == 0 || D.ctl[w].iter == 0);
    vi = ix.w;
    sink.has_vp = vi >= 0;
    sink.vp.half_rho = 0.0;
    if (vi >= 0) {
      const double *vp = D.vp_dir + 3 * (size_t)vi;
      sink.vp.vp = mk3(__ldg(vp), __ldg(vp + 1), __ldg(vp + 2));
      sink.vp.vf = P.vp_factor; sink.vp.loss_a = P.cauchy_vp; sink.vp.correct = true;
    }
    double rloc[3];
    if (kJac) {
      double *t = tile + threadIdx.x * (REC_LINE + 1), *v = vtile + threadIdx.x * (REC_VP + 1);
      sink.ln.out_r = t; sink.ln.out_jp = t + 2; sink.ln.out_jl = t + 14;
      sink.vp.out_r = v; sink.vp.out_jp = v + 1; sink.vp.out_jl = v + 7;
    } else {
      sink.ln.out_r = rloc; sink.ln.out_jp = nullptr; sink.ln.out_jl = nullptr;
      sink.vp.out_r = rloc + 2; sink.vp.out_jp = nullptr; sink.vp.out_jl = nullptr;
    }
    line_obs_eval<kJac, true>(ftab + (ix.x - fo) * FT_STRIDE, LT, sink);
    half_rho = sink.ln.half_rho + sink.vp.half_rho;
  }
  if (cost) add_window_scalar(cost, cost_stride, w, half_rho, valid);
  if (kJac) {
    ok[threadIdx.x] = valid;
    vslot[threadIdx.x] = valid ? vi : -1;
    const bool all_ok = __syncthreads_and(valid || f >= a1) != 0;
    flush_tile<REC_LINE>(tile, ok, out_line, first, min(NT, a1 - first), all_ok);
    {
#pragma unroll
      for (int i = 0; i < REC_VP; i++) {
        const int e = threadIdx.x + NT * i, slot = e / REC_VP, c = e - slot * REC_VP;
        const int v = vslot[slot];
        if (v >= 0) out_vp[(size_t)v * REC_VP + c] = vtile[e + slot];
      }
    }
  }
}

// Per-frame and per-line tables of the line / VP factors for the state buffer cur ^ cand of every window: thread per
// (frame, quaternion coordinate) and per (line, parameter).  The tables live per STATE BUFFER, so an accepted step needs
// no new ones (the candidate's tables become the current ones with the buffer flip): one launch per LM iteration.
__global__ void __launch_bounds__(128) k_line_tables(Dev D, int cand) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < 3 * D.nF) {
    const int f = t / 3, m = t - 3 * f;
    const int w = D.fr_win[f], buf = D.cur[w] ^ cand;
    line_frame_table(D.pose[buf] + 7 * (size_t)f, D.ric + 9 * (size_t)w, D.tic + 3 * (size_t)w, m, D.ftab[buf] + (size_t)f * FT_DOUBLES);
  } else if (t < 3 * D.nF + 4 * D.nL) {
    const int e = t - 3 * D.nF, l = e >> 2, c = e & 3;
    const int buf = D.cur[D.ln_win[l]] ^ cand;
    double sv, cv;
    sincos(D.ortho[buf][4 * (size_t)l + c], &sv, &cv);
    D.lsc[buf][8 * (size_t)l + 2 * c] = sv; D.lsc[buf][8 * (size_t)l + 2 * c + 1] = cv;
  }
}

// ------------------------------------------------------------------------------------------------
// IMU, two kernels.
//   k_imu_geom    one THREAD per factor: frame geometry (unweighted residual, the twelve 3x3 blocks the unweighted
//                 Jacobian is made of -> D.imu_comp, struct-of-arrays so the stores coalesce), weighted residual, cost.
//                 Every lane of a warp carries a factor (the dependent chain of quaternion algebra is ~1000
//                 instructions long; it is latency, not work).
//   k_imu_weight  one WARP per factor: [sqrt_info (15x15 upper) x J (15x30)] as a 16x16 by 16x32 product on the FP64
//                 tensor cores (mma.sync m8n8k4, 24 DMMAs - the all-zero lower-left tiles are skipped); J is expanded
//                 into shared memory with row stride 40 (= 8 mod 16: conflict-free B fragments), the product goes
//                 back through shared memory so that consecutive lanes write consecutive doubles of the record.
// Record = [r(15) | sqrt_info * J (15x30 row-major)].
constexpr int GT_IMU = 64;      // threads (= factors) per CTA of k_imu_geom
constexpr int IMU_WPC = 4;      // warps (= factors) per CTA of k_imu_weight
constexpr int JS = 40;          // row stride of the staged unweighted Jacobian

template <bool kJac>
__global__ void __launch_bounds__(GT_IMU) k_imu_geom(Dev D, Params P, int mode, int cand, double *__restrict__ out,
                                                    double *__restrict__ res_out, double *cost, int cost_stride) {
  const int f = blockIdx.x * GT_IMU + threadIdx.x;
  bool valid = f < D.nImu;
  int2 ix = make_int2(0, 0);
  if (valid) {
    ix = D.imu_idx[f];
    valid = wants<kJac>(D.ctl[ix.y].state, mode) && !(D.nranks > 1 && mode != 0 && D.rank != 0);
  }
  double half = 0.0;
  if (valid) {
    const int buf = D.cur[ix.y] ^ cand;
    ImuIn in;
    in.pose_i = D.pose[buf] + 7 * (size_t)ix.x; in.pose_j = in.pose_i + 7;
    in.sb_i = D.sb[buf] + 9 * (size_t)ix.x; in.sb_j = in.sb_i + 9;
    in.dp = D.imu_dp + 3 * (size_t)f; in.dq = D.imu_dq + 4 * (size_t)f; in.dv = D.imu_dv + 3 * (size_t)f;
    in.lin_ba = D.imu_lin_ba + 3 * (size_t)f; in.lin_bg = D.imu_lin_bg + 3 * (size_t)f;
    in.sum_dt = __ldg(D.imu_sum_dt + f);
    in.jac = D.imu_jac + 225 * (size_t)f;
    in.sqrt_info = D.imu_sqrt_info + 225 * (size_t)f;
    double raw[15];
    imu_geometry<kJac>(in, P.g, raw, D.imu_comp + f, D.nImu);
    if (kJac) {
      // Jacobian mode: the unweighted residual rides to k_imu_weight as a 31st column of the unweighted Jacobian, so the
      // sqrt_info product of residual and Jacobian is ONE tensor-core product there (and this thread never reads sqrt_info)
#pragma unroll
      for (int i = 0; i < 15; i++) D.imu_comp[(size_t)(IMU_COMP + i) * D.nImu + f] = raw[i];
    } else {
      const double *SI = in.sqrt_info;
      double *rdst = res_out ? res_out + 15 * (size_t)f : nullptr;
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < 15; i++) {
        double r = 0.0;
#pragma unroll
        for (int k = 0; k < 15; k++) if (k >= i) r += __ldg(SI + i * 15 + k) * raw[k];
        if (rdst) rdst[i] = r;
        s += r * r;
      }
      half = 0.5 * s;
    }
  }
  if (!kJac && cost) add_window_scalar(cost, cost_stride, ix.y, half, valid);
}

// [r | J] = sqrt_info x [raw | J_raw]: a CTA takes IMU_WPC consecutive factors at a time (their compact blocks are
// IMU_WPC consecutive doubles of every row of imu_comp = one 32-byte sector, loaded cooperatively), a warp per factor.
// The zero pattern of J_raw is the same for every factor, so the staging tile is cleared once per CTA.
constexpr int IMU_CS = IMU_COMP + 15;   // rows of imu_comp: compact Jacobian blocks + unweighted residual
constexpr int IMU_CSP = IMU_CS + 1;     // padded row of the staged copy
__global__ void __launch_bounds__(32 * IMU_WPC) k_imu_weight(Dev D, int mode, double *__restrict__ out, double *cost, int cost_stride) {
  __shared__ __align__(16) double Jraw_all[IMU_WPC][16 * JS];
  __shared__ __align__(16) double prod_all[IMU_WPC][15 * JS];   // row stride 40 = 8 mod 16 doubles: conflict-free double2 stores of the accumulator fragments
  __shared__ double comp_s[IMU_WPC][IMU_CSP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double *Jraw = Jraw_all[warp], *prod = prod_all[warp];
  for (int e = lane; e < 16 * JS; e += 32) Jraw[e] = 0.0;
  const int fr = lane >> 2, fc = lane & 3;
  const int ce = threadIdx.x / IMU_WPC, ck = threadIdx.x % IMU_WPC;   // this thread's share of a group's compact blocks
  // both inputs of a group are fetched one group ahead (registers), so a group costs one memory round trip that overlaps
  // the previous group's product
  double cv[4], a[2][4];
  auto fetch_comp = [&](int f0) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int e = ce + 32 * i;
      cv[i] = (e < IMU_CS && f0 + ck < D.nImu) ? __ldg(D.imu_comp + (size_t)e * D.nImu + f0 + ck) : 0.0;
    }
  };
  auto fetch_info = [&](int f) {   // A fragments: sqrt_info[8 mt + fr][4 ks + fc] (upper triangular, row / column 15 = padding)
    const double *SI = D.imu_sqrt_info + 225 * (size_t)min(f, D.nImu - 1);
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
      for (int ks = 0; ks < 4; ks++) {
        const int row = 8 * mt + fr, col = 4 * ks + fc;
        a[mt][ks] = (row < 15 && col < 15 && col >= row) ? __ldg(SI + row * 15 + col) : 0.0;
      }
  };
  int f0 = blockIdx.x * IMU_WPC;
  if (f0 < D.nImu) { fetch_comp(f0); fetch_info(f0 + warp); }
  for (; f0 < D.nImu; f0 += gridDim.x * IMU_WPC) {
    __syncthreads();   // comp_s of the previous group is consumed
#pragma unroll
    for (int i = 0; i < 4; i++) if (ce + 32 * i < IMU_CS) comp_s[ck][ce + 32 * i] = cv[i];
    double a0[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
      for (int ks = 0; ks < 4; ks++) a0[mt][ks] = a[mt][ks];
    __syncthreads();
    const int fn = f0 + gridDim.x * IMU_WPC;
    if (fn < D.nImu) { fetch_comp(fn); fetch_info(fn + warp); }
    const int f = f0 + warp;
    if (f >= D.nImu) continue;
    const int2 ix = D.imu_idx[f];
    if (!wants<true>(D.ctl[ix.y].state, mode) || (D.nranks > 1 && mode != 0 && D.rank != 0)) continue;   // uniform over the warp
    const double *cs = comp_s[warp];
    for (int e = lane; e < 18 * 9; e += 32) {
      const int b = e / 9, k = e - 9 * b;
      const ImuPut p = c_imu_puts[b];
      Jraw[(p.r0 + k / 3) * JS + p.c0 + k % 3] = (double)p.sgn * cs[9 * p.blk + k];
    }
    if (lane < 15) Jraw[lane * JS + 30] = cs[IMU_COMP + lane];
    __syncwarp();
    double acc[2][4][2];
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
      for (int nt = 0; nt < 4; nt++) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
#pragma unroll
    for (int ks = 0; ks < 4; ks++)
#pragma unroll
      for (int nt = 0; nt < 4; nt++) {
        const double b = Jraw[(4 * ks + fc) * JS + 8 * nt + fr];   // B[k][n] = J[k][n]
        dmma884(acc[0][nt], a0[0][ks], b);
        if (ks >= 2) dmma884(acc[1][nt], a0[1][ks], b);              // rows 8..14 of sqrt_info start at column 8
      }
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
      for (int nt = 0; nt < 4; nt++) {
        const int row = 8 * mt + fr;
        if (row < 15) *reinterpret_cast<double2 *>(prod + row * JS + 8 * nt + 2 * fc) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
      }
    __syncwarp();
    // record = [r (15) | J (15 x 30 row-major)]: consecutive lanes write consecutive doubles
    double *rec = out + (size_t)f * REC_IMU;
    double rr = 0.0;
    if (lane < 15) { rr = prod[lane * JS + 30]; rec[lane] = rr; }
    for (int e = lane; e < 450; e += 32) { const int row = e / 30, col = e - 30 * row; rec[15 + e] = prod[row * JS + col]; }
    if (cost) {
      double s = rr * rr;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
      if (lane == 0) atomicAdd(cost + (size_t)ix.y * cost_stride, 0.5 * s);
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// Prior residual r = r0 + J0 dx (its Jacobian is J0 itself).  One CTA per window, a warp per row.
__global__ void __launch_bounds__(1024) k_prior(Dev D, int jac_phase, int mode, int cand, double *__restrict__ res_out,
                                              double *cost, int cost_stride) {
  extern __shared__ double dx[];
  const int w = blockIdx.x;
  const int n = D.prior_off[w + 1] - D.prior_off[w];
  if (n <= 0) return;
  if (jac_phase ? !wants<true>(D.ctl[w].state, mode) : !wants<false>(D.ctl[w].state, mode)) return;
  if (D.nranks > 1 && D.rank != 0) return;
  const int buf = D.cur[w] ^ cand;
  const int b0 = D.pblk_off[w], b1 = D.pblk_off[w + 1];
  const int nthr = blockDim.x;   // 128 for batches that fill the GPU, 1024 for a handful of windows (one round of rows instead of six)
  for (int b = b0 + threadIdx.x; b < b1; b += nthr) {
    const int kind = D.pblk_kind[b], row = D.pblk_row[b], col = D.pblk_col[b];
    const double *x0 = D.prior_x0 + 9 * (size_t)b;
    if (kind == 0 || kind == 2) {   // pose / ex-pose: [dp ; 2 vec(q0^-1 q)] with sign fix
      const double *x = (kind == 0 ? D.pose[buf] : D.ex[buf]) + 7 * (size_t)row;
      for (int k = 0; k < 3; k++) dx[col + k] = x[k] - x0[k];
      const q4 dq = qmul(qinv(mkq(x0[3], x0[4], x0[5], x0[6])), mkq(x[3], x[4], x[5], x[6]));
      const double sgn = (dq.w >= 0.0) ? 2.0 : -2.0;
      dx[col + 3] = sgn * dq.x; dx[col + 4] = sgn * dq.y; dx[col + 5] = sgn * dq.z;
    } else if (kind == 1) {
      const double *x = D.sb[buf] + 9 * (size_t)row;
      for (int k = 0; k < 9; k++) dx[col + k] = x[k] - x0[k];
    } else {
      dx[col] = D.td[buf][row] - x0[0];
    }
  }
  __syncthreads();
  const double *J0 = D.prior_J + D.priorJ_off[w];
  const double *r0 = D.prior_r0 + D.prior_off[w];
  double *rout = res_out ? res_out + D.prior_off[w] : nullptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double csum = 0.0;
  // a warp per row, four rows in flight per warp (independent accumulators keep the loads overlapped)
  constexpr int RB = 4;
  const int NW = nthr / 32;
  for (int i0 = warp * RB; i0 < n; i0 += NW * RB) {
    double acc[RB];
#pragma unroll
    for (int u = 0; u < RB; u++) acc[u] = 0.0;
    for (int k = lane; k < n; k += 32) {
      const double x = dx[k];
#pragma unroll
      for (int u = 0; u < RB; u++) if (i0 + u < n) acc[u] += __ldg(J0 + (size_t)(i0 + u) * n + k) * x;
    }
#pragma unroll
    for (int u = 0; u < RB; u++) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[u] += __shfl_down_sync(0xffffffffu, acc[u], o);
      if (lane == 0 && i0 + u < n) {
        const double r = __ldg(r0 + i0 + u) + acc[u];
        if (rout) rout[i0 + u] = r;
        csum += r * r;
      }
    }
  }
  if (lane == 0 && cost) atomicAdd(cost + (size_t)w * cost_stride, 0.5 * csum);
}

// ------------------------------------------------------------------------------------------------
// launch wrappers
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

template <int REC>
static size_t tile_bytes() { return (size_t)NT * (REC + 1) * sizeof(double) + NT; }

// tiles of the widest records exceed the 48 KB default of dynamic shared memory
// developer switch: resident CTAs per SM the register allocation of a sweep kernel is bounded for (A/B runs)
static int sweep_occ(const char *name, int dflt) {
  const char *e = std::getenv(name);
  return e ? std::atoi(e) : dflt;
}

static void raise_smem_limits() {
  static bool done = false;
  if (done) return;
  done = true;
  cudaFuncSetAttribute(k_proj<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_bytes<CREC_PROJ_TD>());
  cudaFuncSetAttribute(k_proj<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_bytes<CREC_PROJ>());
  cudaFuncSetAttribute(k_proj<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_bytes<REC_PROJ_TD>());
  cudaFuncSetAttribute(k_proj<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_bytes<REC_PROJ>());
  cudaFuncSetAttribute(k_proj<true, false, false, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_bytes<REC_PROJ>());
}

int launch_proj(const Dev &D, const Params &P, bool jac, bool ceres, int mode, int cand, double *out, double *res_out,
                double *cost, int cost_stride, cudaStream_t st) {
  if (D.nProj == 0) return 0;
  raise_smem_limits();
  const int grid = cdiv(D.nProj, NT);
  const bool td = D.estimate_td != 0;
  if (!jac) {
    if (td) k_proj<false, true, false><<<grid, NT, 0, st>>>(D, P, mode, cand, out, res_out, cost, cost_stride);
    else k_proj<false, false, false><<<grid, NT, 0, st>>>(D, P, mode, cand, out, res_out, cost, cost_stride);
    return 1;
  }
  if (ceres) {
    if (td) k_proj<true, true, true><<<grid, NT, tile_bytes<CREC_PROJ_TD>(), st>>>(D, P, mode, cand, out, res_out, cost, cost_stride);
    else k_proj<true, false, true><<<grid, NT, tile_bytes<CREC_PROJ>(), st>>>(D, P, mode, cand, out, res_out, cost, cost_stride);
  } else {
    if (td) k_proj<true, true, false><<<grid, NT, tile_bytes<REC_PROJ_TD>(), st>>>(D, P, mode, cand, out, res_out, cost, cost_stride);
    else if (sweep_occ("UVS_PROJ_OCC", 4) == 5) k_proj<true, false, false, 5><<<grid, NT, tile_bytes<REC_PROJ>(), st>>>(D, P, mode, cand, out, res_out, cost, cost_stride);
    else k_proj<true, false, false><<<grid, NT, tile_bytes<REC_PROJ>(), st>>>(D, P, mode, cand, out, res_out, cost, cost_stride);
  }
  return 1;
}

int launch_line(const Dev &D, const Params &P, bool jac, bool ceres, int mode, int cand, double *out, double *res_out,
                double *cost, int cost_stride, cudaStream_t st) {
  if (D.nLobs == 0) return 0;
  const int grid = cdiv(D.nLobs, NT);
  if (!jac) k_line<false, false><<<grid, NT, 0, st>>>(D, P, mode, cand, out, res_out, cost, cost_stride);
  else if (ceres) k_line<true, true><<<grid, NT, tile_bytes<CREC_LINE>(), st>>>(D, P, mode, cand, out, res_out, cost, cost_stride);
  else k_line<true, false><<<grid, NT, tile_bytes<REC_LINE>(), st>>>(D, P, mode, cand, out, res_out, cost, cost_stride);
  return 1;
}

int launch_vp(const Dev &D, const Params &P, bool jac, bool ceres, int mode, int cand, double *out, double *res_out,
              double *cost, int cost_stride, cudaStream_t st) {
  if (D.nVobs == 0) return 0;
  const int grid = cdiv(D.nVobs, NT);
  if (!jac) k_vp<false, false><<<grid, NT, 0, st>>>(D, P, mode, cand, out, res_out, cost, cost_stride);
  else if (ceres) k_vp<true, true><<<grid, NT, tile_bytes<CREC_VP>(), st>>>(D, P, mode, cand, out, res_out, cost, cost_stride);
  else k_vp<true, false><<<grid, NT, tile_bytes<REC_VP>(), st>>>(D, P, mode, cand, out, res_out, cost, cost_stride);
  return 1;
}

int launch_line_tables(const Dev &D, int cand, cudaStream_t st) {
  const int n = 3 * D.nF + 4 * D.nL;
  if (D.nLobs == 0 || n == 0) return 0;
  k_line_tables<<<cdiv(n, 128), 128, 0, st>>>(D, cand);
  return 1;
}

static size_t line_vp_smem(int max_frames, bool jac) {
  size_t dbl = (size_t)((max_frames * FT_STRIDE + 1) & ~1);
  if (jac) return (dbl + (size_t)NT * (REC_LINE + 1 + REC_VP + 1)) * sizeof(double) + NT * sizeof(int) + NT;
  return dbl * sizeof(double);
}

int launch_line_vp(const Dev &D, const Params &P, bool jac, int mode, int cand, double *out_line, double *out_vp, double *cost,
                   int cost_stride, cudaStream_t st) {
  if (D.nLobs == 0 || D.max_lobs == 0) return 0;
  const dim3 grid(cdiv(D.max_lobs, NT), D.B);
  const size_t smem = line_vp_smem(D.max_frames, jac);
  static size_t raised = 0;
  if (jac && smem > raised) {
    cudaFuncSetAttribute(k_line_vp<true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_line_vp<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_line_vp<true, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_line_vp<true, 3>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(k_line_vp<true, 4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(k_line_vp<true, 5>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    raised = smem;
  }
  // resident CTAs per SM the registers are bounded for: 4 measured best (99 -> 85 us on the C2 x 1184 batch; 3 = 134 registers)
  const int occ = jac ? sweep_occ("UVS_LINE_OCC", 4) : 0;
  if (jac && occ == 3) k_line_vp<true, 3><<<grid, NT, smem, st>>>(D, P, mode, cand, out_line, out_vp, cost, cost_stride);
  else if (jac && occ == 5) k_line_vp<true, 5><<<grid, NT, smem, st>>>(D, P, mode, cand, out_line, out_vp, cost, cost_stride);
  else if (jac) k_line_vp<true, 4><<<grid, NT, smem, st>>>(D, P, mode, cand, out_line, out_vp, cost, cost_stride);
  else k_line_vp<false, 4><<<grid, NT, smem, st>>>(D, P, mode, cand, out_line, out_vp, cost, cost_stride);
  return 1;
}

int launch_imu(const Dev &D, const Params &P, bool jac, int mode, int cand, double *out, double *res_out, double *cost,
               int cost_stride, cudaStream_t st) {
  if (D.nImu == 0) return 0;
  const int grid = cdiv(D.nImu, GT_IMU);
  if (!jac) { k_imu_geom<false><<<grid, GT_IMU, 0, st>>>(D, P, mode, cand, out, res_out, cost, cost_stride); return 1; }
  k_imu_geom<true><<<grid, GT_IMU, 0, st>>>(D, P, mode, cand, out, res_out, cost, cost_stride);
  k_imu_weight<<<std::min(cdiv(D.nImu, IMU_WPC), 148 * 5), 32 * IMU_WPC, 0, st>>>(D, mode, out, cost, cost_stride);
  return 2;
}

int launch_prior(const Dev &D, int max_prior_n, bool jac_phase, int mode, int cand, double *res_out, double *cost,
                 int cost_stride, cudaStream_t st) {
  if (D.nPriorR == 0) return 0;
  k_prior<<<D.B, D.B >= 74 ? NT : 1024, (size_t)max_prior_n * sizeof(double), st>>>(D, jac_phase ? 1 : 0, mode, cand, res_out, cost, cost_stride);
  return 1;
}

}  // namespace uvs

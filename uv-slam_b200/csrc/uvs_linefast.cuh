// uvs_linefast.cuh — line / vanishing-point factor evaluation through per-frame and per-line tables.
//
// Same functions as line_to_camera() in uvs_factors.cuh (LineProjectionFactor / VPProjectionFactor,
// factor/line_projection_factor.h:21-57, factor/vp_projection_factor.h:24-61; derivatives w.r.t. the RAW quaternion
// coordinates = what Ceres AutoDiff yields), factored so that everything that depends only on the FRAME or only on
// the LINE is computed once per CTA / lane group and kept in shared memory / registers:
//
//   frame table (FT_DOUBLES per frame):  A = ric^T R^T (9) | av = A t_wc (3) | for m = x, y, z:
//                                        M_m = ric^T (dR/dq_m)^T (9) | da_m = M_m t_wc + A (dR/dq_m) t_ic (3)
//   line table:                          u0, u1 (columns 0, 1 of U = Rx Ry Rz), their psi_x / psi_y derivatives,
//                                        sin(phi), cos(phi)
//
// per observation:  d_c = sin(phi) A u1,  n_c = cos(phi) A u0 - av x d_c  and ten partials, ~500 FP64 instructions
// instead of ~1500.  The raw-qw column of the Ceres layout is not produced here (API evaluation in Ceres layout keeps
// line_to_camera).
#pragma once
#include "uvs_math.cuh"

namespace uvs {

constexpr int FT_DOUBLES = 48;
constexpr int FT_STRIDE = 49;   // odd row stride: lanes that read different frames hit different banks

// part `m` (0..2) of the frame table of one frame; the thread with m == 0 also writes A and av
__device__ __forceinline__ void line_frame_table(const double *__restrict__ pose, const double *__restrict__ ric_rm,
                                                 const double *__restrict__ tic3, int m, double *T) {
  d3 p; q4 q;
  load_pose(pose, p, q);
  m33 ric;
#pragma unroll
  for (int k = 0; k < 9; k++) ric.a[k] = __ldg(ric_rm + k);
  const d3 tic = mk3(__ldg(tic3), __ldg(tic3 + 1), __ldg(tic3 + 2));
  const m33 R = qmat(q);
  const m33 A = mtmul(ric, mtrans(R));
  const d3 twc = mvec(R, tic) + p;
  if (m == 0) {
#pragma unroll
    for (int k = 0; k < 9; k++) T[k] = A.a[k];
    const d3 av = mvec(A, twc);
    T[9] = av.x; T[10] = av.y; T[11] = av.z;
  }
  const double x2 = 2 * q.x, y2 = 2 * q.y, z2 = 2 * q.z, w2 = 2 * q.w;
  // G = dR/dq_m applied as G^T v and G v (R(q) = I + 2 w [u]x + 2 [u]x^2, not normalised)
  auto GT = [&](d3 v) -> d3 {
    if (m == 0) return mk3(y2 * v.y + z2 * v.z, y2 * v.x - 2 * x2 * v.y + w2 * v.z, z2 * v.x - w2 * v.y - 2 * x2 * v.z);
    if (m == 1) return mk3(-2 * y2 * v.x + x2 * v.y - w2 * v.z, x2 * v.x + z2 * v.z, w2 * v.x + z2 * v.y - 2 * y2 * v.z);
    return mk3(-2 * z2 * v.x + w2 * v.y + x2 * v.z, -w2 * v.x - 2 * z2 * v.y + y2 * v.z, x2 * v.x + y2 * v.y);
  };
  auto G = [&](d3 v) -> d3 {
    if (m == 0) return mk3(y2 * v.y + z2 * v.z, y2 * v.x - 2 * x2 * v.y - w2 * v.z, z2 * v.x + w2 * v.y - 2 * x2 * v.z);
    if (m == 1) return mk3(-2 * y2 * v.x + x2 * v.y + w2 * v.z, x2 * v.x + z2 * v.z, -w2 * v.x + z2 * v.y - 2 * y2 * v.z);
    return mk3(-2 * z2 * v.x - w2 * v.y + x2 * v.z, w2 * v.x - 2 * z2 * v.y + y2 * v.z, x2 * v.x + y2 * v.y);
  };
  double *Mo = T + 12 + 12 * m;
  // column c of M_m = ric^T G^T e_c
#pragma unroll
  for (int c = 0; c < 3; c++) {
    const d3 col = mtvec(ric, GT(mk3(c == 0 ? 1.0 : 0.0, c == 1 ? 1.0 : 0.0, c == 2 ? 1.0 : 0.0)));
    Mo[c] = col.x; Mo[3 + c] = col.y; Mo[6 + c] = col.z;
  }
  const d3 da = mtvec(ric, GT(twc)) + mvec(A, G(tic));
  Mo[9] = da.x; Mo[10] = da.y; Mo[11] = da.z;
}

// copy of one window's frame tables (contiguous in global memory, FT_DOUBLES per frame) into shared memory with the
// padded row stride; call with all threads of the CTA, then __syncthreads()
__device__ __forceinline__ void load_frame_tables(const double *__restrict__ src, int F, double *ftab, int nthreads) {
  for (int e = threadIdx.x; e < F * FT_DOUBLES; e += nthreads) {
    const int f = e / FT_DOUBLES;
    ftab[e + f * (FT_STRIDE - FT_DOUBLES)] = __ldg(src + e);
  }
}

struct LineTab {
  d3 u0, u1, du0a, du1a, du0b, du1b;
  double sp, cp;
};

// from the eight sines / cosines of the orthonormal line parameters (psi_x, psi_y, psi_z, phi)
__device__ __forceinline__ void line_table(double sa, double ca, double sb, double cb, double sc, double cc, double sp, double cp,
                                           LineTab &L) {
  L.u0 = mk3(cb * cc, ca * sc + sa * sb * cc, sa * sc - ca * sb * cc);
  L.u1 = mk3(-cb * sc, ca * cc - sa * sb * sc, sa * cc + ca * sb * sc);
  L.du0a = mk3(0.0, -sa * sc + ca * sb * cc, ca * sc + sa * sb * cc);
  L.du1a = mk3(0.0, -sa * cc - ca * sb * sc, ca * cc - sa * sb * sc);
  L.du0b = mk3(-sb * cc, sa * cb * cc, -ca * cb * cc);
  L.du1b = mk3(sb * sc, -sa * cb * sc, ca * cb * sc);
  L.sp = sp; L.cp = cp;
}
__device__ __forceinline__ void line_table(const double *__restrict__ line, LineTab &L) {
  double sa, ca, sb, cb, sc, cc, sp, cp;
  sincos(__ldg(line), &sa, &ca); sincos(__ldg(line + 1), &sb, &cb); sincos(__ldg(line + 2), &sc, &cc); sincos(__ldg(line + 3), &sp, &cp);
  line_table(sa, ca, sb, cb, sc, cc, sp, cp, L);
}

__device__ __forceinline__ d3 mv9(const double *M, d3 v) {
  return mk3(M[0] * v.x + M[1] * v.y + M[2] * v.z, M[3] * v.x + M[4] * v.y + M[5] * v.z, M[6] * v.x + M[7] * v.y + M[8] * v.z);
}

// One observation: the same producer protocol as line_to_camera (sink.base(n_c, d_c), then sink.partial(k, dn, du) for
// k = 0-2 (p), 3-5 (raw qx, qy, qz), 6-8 (psi), 9 (phi)).  T = frame table of the observing frame (shared memory).
template <bool kJac, bool kNeedN, class Sink>
__device__ __forceinline__ void line_obs_eval(const double *T, const LineTab &L, Sink &sink) {
  const d3 av = mk3(T[9], T[10], T[11]);
  const d3 a0 = mv9(T, L.u0), a1 = mv9(T, L.u1);
  const d3 u = L.sp * a1;
  const d3 zero = mk3(0, 0, 0);
  sink.base(kNeedN ? L.cp * a0 - cross(av, u) : zero, u);
  if (!kJac) return;
#pragma unroll
  for (int k = 0; k < 3; k++) sink.partial(k, kNeedN ? -cross(mk3(T[k], T[3 + k], T[6 + k]), u) : zero, zero);
#pragma unroll
  for (int m = 0; m < 3; m++) {
    const double *M = T + 12 + 12 * m;
    const d3 du = L.sp * mv9(M, L.u1);
    d3 dn = zero;
    if (kNeedN) dn = L.cp * mv9(M, L.u0) - cross(mk3(M[9], M[10], M[11]), u) - cross(av, du);
    sink.partial(3 + m, dn, du);
  }
  { const d3 du = L.sp * mv9(T, L.du1a); sink.partial(6, kNeedN ? L.cp * mv9(T, L.du0a) - cross(av, du) : zero, du); }
  { const d3 du = L.sp * mv9(T, L.du1b); sink.partial(7, kNeedN ? L.cp * mv9(T, L.du0b) - cross(av, du) : zero, du); }
  { const d3 du = (-L.sp) * a0; sink.partial(8, kNeedN ? L.cp * a1 - cross(av, du) : zero, du); }
  { const d3 du = L.cp * a1; sink.partial(9, kNeedN ? (-L.sp) * a0 - cross(av, du) : zero, du); }
}

}  // namespace uvs

// uvs_factors.cuh — device-side residual / Jacobian evaluation of the five factor types.
//
// Jacobians are produced directly in LOCAL layout (tangent columns: first 6 columns of every pose
// block, pose_local_parameterization.cpp:20-27), raw (no loss correction).  Line / VP Jacobians are
// hand-derived analytic derivatives w.r.t. the RAW quaternion coordinates (qx,qy,qz), which is what
// Ceres AutoDiff yields for the reference's functors (SURVEY.md 8a "rotation-column quirk"); the CPU
// oracle obtains the same numbers by forward-mode dual numbers, so the two are independent.
#pragma once
#include "uvs_math.cuh"

namespace uvs {

// ---------------------------------------------------------------------------------------------
// ProjectionFactor::Evaluate (factor/projection_factor.cpp:22-175) and the td variant
// (factor/projection_td_factor.cpp:34-145).
//   J layout: Ji[2x6] Jj[2x6] Jex[2x6] Jl[2] (Jtd[2])   row-major inside each block
struct ProjTd {
  double td, td_i, td_j, row_i, row_j, vix, viy, vjx, vjy, tr_over_row, half_row;
};

// Writes r[2] and the Jacobian blocks through the given pointers (row stride `ld` inside a pose block:
// 6 in the tangent layout, 7 in the Ceres layout; 7th column is the caller's business).  When
// loss_a > 0 the Cauchy corrector (sqrt(rho') scaling, marginalization_factor.cpp:37-68) is folded
// into the outputs; *half_rho returns 1/2 rho(|r|^2).  Outputs may live in shared memory: every value
// is written once, as soon as it is known, so nothing stays live in registers.
template <bool kJac, bool kTd>
__device__ __forceinline__ void proj_eval(const double *__restrict__ pose_i, const double *__restrict__ pose_j,
                                          const double *__restrict__ ex, double inv_dep, d3 pts_i, d3 pts_j, double S,
                                          const ProjTd *tdp, bool want_ex, double loss_a, bool correct, int ld, double *r,
                                          double *Ji, double *Jj, double *Jex, double *Jl, double *Jtd, double *half_rho) {
  d3 Pi, Pj, tic; q4 Qi, Qj, qic;
  load_pose(pose_i, Pi, Qi);
  load_pose(pose_j, Pj, Qj);
  load_pose(ex, tic, qic);
  d3 vel_i = mk3(0, 0, 0);
  if (kTd) {
    const double si = tdp->td - tdp->td_i + tdp->tr_over_row * (tdp->row_i - tdp->half_row);
    const double sj = tdp->td - tdp->td_j + tdp->tr_over_row * (tdp->row_j - tdp->half_row);
    vel_i = mk3(tdp->vix, tdp->viy, 0.0);
    pts_i = pts_i - si * vel_i;
    pts_j = pts_j - sj * mk3(tdp->vjx, tdp->vjy, 0.0);
  }
  const double depth = 1.0 / inv_dep;
  const d3 pci = depth * pts_i;
  const d3 pbi = qrot(qic, pci) + tic;
  const d3 pw = qrot(Qi, pbi) + Pi;
  const d3 pbj = qrot(qinv(Qj), pw - Pj);
  const d3 pcj = qrot(qinv(qic), pbj - tic);
  const double iz = 1.0 / pcj.z;
  double r0 = S * (pcj.x * iz - pts_j.x), r1 = S * (pcj.y * iz - pts_j.y);
  const double s = r0 * r0 + r1 * r1;
  double sq = 1.0;
  // half_rho == nullptr: the caller does not need the cost (see cauchy_rho1)
  if (correct && loss_a > 0.0) {
    if (half_rho) {
      double rho0, rho1;
      cauchy(loss_a, s, rho0, rho1);
      *half_rho = 0.5 * rho0;
      sq = sqrt(rho1);
    } else {
      sq = sqrt(cauchy_rho1(loss_a, s));
    }
  } else if (half_rho) {
    *half_rho = 0.5 * s;
  }
  r[0] = sq * r0; r[1] = sq * r1;
  if (!kJac) return;

  const m33 Ri = qmat(Qi), Rj = qmat(Qj), Ric = qmat(qic);
  const double Ss = S * sq;
  const double r00 = Ss * iz, r02 = -Ss * pcj.x * iz * iz, r12 = -Ss * pcj.y * iz * iz;
  // out[2x3] = reduce * M
  auto red = [&](const m33 &M, double *o, double sgn) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
      o[c] = sgn * (r00 * M.a[c] + r02 * M.a[6 + c]);
      o[ld + c] = sgn * (r00 * M.a[3 + c] + r12 * M.a[6 + c]);
    }
  };
  const m33 A = mtmul(Ric, mtrans(Rj));  // ric^T Rj^T
  const m33 B = mmul(A, Ri);             // ric^T Rj^T Ri
  red(A, Ji, 1.0);
  red(mmul(B, skew(pbi)), Ji + 3, -1.0);
  red(A, Jj, -1.0);
  red(mtmul(Ric, skew(pbj)), Jj + 3, 1.0);
  const m33 T = mmul(B, Ric);  // tmp_r
  if (!want_ex) {
    if (Jex) {   // callers that never use the extrinsic block pass no storage for it
#pragma unroll
      for (int k = 0; k < 6; k++) { Jex[k] = 0.0; Jex[ld + k] = 0.0; }
    }
  } else {
    m33 L = B;  // ric^T (Rj^T Ri - I) = B - ric^T
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) L.a[3 * i + j] -= Ric.a[3 * j + i];
    red(L, Jex, 1.0);
    const d3 tp = mvec(T, pci);
    const d3 e = mtvec(Ric, mtvec(Rj, mvec(Ri, tic) + Pi - Pj) - tic);
    m33 Rr = mmul(T, skew(pci));
    const m33 s1 = skew(tp + e);
#pragma unroll
    for (int k = 0; k < 9; k++) Rr.a[k] = s1.a[k] - Rr.a[k];
    red(Rr, Jex + 3, 1.0);
  }
  {
    const double f = -depth * depth;
    const d3 v = mvec(T, pts_i);
    Jl[0] = (r00 * v.x + r02 * v.z) * f;
    Jl[1] = (r00 * v.y + r12 * v.z) * f;
  }
  if (kTd) {
    const double f = -depth;
    const d3 v = mvec(T, vel_i);
    Jtd[0] = (r00 * v.x + r02 * v.z) * f + Ss * tdp->vjx;
    Jtd[1] = (r00 * v.y + r12 * v.z) * f + Ss * tdp->vjy;
  }
}

// ---------------------------------------------------------------------------------------------
// Shared transform of LineProjectionFactor / VPProjectionFactor (line_projection_factor.h:21-53).
// Produces n_c, d_c and their 10 partials: columns 0-2 = p, 3-5 = raw (qx,qy,qz), 6-8 = psi, 9 = phi.
// The transform is written as a producer: `sink.base(n_c, d_c)` is called once, then
// `sink.partial(k, dn_c/dtheta_k, dd_c/dtheta_k)` for k = 0-2 (p), 3-5 (raw qx,qy,qz), 6-8 (psi), 9 (phi) and, when kQw,
// 10 (raw qw).  The consumer turns every partial into its Jacobian entries immediately, so no array of partials
// stays live in registers.
template <bool kJac, bool kNeedN, bool kQw, class Sink>
__device__ __forceinline__ void line_to_camera(const double *__restrict__ pose, const double *__restrict__ line,
                                               const double *__restrict__ ric_rm, const double *__restrict__ tic3,
                                               Sink &sink) {
  d3 p; q4 q;
  load_pose(pose, p, q);
  m33 ric;
#pragma unroll
  for (int k = 0; k < 9; k++) ric.a[k] = __ldg(ric_rm + k);
  const d3 tic = mk3(__ldg(tic3), __ldg(tic3 + 1), __ldg(tic3 + 2));
  const double a = __ldg(line), b = __ldg(line + 1), c = __ldg(line + 2), phi = __ldg(line + 3);
  double sa, ca, sb, cb, sc, cc, sp, cp;
  sincos(a, &sa, &ca); sincos(b, &sb, &cb); sincos(c, &sc, &cc); sincos(phi, &sp, &cp);
  // U = Rx(a) Ry(b) Rz(c): columns 0 and 1
  const d3 u0 = mk3(cb * cc, ca * sc + sa * sb * cc, sa * sc - ca * sb * cc);
  const d3 u1 = mk3(-cb * sc, ca * cc - sa * sb * sc, sa * cc + ca * sb * sc);
  const d3 nw = cp * u0, dw = sp * u1;
  const m33 R = qmat(q);
  const m33 A = mtmul(ric, mtrans(R));  // R_wc^T = ric^T R^T
  const d3 twc = mvec(R, tic) + p;
  const d3 av = mvec(A, twc);           // -t_cw
  const d3 u = mvec(A, dw);             // d_c
  const d3 zero = mk3(0, 0, 0);
  sink.base(kNeedN ? mvec(A, nw) - cross(av, u) : zero, u);
  if (!kJac) return;
  // translation
#pragma unroll
  for (int k = 0; k < 3; k++) sink.partial(k, kNeedN ? -cross(mcol(A, k), u) : zero, zero);
  // raw quaternion coordinates: G_m = dR/dq_m, applied as G^T v (m = 0,1,2: x,y,z; 3: w)
  {
    const double x2 = 2 * q.x, y2 = 2 * q.y, z2 = 2 * q.z, w2 = 2 * q.w;
    auto GT = [&](int m, d3 v) -> d3 {   // G_m^T v
      if (m == 0) return mk3(y2 * v.y + z2 * v.z, y2 * v.x - 2 * x2 * v.y + w2 * v.z, z2 * v.x - w2 * v.y - 2 * x2 * v.z);
      if (m == 1) return mk3(-2 * y2 * v.x + x2 * v.y - w2 * v.z, x2 * v.x + z2 * v.z, w2 * v.x + z2 * v.y - 2 * y2 * v.z);
      if (m == 2) return mk3(-2 * z2 * v.x + w2 * v.y + x2 * v.z, -w2 * v.x - 2 * z2 * v.y + y2 * v.z, x2 * v.x + y2 * v.y);
      return mk3(z2 * v.y - y2 * v.z, -z2 * v.x + x2 * v.z, y2 * v.x - x2 * v.y);
    };
    auto G = [&](int m, d3 v) -> d3 {    // G_m v
      if (m == 0) return mk3(y2 * v.y + z2 * v.z, y2 * v.x - 2 * x2 * v.y - w2 * v.z, z2 * v.x + w2 * v.y - 2 * x2 * v.z);
      if (m == 1) return mk3(-2 * y2 * v.x + x2 * v.y + w2 * v.z, x2 * v.x + z2 * v.z, -w2 * v.x + z2 * v.y - 2 * y2 * v.z);
      if (m == 2) return mk3(-2 * z2 * v.x - w2 * v.y + x2 * v.z, w2 * v.x - 2 * z2 * v.y + y2 * v.z, x2 * v.x + y2 * v.y);
      return mk3(-z2 * v.y + y2 * v.z, z2 * v.x - x2 * v.z, -y2 * v.x + x2 * v.y);
    };
#pragma unroll
    for (int m = 0; m < (kQw ? 4 : 3); m++) {
      const int slot = m < 3 ? 3 + m : 10;
      // A' v = ric^T (G^T v)
      const d3 du = mtvec(ric, GT(m, dw));
      d3 dn = zero;
      if (kNeedN) {
        const d3 dm = mtvec(ric, GT(m, nw));
        const d3 da = mtvec(ric, GT(m, twc)) + mvec(A, G(m, tic));
        dn = dm - cross(da, u) - cross(av, du);
      }
      sink.partial(slot, dn, du);
    }
  }
  // line parameters
  {
    const d3 du0a = mk3(0.0, -sa * sc + ca * sb * cc, ca * sc + sa * sb * cc), du1a = mk3(0.0, -sa * cc - ca * sb * sc, ca * cc - sa * sb * sc);
    const d3 du0b = mk3(-sb * cc, sa * cb * cc, -ca * cb * cc), du1b = mk3(sb * sc, -sa * cb * sc, ca * cb * sc);
    { const d3 du = mvec(A, sp * du1a); sink.partial(6, kNeedN ? mvec(A, cp * du0a) - cross(av, du) : zero, du); }
    { const d3 du = mvec(A, sp * du1b); sink.partial(7, kNeedN ? mvec(A, cp * du0b) - cross(av, du) : zero, du); }
    { const d3 du = mvec(A, (-sp) * u0); sink.partial(8, kNeedN ? mvec(A, cp * u1) - cross(av, du) : zero, du); }
    { const d3 du = mvec(A, cp * u1); sink.partial(9, kNeedN ? mvec(A, (-sp) * u0) - cross(av, du) : zero, du); }
  }
}

// LineProjectionFactor: residual r[2] -> out_r, Jacobian entries -> out_jp (2 x PW pose block, row stride PW; the raw qw
// derivative goes to column 6 when kQw) and out_jl (2 x 4).  When `correct`, the Cauchy corrector is folded in.
template <bool kJac, bool kQw>
struct LineSink {
  double spx, spy, epx, epy, lf, loss_a;
  bool correct;
  bool want_cost = true;   // false: rho' only, half_rho stays 0 (linearisation after iteration 0, cost-free sweeps)
  int PW;
  double *out_r, *out_jp, *out_jl;
  double half_rho;
  // state between base() and partial()
  double nx, ny, irho, irho3, ds, de, sq;
  __device__ __forceinline__ void base(d3 n, d3) {
    const double rho2 = n.x * n.x + n.y * n.y;
    irho = rsqrt(rho2);
    irho3 = irho * irho * irho;
    nx = n.x; ny = n.y;
    ds = spx * n.x + spy * n.y + n.z;
    de = epx * n.x + epy * n.y + n.z;
    const double r0 = lf * ds * irho, r1 = lf * de * irho;
    const double s = r0 * r0 + r1 * r1;
    sq = 1.0;
    if (correct && loss_a > 0.0) {
      if (want_cost) { double rho0, rho1; cauchy(loss_a, s, rho0, rho1); half_rho = 0.5 * rho0; sq = sqrt(rho1); }
      else { half_rho = 0.0; sq = sqrt(cauchy_rho1(loss_a, s)); }
    } else half_rho = 0.5 * s;
    out_r[0] = sq * r0; out_r[1] = sq * r1;
  }
  __device__ __forceinline__ void partial(int k, d3 dn, d3) {
    const double drho = (nx * dn.x + ny * dn.y) * irho3;
    const double f = lf * sq;
    const double j0 = f * ((spx * dn.x + spy * dn.y + dn.z) * irho - ds * drho);
    const double j1 = f * ((epx * dn.x + epy * dn.y + dn.z) * irho - de * drho);
    if (k < 6) { out_jp[k] = j0; out_jp[PW + k] = j1; }
    else if (k < 10) { out_jl[k - 6] = j0; out_jl[4 + k - 6] = j1; }
    else { out_jp[6] = j0; out_jp[PW + 6] = j1; }
  }
};

template <bool kJac, bool kQw>
struct VpSink {
  d3 vp;
  double vf, loss_a;
  bool correct;
  bool want_cost = true;
  double *out_r, *out_jp, *out_jl;
  double half_rho;
  d3 dvec;
  double g, i1, i3;
  __device__ __forceinline__ void base(d3, d3 d) {
    dvec = d;
    const double un2 = dot(d, d), vn2 = dot(vp, vp);
    const double iuv = rsqrt(un2 * vn2);
    const double uv = dot(d, vp);
    const double c = uv * iuv;
    const double ac = fabs(c);
    const double r0 = vf * acos(ac);
    const double s = r0 * r0;
    double sq = 1.0;
    if (correct && loss_a > 0.0) {
      if (want_cost) { double rho0, rho1; cauchy(loss_a, s, rho0, rho1); half_rho = 0.5 * rho0; sq = sqrt(rho1); }
      else { half_rho = 0.0; sq = sqrt(cauchy_rho1(loss_a, s)); }
    } else half_rho = 0.5 * s;
    out_r[0] = sq * r0;
    if (kJac) {
      g = sq * vf * (c < 0.0 ? 1.0 : -1.0) * rsqrt(1.0 - ac * ac);
      i1 = iuv; i3 = uv * iuv / un2;
    }
  }
  __device__ __forceinline__ void partial(int k, d3, d3 du) {
    const double j = k < 3 ? 0.0 : g * (dot(vp, du) * i1 - dot(dvec, du) * i3);
    if (k < 6) out_jp[k] = j;
    else if (k < 10) out_jl[k - 6] = j;
    else out_jp[6] = j;
  }
};

}  // namespace uvs

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
  if (correct && loss_a > 0.0) {
    double rho0, rho1;
    cauchy(loss_a, s, rho0, rho1);
    *half_rho = 0.5 * rho0;
    sq = sqrt(rho1);
  } else {
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
#pragma unroll
    for (int k = 0; k < 6; k++) { Jex[k] = 0.0; Jex[ld + k] = 0.0; }
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
struct LineCam {
  d3 n, d;        // n_c, d_c
  d3 dn[11];      // d n_c / d theta_k      (k = 10: raw qw, only when kQw)
  d3 dd[11];      // d d_c / d theta_k   (zero for k < 3)
};

template <bool kJac, bool kNeedN, bool kQw>
__device__ __forceinline__ void line_to_camera(const double *__restrict__ pose, const double *__restrict__ line,
                                               const double *__restrict__ ric_rm, const double *__restrict__ tic3,
                                               LineCam &o) {
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
  o.d = u;
  if (kNeedN) o.n = mvec(A, nw) - cross(av, u);
  if (!kJac) return;
  // translation
#pragma unroll
  for (int k = 0; k < 3; k++) {
    o.dd[k] = mk3(0, 0, 0);
    if (kNeedN) o.dn[k] = -cross(mcol(A, k), u);
  }
  // raw quaternion coordinates: G_m = dR/dq_m
  {
    const double x2 = 2 * q.x, y2 = 2 * q.y, z2 = 2 * q.z, w2 = 2 * q.w;
    m33 G[4];
    G[0].a[0] = 0;        G[0].a[1] = y2;       G[0].a[2] = z2;
    G[0].a[3] = y2;       G[0].a[4] = -2 * x2;  G[0].a[5] = -w2;
    G[0].a[6] = z2;       G[0].a[7] = w2;       G[0].a[8] = -2 * x2;
    G[1].a[0] = -2 * y2;  G[1].a[1] = x2;       G[1].a[2] = w2;
    G[1].a[3] = x2;       G[1].a[4] = 0;        G[1].a[5] = z2;
    G[1].a[6] = -w2;      G[1].a[7] = z2;       G[1].a[8] = -2 * y2;
    G[2].a[0] = -2 * z2;  G[2].a[1] = -w2;      G[2].a[2] = x2;
    G[2].a[3] = w2;       G[2].a[4] = -2 * z2;  G[2].a[5] = y2;
    G[2].a[6] = x2;       G[2].a[7] = y2;       G[2].a[8] = 0;
    // dR/dqw = 2 [u]x
    G[3].a[0] = 0;        G[3].a[1] = -z2;      G[3].a[2] = y2;
    G[3].a[3] = z2;       G[3].a[4] = 0;        G[3].a[5] = -x2;
    G[3].a[6] = -y2;      G[3].a[7] = x2;       G[3].a[8] = 0;
#pragma unroll
    for (int m = 0; m < (kQw ? 4 : 3); m++) {
      const int slot = m < 3 ? 3 + m : 10;
      // A' v = ric^T (G^T v)
      const d3 du = mtvec(ric, mtvec(G[m], dw));
      o.dd[slot] = du;
      if (kNeedN) {
        const d3 dm = mtvec(ric, mtvec(G[m], nw));
        const d3 da = mtvec(ric, mtvec(G[m], twc)) + mvec(A, mvec(G[m], tic));
        o.dn[slot] = dm - cross(da, u) - cross(av, du);
      }
    }
  }
  // line parameters
  {
    d3 du0[3], du1[3];
    du0[0] = mk3(0.0, -sa * sc + ca * sb * cc, ca * sc + sa * sb * cc);
    du0[1] = mk3(-sb * cc, sa * cb * cc, -ca * cb * cc);
    du0[2] = u1;
    du1[0] = mk3(0.0, -sa * cc - ca * sb * sc, ca * cc - sa * sb * sc);
    du1[1] = mk3(sb * sc, -sa * cb * sc, ca * cb * sc);
    du1[2] = -u0;
#pragma unroll
    for (int m = 0; m < 3; m++) {
      const d3 du = mvec(A, sp * du1[m]);
      o.dd[6 + m] = du;
      if (kNeedN) o.dn[6 + m] = mvec(A, cp * du0[m]) - cross(av, du);
    }
    const d3 du = mvec(A, cp * u1);
    o.dd[9] = du;
    if (kNeedN) o.dn[9] = mvec(A, (-sp) * u0) - cross(av, du);
  }
}

// LineProjectionFactor residual + Jacobian [2 x NP]: cols 0-5 pose tangent, 6-9 line (10: raw qw).
template <bool kJac, bool kQw>
__device__ __forceinline__ void line_eval(const double *__restrict__ pose, const double *__restrict__ line,
                                          const double *__restrict__ ric, const double *__restrict__ tic, double spx,
                                          double spy, double epx, double epy, double line_factor, double r[2],
                                          double *J /*[2][NP]*/) {
  constexpr int NP = kQw ? 11 : 10;
  LineCam lc;
  line_to_camera<kJac, true, kQw>(pose, line, ric, tic, lc);
  const double rho2 = lc.n.x * lc.n.x + lc.n.y * lc.n.y;
  const double rho = sqrt(rho2);
  const double ds = spx * lc.n.x + spy * lc.n.y + lc.n.z;
  const double de = epx * lc.n.x + epy * lc.n.y + lc.n.z;
  r[0] = line_factor * ds / rho;
  r[1] = line_factor * de / rho;
  if (!kJac) return;
  const double irho = 1.0 / rho, irho3 = irho / rho2;
#pragma unroll
  for (int k = 0; k < NP; k++) {
    const d3 dn = lc.dn[k];
    const double drho = (lc.n.x * dn.x + lc.n.y * dn.y) * irho3;
    J[k] = line_factor * ((spx * dn.x + spy * dn.y + dn.z) * irho - ds * drho);
    J[NP + k] = line_factor * ((epx * dn.x + epy * dn.y + dn.z) * irho - de * drho);
  }
}

// VPProjectionFactor residual + Jacobian [1 x NP]
template <bool kJac, bool kQw>
__device__ __forceinline__ void vp_eval(const double *__restrict__ pose, const double *__restrict__ line,
                                        const double *__restrict__ ric, const double *__restrict__ tic, d3 vp,
                                        double vp_factor, double r[1], double *J /*[NP]*/) {
  constexpr int NP = kQw ? 11 : 10;
  LineCam lc;
  line_to_camera<kJac, false, kQw>(pose, line, ric, tic, lc);
  const double un = sqrt(dot(lc.d, lc.d)), vn = sqrt(dot(vp, vp));
  const double uv = dot(lc.d, vp);
  const double c = uv / (un * vn);
  const double ac = fabs(c);
  r[0] = vp_factor * acos(ac);
  if (!kJac) return;
  const double sgn = c < 0.0 ? -1.0 : 1.0;
  const double g = vp_factor * sgn * (-1.0 / sqrt(1.0 - ac * ac));
  const double i1 = 1.0 / (un * vn), i3 = uv / (un * un * un * vn);
#pragma unroll
  for (int k = 0; k < NP; k++) {
    if (k < 3) { J[k] = 0.0; continue; }
    const d3 du = lc.dd[k];
    J[k] = g * (dot(vp, du) * i1 - dot(lc.d, du) * i3);
  }
}

}  // namespace uvs

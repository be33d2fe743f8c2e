// uvs_tri.cu — point and line triangulation on the device: the step that seeds the inverse depths and the line
// orthonormal parameters before Estimator::optimization().
//
// Reference functions replaced (per feature, on the CPU):
//   FeatureManager::triangulate       vins_estimator/src/feature_manager.cpp:427-481   (DLT, JacobiSVD of a 2n x 4 system)
//   FeatureManager::triangulateLine   vins_estimator/src/feature_manager.cpp:504-589   (plane intersection, calcPluckerLine
//                                     :827-902, world Pluecker line -> orthonormal parameters via eulerAngles(0,1,2) + atan2)
// One thread per feature; features are independent, inputs are read once (coalesced over features where the layout
// allows), so both kernels are a single launch over the whole list.
#include <cstring>

#include "uvs_device.cuh"
#include "uvs_handle.h"
#include "uvs_math.cuh"

namespace uvs {

struct TriPointArgs {
  int n_frames, n_tracks;
  const double *Rs, *Ps;        // [n_frames][9] row-major, [n_frames][3]  (Estimator::Rs / Ps)
  double ric[9], tic[3];        // ric[0] row-major, tic[0]
  const int *start_frame, *obs_off;
  const double *pts;            // [n_obs][3] FeaturePerFrame::point (un-normalised)
  double init_depth;
  double *depth;                // [n_tracks]
};

__device__ __forceinline__ m33 ldm(const double *p) { m33 m; for (int k = 0; k < 9; k++) m.a[k] = p[k]; return m; }

// Smallest right singular vector of the rows streamed through `add_row`: Givens QR keeps the 4x4 triangular factor
// (no normal equations, so the conditioning is that of the 2n x 4 system itself), then a one-sided Jacobi SVD of it.
struct Tri4 {
  double r[4][4];
  __device__ void init() { for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r[i][j] = 0.0; }
  __device__ void add_row(double a0, double a1, double a2, double a3) {
    double a[4] = {a0, a1, a2, a3};
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (a[i] == 0.0) continue;
      const double g = hypot(r[i][i], a[i]);
      const double c = r[i][i] / g, s = a[i] / g;
      r[i][i] = g;
#pragma unroll
      for (int j = i + 1; j < 4; j++) {
        const double t = c * r[i][j] + s * a[j];
        a[j] = c * a[j] - s * r[i][j];
        r[i][j] = t;
      }
    }
  }
  // V column of the smallest singular value -> v[4]
  __device__ void smallest(double v[4]) {
    double B[4][4], V[4][4];
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { B[i][j] = j >= i ? r[i][j] : 0.0; V[i][j] = i == j ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 30; sweep++) {
      double off = 0.0;
#pragma unroll
      for (int p = 0; p < 3; p++)
#pragma unroll
        for (int q = p + 1; q < 4; q++) {
          double al = 0.0, be = 0.0, ga = 0.0;
#pragma unroll
          for (int k = 0; k < 4; k++) { al += B[k][p] * B[k][p]; be += B[k][q] * B[k][q]; ga += B[k][p] * B[k][q]; }
          if (fabs(ga) <= 1e-300 || fabs(ga) <= 1e-17 * sqrt(al * be)) continue;
          off = fmax(off, fabs(ga) / sqrt(al * be));
          const double zeta = (be - al) / (2.0 * ga);
          const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const double bp = B[k][p], bq = B[k][q];
            B[k][p] = c * bp - s * bq; B[k][q] = s * bp + c * bq;
            const double vp = V[k][p], vq = V[k][q];
            V[k][p] = c * vp - s * vq; V[k][q] = s * vp + c * vq;
          }
        }
      if (off < 1e-15) break;
    }
    int best = 0;
    double bn = 1e300;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      double n2 = 0.0;
#pragma unroll
      for (int k = 0; k < 4; k++) n2 += B[k][j] * B[k][j];
      if (n2 < bn) { bn = n2; best = j; }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = best == 0 ? V[k][0] : (best == 1 ? V[k][1] : (best == 2 ? V[k][2] : V[k][3]));
  }
};

__global__ void __launch_bounds__(128) k_triangulate_points(TriPointArgs A) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= A.n_tracks) return;
  const int i = A.start_frame[t], o0 = A.obs_off[t], n = A.obs_off[t + 1] - o0;
  if (i < 0 || n < 1 || i + n > A.n_frames) { A.depth[t] = A.init_depth; return; }
  const m33 ric = ldm(A.ric);
  const d3 tic = mk3(A.tic[0], A.tic[1], A.tic[2]);
  const m33 Ri = ldm(A.Rs + 9 * (size_t)i);
  const d3 t0 = mk3(A.Ps[3 * i], A.Ps[3 * i + 1], A.Ps[3 * i + 2]) + mvec(Ri, tic);
  const m33 R0 = mmul(Ri, ric);
  Tri4 T;
  T.init();
  for (int k = 0; k < n; k++) {
    const int j = i + k;
    const m33 Rj = ldm(A.Rs + 9 * (size_t)j);
    const d3 t1 = mk3(A.Ps[3 * j], A.Ps[3 * j + 1], A.Ps[3 * j + 2]) + mvec(Rj, tic);
    const m33 R1 = mmul(Rj, ric);
    const d3 tt = mtvec(R0, t1 - t0);          // R0^T (t1 - t0)
    const m33 R = mtmul(R0, R1);               // R0^T R1
    // P = [R^T | -R^T t]: row k of P = (column k of R, -column k of R . t)
    const d3 c0 = mcol(R, 0), c1 = mcol(R, 1), c2 = mcol(R, 2);
    const double p0 = -dot(c0, tt), p1 = -dot(c1, tt), p2 = -dot(c2, tt);
    const double *pt = A.pts + 3 * (size_t)(o0 + k);
    const double nf = sqrt(pt[0] * pt[0] + pt[1] * pt[1] + pt[2] * pt[2]);
    const double f0 = pt[0] / nf, f1 = pt[1] / nf, f2 = pt[2] / nf;
    T.add_row(f0 * c2.x - f2 * c0.x, f0 * c2.y - f2 * c0.y, f0 * c2.z - f2 * c0.z, f0 * p2 - f2 * p0);
    T.add_row(f1 * c2.x - f2 * c1.x, f1 * c2.y - f2 * c1.y, f1 * c2.z - f2 * c1.z, f1 * p2 - f2 * p1);
  }
  double v[4];
  T.smallest(v);
  double depth = v[2] / v[3];
  if (!(depth >= 0.1)) depth = A.init_depth;   // also catches NaN (the reference keeps NaN; a NaN depth is unusable)
  A.depth[t] = depth;
}

struct TriLineArgs {
  int n_frames, n_lines;
  const double *Rs, *Ps;
  double ric[9], tic[3];
  const int *frame_first, *frame_last;
  const double *sp_first, *ep_first, *sp_last, *ep_last;   // [n_lines][3] start / end point in the first / last frame
  double *ortho;                                            // [n_lines][4]
};

__global__ void __launch_bounds__(128) k_triangulate_lines(TriLineArgs A) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= A.n_lines) return;
  const int i = A.frame_first[t], j = A.frame_last[t];
  double *out = A.ortho + 4 * (size_t)t;
  if (i < 0 || j < 0 || i >= A.n_frames || j >= A.n_frames) { out[0] = out[1] = out[2] = out[3] = 0.0; return; }
  const m33 ric = ldm(A.ric);
  const d3 tic = mk3(A.tic[0], A.tic[1], A.tic[2]);
  const m33 Ri = ldm(A.Rs + 9 * (size_t)i), Rj = ldm(A.Rs + 9 * (size_t)j);
  const m33 Rl = mmul(Ri, ric), Rr = mmul(Rj, ric);
  const d3 tl = mvec(Ri, tic) + mk3(A.Ps[3 * i], A.Ps[3 * i + 1], A.Ps[3 * i + 2]);
  const d3 tr = mvec(Rj, tic) + mk3(A.Ps[3 * j], A.Ps[3 * j + 1], A.Ps[3 * j + 2]);
  const m33 Rrel = mtmul(Rl, Rr);              // q_left^-1 q_right
  const d3 trel = mtvec(Rl, tr - tl);
  auto ld3 = [](const double *p) { return mk3(p[0], p[1], p[2]); };
  const d3 lsp = ld3(A.sp_first + 3 * (size_t)t), lep = ld3(A.ep_first + 3 * (size_t)t);
  const d3 rsp = mvec(Rrel, ld3(A.sp_last + 3 * (size_t)t)), rep = mvec(Rrel, ld3(A.ep_last + 3 * (size_t)t));
  // calcPluckerLine: planes through the camera centres (origin_prev = 0, origin_curr = trel)
  const d3 pn = cross(lsp, lep), cn = cross(rsp, rep);
  const double pd = -0.0, cd = -dot(cn, trel);
  (void)pd;
  // dual Pluecker matrix  L* = pi_1 pi_2^T - pi_2 pi_1^T;  direction = (L*(2,1), L*(0,2), L*(1,0)), normal = L*(0..2, 3)
  const d3 dir = mk3(pn.z * cn.y - cn.z * pn.y, pn.x * cn.z - cn.x * pn.z, pn.y * cn.x - cn.y * pn.x);
  const d3 nor = mk3(pn.x * cd, pn.y * cd, pn.z * cd);   // pi_1(3) = 0
  // to the world frame: n_w = R n + [t]x R d,  d_w = R d
  const d3 Rd = mvec(Rl, dir);
  const d3 nw = mvec(Rl, nor) + cross(tl, Rd), dw = Rd;
  const double nn = sqrt(dot(nw, nw)), dn = sqrt(dot(dw, dw));
  const d3 cx = cross(nw, dw);
  const double cn3 = sqrt(dot(cx, cx));
  // Rotation_psi = [n/|n|  d/|d|  (n x d)/|n x d|] (columns); Eigen eulerAngles(0, 1, 2)
  const double m00 = nw.x / nn, m10 = nw.y / nn, m20 = nw.z / nn;
  const double m01 = dw.x / dn, m11 = dw.y / dn, m21 = dw.z / dn;
  const double m02 = cx.x / cn3, m12 = cx.y / cn3, m22 = cx.z / cn3;
  double e0 = atan2(m12, m22), e1;
  const double c2 = sqrt(m00 * m00 + m01 * m01);
  if (e0 > 0.0) { e0 -= 3.14159265358979323846; e1 = atan2(-m02, -c2); }
  else e1 = atan2(-m02, c2);
  const double s1 = sin(e0), c1 = cos(e0);
  const double e2 = atan2(s1 * m20 - c1 * m10, c1 * m11 - s1 * m21);
  out[0] = -e0; out[1] = -e1; out[2] = -e2;
  out[3] = atan2(dn, nn);
}

// FeatureManager::setLineOrtho, the validity test (feature_manager.cpp:333-423): the line - from the orthonormal
// parameters the FEATURE holds, i.e. before the solved ones are written back - is taken into the camera of its first
// frame (Pluecker transform T_cw), intersected with the two planes through the viewing rays of the first observation's
// end points, and is invalid (solve_flag 2) when either 3-D end point lies behind that camera.
struct LineFlagArgs {
  int n_frames, n_lines;
  const double *Rs, *Ps;
  double ric[9], tic[3];
  const int *start_frame;
  const double *ortho;             // [n_lines][4] psi_x, psi_y, psi_z, phi
  const double *sp, *ep;           // [n_lines][3] end points of the first observation (z = 1)
  int *flag;                       // [n_lines] 1 valid, 2 behind the camera
  double *ends;                    // [n_lines][6] D_s_w, D_e_w (world end points, as the reference computes them) or null
};

__global__ void __launch_bounds__(128) k_line_flags(LineFlagArgs A) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= A.n_lines) return;
  const int i = A.start_frame[t];
  if (i < 0 || i >= A.n_frames) { A.flag[t] = 2; return; }
  const double *o = A.ortho + 4 * (size_t)t;
  double sa, ca, sb, cb, sc, cc, sp, cp;
  sincos(o[0], &sa, &ca); sincos(o[1], &sb, &cb); sincos(o[2], &sc, &cc); sincos(o[3], &sp, &cp);
  // Rx(a) Ry(b) Rz(c): column 0 and column 1
  const d3 u0 = mk3(cb * cc, ca * sc + sa * sb * cc, sa * sc - ca * sb * cc);
  const d3 u1 = mk3(-cb * sc, ca * cc - sa * sb * sc, sa * cc + ca * sb * sc);
  const d3 nw = cp * u0, dw = sp * u1;
  const m33 ric = ldm(A.ric);
  const d3 tic = mk3(A.tic[0], A.tic[1], A.tic[2]);
  const m33 Ri = ldm(A.Rs + 9 * (size_t)i);
  const m33 Rwc = mmul(Ri, ric);
  const d3 twc = mvec(Ri, tic) + mk3(A.Ps[3 * i], A.Ps[3 * i + 1], A.Ps[3 * i + 2]);
  // l_c = T_cw l_w:  n_c = R^T n_w + [-R^T t]x R^T d_w,  d_c = R^T d_w
  const d3 dc = mtvec(Rwc, dw);
  const d3 mt = mtvec(Rwc, twc);
  const d3 nc = mtvec(Rwc, nw) - cross(mt, dc);
  const double *ps = A.sp + 3 * (size_t)t, *pe = A.ep + 3 * (size_t)t;
  const d3 s2 = mk3(ps[0], ps[1], ps[2]), e2 = mk3(pe[0], pe[1], pe[2]);
  const double slope = -(e2.x - s2.x) / (e2.y - s2.y);     // scale = 1 (a horizontal segment divides by zero, as in the reference)
  const d3 s2p = mk3(s2.x + 1.0, slope + s2.y, 1.0), e2p = mk3(e2.x + 1.0, slope + e2.y, 1.0);
  const d3 pis = cross(s2, s2p), pie = cross(e2, e2p);
  // D = L_c [pi; 0] with L_c = [[n_c]x d_c; -d_c^T 0]:  D.xyz = n_c x pi, D.w = -d_c . pi
  const d3 ds = cross(nc, pis), de = cross(nc, pie);
  const double ws = -dot(dc, pis), we = -dot(dc, pie);
  const d3 Ds = mk3(ds.x / ws, ds.y / ws, ds.z / ws), De = mk3(de.x / we, de.y / we, de.z / we);
  A.flag[t] = (Ds.z < 0.0 || De.z < 0.0) ? 2 : 1;
  if (A.ends) {
    const d3 a = mvec(Rwc, Ds) + twc, b = mvec(Rwc, De) + twc;
    double *q = A.ends + 6 * (size_t)t;
    q[0] = a.x; q[1] = a.y; q[2] = a.z; q[3] = b.x; q[4] = b.y; q[5] = b.z;
  }
}

}  // namespace uvs

using namespace uvs;

namespace {
size_t al256(size_t v) { return (v + 255) / 256 * 256; }
}

extern "C" int uvs_triangulate_points(UvsHandle *h, int32_t n_frames, const double *Rs, const double *Ps, const double *ric,
                                      const double *tic, int32_t n_tracks, const int32_t *start_frame, const int32_t *obs_off,
                                      const double *obs_pts, double init_depth, double *depth_out) {
  if (!h || n_frames <= 0 || n_tracks <= 0 || !Rs || !Ps || !ric || !tic || !start_frame || !obs_off || !obs_pts || !depth_out)
    return handle_fail(h, UVS_ERR_INVALID_ARG, "uvs_triangulate_points: bad arguments");
  if (cudaSetDevice(h->device) != cudaSuccess) return handle_fail(h, UVS_ERR_CUDA, "cudaSetDevice");
  const size_t Dd = sizeof(double), nobs = (size_t)obs_off[n_tracks];
  size_t o = 0;
  const size_t o_R = o; o += al256((size_t)n_frames * 9 * Dd);
  const size_t o_P = o; o += al256((size_t)n_frames * 3 * Dd);
  const size_t o_sf = o; o += al256((size_t)n_tracks * sizeof(int));
  const size_t o_off = o; o += al256((size_t)(n_tracks + 1) * sizeof(int));
  const size_t o_pts = o; o += al256(nobs * 3 * Dd);
  const size_t in_end = o;
  const size_t o_out = o; o += al256((size_t)n_tracks * Dd);
  int rc = handle_ensure_scratch(h, o); if (rc) return rc;
  rc = handle_ensure_hscratch(h, o); if (rc) return rc;
  char *hs = h->hscratch.base, *ds = h->scratch.base;
  std::memcpy(hs + o_R, Rs, (size_t)n_frames * 9 * Dd); std::memcpy(hs + o_P, Ps, (size_t)n_frames * 3 * Dd);
  std::memcpy(hs + o_sf, start_frame, (size_t)n_tracks * sizeof(int)); std::memcpy(hs + o_off, obs_off, (size_t)(n_tracks + 1) * sizeof(int));
  std::memcpy(hs + o_pts, obs_pts, nobs * 3 * Dd);
  cudaStream_t st = h->stream;
  if (cudaMemcpyAsync(ds, hs, in_end, cudaMemcpyHostToDevice, st) != cudaSuccess) return handle_fail(h, UVS_ERR_CUDA, "uvs_triangulate_points: H2D");
  TriPointArgs A;
  A.n_frames = n_frames; A.n_tracks = n_tracks;
  A.Rs = (const double *)(ds + o_R); A.Ps = (const double *)(ds + o_P);
  for (int k = 0; k < 9; k++) A.ric[k] = ric[k];
  for (int k = 0; k < 3; k++) A.tic[k] = tic[k];
  A.start_frame = (const int *)(ds + o_sf); A.obs_off = (const int *)(ds + o_off); A.pts = (const double *)(ds + o_pts);
  A.init_depth = init_depth; A.depth = (double *)(ds + o_out);
  k_triangulate_points<<<(n_tracks + 127) / 128, 128, 0, st>>>(A);
  h->launches++;
  if (cudaGetLastError() != cudaSuccess) return handle_fail(h, UVS_ERR_CUDA, "uvs_triangulate_points: launch");
  if (cudaMemcpyAsync(hs + o_out, ds + o_out, (size_t)n_tracks * Dd, cudaMemcpyDeviceToHost, st) != cudaSuccess)
    return handle_fail(h, UVS_ERR_CUDA, "uvs_triangulate_points: D2H");
  if (cudaStreamSynchronize(st) != cudaSuccess) return handle_fail(h, UVS_ERR_CUDA, "uvs_triangulate_points: sync");
  std::memcpy(depth_out, hs + o_out, (size_t)n_tracks * Dd);
  return UVS_OK;
}

extern "C" int uvs_triangulate_lines(UvsHandle *h, int32_t n_frames, const double *Rs, const double *Ps, const double *ric,
                                     const double *tic, int32_t n_lines, const int32_t *frame_first, const int32_t *frame_last,
                                     const double *sp_first, const double *ep_first, const double *sp_last, const double *ep_last,
                                     double *ortho_out) {
  if (!h || n_frames <= 0 || n_lines <= 0 || !Rs || !Ps || !ric || !tic || !frame_first || !frame_last || !sp_first || !ep_first ||
      !sp_last || !ep_last || !ortho_out)
    return handle_fail(h, UVS_ERR_INVALID_ARG, "uvs_triangulate_lines: bad arguments");
  if (cudaSetDevice(h->device) != cudaSuccess) return handle_fail(h, UVS_ERR_CUDA, "cudaSetDevice");
  const size_t Dd = sizeof(double), L3 = (size_t)n_lines * 3 * Dd;
  size_t o = 0;
  const size_t o_R = o; o += al256((size_t)n_frames * 9 * Dd);
  const size_t o_P = o; o += al256((size_t)n_frames * 3 * Dd);
  const size_t o_ff = o; o += al256((size_t)n_lines * sizeof(int));
  const size_t o_fl = o; o += al256((size_t)n_lines * sizeof(int));
  const size_t o_a = o; o += al256(L3);
  const size_t o_b = o; o += al256(L3);
  const size_t o_c = o; o += al256(L3);
  const size_t o_d = o; o += al256(L3);
  const size_t in_end = o;
  const size_t o_out = o; o += al256((size_t)n_lines * 4 * Dd);
  int rc = handle_ensure_scratch(h, o); if (rc) return rc;
  rc = handle_ensure_hscratch(h, o); if (rc) return rc;
  char *hs = h->hscratch.base, *ds = h->scratch.base;
  std::memcpy(hs + o_R, Rs, (size_t)n_frames * 9 * Dd); std::memcpy(hs + o_P, Ps, (size_t)n_frames * 3 * Dd);
  std::memcpy(hs + o_ff, frame_first, (size_t)n_lines * sizeof(int)); std::memcpy(hs + o_fl, frame_last, (size_t)n_lines * sizeof(int));
  std::memcpy(hs + o_a, sp_first, L3); std::memcpy(hs + o_b, ep_first, L3); std::memcpy(hs + o_c, sp_last, L3); std::memcpy(hs + o_d, ep_last, L3);
  cudaStream_t st = h->stream;
  if (cudaMemcpyAsync(ds, hs, in_end, cudaMemcpyHostToDevice, st) != cudaSuccess) return handle_fail(h, UVS_ERR_CUDA, "uvs_triangulate_lines: H2D");
  TriLineArgs A;
  A.n_frames = n_frames; A.n_lines = n_lines;
  A.Rs = (const double *)(ds + o_R); A.Ps = (const double *)(ds + o_P);
  for (int k = 0; k < 9; k++) A.ric[k] = ric[k];
  for (int k = 0; k < 3; k++) A.tic[k] = tic[k];
  A.frame_first = (const int *)(ds + o_ff); A.frame_last = (const int *)(ds + o_fl);
  A.sp_first = (const double *)(ds + o_a); A.ep_first = (const double *)(ds + o_b);
  A.sp_last = (const double *)(ds + o_c); A.ep_last = (const double *)(ds + o_d);
  A.ortho = (double *)(ds + o_out);
  k_triangulate_lines<<<(n_lines + 127) / 128, 128, 0, st>>>(A);
  h->launches++;
  if (cudaGetLastError() != cudaSuccess) return handle_fail(h, UVS_ERR_CUDA, "uvs_triangulate_lines: launch");
  if (cudaMemcpyAsync(hs + o_out, ds + o_out, (size_t)n_lines * 4 * Dd, cudaMemcpyDeviceToHost, st) != cudaSuccess)
    return handle_fail(h, UVS_ERR_CUDA, "uvs_triangulate_lines: D2H");
  if (cudaStreamSynchronize(st) != cudaSuccess) return handle_fail(h, UVS_ERR_CUDA, "uvs_triangulate_lines: sync");
  std::memcpy(ortho_out, hs + o_out, (size_t)n_lines * 4 * Dd);
  return UVS_OK;
}

extern "C" int uvs_validate_lines(UvsHandle *h, int32_t n_frames, const double *Rs, const double *Ps, const double *ric, const double *tic,
                                  int32_t n_lines, const int32_t *start_frame, const double *ortho, const double *sp_first,
                                  const double *ep_first, int32_t *solve_flag, double *end_points) {
  if (!h || n_frames <= 0 || n_lines <= 0 || !Rs || !Ps || !ric || !tic || !start_frame || !ortho || !sp_first || !ep_first || !solve_flag)
    return handle_fail(h, UVS_ERR_INVALID_ARG, "uvs_validate_lines: bad arguments");
  if (cudaSetDevice(h->device) != cudaSuccess) return handle_fail(h, UVS_ERR_CUDA, "cudaSetDevice");
  const size_t Dd = sizeof(double), L3 = (size_t)n_lines * 3 * Dd;
  size_t o = 0;
  const size_t o_R = o; o += al256((size_t)n_frames * 9 * Dd);
  const size_t o_P = o; o += al256((size_t)n_frames * 3 * Dd);
  const size_t o_sf = o; o += al256((size_t)n_lines * sizeof(int));
  const size_t o_or = o; o += al256((size_t)n_lines * 4 * Dd);
  const size_t o_a = o; o += al256(L3);
  const size_t o_b = o; o += al256(L3);
  const size_t in_end = o;
  const size_t o_fl = o; o += al256((size_t)n_lines * sizeof(int));
  const size_t o_en = o; o += al256((size_t)n_lines * 6 * Dd);
  int rc = handle_ensure_scratch(h, o); if (rc) return rc;
  rc = handle_ensure_hscratch(h, o); if (rc) return rc;
  char *hs = h->hscratch.base, *ds = h->scratch.base;
  std::memcpy(hs + o_R, Rs, (size_t)n_frames * 9 * Dd); std::memcpy(hs + o_P, Ps, (size_t)n_frames * 3 * Dd);
  std::memcpy(hs + o_sf, start_frame, (size_t)n_lines * sizeof(int)); std::memcpy(hs + o_or, ortho, (size_t)n_lines * 4 * Dd);
  std::memcpy(hs + o_a, sp_first, L3); std::memcpy(hs + o_b, ep_first, L3);
  cudaStream_t st = h->stream;
  if (cudaMemcpyAsync(ds, hs, in_end, cudaMemcpyHostToDevice, st) != cudaSuccess) return handle_fail(h, UVS_ERR_CUDA, "uvs_validate_lines: H2D");
  LineFlagArgs A;
  A.n_frames = n_frames; A.n_lines = n_lines;
  A.Rs = (const double *)(ds + o_R); A.Ps = (const double *)(ds + o_P);
  for (int k = 0; k < 9; k++) A.ric[k] = ric[k];
  for (int k = 0; k < 3; k++) A.tic[k] = tic[k];
  A.start_frame = (const int *)(ds + o_sf); A.ortho = (const double *)(ds + o_or);
  A.sp = (const double *)(ds + o_a); A.ep = (const double *)(ds + o_b);
  A.flag = (int *)(ds + o_fl); A.ends = end_points ? (double *)(ds + o_en) : nullptr;
  k_line_flags<<<(n_lines + 127) / 128, 128, 0, st>>>(A);
  h->launches++;
  if (cudaGetLastError() != cudaSuccess) return handle_fail(h, UVS_ERR_CUDA, "uvs_validate_lines: launch");
  if (cudaMemcpyAsync(hs + o_fl, ds + o_fl, o - o_fl, cudaMemcpyDeviceToHost, st) != cudaSuccess)
    return handle_fail(h, UVS_ERR_CUDA, "uvs_validate_lines: D2H");
  if (cudaStreamSynchronize(st) != cudaSuccess) return handle_fail(h, UVS_ERR_CUDA, "uvs_validate_lines: sync");
  std::memcpy(solve_flag, hs + o_fl, (size_t)n_lines * sizeof(int));
  if (end_points) std::memcpy(end_points, hs + o_en, (size_t)n_lines * 6 * Dd);
  return UVS_OK;
}

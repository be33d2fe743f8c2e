// uvs_math.cuh — fixed-size FP64 algebra for the sm_100a factor kernels.
//
// Product code: written independently of the CPU checker (test infrastructure).  Formulas follow
// the reference's Eigen usage: R(q) = I + 2w[u]x + 2[u]x^2 is NOT
// normalised, q^-1 = conj(q)/|q|^2, deltaQ(theta) = (theta/2, 1) (utility/utility.h:11-24).
#pragma once
#include <cuda_runtime.h>

namespace uvs {

struct d3 { double x, y, z; };
struct q4 { double x, y, z, w; };      // storage order of the parameter blocks
struct m33 { double a[9]; };           // row-major

__device__ __forceinline__ d3 mk3(double x, double y, double z) { d3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ d3 operator+(d3 a, d3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ d3 operator-(d3 a, d3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ d3 operator-(d3 a) { return mk3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ d3 operator*(double s, d3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ double dot(d3 a, d3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ d3 cross(d3 a, d3 b) {
  return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ double comp(const d3 &v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }

__device__ __forceinline__ q4 mkq(double x, double y, double z, double w) { q4 q; q.x = x; q.y = y; q.z = z; q.w = w; return q; }
__device__ __forceinline__ q4 qmul(q4 a, q4 b) {
  return mkq(a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
             a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
             a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x,
             a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z);
}
__device__ __forceinline__ q4 qinv(q4 q) {
  const double in2 = 1.0 / (q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  return mkq(-q.x * in2, -q.y * in2, -q.z * in2, q.w * in2);
}
__device__ __forceinline__ d3 qvec(q4 q) { return mk3(q.x, q.y, q.z); }
// v + w (2 u x v) + u x (2 u x v)
__device__ __forceinline__ d3 qrot(q4 q, d3 v) {
  d3 u = qvec(q);
  d3 uv = cross(u, v);
  uv = uv + uv;
  return v + q.w * uv + cross(u, uv);
}
__device__ __forceinline__ m33 qmat(q4 q) {
  m33 r;
  const double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  r.a[0] = 1.0 - (tyy + tzz); r.a[1] = txy - twz;         r.a[2] = txz + twy;
  r.a[3] = txy + twz;         r.a[4] = 1.0 - (txx + tzz); r.a[5] = tyz - twx;
  r.a[6] = txz - twy;         r.a[7] = tyz + twx;         r.a[8] = 1.0 - (txx + tyy);
  return r;
}
__device__ __forceinline__ m33 mmul(const m33 &A, const m33 &B) {
  m33 r;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) r.a[3 * i + j] = A.a[3 * i] * B.a[j] + A.a[3 * i + 1] * B.a[3 + j] + A.a[3 * i + 2] * B.a[6 + j];
  return r;
}
// A^T B
__device__ __forceinline__ m33 mtmul(const m33 &A, const m33 &B) {
  m33 r;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) r.a[3 * i + j] = A.a[i] * B.a[j] + A.a[3 + i] * B.a[3 + j] + A.a[6 + i] * B.a[6 + j];
  return r;
}
__device__ __forceinline__ m33 mtrans(const m33 &A) {
  m33 r;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) r.a[3 * i + j] = A.a[3 * j + i];
  return r;
}
__device__ __forceinline__ d3 mvec(const m33 &A, d3 v) {
  return mk3(A.a[0] * v.x + A.a[1] * v.y + A.a[2] * v.z, A.a[3] * v.x + A.a[4] * v.y + A.a[5] * v.z,
             A.a[6] * v.x + A.a[7] * v.y + A.a[8] * v.z);
}
// A^T v
__device__ __forceinline__ d3 mtvec(const m33 &A, d3 v) {
  return mk3(A.a[0] * v.x + A.a[3] * v.y + A.a[6] * v.z, A.a[1] * v.x + A.a[4] * v.y + A.a[7] * v.z,
             A.a[2] * v.x + A.a[5] * v.y + A.a[8] * v.z);
}
__device__ __forceinline__ m33 skew(d3 v) {
  m33 r;
  r.a[0] = 0.0;  r.a[1] = -v.z; r.a[2] = v.y;
  r.a[3] = v.z;  r.a[4] = 0.0;  r.a[5] = -v.x;
  r.a[6] = -v.y; r.a[7] = v.x;  r.a[8] = 0.0;
  return r;
}
__device__ __forceinline__ d3 mcol(const m33 &A, int j) { return mk3(A.a[j], A.a[3 + j], A.a[6 + j]); }

// pose block [p(3), qx,qy,qz,qw]
__device__ __forceinline__ void load_pose(const double *__restrict__ b, d3 &p, q4 &q) {
  p = mk3(__ldg(b), __ldg(b + 1), __ldg(b + 2));
  q = mkq(__ldg(b + 3), __ldg(b + 4), __ldg(b + 5), __ldg(b + 6));
}

// ceres::CauchyLoss(a) evaluated at s: rho[0..2]
__device__ __forceinline__ void cauchy(double a, double s, double &rho0, double &rho1) {
  const double b = a * a, c = 1.0 / b;
  const double sum = 1.0 + s * c;
  const double inv = 1.0 / sum;
  rho0 = b * log(sum);
  rho1 = fmax(inv, 2.2250738585072014e-308);
}
// rho' alone: what the loss correction of a Jacobian needs.  The cost rho (one FP64 log per factor, ~45 instructions) is only
// used from a linearisation at iteration 0 (afterwards the cost of an iterate is the candidate cost that accepted it)
__device__ __forceinline__ double cauchy_rho1(double a, double s) {
  const double c = 1.0 / (a * a);
  return fmax(1.0 / (1.0 + s * c), 2.2250738585072014e-308);
}

}  // namespace uvs

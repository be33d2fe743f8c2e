// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into, imported by or called from the product
// path (uv-slam_b200/); only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use it.
//
// smallmat.h: the handful of Eigen formulas the UV-SLAM hot path relies on, restated for a
// generic scalar T (double or Jet).  PARITY UNPINNED at the Eigen/Ceres boundary: neither
// library is vendored or version-pinned by the reference (vins_estimator/CMakeLists.txt:22,29).
//
// Conventions restated:
//   Quaternion q = (w; x,y,z), u = (x,y,z).
//   toRotationMatrix()/q*v are the un-normalised polynomial R(q) = I + 2w[u]x + 2[u]x^2.
//   q.inverse() = conj(q)/|q|^2.
#pragma once
#include <cmath>

namespace orc {

template <class T>
struct V3 {
  T x, y, z;
  V3() : x(T(0.0)), y(T(0.0)), z(T(0.0)) {}
  V3(const T &a, const T &b, const T &c) : x(a), y(b), z(c) {}
  T &operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
  const T &operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
template <class T> V3<T> operator+(const V3<T> &a, const V3<T> &b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <class T> V3<T> operator-(const V3<T> &a, const V3<T> &b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <class T> V3<T> operator-(const V3<T> &a) { return {-a.x, -a.y, -a.z}; }
template <class T> V3<T> operator*(const V3<T> &a, const T &s) { return {a.x * s, a.y * s, a.z * s}; }
template <class T> V3<T> operator*(const T &s, const V3<T> &a) { return {s * a.x, s * a.y, s * a.z}; }
template <class T> V3<T> operator/(const V3<T> &a, const T &s) { return {a.x / s, a.y / s, a.z / s}; }
template <class T> T dot(const V3<T> &a, const V3<T> &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class T> V3<T> cross(const V3<T> &a, const V3<T> &b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

template <class T>
struct M3 {
  T m[3][3];
  M3() { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m[i][j] = T(0.0); }
  T &operator()(int i, int j) { return m[i][j]; }
  const T &operator()(int i, int j) const { return m[i][j]; }
  static M3 Identity() { M3 r; r.m[0][0] = r.m[1][1] = r.m[2][2] = T(1.0); return r; }
  V3<T> col(int j) const { return {m[0][j], m[1][j], m[2][j]}; }
};
template <class T> M3<T> operator*(const M3<T> &a, const M3<T> &b) {
  M3<T> r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
  return r;
}
template <class T> V3<T> operator*(const M3<T> &a, const V3<T> &v) {
  return {a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z,
          a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z,
          a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z};
}
template <class T> M3<T> operator+(const M3<T> &a, const M3<T> &b) {
  M3<T> r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][j] + b.m[i][j]; return r;
}
template <class T> M3<T> operator-(const M3<T> &a, const M3<T> &b) {
  M3<T> r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][j] - b.m[i][j]; return r;
}
template <class T> M3<T> operator-(const M3<T> &a) {
  M3<T> r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = -a.m[i][j]; return r;
}
template <class T> M3<T> operator*(const M3<T> &a, const T &s) {
  M3<T> r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][j] * s; return r;
}
template <class T> M3<T> transpose(const M3<T> &a) {
  M3<T> r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = a.m[j][i]; return r;
}
// Utility::skewSymmetric, utility/utility.h:26-35
template <class T> M3<T> skew(const V3<T> &q) {
  M3<T> a;
  a.m[0][0] = T(0.0); a.m[0][1] = -q.z;   a.m[0][2] = q.y;
  a.m[1][0] = q.z;    a.m[1][1] = T(0.0); a.m[1][2] = -q.x;
  a.m[2][0] = -q.y;   a.m[2][1] = q.x;    a.m[2][2] = T(0.0);
  return a;
}

template <class T>
struct Quat {
  T w, x, y, z;
  Quat() : w(T(1.0)), x(T(0.0)), y(T(0.0)), z(T(0.0)) {}
  Quat(const T &w_, const T &x_, const T &y_, const T &z_) : w(w_), x(x_), y(y_), z(z_) {}
  V3<T> vec() const { return {x, y, z}; }
};
// Hamilton product (Eigen quat_product<Arch::None>)
template <class T> Quat<T> operator*(const Quat<T> &a, const Quat<T> &b) {
  return Quat<T>(a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z,
                 a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
                 a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
                 a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x);
}
template <class T> T squaredNorm(const Quat<T> &q) { return q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z; }
// Eigen QuaternionBase::inverse(): conjugate / squaredNorm
template <class T> Quat<T> inverse(const Quat<T> &q) {
  T n2 = squaredNorm(q);
  return Quat<T>(q.w / n2, -q.x / n2, -q.y / n2, -q.z / n2);
}
template <class T> Quat<T> normalized(const Quat<T> &q) {
  using std::sqrt;
  T n = sqrt(squaredNorm(q));
  return Quat<T>(q.w / n, q.x / n, q.y / n, q.z / n);
}
// Eigen QuaternionBase::_transformVector: v + w*(2 u x v) + u x (2 u x v)
template <class T> V3<T> rotate(const Quat<T> &q, const V3<T> &v) {
  V3<T> u = q.vec();
  V3<T> uv = cross(u, v);
  uv = uv + uv;
  return v + q.w * uv + cross(u, uv);
}
// Eigen QuaternionBase::toRotationMatrix (not normalised)
template <class T> M3<T> toRotationMatrix(const Quat<T> &q) {
  M3<T> r;
  const T tx = T(2.0) * q.x, ty = T(2.0) * q.y, tz = T(2.0) * q.z;
  const T twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const T txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const T tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  r.m[0][0] = T(1.0) - (tyy + tzz); r.m[0][1] = txy - twz;            r.m[0][2] = txz + twy;
  r.m[1][0] = txy + twz;            r.m[1][1] = T(1.0) - (txx + tzz); r.m[1][2] = tyz - twx;
  r.m[2][0] = txz - twy;            r.m[2][1] = tyz + twx;            r.m[2][2] = T(1.0) - (txx + tyy);
  return r;
}
// Utility::deltaQ, utility/utility.h:11-24 — (theta/2, 1), NOT normalised
template <class T> Quat<T> deltaQ(const V3<T> &theta) {
  return Quat<T>(T(1.0), theta.x / T(2.0), theta.y / T(2.0), theta.z / T(2.0));
}
// bottom-right 3x3 of Utility::Qleft / Qright, utility/utility.h:46-64
template <class T> M3<T> QleftBR(const Quat<T> &q) { return M3<T>::Identity() * q.w + skew(q.vec()); }
template <class T> M3<T> QrightBR(const Quat<T> &q) { return M3<T>::Identity() * q.w - skew(q.vec()); }
// full 4x4 Qleft/Qright in Eigen order rows/cols (w,x,y,z)
template <class T> void Qleft4(const Quat<T> &q, T out[4][4]) {
  out[0][0] = q.w; out[0][1] = -q.x; out[0][2] = -q.y; out[0][3] = -q.z;
  M3<T> br = QleftBR(q);
  V3<T> u = q.vec();
  for (int i = 0; i < 3; i++) { out[i + 1][0] = u[i]; for (int j = 0; j < 3; j++) out[i + 1][j + 1] = br.m[i][j]; }
}
template <class T> void Qright4(const Quat<T> &q, T out[4][4]) {
  out[0][0] = q.w; out[0][1] = -q.x; out[0][2] = -q.y; out[0][3] = -q.z;
  M3<T> br = QrightBR(q);
  V3<T> u = q.vec();
  for (int i = 0; i < 3; i++) { out[i + 1][0] = u[i]; for (int j = 0; j < 3; j++) out[i + 1][j + 1] = br.m[i][j]; }
}
// Eigen Quaternion(AngleAxis): w = cos(a/2), vec = sin(a/2)*axis
template <class T> Quat<T> fromAngleAxis(const T &angle, int axis) {
  using std::cos; using std::sin;
  T ha = T(0.5) * angle;
  T s = sin(ha), c = cos(ha);
  return Quat<T>(c, axis == 0 ? s : T(0.0), axis == 1 ? s : T(0.0), axis == 2 ? s : T(0.0));
}

}  // namespace orc

// ORACLE — TEST INFRASTRUCTURE ONLY (see smallmat.h).
//
// jet.h: forward-mode dual number mirroring the documented arithmetic rules of ceres::Jet<T,N>
// (Ceres is NOT vendored by the reference — `find_package(Ceres REQUIRED)`,
// vins_estimator/CMakeLists.txt:22, version unpinned, expected 1.13/1.14 from README.md:32-35).
// Only the operations used by line_projection_factor.h:16-60 and vp_projection_factor.h:19-66 are
// provided: + - * /, sqrt, sin, cos, acos, abs, pow(f, p).
#pragma once
#include <cmath>

namespace orc {

template <int N>
struct Jet {
  double a;
  double v[N];
  Jet() : a(0.0) { for (int i = 0; i < N; i++) v[i] = 0.0; }
  explicit Jet(double s) : a(s) { for (int i = 0; i < N; i++) v[i] = 0.0; }
  Jet(double s, int k) : a(s) { for (int i = 0; i < N; i++) v[i] = 0.0; v[k] = 1.0; }
};

template <int N> Jet<N> operator+(const Jet<N> &f, const Jet<N> &g) {
  Jet<N> r; r.a = f.a + g.a; for (int i = 0; i < N; i++) r.v[i] = f.v[i] + g.v[i]; return r;
}
template <int N> Jet<N> operator-(const Jet<N> &f, const Jet<N> &g) {
  Jet<N> r; r.a = f.a - g.a; for (int i = 0; i < N; i++) r.v[i] = f.v[i] - g.v[i]; return r;
}
template <int N> Jet<N> operator-(const Jet<N> &f) {
  Jet<N> r; r.a = -f.a; for (int i = 0; i < N; i++) r.v[i] = -f.v[i]; return r;
}
// (f g)' = f.a g' + f' g.a
template <int N> Jet<N> operator*(const Jet<N> &f, const Jet<N> &g) {
  Jet<N> r; r.a = f.a * g.a; for (int i = 0; i < N; i++) r.v[i] = f.a * g.v[i] + f.v[i] * g.a; return r;
}
// f/g: value f.a/g.a, derivative (f' - (f.a/g.a) g') / g.a
template <int N> Jet<N> operator/(const Jet<N> &f, const Jet<N> &g) {
  Jet<N> r;
  const double g_a_inverse = 1.0 / g.a;
  const double f_a_by_g_a = f.a * g_a_inverse;
  r.a = f_a_by_g_a;
  for (int i = 0; i < N; i++) r.v[i] = (f.v[i] - f_a_by_g_a * g.v[i]) * g_a_inverse;
  return r;
}
template <int N> Jet<N> sqrt(const Jet<N> &f) {
  Jet<N> r; const double t = std::sqrt(f.a); const double two_a_inverse = 1.0 / (2.0 * t);
  r.a = t; for (int i = 0; i < N; i++) r.v[i] = f.v[i] * two_a_inverse; return r;
}
template <int N> Jet<N> cos(const Jet<N> &f) {
  Jet<N> r; r.a = std::cos(f.a); const double s = -std::sin(f.a);
  for (int i = 0; i < N; i++) r.v[i] = s * f.v[i]; return r;
}
template <int N> Jet<N> sin(const Jet<N> &f) {
  Jet<N> r; r.a = std::sin(f.a); const double c = std::cos(f.a);
  for (int i = 0; i < N; i++) r.v[i] = c * f.v[i]; return r;
}
// acos'(x) = -1/sqrt(1-x^2): infinite at |x| = 1 (SURVEY.md 8a "Jet rules to mirror")
template <int N> Jet<N> acos(const Jet<N> &f) {
  Jet<N> r; r.a = std::acos(f.a); const double t = -1.0 / std::sqrt(1.0 - f.a * f.a);
  for (int i = 0; i < N; i++) r.v[i] = t * f.v[i]; return r;
}
template <int N> Jet<N> abs(const Jet<N> &f) { return f.a < 0.0 ? -f : f; }
// pow(f, p) with constant exponent: p f.a^(p-1) f'
template <int N> Jet<N> pow(const Jet<N> &f, double p) {
  Jet<N> r; r.a = std::pow(f.a, p); const double t = p * std::pow(f.a, p - 1.0);
  for (int i = 0; i < N; i++) r.v[i] = t * f.v[i]; return r;
}

}  // namespace orc

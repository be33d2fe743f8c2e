// ORACLE — TEST INFRASTRUCTURE ONLY (see smallmat.h).  PARITY UNPINNED at the Ceres/Eigen boundary.
//
// marg.h: restatement of the prior construction at the end of Estimator::optimization()
// (vins_estimator/src/estimator.cpp:1003-1228) and of MarginalizationInfo::{addResidualBlockInfo,
// preMarginalize,marginalize,getParameterBlocks} (factor/marginalization_factor.cpp:89-319).
//
// Deliberate difference: the reference orders parameter blocks by iterating unordered_maps keyed by
// host addresses (marginalization_factor.cpp:176-194), which is not deterministic.  Here the order is
// fixed: dropped = [pose, speed-bias, points, lines], kept = [pose_f, speedbias_f by frame, ex, td].
// A' and b' are order-equivariant; J0/r0 are only defined up to an orthogonal transform, so parity
// is checked on J0'J0 = A' and J0'r0 = b' (the identities the authors left commented at :295-296).
#pragma once
#include <algorithm>
#include <cmath>
#include <map>
#include <vector>

#include "problem.h"

namespace orc {

// Symmetric eigen-decomposition (Householder tridiagonalisation + implicit QL), the same class of
// algorithm as Eigen::SelfAdjointEigenSolver.  A row-major n x n (symmetric); V columns = vectors.
inline void sym_eig(int n, const std::vector<double> &A, std::vector<double> &evals, std::vector<double> &V) {
  V = A;
  evals.assign(n, 0.0);
  std::vector<double> e(n, 0.0);
  std::vector<double> &d = evals;
  auto v = [&](int i, int j) -> double & { return V[(size_t)i * n + j]; };
  // tred2
  for (int j = 0; j < n; j++) d[j] = v(n - 1, j);
  for (int i = n - 1; i > 0; i--) {
    double scale = 0.0, h = 0.0;
    for (int k = 0; k < i; k++) scale += std::fabs(d[k]);
    if (scale == 0.0) {
      e[i] = d[i - 1];
      for (int j = 0; j < i; j++) { d[j] = v(i - 1, j); v(i, j) = 0.0; v(j, i) = 0.0; }
    } else {
      for (int k = 0; k < i; k++) { d[k] /= scale; h += d[k] * d[k]; }
      double f = d[i - 1];
      double g = std::sqrt(h);
      if (f > 0) g = -g;
      e[i] = scale * g;
      h = h - f * g;
      d[i - 1] = f - g;
      for (int j = 0; j < i; j++) e[j] = 0.0;
      for (int j = 0; j < i; j++) {
        f = d[j];
        v(j, i) = f;
        g = e[j] + v(j, j) * f;
        for (int k = j + 1; k <= i - 1; k++) { g += v(k, j) * d[k]; e[k] += v(k, j) * f; }
        e[j] = g;
      }
      f = 0.0;
      for (int j = 0; j < i; j++) { e[j] /= h; f += e[j] * d[j]; }
      double hh = f / (h + h);
      for (int j = 0; j < i; j++) e[j] -= hh * d[j];
      for (int j = 0; j < i; j++) {
        f = d[j]; g = e[j];
        for (int k = j; k <= i - 1; k++) v(k, j) -= (f * e[k] + g * d[k]);
        d[j] = v(i - 1, j);
        v(i, j) = 0.0;
      }
    }
    d[i] = h;
  }
  for (int i = 0; i < n - 1; i++) {
    v(n - 1, i) = v(i, i);
    v(i, i) = 1.0;
    double h = d[i + 1];
    if (h != 0.0) {
      for (int k = 0; k <= i; k++) d[k] = v(k, i + 1) / h;
      for (int j = 0; j <= i; j++) {
        double g = 0.0;
        for (int k = 0; k <= i; k++) g += v(k, i + 1) * v(k, j);
        for (int k = 0; k <= i; k++) v(k, j) -= g * d[k];
      }
    }
    for (int k = 0; k <= i; k++) v(k, i + 1) = 0.0;
  }
  for (int j = 0; j < n; j++) { d[j] = v(n - 1, j); v(n - 1, j) = 0.0; }
  v(n - 1, n - 1) = 1.0;
  e[0] = 0.0;
  // tql2
  for (int i = 1; i < n; i++) e[i - 1] = e[i];
  e[n - 1] = 0.0;
  double f = 0.0, tst1 = 0.0;
  const double eps = std::pow(2.0, -52.0);
  for (int l = 0; l < n; l++) {
    tst1 = std::max(tst1, std::fabs(d[l]) + std::fabs(e[l]));
    int m = l;
    while (m < n) { if (std::fabs(e[m]) <= eps * tst1) break; m++; }
    if (m > l) {
      int iter = 0;
      do {
        iter++;
        double g = d[l];
        double p = (d[l + 1] - g) / (2.0 * e[l]);
        double r = std::hypot(p, 1.0);
        if (p < 0) r = -r;
        d[l] = e[l] / (p + r);
        d[l + 1] = e[l] * (p + r);
        double dl1 = d[l + 1];
        double h = g - d[l];
        for (int i = l + 2; i < n; i++) d[i] -= h;
        f += h;
        p = d[m];
        double c = 1.0, c2 = c, c3 = c, el1 = e[l + 1], s = 0.0, s2 = 0.0;
        for (int i = m - 1; i >= l; i--) {
          c3 = c2; c2 = c; s2 = s;
          g = c * e[i];
          h = c * p;
          r = std::hypot(p, e[i]);
          e[i + 1] = s * r;
          s = e[i] / r;
          c = p / r;
          p = c * d[i] - s * g;
          d[i + 1] = h + s * (c * g + s * d[i]);
          for (int k = 0; k < n; k++) { h = v(k, i + 1); v(k, i + 1) = s * v(k, i) + c * h; v(k, i) = c * v(k, i) - s * h; }
        }
        p = -s * s2 * c3 * el1 * e[l] / dl1;
        e[l] = s * p;
        d[l] = c * p;
      } while (std::fabs(e[l]) > eps * tst1 && iter < 200);
    }
    d[l] = d[l] + f;
    e[l] = 0.0;
  }
}

struct MargBlockKey {
  int kind;  // 0 pose, 1 speedbias, 2 ex, 3 td, 4 point, 5 line
  int id;
  bool operator<(const MargBlockKey &o) const { return kind != o.kind ? kind < o.kind : id < o.id; }
};

struct MargResult {
  int m = 0, n = 0;
  std::vector<double> A, b;     // Schur complement A' (n x n), b'
  std::vector<double> J, r;     // linearized_jacobians (n x n), linearized_residuals
  std::vector<int> block_kind, block_id;   // kept blocks, ids after the window shift
  std::vector<double> x0;       // global-size data of kept blocks
  std::vector<double> A_full, b_full;   // the system before the Schur complement, (m + n) x (m + n), dropped blocks first (tests only)
};

// Build the next prior from the window's state `s`.  flag: UVS_MARGIN_OLD / UVS_MARGIN_SECOND_NEW.
// Returns false when the reference would build nothing (estimator.cpp:1162-1164).
inline bool marginalize(const Problem &P, const State &s, int flag, MargResult &out, double eps = 1e-8) {
  const UvsWindow &w = P.w;
  const int F = w.n_frames;
  struct Fac { int nr; int nb; MargBlockKey key[40]; int gs[40]; std::vector<double> r; std::vector<std::vector<double>> J; std::vector<int> drop; };
  std::vector<Fac> facs;

  auto gsize = [](int kind) { return kind == 0 || kind == 2 ? 7 : (kind == 1 ? 9 : (kind == 5 ? 4 : 1)); };
  auto add_prior = [&](const std::vector<MargBlockKey> &drop_keys) {
    Fac f; f.nr = w.prior_n; f.nb = w.prior_n_blocks;
    f.r.resize(f.nr);
    P.evaluate_prior(s, f.r.data());
    f.J.resize(f.nb);
    for (int b = 0; b < f.nb; b++) {
      f.key[b] = {w.prior_block_kind[b], w.prior_block_id[b]};
      f.gs[b] = gsize(f.key[b].kind);
      const int ls = f.gs[b] == 7 ? 6 : f.gs[b];
      f.J[b].assign((size_t)f.nr * f.gs[b], 0.0);
      for (int i = 0; i < f.nr; i++) for (int c = 0; c < ls; c++) f.J[b][(size_t)i * f.gs[b] + c] = w.prior_J[(size_t)i * w.prior_n + P.prior_col_[b] + c];
      for (const auto &dk : drop_keys) if (!(dk < f.key[b]) && !(f.key[b] < dk)) f.drop.push_back(b);
    }
    facs.push_back(std::move(f));
  };
  auto add_small = [&](int type, int idx, const MargBlockKey *keys, std::initializer_list<int> drop) {
    BlockEval e;
    P.evaluate_raw(type, idx, s, e, true);
    double *jp[BlockEval::MAXB];
    for (int b = 0; b < e.nb; b++) jp[b] = e.J[b];
    apply_corrector(P.loss_scale(type), e.nr, e.r, e.nb, jp, e.gs);   // ResidualBlockInfo::Evaluate
    Fac f; f.nr = e.nr; f.nb = e.nb;
    f.r.assign(e.r, e.r + e.nr);
    f.J.resize(e.nb);
    for (int b = 0; b < e.nb; b++) { f.key[b] = keys[b]; f.gs[b] = e.gs[b]; f.J[b].assign(e.J[b], e.J[b] + e.nr * e.gs[b]); }
    f.drop.assign(drop.begin(), drop.end());
    facs.push_back(std::move(f));
  };

  if (flag == UVS_MARGIN_OLD) {
    if (w.prior_n > 0) add_prior({{0, 0}, {1, 0}});                                   // estimator.cpp:1008-1024
    for (int k = 0; k < w.n_imu; k++) {                                               // :1026-1035
      if (w.imu_frame_i[k] != 0 || !(w.imu_sum_dt[k] < 10.0)) continue;
      MargBlockKey keys[4] = {{0, 0}, {1, 0}, {0, 1}, {1, 1}};
      add_small(F_IMU, k, keys, {0, 1});
    }
    for (int k = 0; k < w.n_proj; k++) {                                              // :1037-1079
      if (w.proj_frame_i[k] != 0) continue;
      MargBlockKey keys[5] = {{0, 0}, {0, w.proj_frame_j[k]}, {2, 0}, {4, w.proj_point[k]}, {3, 0}};
      add_small(F_PROJ, k, keys, {0, 3});
    }
    std::vector<int> line_start(w.n_lines, 1 << 30);                                  // :1081-1129
    for (int k = 0; k < w.n_line_obs; k++) line_start[w.line_idx[k]] = std::min(line_start[w.line_idx[k]], w.line_frame[k]);
    for (int k = 0; k < w.n_line_obs; k++) {
      const int lk = w.line_idx[k], fj = w.line_frame[k];
      if (line_start[lk] != 0 || fj == 0) continue;
      MargBlockKey keys[2] = {{0, fj}, {5, lk}};
      add_small(F_LINE, k, keys, {1});
    }
    for (int k = 0; k < w.n_vp_obs; k++) {
      const int lk = w.vp_line[k], fj = w.vp_frame[k];
      if (line_start[lk] != 0 || fj == 0) continue;
      MargBlockKey keys[2] = {{0, fj}, {5, lk}};
      add_small(F_VP, k, keys, {1});
    }
  } else {
    bool has = false;                                                                 // :1162-1164
    for (int b = 0; b < (w.prior_n > 0 ? w.prior_n_blocks : 0); b++) if (w.prior_block_kind[b] == UVS_BLOCK_POSE && w.prior_block_id[b] == F - 2) has = true;
    if (!has) return false;
    add_prior({{0, F - 2}});
  }
  if (facs.empty()) return false;

  // addResidualBlockInfo + marginalize(): index assignment, dropped blocks first
  std::map<MargBlockKey, int> size_of, idx_of;
  std::map<MargBlockKey, bool> dropped;
  for (const Fac &f : facs) {
    for (int b = 0; b < f.nb; b++) size_of[f.key[b]] = f.gs[b];
    for (int di : f.drop) dropped[f.key[di]] = true;
  }
  auto local = [](int gs) { return gs == 7 ? 6 : gs; };
  int pos = 0;
  const int drop_order[6] = {0, 1, 4, 5, 2, 3};
  for (int ko = 0; ko < 6; ko++) for (auto &kv : size_of) if (kv.first.kind == drop_order[ko] && dropped.count(kv.first)) { idx_of[kv.first] = pos; pos += local(kv.second); }
  const int m = pos;
  std::vector<MargBlockKey> kept;
  for (int f = 0; f < F; f++) for (int kind = 0; kind < 2; kind++) { MargBlockKey k{kind, f}; if (size_of.count(k) && !dropped.count(k)) kept.push_back(k); }
  for (int kind = 2; kind < 6; kind++) for (auto &kv : size_of) if (kv.first.kind == kind && !dropped.count(kv.first)) kept.push_back(kv.first);
  for (const auto &k : kept) { idx_of[k] = pos; pos += local(size_of[k]); }
  const int n = pos - m;

  // ThreadsConstructA: A = sum J'J, b = sum J'r over tangent columns
  std::vector<double> A((size_t)pos * pos, 0.0), b(pos, 0.0);
  for (const Fac &f : facs) {
    for (int i = 0; i < f.nb; i++) {
      const int idx_i = idx_of[f.key[i]], size_i = local(f.gs[i]), gi = f.gs[i];
      for (int j = i; j < f.nb; j++) {
        const int idx_j = idx_of[f.key[j]], size_j = local(f.gs[j]), gj = f.gs[j];
        for (int p = 0; p < size_i; p++) for (int q = 0; q < size_j; q++) {
          double h = 0.0;
          for (int t = 0; t < f.nr; t++) h += f.J[i][(size_t)t * gi + p] * f.J[j][(size_t)t * gj + q];
          A[(size_t)(idx_i + p) * pos + idx_j + q] += h;
          if (i != j) A[(size_t)(idx_j + q) * pos + idx_i + p] = A[(size_t)(idx_i + p) * pos + idx_j + q];
        }
      }
      for (int p = 0; p < size_i; p++) { double g = 0.0; for (int t = 0; t < f.nr; t++) g += f.J[i][(size_t)t * gi + p] * f.r[t]; b[idx_i + p] += g; }
    }
  }
  out.A_full = A; out.b_full = b;
  // Amm^-1 by eigendecomposition with eigenvalues <= eps zeroed          marginalization_factor.cpp:266-272
  std::vector<double> Amm((size_t)m * m), ev, V;
  for (int i = 0; i < m; i++) for (int j = 0; j < m; j++) Amm[(size_t)i * m + j] = 0.5 * (A[(size_t)i * pos + j] + A[(size_t)j * pos + i]);
  std::vector<double> Amm_inv((size_t)m * m, 0.0);
  if (m > 0) {
    sym_eig(m, Amm, ev, V);
    for (int k = 0; k < m; k++) {
      if (!(ev[k] > eps)) continue;
      const double inv = 1.0 / ev[k];
      for (int i = 0; i < m; i++) { const double vi = V[(size_t)i * m + k] * inv; for (int j = 0; j < m; j++) Amm_inv[(size_t)i * m + j] += vi * V[(size_t)j * m + k]; }
    }
  }
  // A' = Arr - Arm Amm^-1 Amr ; b' = brr - Arm Amm^-1 bmm                 :274-281
  std::vector<double> T((size_t)n * m, 0.0);  // Arm * Amm_inv
  for (int i = 0; i < n; i++) for (int k = 0; k < m; k++) { const double a = A[(size_t)(m + i) * pos + k]; if (a == 0.0) continue; for (int j = 0; j < m; j++) T[(size_t)i * m + j] += a * Amm_inv[(size_t)k * m + j]; }
  out.A.assign((size_t)n * n, 0.0); out.b.assign(n, 0.0);
  for (int i = 0; i < n; i++) {
    for (int j = 0; j < n; j++) { double sacc = 0.0; for (int k = 0; k < m; k++) sacc += T[(size_t)i * m + k] * A[(size_t)k * pos + m + j]; out.A[(size_t)i * n + j] = A[(size_t)(m + i) * pos + m + j] - sacc; }
    double sacc = 0.0; for (int k = 0; k < m; k++) sacc += T[(size_t)i * m + k] * b[k];
    out.b[i] = b[m + i] - sacc;
  }
  // second eigendecomposition: J0 = sqrt(S) V', r0 = S^-1/2 V' b'           :283-291
  std::vector<double> ev2, V2;
  sym_eig(n, out.A, ev2, V2);
  out.J.assign((size_t)n * n, 0.0); out.r.assign(n, 0.0);
  for (int k = 0; k < n; k++) {
    const double S = ev2[k] > eps ? ev2[k] : 0.0, Sinv = ev2[k] > eps ? 1.0 / ev2[k] : 0.0;
    const double ssq = std::sqrt(S), sisq = std::sqrt(Sinv);
    double vb = 0.0;
    for (int j = 0; j < n; j++) { out.J[(size_t)k * n + j] = ssq * V2[(size_t)j * n + k]; vb += V2[(size_t)j * n + k] * out.b[j]; }
    out.r[k] = sisq * vb;
  }
  out.m = m; out.n = n;
  // getParameterBlocks with addr_shift (estimator.cpp:1139-1153 / 1199-1222)
  out.block_kind.clear(); out.block_id.clear(); out.x0.clear();
  for (const auto &k : kept) {
    int id = k.id;
    if (k.kind <= 1) {
      if (flag == UVS_MARGIN_OLD) id = k.id - 1;
      else if (k.id == F - 1) id = k.id - 1;
    }
    out.block_kind.push_back(k.kind);
    out.block_id.push_back(id);
    const double *src = k.kind == 0 ? &s.pose[7 * k.id] : (k.kind == 1 ? &s.sb[9 * k.id] : (k.kind == 2 ? s.ex.data() : s.td.data()));
    out.x0.insert(out.x0.end(), src, src + size_of[k]);
  }
  return true;
}

}  // namespace orc

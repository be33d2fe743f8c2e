// ORACLE — TEST INFRASTRUCTURE ONLY (see smallmat.h).  PARITY UNPINNED: Ceres is not vendored and
// its version is not pinned by the reference; this file restates Ceres' DOCUMENTED trust-region
// behaviour for the options Estimator::optimization() sets (estimator.cpp:982-994):
//   linear_solver_type = SPARSE_SCHUR, trust_region_strategy_type = LEVENBERG_MARQUARDT,
//   max_num_iterations = NUM_ITERATIONS, max_solver_time_in_seconds, everything else default.
//
// Semantics restated:
//   * cost = 1/2 sum rho(|r|^2); residual blocks corrected as marginalization_factor.cpp:37-68
//   * Jacobi scaling  s_c = 1/(1 + ||J0[:,c]||)   from the FIRST Jacobian only
//   * LM diagonal     D_c^2 = clamp(||J~[:,c]||^2, 1e-6, 1e32) / radius     (J~ = J diag(s))
//   * (J~'J~ + D^2) y = -J~'r solved exactly by Schur elimination of all point (1x1) and line (4x4)
//     blocks followed by dense Cholesky of the reduced camera system; delta = s .* y
//   * model_cost_change = -(J~y)'(r + J~y/2); invalid step if <= 0
//   * x+ = Plus(x, delta); relative_decrease = (cost - cost+)/model_cost_change
//   * accept iff > min_relative_decrease: radius /= max(1/3, 1-(2 rho-1)^3), decrease_factor = 2
//     else radius /= decrease_factor, decrease_factor *= 2
//   * termination: iteration cap, |dcost| <= ftol*cost, |step| <= ptol(|x|+ptol), max|g| <= gtol
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <vector>

#include "problem.h"

namespace orc {

// Normal equations in Schur-ready form (all in Jacobi-SCALED variables when scale != nullptr).
struct NormalEq {
  int d = 0, Np = 0, Nl = 0;
  std::vector<double> Hcc;  // d x d row-major, full symmetric
  std::vector<double> gc;   // d          (J'r)
  std::vector<double> Hpp, gp;          // [Np]
  std::vector<double> Hll, gl;          // [Nl][16], [Nl][4]
  // landmark-camera coupling rows: per landmark a short list of camera blocks
  struct Row {
    int nblk = 0;
    int off[40];
    int sz[40];
    double v[40][24];  // sz x lsz row-major (lsz = 1 for points, 4 for lines)
  };
  std::vector<Row> Pc, Lc;
  void reset(const Layout &lay) {
    d = lay.d; Np = lay.Np; Nl = lay.Nl;
    Hcc.assign((size_t)d * d, 0.0); gc.assign(d, 0.0);
    Hpp.assign(Np, 0.0); gp.assign(Np, 0.0);
    Hll.assign((size_t)Nl * 16, 0.0); gl.assign((size_t)Nl * 4, 0.0);
    Pc.resize(Np); Lc.resize(Nl);
    for (auto &r : Pc) r.nblk = 0;
    for (auto &r : Lc) r.nblk = 0;
  }
  static double *row_block(Row &row, int off, int sz, int lsz) {
    for (int i = 0; i < row.nblk; i++) if (row.off[i] == off) return row.v[i];
    const int i = row.nblk++;
    row.off[i] = off; row.sz[i] = sz;
    std::fill(row.v[i], row.v[i] + sz * lsz, 0.0);
    return row.v[i];
  }
};

struct IterLog {
  double cost, radius, relative_decrease, step_norm, gradient_max_norm;
  int accepted;
};

class Solver {
 public:
  const Problem &P;
  const Layout &lay;
  std::vector<double> scale;   // Jacobi scaling per tangent column
  std::vector<double> diag;    // squared column norms of the scaled Jacobian
  std::vector<double> grad;    // unscaled gradient J'r
  NormalEq ne;
  std::vector<double> prior_r;
  bool have_scale = false;
  bool dense_check = false;    // solve the full normal equations densely instead of by Schur (tests)

  explicit Solver(const Problem &p) : P(p), lay(p.lay) {
    scale.assign(lay.total, 1.0);
    diag.assign(lay.total, 0.0);
    grad.assign(lay.total, 0.0);
    prior_r.resize(std::max(1, P.w.prior_n));
  }

  // Evaluate all residual blocks with Jacobians at s, build the (scaled) normal equations.
  // Returns the cost.  First call fixes the Jacobi scaling.
  double linearize(const State &s) {
    const int d = lay.d;
    // pass 1 (first call only): column norms of the unscaled Jacobian
    double cost = 0.0;
    std::vector<BlockEval> &ev = evals_;
    size_t nf = 0;
    for (int t = F_IMU; t <= F_LAST; t++) nf += P.num_factors(t);
    ev.resize(nf);
    size_t k = 0;
    for (int t = F_IMU; t <= F_LAST; t++)
      for (int i = 0; i < P.num_factors(t); i++) { P.evaluate_block(t, i, s, ev[k], true); cost += ev[k].cost; k++; }
    const int n = P.w.prior_n;
    if (n > 0) {
      P.evaluate_prior(s, prior_r.data());
      double sq = 0; for (int i = 0; i < n; i++) sq += prior_r[i] * prior_r[i];
      cost += 0.5 * sq;
    }
    if (!have_scale) {
      std::vector<double> cn(lay.total, 0.0);
      for (const BlockEval &e : ev)
        for (int b = 0; b < e.nb; b++) {
          if (e.off[b] < 0) continue;
          for (int i = 0; i < e.nr; i++) for (int c = 0; c < e.ls[b]; c++) { double v = e.J[b][i * e.ls[b] + c]; cn[e.off[b] + c] += v * v; }
        }
      for (int b = 0; b < (n > 0 ? P.w.prior_n_blocks : 0); b++) {
        if (P.prior_off_[b] < 0) continue;
        const int ls = P.prior_gs_[b] == 7 ? 6 : P.prior_gs_[b];
        for (int i = 0; i < n; i++) for (int c = 0; c < ls; c++) { double v = P.w.prior_J[(size_t)i * n + P.prior_col_[b] + c]; cn[P.prior_off_[b] + c] += v * v; }
      }
      for (int c = 0; c < lay.total; c++) scale[c] = 1.0 / (1.0 + std::sqrt(cn[c]));
      have_scale = true;
    }
    // pass 2: accumulate scaled J'J and J'r
    ne.reset(lay);
    std::fill(grad.begin(), grad.end(), 0.0);
    for (const BlockEval &e : ev) accumulate(e);
    if (n > 0) accumulate_prior();
    // squared column norms of the scaled Jacobian = diagonal of the scaled Hessian
    for (int c = 0; c < d; c++) diag[c] = ne.Hcc[(size_t)c * d + c];
    for (int kp = 0; kp < lay.Np; kp++) diag[lay.point(kp)] = ne.Hpp[kp];
    for (int kl = 0; kl < lay.Nl; kl++) for (int c = 0; c < 4; c++) diag[lay.line(kl) + c] = ne.Hll[(size_t)kl * 16 + c * 5];
    return cost;
  }

  // Solve (H + D^2) y = -g in scaled variables; delta = scale .* y.  Returns false if not PD.
  bool compute_step(double radius, std::vector<double> &delta, double *model_cost_change) {
    const int d = lay.d, T = lay.total;
    std::vector<double> D2(T);
    for (int c = 0; c < T; c++) D2[c] = std::min(std::max(diag[c], P.o.min_lm_diagonal), P.o.max_lm_diagonal) / radius;
    std::vector<double> y(T, 0.0);
    bool ok = dense_check ? solve_dense(D2, y) : solve_schur(D2, y);
    if (!ok) return false;
    // model_cost_change = -(J y)'(r + J y / 2) = -y'g - 1/2 y'Hy   (explicit product, as Ceres does)
    double yg = 0.0, yHy = 0.0;
    {
      std::vector<double> Hy(T, 0.0);
      multiply_H(y, Hy);
      for (int c = 0; c < T; c++) yHy += y[c] * Hy[c];
      for (int c = 0; c < d; c++) yg += y[c] * ne.gc[c];
      for (int kp = 0; kp < lay.Np; kp++) yg += y[lay.point(kp)] * ne.gp[kp];
      for (int kl = 0; kl < lay.Nl; kl++) for (int c = 0; c < 4; c++) yg += y[lay.line(kl) + c] * ne.gl[(size_t)kl * 4 + c];
    }
    *model_cost_change = -yg - 0.5 * yHy;
    delta.resize(T);
    for (int c = 0; c < T; c++) delta[c] = y[c] * scale[c];
    return true;
  }

  // ceres::Solve restated.  Returns final state in `s`.
  void solve(State &s, UvsSummary *sum) {
    const UvsOptions &o = P.o;
    auto t0 = std::chrono::steady_clock::now();
    double radius = o.initial_radius, decrease_factor = 2.0;
    double cost = linearize(s);
    double x_norm = std::sqrt(P.ambient_sqnorm(s));
    std::vector<IterLog> log;
    auto gmax = [&]() { double m = 0; for (double v : grad) m = std::max(m, std::fabs(v)); return m; };
    log.push_back({cost, radius, 0.0, 0.0, gmax(), 1});
    int termination = UVS_TERM_NO_CONVERGENCE, n_success = 0, n_invalid = 0;
    const double initial_cost = cost;
    State cand;
    std::vector<double> delta;
    const bool fixed = o.fixed_iterations != 0;
    if (!fixed && log.back().gradient_max_norm <= o.gradient_tolerance) termination = UVS_TERM_GRADIENT_TOL;
    else
      for (int it = 1; it <= o.max_num_iterations; it++) {
        if (o.max_solver_time > 0.0 && !fixed) {
          double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
          if (el >= o.max_solver_time) { termination = UVS_TERM_TIME; break; }
        }
        double model_change = 0.0;
        bool ok = compute_step(radius, delta, &model_change);
        bool valid = ok && model_change > 0.0;
        for (double v : delta) if (!std::isfinite(v)) valid = false;
        if (!valid) {  // invalid step: treated as a rejected step by the LM strategy
          radius /= decrease_factor; decrease_factor *= 2.0;
          log.push_back({cost, radius, 0.0, 0.0, log.back().gradient_max_norm, 0});
          if (++n_invalid >= 5 && !fixed) { termination = UVS_TERM_FAILURE; break; }
          if (!fixed && radius < o.min_radius) { termination = UVS_TERM_MIN_RADIUS; break; }
          continue;
        }
        n_invalid = 0;
        P.plus(s, delta.data(), cand);
        const double cand_cost = P.total_cost(cand);
        const double step_norm = std::sqrt(P.ambient_sqdist(s, cand));
        if (!fixed && step_norm <= o.parameter_tolerance * (x_norm + o.parameter_tolerance)) {
          log.push_back({cost, radius, 0.0, step_norm, log.back().gradient_max_norm, 0});
          termination = UVS_TERM_PARAMETER_TOL; break;
        }
        const double cost_change = cost - cand_cost;
        if (!fixed && std::fabs(cost_change) <= o.function_tolerance * cost) {
          log.push_back({cost, radius, 0.0, step_norm, log.back().gradient_max_norm, 0});
          termination = UVS_TERM_FUNCTION_TOL; break;
        }
        const double rel = cost_change / model_change;
        const bool accept = std::isfinite(cand_cost) && rel > o.min_relative_decrease;
        if (accept) {
          s = cand;
          x_norm = std::sqrt(P.ambient_sqnorm(s));
          cost = linearize(s);
          radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rel - 1.0, 3));
          radius = std::min(o.max_radius, radius);
          decrease_factor = 2.0;
          n_success++;
        } else {
          radius /= decrease_factor; decrease_factor *= 2.0;
        }
        log.push_back({cost, radius, rel, step_norm, gmax(), accept ? 1 : 0});
        if (!fixed) {
          if (accept && log.back().gradient_max_norm <= o.gradient_tolerance) { termination = UVS_TERM_GRADIENT_TOL; break; }
          if (radius < o.min_radius) { termination = UVS_TERM_MIN_RADIUS; break; }
        }
      }
    if (sum) {
      std::memset(sum, 0, sizeof(*sum));
      sum->num_iterations = (int)log.size();
      sum->num_successful_steps = n_success;
      sum->termination = termination;
      sum->status = std::isfinite(cost) ? UVS_OK : UVS_ERR_NOT_FINITE;
      sum->initial_cost = initial_cost;
      sum->final_cost = cost;
      for (size_t i = 0; i < log.size() && i < UVS_MAX_ITER_LOG; i++) {
        sum->cost[i] = log[i].cost; sum->radius[i] = log[i].radius; sum->relative_decrease[i] = log[i].relative_decrease;
        sum->step_norm[i] = log[i].step_norm; sum->gradient_max_norm[i] = log[i].gradient_max_norm; sum->step_accepted[i] = log[i].accepted;
      }
    }
  }

  // Reduced camera system of the last compute_step (for tests): S (d x d), g_S (d), scaled variables.
  std::vector<double> last_S, last_gS;

 private:
  std::vector<BlockEval> evals_;

  void accumulate(const BlockEval &e) {
    const int d = lay.d;
    // scaled copies of the blocks
    double Js[BlockEval::MAXB][BlockEval::MAXJ];
    for (int b = 0; b < e.nb; b++) {
      if (e.off[b] < 0) continue;
      const int ls = e.ls[b];
      for (int i = 0; i < e.nr; i++) for (int c = 0; c < ls; c++) Js[b][i * ls + c] = e.J[b][i * ls + c] * scale[e.off[b] + c];
      for (int c = 0; c < ls; c++) { double g = 0; for (int i = 0; i < e.nr; i++) g += e.J[b][i * ls + c] * e.r[i]; grad[e.off[b] + c] += g; }
    }
    for (int a = 0; a < e.nb; a++) {
      if (e.off[a] < 0) continue;
      const int la = e.ls[a], oa = e.off[a];
      const bool a_cam = oa < d;
      // gradient (scaled)
      for (int c = 0; c < la; c++) {
        double g = 0; for (int i = 0; i < e.nr; i++) g += Js[a][i * la + c] * e.r[i];
        if (a_cam) ne.gc[oa + c] += g;
        else if (oa < d + lay.Np) ne.gp[oa - d] += g;
        else ne.gl[oa - d - lay.Np + c] += g;   // (oa-d-Np) = 4*k
      }
      for (int b = 0; b < e.nb; b++) {
        if (e.off[b] < 0) continue;
        const int lb = e.ls[b], ob = e.off[b];
        const bool b_cam = ob < d;
        if (a_cam && b_cam) {
          for (int p = 0; p < la; p++) for (int q = 0; q < lb; q++) { double h = 0; for (int i = 0; i < e.nr; i++) h += Js[a][i * la + p] * Js[b][i * lb + q]; ne.Hcc[(size_t)(oa + p) * d + ob + q] += h; }
        } else if (!a_cam && !b_cam) {
          if (a != b) continue;  // a factor touches at most one landmark
          if (oa < d + lay.Np) { double h = 0; for (int i = 0; i < e.nr; i++) h += Js[a][i] * Js[a][i]; ne.Hpp[oa - d] += h; }
          else {
            double *H = &ne.Hll[(size_t)((oa - d - lay.Np) / 4) * 16];
            for (int p = 0; p < 4; p++) for (int q = 0; q < 4; q++) { double h = 0; for (int i = 0; i < e.nr; i++) h += Js[a][i * 4 + p] * Js[a][i * 4 + q]; H[p * 4 + q] += h; }
          }
        } else if (a_cam && !b_cam) {  // coupling block camera(a) x landmark(b), stored sz x lsz
          const bool is_pt = ob < d + lay.Np;
          NormalEq::Row &row = is_pt ? ne.Pc[ob - d] : ne.Lc[(ob - d - lay.Np) / 4];
          double *v = NormalEq::row_block(row, oa, la, lb);
          for (int p = 0; p < la; p++) for (int q = 0; q < lb; q++) { double h = 0; for (int i = 0; i < e.nr; i++) h += Js[a][i * la + p] * Js[b][i * lb + q]; v[p * lb + q] += h; }
        }
      }
    }
  }

  void accumulate_prior() {
    const int n = P.w.prior_n, d = lay.d, nb = P.w.prior_n_blocks;
    // J = J0 column slices scaled; H += J'J ; g += J' r
    std::vector<int> cmap(n, -1);   // J0 column -> tangent offset
    for (int b = 0; b < nb; b++) {
      if (P.prior_off_[b] < 0) continue;
      const int ls = P.prior_gs_[b] == 7 ? 6 : P.prior_gs_[b];
      for (int c = 0; c < ls; c++) cmap[P.prior_col_[b] + c] = P.prior_off_[b] + c;
    }
    const double *J0 = P.w.prior_J;
    std::vector<double> Jt((size_t)n * n);  // transposed + scaled for contiguous dot products
    for (int c = 0; c < n; c++) {
      const double sc = cmap[c] >= 0 ? scale[cmap[c]] : 0.0;
      for (int i = 0; i < n; i++) Jt[(size_t)c * n + i] = J0[(size_t)i * n + c] * sc;
    }
    for (int p = 0; p < n; p++) {
      if (cmap[p] < 0) continue;
      double g = 0, gu = 0;
      for (int i = 0; i < n; i++) { g += Jt[(size_t)p * n + i] * prior_r[i]; gu += J0[(size_t)i * n + p] * prior_r[i]; }
      ne.gc[cmap[p]] += g;
      grad[cmap[p]] += gu;
      for (int q = 0; q < n; q++) {
        if (cmap[q] < 0) continue;
        double h = 0;
        const double *a = &Jt[(size_t)p * n], *b = &Jt[(size_t)q * n];
        for (int i = 0; i < n; i++) h += a[i] * b[i];
        ne.Hcc[(size_t)cmap[p] * d + cmap[q]] += h;
      }
    }
  }

  static void inv4_chol(const double *A, double *Ainv, bool *ok) {
    double L[16];
    *ok = cholesky_lower(4, A, L);
    if (!*ok) return;
    // Ainv = L^-T L^-1 by solving for identity columns
    for (int c = 0; c < 4; c++) {
      double x[4] = {0, 0, 0, 0};
      x[c] = 1.0;
      for (int i = 0; i < 4; i++) { double s = x[i]; for (int j = 0; j < i; j++) s -= L[i * 4 + j] * x[j]; x[i] = s / L[i * 4 + i]; }
      for (int i = 3; i >= 0; i--) { double s = x[i]; for (int j = i + 1; j < 4; j++) s -= L[j * 4 + i] * x[j]; x[i] = s / L[i * 4 + i]; }
      for (int i = 0; i < 4; i++) Ainv[i * 4 + c] = x[i];
    }
  }

  bool solve_schur(const std::vector<double> &D2, std::vector<double> &y) {
    const int d = lay.d;
    std::vector<double> S(ne.Hcc), g(ne.gc);
    for (int c = 0; c < d; c++) S[(size_t)c * d + c] += D2[c];
    std::vector<double> hinv_p(lay.Np);
    std::vector<double> hinv_l((size_t)lay.Nl * 16);
    for (int k = 0; k < lay.Np; k++) {
      const double h = ne.Hpp[k] + D2[lay.point(k)];
      hinv_p[k] = 1.0 / h;
      const NormalEq::Row &row = ne.Pc[k];
      for (int a = 0; a < row.nblk; a++) {
        for (int p = 0; p < row.sz[a]; p++) {
          const double wa = row.v[a][p] * hinv_p[k];
          g[row.off[a] + p] -= wa * ne.gp[k];
          for (int b = 0; b < row.nblk; b++) for (int q = 0; q < row.sz[b]; q++) S[(size_t)(row.off[a] + p) * d + row.off[b] + q] -= wa * row.v[b][q];
        }
      }
    }
    for (int k = 0; k < lay.Nl; k++) {
      double H[16];
      for (int i = 0; i < 16; i++) H[i] = ne.Hll[(size_t)k * 16 + i];
      for (int c = 0; c < 4; c++) H[c * 5] += D2[lay.line(k) + c];
      bool ok;
      inv4_chol(H, &hinv_l[(size_t)k * 16], &ok);
      if (!ok) return false;
      const double *Hi = &hinv_l[(size_t)k * 16];
      const NormalEq::Row &row = ne.Lc[k];
      for (int a = 0; a < row.nblk; a++) {
        double WH[6 * 4];  // V_a Hinv  (sz x 4)
        for (int p = 0; p < row.sz[a]; p++) for (int q = 0; q < 4; q++) { double s = 0; for (int t = 0; t < 4; t++) s += row.v[a][p * 4 + t] * Hi[t * 4 + q]; WH[p * 4 + q] = s; }
        for (int p = 0; p < row.sz[a]; p++) {
          double s = 0; for (int t = 0; t < 4; t++) s += WH[p * 4 + t] * ne.gl[(size_t)k * 4 + t];
          g[row.off[a] + p] -= s;
          for (int b = 0; b < row.nblk; b++) for (int q = 0; q < row.sz[b]; q++) {
            double h = 0; for (int t = 0; t < 4; t++) h += WH[p * 4 + t] * row.v[b][q * 4 + t];
            S[(size_t)(row.off[a] + p) * d + row.off[b] + q] -= h;
          }
        }
      }
    }
    last_S = S; last_gS = g;
    // dense Cholesky of the reduced camera system
    std::vector<double> L((size_t)d * d);
    if (!cholesky_lower(d, S.data(), L.data())) return false;
    std::vector<double> yc(d);
    for (int i = 0; i < d; i++) { double s = -g[i]; for (int j = 0; j < i; j++) s -= L[(size_t)i * d + j] * yc[j]; yc[i] = s / L[(size_t)i * d + i]; }
    for (int i = d - 1; i >= 0; i--) { double s = yc[i]; for (int j = i + 1; j < d; j++) s -= L[(size_t)j * d + i] * yc[j]; yc[i] = s / L[(size_t)i * d + i]; }
    for (int c = 0; c < d; c++) y[c] = yc[c];
    // back-substitution: y_k = -Hkk^-1 (g_k + H_kC y_C)
    for (int k = 0; k < lay.Np; k++) {
      double t = ne.gp[k];
      const NormalEq::Row &row = ne.Pc[k];
      for (int a = 0; a < row.nblk; a++) for (int p = 0; p < row.sz[a]; p++) t += row.v[a][p] * yc[row.off[a] + p];
      y[lay.point(k)] = -hinv_p[k] * t;
    }
    for (int k = 0; k < lay.Nl; k++) {
      double t[4];
      for (int c = 0; c < 4; c++) t[c] = ne.gl[(size_t)k * 4 + c];
      const NormalEq::Row &row = ne.Lc[k];
      for (int a = 0; a < row.nblk; a++) for (int p = 0; p < row.sz[a]; p++) for (int c = 0; c < 4; c++) t[c] += row.v[a][p * 4 + c] * yc[row.off[a] + p];
      const double *Hi = &hinv_l[(size_t)k * 16];
      for (int c = 0; c < 4; c++) { double s = 0; for (int q = 0; q < 4; q++) s += Hi[c * 4 + q] * t[q]; y[lay.line(k) + c] = -s; }
    }
    return true;
  }

  void full_matrix(std::vector<double> &H, std::vector<double> &g) const {
    const int d = lay.d, T = lay.total;
    H.assign((size_t)T * T, 0.0); g.assign(T, 0.0);
    for (int i = 0; i < d; i++) { g[i] = ne.gc[i]; for (int j = 0; j < d; j++) H[(size_t)i * T + j] = ne.Hcc[(size_t)i * d + j]; }
    for (int k = 0; k < lay.Np; k++) {
      const int o = lay.point(k);
      H[(size_t)o * T + o] = ne.Hpp[k]; g[o] = ne.gp[k];
      const NormalEq::Row &row = ne.Pc[k];
      for (int a = 0; a < row.nblk; a++) for (int p = 0; p < row.sz[a]; p++) { H[(size_t)(row.off[a] + p) * T + o] = row.v[a][p]; H[(size_t)o * T + row.off[a] + p] = row.v[a][p]; }
    }
    for (int k = 0; k < lay.Nl; k++) {
      const int o = lay.line(k);
      for (int p = 0; p < 4; p++) { g[o + p] = ne.gl[(size_t)k * 4 + p]; for (int q = 0; q < 4; q++) H[(size_t)(o + p) * T + o + q] = ne.Hll[(size_t)k * 16 + p * 4 + q]; }
      const NormalEq::Row &row = ne.Lc[k];
      for (int a = 0; a < row.nblk; a++) for (int p = 0; p < row.sz[a]; p++) for (int c = 0; c < 4; c++) {
        H[(size_t)(row.off[a] + p) * T + o + c] = row.v[a][p * 4 + c]; H[(size_t)(o + c) * T + row.off[a] + p] = row.v[a][p * 4 + c];
      }
    }
  }

  bool solve_dense(const std::vector<double> &D2, std::vector<double> &y) {
    const int T = lay.total;
    std::vector<double> H, g;
    full_matrix(H, g);
    for (int c = 0; c < T; c++) H[(size_t)c * T + c] += D2[c];
    std::vector<double> L((size_t)T * T);
    if (!cholesky_lower(T, H.data(), L.data())) return false;
    for (int i = 0; i < T; i++) { double s = -g[i]; for (int j = 0; j < i; j++) s -= L[(size_t)i * T + j] * y[j]; y[i] = s / L[(size_t)i * T + i]; }
    for (int i = T - 1; i >= 0; i--) { double s = y[i]; for (int j = i + 1; j < T; j++) s -= L[(size_t)j * T + i] * y[j]; y[i] = s / L[(size_t)i * T + i]; }
    return true;
  }

  // Hy = H y using the block structure (no damping)
  void multiply_H(const std::vector<double> &y, std::vector<double> &Hy) const {
    const int d = lay.d;
    for (int i = 0; i < d; i++) { double s = 0; const double *r = &ne.Hcc[(size_t)i * d]; for (int j = 0; j < d; j++) s += r[j] * y[j]; Hy[i] = s; }
    for (int k = 0; k < lay.Np; k++) {
      const int o = lay.point(k);
      double s = ne.Hpp[k] * y[o];
      const NormalEq::Row &row = ne.Pc[k];
      for (int a = 0; a < row.nblk; a++) for (int p = 0; p < row.sz[a]; p++) { s += row.v[a][p] * y[row.off[a] + p]; Hy[row.off[a] + p] += row.v[a][p] * y[o]; }
      Hy[o] = s;
    }
    for (int k = 0; k < lay.Nl; k++) {
      const int o = lay.line(k);
      double s[4];
      for (int p = 0; p < 4; p++) { s[p] = 0; for (int q = 0; q < 4; q++) s[p] += ne.Hll[(size_t)k * 16 + p * 4 + q] * y[o + q]; }
      const NormalEq::Row &row = ne.Lc[k];
      for (int a = 0; a < row.nblk; a++) for (int p = 0; p < row.sz[a]; p++) for (int c = 0; c < 4; c++) {
        s[c] += row.v[a][p * 4 + c] * y[row.off[a] + p];
        Hy[row.off[a] + p] += row.v[a][p * 4 + c] * y[o + c];
      }
      for (int p = 0; p < 4; p++) Hy[o + p] = s[p];
    }
  }
};

}  // namespace orc

// ORACLE — TEST INFRASTRUCTURE ONLY.  C entry points (ctypes) over the CPU restatement.
// Allowed callers: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs.
// The product library (uv-slam_b200/csrc) never links or loads this file.
// PARITY UNPINNED at the Ceres/Eigen boundary (see smallmat.h, solver.h).
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "marg.h"
#include "solver.h"

using namespace orc;

extern "C" {

int orc_abi_version() { return UVS_ABI_VERSION; }

// number of doubles per factor for residuals / Jacobians in the two layouts
static void factor_dims(const UvsWindow *w, int type, int local, int *nr, int *jdoubles) {
  const int P = local ? 6 : 7;
  switch (type) {
    case F_PROJ: *nr = 2; *jdoubles = 2 * (3 * P + 1 + (w->estimate_td ? 1 : 0)); break;
    case F_LINE: *nr = 2; *jdoubles = 2 * (P + 4); break;
    case F_VP: *nr = 1; *jdoubles = P + 4; break;
    case F_IMU: *nr = 15; *jdoubles = 15 * (2 * P + 18); break;
    default: {
      *nr = w->prior_n;
      int cols = 0;
      for (int b = 0; b < w->prior_n_blocks; b++) {
        const int k = w->prior_block_kind[b];
        cols += (k == UVS_BLOCK_POSE || k == UVS_BLOCK_EXPOSE) ? P : (k == UVS_BLOCK_SPEEDBIAS ? 9 : 1);
      }
      *jdoubles = w->prior_n * cols;
    }
  }
}

int orc_factor_dims(const UvsWindow *w, int type, int local, int *nr, int *jdoubles) { factor_dims(w, type, local, nr, jdoubles); return 0; }

// Evaluate every factor of `type`.  local = 0: raw Evaluate(), Ceres layout.  local = 1: tangent
// columns + loss correction (what Ceres hands the linear solver).  J nullable.  cost nullable [n].
int orc_eval(const UvsWindow *w, const UvsOptions *o, int type, int local, double *r, double *J, double *cost) {
  Problem P(*w, *o);
  State s = P.initial_state();
  int nr, jd;
  factor_dims(w, type, local, &nr, &jd);
  if (type == F_PRIOR) {
    if (w->prior_n <= 0) return 0;
    P.evaluate_prior(s, r);
    if (cost) { double sq = 0; for (int i = 0; i < nr; i++) sq += r[i] * r[i]; cost[0] = 0.5 * sq; }
    if (J) {
      const int n = w->prior_n;
      size_t o2 = 0;
      for (int b = 0; b < w->prior_n_blocks; b++) {
        const int gs = P.prior_gs_[b], ls = gs == 7 ? 6 : gs, width = local ? ls : gs;
        for (int i = 0; i < n; i++) for (int c = 0; c < width; c++) J[o2 + (size_t)i * width + c] = c < ls ? w->prior_J[(size_t)i * n + P.prior_col_[b] + c] : 0.0;
        o2 += (size_t)n * width;
      }
    }
    return 0;
  }
  BlockEval e;
  const int nf = P.num_factors(type);
  for (int i = 0; i < nf; i++) {
    P.evaluate_raw(type, i, s, e, J != nullptr || local);
    if (local) {
      double *jp[BlockEval::MAXB];
      for (int b = 0; b < e.nb; b++) {
        if (e.gs[b] != e.ls[b]) for (int q = 0; q < e.nr; q++) for (int c = 0; c < e.ls[b]; c++) e.J[b][q * e.ls[b] + c] = e.J[b][q * e.gs[b] + c];
        jp[b] = e.J[b];
      }
      const double c = apply_corrector(P.loss_scale(type), e.nr, e.r, e.nb, jp, e.ls);
      if (cost) cost[i] = c;
    } else if (cost) {
      double sq = 0; for (int q = 0; q < e.nr; q++) sq += e.r[q] * e.r[q];
      cost[i] = 0.5 * sq;
    }
    std::memcpy(r + (size_t)i * nr, e.r, sizeof(double) * nr);
    if (J) {
      double *dst = J + (size_t)i * jd;
      for (int b = 0; b < e.nb; b++) { const int width = local ? e.ls[b] : e.gs[b]; std::memcpy(dst, e.J[b], sizeof(double) * e.nr * width); dst += e.nr * width; }
    }
  }
  return 0;
}

int orc_total_cost(const UvsWindow *w, const UvsOptions *o, double *cost) {
  Problem P(*w, *o);
  *cost = P.total_cost(P.initial_state());
  return 0;
}

// ceres::Solve restatement; state arrays of `w` are updated in place.
int orc_solve(UvsWindow *w, const UvsOptions *o, UvsSummary *sum, int dense_check) {
  Problem P(*w, *o);
  State s = P.initial_state();
  Solver S(P);
  S.dense_check = dense_check != 0;
  S.solve(s, sum);
  P.store_state(s, *w);
  return 0;
}

// Window-parallel batch: windows are independent, one std::thread per slice.
int orc_solve_batch(int n, UvsWindow *w, const UvsOptions *o, UvsSummary *sums, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  std::atomic<int> next(0);
  auto work = [&]() {
    for (;;) {
      const int i = next.fetch_add(1);
      if (i >= n) break;
      orc_solve(&w[i], o, sums ? &sums[i] : nullptr, 0);
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < n_threads; t++) th.emplace_back(work);
  work();
  for (auto &t : th) t.join();
  return 0;
}

// First LM step from the window's state at the given radius: delta over the oracle's tangent layout
// [15F (+6)(+1) | Np | 4 Nl], plus the reduced camera system in UNSCALED variables
// (S_unscaled = diag(1/s) S diag(1/s)), for parity tests of the Schur stage.
int orc_first_step(const UvsWindow *w, const UvsOptions *o, double radius, double *delta, double *S_out, double *g_out,
                   double *model_change, double *cost, double *scale_out) {
  Problem P(*w, *o);
  State s = P.initial_state();
  Solver S(P);
  *cost = S.linearize(s);
  std::vector<double> d;
  double mc = 0;
  if (!S.compute_step(radius, d, &mc)) return UVS_ERR_NOT_PD;
  if (model_change) *model_change = mc;
  std::memcpy(delta, d.data(), sizeof(double) * d.size());
  const int dd = P.lay.d;
  if (S_out) for (int i = 0; i < dd; i++) for (int j = 0; j < dd; j++) S_out[(size_t)i * dd + j] = S.last_S[(size_t)i * dd + j] / (S.scale[i] * S.scale[j]);
  if (g_out) for (int i = 0; i < dd; i++) g_out[i] = S.last_gS[i] / S.scale[i];
  if (scale_out) std::memcpy(scale_out, S.scale.data(), sizeof(double) * S.scale.size());
  return 0;
}

int orc_marginalize(const UvsWindow *w, const UvsOptions *o, int flag, UvsPrior *out) {
  Problem P(*w, *o);
  State s = P.initial_state();
  MargResult R;
  if (!marginalize(P, s, flag, R)) { out->n = 0; out->n_blocks = 0; out->m = 0; return 0; }
  if (R.n > out->cap_n || (int)R.block_kind.size() > out->cap_blocks) return UVS_ERR_CAPACITY;
  out->n = R.n; out->m = R.m; out->n_blocks = (int)R.block_kind.size();
  std::memcpy(out->J, R.J.data(), sizeof(double) * R.n * R.n);
  std::memcpy(out->r, R.r.data(), sizeof(double) * R.n);
  std::memcpy(out->block_kind, R.block_kind.data(), sizeof(int32_t) * R.block_kind.size());
  std::memcpy(out->block_id, R.block_id.data(), sizeof(int32_t) * R.block_id.size());
  std::memcpy(out->x0, R.x0.data(), sizeof(double) * R.x0.size());
  if (out->A) std::memcpy(out->A, R.A.data(), sizeof(double) * R.n * R.n);
  if (out->b) std::memcpy(out->b, R.b.data(), sizeof(double) * R.n);
  return 0;
}

// the (m + n)-dimensional system A, b the prior is built from (before the Schur complement), for the high-precision
// reference of tests/margref.py.  A_full: [cap x cap] row-major, first (m + n)^2 entries used.
int orc_marginalize_system(const UvsWindow *w, const UvsOptions *o, int flag, double *A_full, double *b_full, int cap, int *m, int *n) {
  Problem P(*w, *o);
  State s = P.initial_state();
  MargResult R;
  *m = 0; *n = 0;
  if (!marginalize(P, s, flag, R)) return 0;
  const int pos = R.m + R.n;
  if (pos > cap) return UVS_ERR_CAPACITY;
  std::memcpy(A_full, R.A_full.data(), sizeof(double) * pos * pos);
  std::memcpy(b_full, R.b_full.data(), sizeof(double) * pos);
  *m = R.m; *n = R.n;
  return 0;
}

// a4: preintegrate n IMU samples (integration_base.h:30-36, 130-158).  noise = {acc_n, gyr_n, acc_w, gyr_w}.
int orc_preintegrate(int n, const double *dt, const double *acc, const double *gyr, const double *acc0, const double *gyr0,
                     const double *ba, const double *bg, const double *noise, double *delta_p, double *delta_q_xyzw,
                     double *delta_v, double *sum_dt, double *jacobian, double *covariance) {
  Preintegration pre(vec_from(acc0), vec_from(gyr0), vec_from(ba), vec_from(bg), noise[0], noise[1], noise[2], noise[3]);
  for (int i = 0; i < n; i++) pre.push_back(dt[i], vec_from(acc + 3 * i), vec_from(gyr + 3 * i));
  for (int k = 0; k < 3; k++) { delta_p[k] = pre.delta_p[k]; delta_v[k] = pre.delta_v[k]; }
  delta_q_xyzw[0] = pre.delta_q.x; delta_q_xyzw[1] = pre.delta_q.y; delta_q_xyzw[2] = pre.delta_q.z; delta_q_xyzw[3] = pre.delta_q.w;
  *sum_dt = pre.sum_dt;
  std::memcpy(jacobian, pre.jacobian, sizeof(pre.jacobian));
  std::memcpy(covariance, pre.covariance, sizeof(pre.covariance));
  return 0;
}

int orc_imu_sqrt_info(const double *cov, double *out) { return imu_sqrt_info(cov, out) ? 0 : UVS_ERR_NOT_PD; }
int orc_pose_plus(const double *x, const double *delta, double *out) { pose_plus(x, delta, out); return 0; }
int orc_cauchy(double a, double s, double *rho) { cauchy_loss(a, s, rho); return 0; }
int orc_sym_eig(int n, const double *A, double *evals, double *V) {
  std::vector<double> a(A, A + (size_t)n * n), e, v;
  sym_eig(n, a, e, v);
  std::memcpy(evals, e.data(), sizeof(double) * n);
  std::memcpy(V, v.data(), sizeof(double) * n * n);
  return 0;
}

}  // extern "C"

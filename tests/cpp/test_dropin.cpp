// test_dropin.cpp — signature-level test of the C++ drop-in classes (uv-slam_b200/host): each Gpu*Factor is
// called exactly like the reference's cost functions (Evaluate(parameters, residuals, jacobians) with
// nullable Jacobian pointers) on the known-answer inputs of SURVEY.md Appendix C, and a small window goes
// through GpuWindowProblem::solve()/marginalize().  Built with g++ against libuvs_b200.so; exit code 0 = ok.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../uv-slam_b200/host/gpu_factors.h"
#include "../../uv-slam_b200/host/optimization_shim.h"

using namespace uvs_host;

static int g_fail = 0;
static void check(bool ok, const char *what) { if (!ok) { std::printf("FAIL %s\n", what); g_fail++; } }
static bool close(double a, double b, double rtol = 1e-9) { return std::fabs(a - b) <= rtol * std::fmax(1.0, std::fabs(b)); }

static void aa(const double axis[3], double ang, double q[4]) {
  const double n = std::sqrt(axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2]), s = std::sin(ang / 2) / n;
  q[0] = axis[0] * s; q[1] = axis[1] * s; q[2] = axis[2] * s; q[3] = std::cos(ang / 2);
}

int main() {
  if (!shared_handle()) { std::printf("no CUDA device: the drop-in classes have no CPU fallback\n"); return 3; }
  // ---- C.1 line + VP ---------------------------------------------------------------------------------
  const double ric[9] = {0.0148655429818, -0.999880929698, 0.00414029679422, 0.999557249008, 0.0149672133247, 0.025715529948,
                         -0.0257744366974, 0.00375618835797, 0.999660727178};
  const double tic[3] = {-0.0216401454975, -0.064676986768, 0.00981073058949};
  double pose[7] = {0.4, -0.3, 1.2, 0, 0, 0, 0};
  { const double ax[3] = {0.2, -0.5, 0.84}; aa(ax, 0.3, pose + 3); }
  double line[4] = {0.35, -0.6, 1.1, 0.45};
  const double sp[2] = {-0.21, 0.13}, ep[2] = {0.32, 0.27}, vp[3] = {0.8, -0.15, 1.0};
  {
    GpuLineProjectionFactor f(ric, tic, sp, ep);
    check(f.num_residuals() == 2 && f.parameter_block_sizes().size() == 2 && f.parameter_block_sizes()[0] == 7, "line sizes");
    const double *params[2] = {pose, line};
    double r[2], Jp[14], Jl[8];
    double *J[2] = {Jp, Jl};
    check(f.Evaluate(params, r, J), "line Evaluate");
    check(close(r[0], 17.293822436935173) && close(r[1], 96.728765415403701), "line residual");
    check(close(Jp[3], -505.14905271817169) && close(Jp[6], -48.270336835244913) && close(Jp[7 + 5], -301.68065117474516), "line pose jacobian (raw quaternion columns)");
    check(close(Jl[0], 161.67443332335478) && close(Jl[7], -113.30779412740195), "line jacobian");
    double r2[2];
    double *Jnull[2] = {nullptr, Jl};
    check(f.Evaluate(params, r2, Jnull) && close(r2[0], r[0]), "line Evaluate with a NULL block");
    check(f.Evaluate(params, r2, nullptr) && close(r2[1], r[1]), "line Evaluate without jacobians");
  }
  {
    GpuVPProjectionFactor f(ric, tic, vp);
    const double *params[2] = {pose, line};
    double r[1], Jp[7], Jl[4];
    double *J[2] = {Jp, Jl};
    check(f.Evaluate(params, r, J), "vp Evaluate");
    check(close(r[0], 13.61897862064066), "vp residual");
    check(close(Jp[3], 14.392456065595809) && close(Jp[6], -1.704136164781298) && Jp[0] == 0.0 && close(Jl[2], 8.3959490942810806), "vp jacobian");
  }
  // ---- C.2 point -----------------------------------------------------------------------------------------
  {
    double pi[7] = {0.1, -0.2, 0.3}, pj[7] = {0.6, 0.1, 0.25}, ex[7] = {-0.02, -0.06, 0.01}, lam[1] = {0.2};
    { const double a1[3] = {0.1, 0.7, -0.2}, a2[3] = {-0.3, 0.5, 0.4}, a3[3] = {0.01, -0.02, 1.0}; aa(a1, 0.25, pi + 3); aa(a2, 0.4, pj + 3); aa(a3, 1.55, ex + 3); }
    const double pts_i[3] = {0.11, -0.07, 1}, pts_j[3] = {-0.05, 0.02, 1};
    GpuProjectionFactor f(pts_i, pts_j);
    const double *params[4] = {pi, pj, ex, lam};
    double r[2], J0[14], J1[14], J2[14], J3[2];
    double *J[4] = {J0, J1, J2, J3};
    check(f.Evaluate(params, r, J), "proj Evaluate");
    check(close(r[0], -38.264559771249328) && close(r[1], 21.82680366598376), "proj residual");
    check(close(J0[3], -295.63248500390809) && J0[6] == 0.0 && close(J1[7 + 4], 290.21541603374538) && close(J2[5], 56.664974431077234) && close(J3[1], 169.08419404677495), "proj jacobians");
  }
  // ---- pose manifold -----------------------------------------------------------------------------------------
  {
    PoseLocalParameterization lp;
    const double d[6] = {0.01, -0.02, 0.03, 0.002, -0.001, 0.004};
    double out[7], Jm[42];
    lp.Plus(pose, d, out); lp.ComputeJacobian(pose, Jm);
    const double n = std::sqrt(out[3] * out[3] + out[4] * out[4] + out[5] * out[5] + out[6] * out[6]);
    check(close(n, 1.0, 1e-14) && close(out[0], 0.41) && Jm[0] == 1.0 && Jm[36] == 0.0 && lp.GlobalSize() == 7 && lp.LocalSize() == 6, "pose manifold");
  }
  // ---- a 2-frame window through the optimization() surface -------------------------------------------------
  {
    double para_Pose[2][7] = {{0, 0, 0, 0, 0, 0, 1}, {0.5, 0.05, 0.0, 0, 0, 0.02, 0.9998}};
    double para_SpeedBias[2][9] = {{0}}, para_Ex[1][7] = {{0, 0, 0, 0, 0, 0, 1}}, para_Td[1] = {0};
    double para_Feature[8][1], para_Ortho[1][4] = {{0, 0, 0, 0}};
    GpuWindowProblem prob(2, para_Pose, para_SpeedBias, para_Ex, para_Td, para_Feature, nullptr);
    (void)para_Ortho;
    // 8 points in front of frame 0, observed from frame 1 which is truly at (0.4, 0, 0)
    for (int k = 0; k < 8; k++) {
      const double X[3] = {-1.0 + 0.3 * k, 0.5 - 0.15 * k, 4.0 + 0.5 * (k % 3)};
      const double pi[3] = {X[0] / X[2], X[1] / X[2], 1.0}, pj[3] = {(X[0] - 0.4) / X[2], X[1] / X[2], 1.0};
      para_Feature[k][0] = 1.0 / X[2] * 1.1;
      prob.addProjection(0, 1, k, pi, pj);
    }
    prob.options().max_num_iterations = 15;
    UvsSummary sm;
    const int rc = prob.solve(&sm);
    check(rc == UVS_OK, "GpuWindowProblem::solve");
    check(sm.final_cost < 1e-3 * sm.initial_cost + 1e-9, "window solve reduces the cost");
    PriorData prior;
    check(prob.marginalize(UVS_MARGIN_OLD, prior) == UVS_OK && prior.n > 0 && prior.J.size() == (size_t)prior.n * prior.n, "GpuWindowProblem::marginalize");
  }
  if (g_fail == 0) std::printf("drop-in OK\n");
  return g_fail ? 1 : 0;
}

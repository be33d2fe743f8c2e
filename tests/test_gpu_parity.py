"""GPU parity tests: libuvs_b200.so (through its C ABI) against the CPU oracle on the same seeded
windows.  Tolerances are the ones BASELINE.json's north_star states: 1e-6 relative on residuals
(and, our addition, on Jacobian blocks), 1e-4 on the solved pose delta."""
import numpy as np
import pytest

import uvs_b200
from tests import orc
from tools import gen_window as gw

pytestmark = pytest.mark.gpu

RTOL = 1e-6          # residuals / Jacobians, relative to the largest magnitude of the block
STEP_TOL = 1e-4      # solved pose delta

KINDS = (("proj", orc.F_PROJ), ("line", orc.F_LINE), ("vp", orc.F_VP), ("imu", orc.F_IMU))


@pytest.fixture(scope="module")
def solver():
    s = uvs_b200.Solver(0)
    yield s
    s.close()


@pytest.fixture(scope="module")
def windows():
    return {k: gw.make_window(k) for k in ("tiny", "C1", "C2")}


def rel_err(a, b):
    scale = max(1.0, float(np.abs(b).max())) if b.size else 1.0
    return float(np.abs(a - b).max()) / scale if b.size else 0.0


def rel_err_rows(a, b):
    """max over factors of |a-b|_inf / max(|b|_inf of that factor, 1e-300) - the 1e-6 *relative* bar"""
    if b.size == 0:
        return 0.0
    scale = np.maximum(np.abs(b).max(axis=1, keepdims=True), 1e-12)
    return float((np.abs(a - b) / scale).max())


@pytest.mark.parametrize("cfg", ["tiny", "C1", "C2"])
@pytest.mark.parametrize("local", [False, True])
def test_factor_sweep_parity(solver, windows, opts, cfg, local):
    w = windows[cfg]
    solver.upload([w], opts)
    for kind, ft in KINDS:
        r, J = solver.eval(kind, local=local)
        r0, J0, _ = orc.eval_factors(w, opts, ft, local=local)
        assert r.shape == r0.shape and J.shape == J0.shape, kind
        assert rel_err_rows(r, r0) < RTOL, (kind, "residual", rel_err_rows(r, r0))
        assert rel_err_rows(J, J0) < RTOL, (kind, "jacobian", rel_err_rows(J, J0))


@pytest.mark.parametrize("local", [False, True])
def test_prior_parity(solver, windows, opts, local):
    w = windows["C1"]
    solver.upload([w], opts)
    rs, Js = solver.eval_prior(local=local)
    r0, J0, _ = orc.eval_factors(w, opts, orc.F_PRIOR, local=local)
    assert rel_err(rs[0], r0.ravel()) < RTOL
    assert rel_err(Js[0], J0.ravel()) < 1e-12   # the prior Jacobian is a copy of J0's columns


def test_extrinsic_and_cost(solver, opts):
    w = gw.make_window("C1", estimate_extrinsic=1)
    solver.upload([w], opts)
    r, J = solver.eval("proj", local=True)
    r0, J0, _ = orc.eval_factors(w, opts, orc.F_PROJ, local=True)
    assert rel_err_rows(J, J0) < RTOL
    c = solver.cost()[0]
    c0 = orc.total_cost(w, opts)
    assert abs(c - c0) <= 1e-9 * abs(c0)


def test_td_factor_parity(solver, opts):
    rng = np.random.default_rng(5)
    w = gw.make_window("C1", with_prior=False)
    n = w.n_proj
    w.estimate_td = 1
    w.td = np.array([0.013])
    w.proj_vel_i = rng.normal(0, 0.3, (n, 2)); w.proj_vel_j = rng.normal(0, 0.3, (n, 2))
    w.proj_td_i = rng.normal(0, 0.005, n); w.proj_td_j = rng.normal(0, 0.005, n)
    w.proj_row_i = rng.uniform(0, 480, n); w.proj_row_j = rng.uniform(0, 480, n)
    w.normalize()
    o = uvs_b200.default_options(tr=0.03, row=480.0)
    solver.upload([w], o)
    for local in (False, True):
        r, J = solver.eval("proj", local=local)
        r0, J0, _ = orc.eval_factors(w, o, orc.F_PROJ, local=local)
        assert J.shape == J0.shape
        assert rel_err_rows(r, r0) < RTOL and rel_err_rows(J, J0) < RTOL
    ref = w.copy()
    sm0 = orc.solve(ref, o)
    sm = solver.solve()[0]
    solver.download()
    assert abs(sm.final_cost - sm0.final_cost) <= 1e-6 * abs(sm0.final_cost)
    assert abs(w.td[0] - ref.td[0]) < STEP_TOL


def _tangent_delta(w0, w1):
    """pose delta between two windows: positions and the quaternion difference (as the north_star's
    'solved pose delta')"""
    dp = np.abs(w1.pose[:, :3] - w0.pose[:, :3]).max()
    dq = np.abs(np.abs((w1.pose[:, 3:] * w0.pose[:, 3:]).sum(axis=1)) - 1.0).max()
    return dp, dq


@pytest.fixture(params=["fused", "record", "default"])
def landmark_path(request, monkeypatch):
    """fused linearisation (uvs_lin.cu) | record path (k_proj / k_line_vp -> k_core_* -> k_direct_fused) | the library's own
    choice by batch size (record path for a single window); read at every upload"""
    monkeypatch.delenv("UVS_NO_FUSE", raising=False)
    if request.param == "record":
        monkeypatch.setenv("UVS_NO_FUSE", "1")
    elif request.param == "default":
        monkeypatch.delenv("UVS_FUSE_MIN", raising=False)
    return request.param


@pytest.mark.parametrize("cfg", ["tiny", "C1", "C2"])
def test_first_step_parity(solver, windows, opts, cfg, landmark_path):
    """one LM iteration (fixed radius 1e4): the GPU candidate equals Plus(x, oracle delta)"""
    w = windows[cfg].copy()
    fs = orc.first_step(w, opts, radius=1e4)
    o = uvs_b200.default_options(max_num_iterations=1, fixed_iterations=1)
    solver.upload([w], o)
    sm = solver.solve()[0]
    assert abs(sm.initial_cost - fs["cost"]) <= 1e-9 * abs(fs["cost"])
    assert sm.step_accepted[1] == 1
    x0 = windows[cfg]
    solver.download()
    d = w.cam_dim
    delta = fs["delta"]
    for f in range(w.n_frames):
        ref = orc.pose_plus(x0.pose[f], delta[15 * f:15 * f + 6])
        assert np.abs(w.pose[f] - ref).max() < STEP_TOL
        assert np.abs(w.speed_bias[f] - (x0.speed_bias[f] + delta[15 * f + 6:15 * f + 15])).max() < STEP_TOL
    scale = np.maximum(np.abs(delta[d:d + w.n_points]), 1e-3)
    assert (np.abs((w.inv_depth - x0.inv_depth) - delta[d:d + w.n_points]) / scale).max() < 1e-4
    dl = delta[d + w.n_points:].reshape(-1, 4)
    assert np.abs((w.ortho - x0.ortho) - dl).max() < 1e-4 * max(1.0, np.abs(dl).max())
    # model cost change through the summary: relative_decrease = (cost - cost+) / model_change
    rel = (sm.initial_cost - sm.cost[1]) / fs["model_change"]
    assert abs(sm.relative_decrease[1] - rel) < 1e-6 * max(1.0, abs(rel))


def _well_conditioned_lines(w, opts, bound=1e8, min_information=1e-3):
    """lines whose landmark block E = sum J_l^T J_l (line + VP factors, loss-corrected, at the state of `w`) has a
    condition number below `bound` and a smallest eigenvalue above `min_information`: their four parameters are
    determined by the data, so they must agree individually.  (What is left out are outlier lines whose residuals the
    Cauchy loss has weighted down to nothing - on C2 one line of 80, smallest eigenvalue 4e-5, where the oracle's own
    Schur and dense solves differ by 1e-3; every other line agrees to 1e-6 there.)"""
    E = np.zeros((w.n_lines, 4, 4))
    _, Jl, _ = orc.eval_factors(w, opts, orc.F_LINE, local=True)
    for k in range(w.n_line_obs):
        J = Jl[k].reshape(2, 10)[:, 6:]
        E[w.line_idx[k]] += J.T @ J
    _, Jv, _ = orc.eval_factors(w, opts, orc.F_VP, local=True)
    for k in range(w.n_vp_obs):
        J = Jv[k].reshape(1, 10)[:, 6:]
        E[w.vp_line[k]] += J.T @ J
    ev = np.linalg.eigvalsh(E)
    return (ev[:, 0] > min_information) & (ev[:, -1] < bound * np.maximum(ev[:, 0], 1e-300))


@pytest.mark.parametrize("cfg", ["tiny", "C1", "C2"])
def test_full_solve_parity(solver, windows, opts, cfg, landmark_path):
    w = windows[cfg].copy()
    ref = windows[cfg].copy()
    sm0 = orc.solve(ref, opts)
    solver.upload([w], opts)
    sm = solver.solve()[0]
    solver.download()
    assert sm.num_iterations == sm0.num_iterations
    assert sm.termination == sm0.termination
    n = sm.num_iterations
    acc = [sm.step_accepted[i] for i in range(n)]
    acc0 = [sm0.step_accepted[i] for i in range(n)]
    assert acc == acc0
    for i in range(n):
        assert abs(sm.cost[i] - sm0.cost[i]) <= 1e-6 * abs(sm0.cost[i]) + 1e-9, (i, sm.cost[i], sm0.cost[i])
        assert abs(sm.radius[i] - sm0.radius[i]) <= 1e-4 * abs(sm0.radius[i])   # radius amplifies the 1e-8 cost noise through (2 rho - 1)^3
    dp, dq = _tangent_delta(ref, w)
    assert dp < STEP_TOL and dq < STEP_TOL
    assert np.abs(w.speed_bias - ref.speed_bias).max() < STEP_TOL
    assert np.abs(w.inv_depth - ref.inv_depth).max() < STEP_TOL
    # line parameters can be weakly observable (the oracle's own Schur and dense paths differ by 1e-3
    # on C2): compare them through what they produce, the cost of the GPU solution under the oracle
    assert abs(orc.total_cost(w, opts) - sm0.final_cost) <= 1e-6 * abs(sm0.final_cost)
    assert np.median(np.abs(w.ortho - ref.ortho)) < STEP_TOL
    # every line whose 4x4 block is well conditioned (cond(J_l^T J_l) < 1e8 at the oracle's solution) individually
    good = _well_conditioned_lines(ref, opts)
    assert good.sum() >= max(1, ref.n_lines // 4), good.sum()
    assert np.abs(w.ortho - ref.ortho)[good].max() < STEP_TOL, np.abs(w.ortho - ref.ortho)[good].max()


def _quat_rot(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _reanchor_at_last_frame(w0, every=3):
    """Every `every`-th point is re-anchored at its LAST observing frame, so its projection factors have
    frame_i > frame_j (the reference never builds these - imu_i is the start frame - but the C ABI takes any pair;
    the direct-term kernel keeps a separate fragment mapping for them).  The inverse depth is recomputed in the new
    anchor's camera (projection_factor.cpp:44-49 run forward), so the window stays geometrically consistent."""
    w = w0.copy()
    Ric, tic = _quat_rot(w.ex_pose[3:]), w.ex_pose[:3]
    for k in range(0, w.n_points, every):
        idx = np.nonzero(w.proj_point == k)[0]
        if idx.size < 2:
            continue
        last = idx[np.argmax(w.proj_frame_j[idx])]
        a_new, a_old = int(w.proj_frame_j[last]), int(w.proj_frame_i[last])
        obs_new, obs_old = w.proj_pts_j[last].copy(), w.proj_pts_i[last].copy()
        Ro, po = _quat_rot(w.pose[a_old, 3:]), w.pose[a_old, :3]
        Rn, pn = _quat_rot(w.pose[a_new, 3:]), w.pose[a_new, :3]
        p_w = Ro @ (Ric @ (obs_old / w.inv_depth[k]) + tic) + po
        p_c = Ric.T @ (Rn.T @ (p_w - pn) - tic)
        w.inv_depth[k] = 1.0 / p_c[2]
        w.proj_frame_i[idx] = a_new
        w.proj_pts_i[idx] = obs_new
        w.proj_frame_j[last] = a_old
        w.proj_pts_j[last] = obs_old
    assert (w.proj_frame_i > w.proj_frame_j).any()
    return w


def test_anchor_frame_after_observing_frame(solver, windows, opts):
    """projection factors with frame_i > frame_j: sweep, first step and the full solve against the oracle"""
    w0 = _reanchor_at_last_frame(windows["C1"])
    solver.upload([w0.copy()], opts)
    r, J = solver.eval("proj", local=True)
    r0, J0, _ = orc.eval_factors(w0, opts, orc.F_PROJ, local=True)
    assert rel_err_rows(r, r0) < RTOL and rel_err_rows(J, J0) < RTOL
    w = w0.copy()
    fs = orc.first_step(w, opts, radius=1e4)
    o = uvs_b200.default_options(max_num_iterations=1, fixed_iterations=1)
    solver.upload([w], o)
    sm = solver.solve()[0]
    assert abs(sm.initial_cost - fs["cost"]) <= 1e-9 * abs(fs["cost"])
    rel = (sm.initial_cost - sm.cost[1]) / fs["model_change"]
    assert abs(sm.relative_decrease[1] - rel) < 1e-6 * max(1.0, abs(rel))
    if sm.step_accepted[1] == 1:
        solver.download()
        delta = fs["delta"]
        for f in range(w.n_frames):
            assert np.abs(w.pose[f] - orc.pose_plus(w0.pose[f], delta[15 * f:15 * f + 6])).max() < STEP_TOL
    w, ref = w0.copy(), w0.copy()
    sm0 = orc.solve(ref, opts)
    solver.upload([w], opts)
    sm = solver.solve()[0]
    solver.download()
    n = sm.num_iterations
    assert n == sm0.num_iterations
    assert [sm.step_accepted[i] for i in range(n)] == [sm0.step_accepted[i] for i in range(n)]
    for i in range(n):
        assert abs(sm.cost[i] - sm0.cost[i]) <= 1e-6 * abs(sm0.cost[i]) + 1e-9, (i, sm.cost[i], sm0.cost[i])
    dp, dq = _tangent_delta(ref, w)
    assert dp < STEP_TOL and dq < STEP_TOL


def test_solve_with_extrinsic(solver, opts):
    w = gw.make_window("C1", estimate_extrinsic=1)
    ref = w.copy()
    sm0 = orc.solve(ref, opts)
    solver.upload([w], opts)
    sm = solver.solve()[0]
    solver.download()
    assert abs(sm.final_cost - sm0.final_cost) <= 1e-6 * abs(sm0.final_cost)
    assert np.abs(w.ex_pose - ref.ex_pose).max() < STEP_TOL
    assert np.abs(w.pose - ref.pose).max() < STEP_TOL


def test_batch_equals_individual(solver, opts):
    """windows of different shapes solved in one batch give the same result as one by one"""
    ws = [gw.make_window("tiny", seed=s) for s in (11, 12)] + [gw.make_window("C1", seed=13), gw.make_window("tiny", seed=14, with_prior=False)]
    single = []
    for w in ws:
        c = w.copy()
        solver.upload([c], opts)
        sm = solver.solve()[0]
        solver.download()
        single.append((c, sm.final_cost, sm.num_iterations))
    batch = [w.copy() for w in ws]
    sums = solver.batch_solve(batch, opts)
    for k, (c, cost, iters) in enumerate(single):
        assert sums[k].num_iterations == iters
        assert abs(sums[k].final_cost - cost) <= 1e-7 * abs(cost)   # FP64 reductions (RED.ADD) are order-dependent
        assert np.abs(batch[k].pose - c.pose).max() < 1e-8
        assert np.median(np.abs(batch[k].ortho - c.ortho)) < 1e-8   # FP64 reductions are order-dependent


def test_converges_from_perturbed_truth(solver, opts):
    """noise-free window perturbed from the truth converges back (SURVEY 8c pin #5)"""
    w, truth = gw.make_window("C1", with_prior=False, return_truth=True)
    solver.upload([w], uvs_b200.default_options(max_num_iterations=30))
    sm = solver.solve()[0]
    assert sm.final_cost < 1e-6 * sm.initial_cost
    n = sm.num_iterations
    costs = [sm.cost[i] for i in range(n)]
    assert all(costs[i + 1] <= costs[i] * (1 + 1e-12) for i in range(n - 1))   # monotone under accepted steps


def test_edge_cases(solver, opts):
    # IMU + prior only (no visual factors)
    w = gw.make_window("tiny")
    bare = uvs_b200.Window(pose=w.pose, speed_bias=w.speed_bias, ex_pose=w.ex_pose, imu_frame_i=w.imu_frame_i,
                           imu_delta_p=w.imu_delta_p, imu_delta_q=w.imu_delta_q, imu_delta_v=w.imu_delta_v,
                           imu_sum_dt=w.imu_sum_dt, imu_lin_ba=w.imu_lin_ba, imu_lin_bg=w.imu_lin_bg,
                           imu_jacobian=w.imu_jacobian, imu_covariance=w.imu_covariance)
    bare.set_prior(w.prior_J, w.prior_r, w.prior_block_kind, w.prior_block_id, w.prior_x0)
    ref = bare.copy()
    sm0 = orc.solve(ref, opts)
    solver.upload([bare], opts)
    sm = solver.solve()[0]
    solver.download()
    assert abs(sm.final_cost - sm0.final_cost) <= 1e-6 * abs(sm0.final_cost) + 1e-9
    assert np.abs(bare.pose - ref.pose).max() < STEP_TOL
    # a landmark without observations keeps its value
    w2 = gw.make_window("tiny")
    w2.inv_depth = np.append(w2.inv_depth, 0.37)
    w2.ortho = np.vstack([w2.ortho, [0.1, 0.2, 0.3, 0.4]])
    w2.normalize()
    solver.upload([w2], opts)
    solver.solve()
    solver.download()
    assert w2.inv_depth[-1] == 0.37 and np.all(w2.ortho[-1] == [0.1, 0.2, 0.3, 0.4])
    # malformed input is rejected, not computed on
    bad = gw.make_window("tiny")
    bad.proj_frame_j = bad.proj_frame_j.copy(); bad.proj_frame_j[0] = 99
    with pytest.raises(uvs_b200.UvsError):
        solver.upload([bad], opts)
    with pytest.raises(uvs_b200.UvsError):
        uvs_b200.Solver(0).solve()   # no window uploaded


def test_stress_window_global_cholesky(solver, opts):
    """C5 (31 frames, d = 465): reduced system too large for shared memory -> blocked Cholesky in a global scratch"""
    w = gw.make_window("C5")
    ref = w.copy()
    sm0 = orc.solve(ref, opts)
    solver.upload([w], opts)
    sm = solver.solve()[0]
    solver.download()
    assert abs(sm.final_cost - sm0.final_cost) <= 1e-6 * abs(sm0.final_cost)
    assert np.abs(w.pose - ref.pose).max() < STEP_TOL


def test_large_window_first_step_and_batch(solver, opts):
    """the blocked global-scratch Cholesky of the large windows (chol_window<0>): one LM step of C5 against the oracle's
    camera step, and a batch of three such windows (every window has its own fragment scratch) against one-by-one solves"""
    w0 = gw.make_window("C5")
    w = w0.copy()
    fs = orc.first_step(w, opts, radius=1e4)
    o = uvs_b200.default_options(max_num_iterations=1, fixed_iterations=1)
    solver.upload([w], o)
    sm = solver.solve()[0]
    solver.download()
    assert abs(sm.initial_cost - fs["cost"]) <= 1e-9 * abs(fs["cost"])
    assert sm.step_accepted[1] == 1
    delta = fs["delta"]
    for f in range(w.n_frames):
        assert np.abs(w.pose[f] - orc.pose_plus(w0.pose[f], delta[15 * f:15 * f + 6])).max() < STEP_TOL
        assert np.abs(w.speed_bias[f] - (w0.speed_bias[f] + delta[15 * f + 6:15 * f + 15])).max() < STEP_TOL
    rng = np.random.default_rng(5)
    batch = [w0.copy() for _ in range(3)]
    for b in batch[1:]:
        b.pose[:, :3] += rng.normal(0, 0.01, b.pose[:, :3].shape)
        b.inv_depth *= 1.0 + rng.normal(0, 0.02, b.n_points)
    single = []
    for b in batch:
        c = b.copy()
        solver.upload([c], opts)
        single.append((solver.solve()[0].final_cost, c))
        solver.download()
    solver.upload(batch, opts)
    sms = solver.solve()
    solver.download()
    for k in range(3):
        assert abs(sms[k].final_cost - single[k][0]) <= 1e-7 * abs(single[k][0])
        assert np.abs(batch[k].pose - single[k][1].pose).max() < 1e-6


@pytest.mark.parametrize("cfg,flag", [("C1", 0), ("C2", 0), ("tiny", 0), ("tiny", 1)])
def test_marginalization_parity(solver, opts, cfg, flag):
    """next prior built on the GPU (after the solve, as estimator.cpp:994-1003) vs the oracle: same kept
    blocks / linearisation points, A' and b' equal, and the identities J0^T J0 = A', J0^T r0 = b'
    (marginalization_factor.cpp:295-296).  J0 itself is only defined up to an orthogonal transform."""
    w = gw.make_window(cfg)
    solver.upload([w], opts)
    solver.solve()
    solver.download()
    g = solver.marginalize(0, flag)
    m = orc.marginalize(w, opts, flag)      # oracle on the SAME (GPU-solved) state
    assert (g is None) == (m is None)
    if m is None:
        return
    assert g["n"] == m["n"] and g["m"] == m["m"]
    assert np.array_equal(g["kinds"], m["kinds"]) and np.array_equal(g["ids"], m["ids"])
    assert np.allclose(g["x0"], m["x0"], atol=1e-12)
    sA, sb = np.abs(m["A"]).max(), max(1.0, np.abs(m["b"]).max())
    # A' = Arr - Arm Amm^+ Amr with eigenvalues <= 1e-8 zeroed: |Amm| reaches 1e10 (IMU information), so
    # eigenvalues near the cut-off are resolved only to ~1e-6 absolute by ANY FP64 eigensolver and the
    # weakly observed landmark directions amplify that; the oracle's own A' is asymmetric at 5e-4
    # relative in such cases (tests/test_oracle.py).  Prior-only marginalization (flag 1) has no such
    # directions and must agree to 1e-6; MARGIN_OLD is held to 2e-3 here and to the 1e-4 pose-delta bar
    # through the next solve below.
    tolA = 1e-6 if flag == 1 else 2e-3
    assert np.abs(g["A"] - m["A"]).max() < tolA * sA, np.abs(g["A"] - m["A"]).max() / sA
    assert np.abs(g["b"] - m["b"]).max() < tolA * sb, np.abs(g["b"] - m["b"]).max() / sb
    J, r = g["J"], g["r"]
    As = np.tril(g["A"]) + np.tril(g["A"], -1).T
    assert np.abs(J.T @ J - As).max() < 1e-7 * sA
    assert np.abs(J.T @ r - g["b"]).max() < 1e-6 * sb
    # the prior is usable: a window carrying it evaluates to the same prior residual on GPU and oracle
    if flag == 0:
        w2, _ = gw.drop_first_frame(w, dict(Rs=[None] * w.n_frames, Ps=[None] * w.n_frames, Vs=[None] * w.n_frames, ba=0, bg=0,
                                            inv_depth=w.inv_depth.copy(), ortho=w.ortho.copy(),
                                            Rwc=[np.eye(3)] * w.n_frames, twc=[np.zeros(3)] * w.n_frames), np.random.default_rng(0))
        w2.set_prior(J, r, g["kinds"], g["ids"], g["x0"])
        w2_init = w2.copy()
        solver.upload([w2], opts)
        rs, _ = solver.eval_prior(local=True)
        r0, _, _ = orc.eval_factors(w2, opts, orc.F_PRIOR, local=True)
        assert np.abs(rs[0] - r0.ravel()).max() < 1e-6 * max(1.0, np.abs(r0).max())
        sm = solver.solve()[0]
        solver.download()
        assert sm.final_cost <= sm.initial_cost
        # functional parity: the same next window solved by the oracle with the ORACLE's prior
        w3 = w2.copy()
        w3.pose, w3.speed_bias, w3.inv_depth, w3.ortho = (a.copy() for a in (w2_init.pose, w2_init.speed_bias, w2_init.inv_depth, w2_init.ortho))
        w3.set_prior(m["J"], m["r"], m["kinds"], m["ids"], m["x0"])
        sm3 = orc.solve(w3, opts)
        if cfg == "tiny":
            # 4 frames / a dozen points after the slide: the problem is held together by the prior's weakest
            # eigenvalues only (the ones the 1e-8 cut-off decides), so the next solve is not a meaningful
            # parity probe here; the A', b' and identity checks above still apply
            assert sm.final_cost < sm.initial_cost and sm3.final_cost < sm3.initial_cost
        else:
            assert np.abs(w2.pose - w3.pose).max() < STEP_TOL, np.abs(w2.pose - w3.pose).max()


def test_marginalization_of_every_window_of_a_batch(solver, windows, opts):
    """uvs_marginalize per window of a solved batch: the factor sweep (whole batch) runs once, the later calls reuse its
    records; every prior equals the one of the same window marginalized on its own."""
    batch = [windows["C1"].copy(), windows["tiny"].copy(), windows["C1"].copy()]
    batch[2].pose[:, :3] += np.random.default_rng(3).normal(0, 0.01, batch[2].pose[:, :3].shape)
    solver.upload(batch, opts)
    solver.solve()
    solver.download()
    counts, priors = [], []
    for i in range(3):
        l0 = solver.launch_count()
        priors.append(solver.marginalize(i, 0))
        counts.append(solver.launch_count() - l0)
    assert counts[1] < counts[0] and counts[2] < counts[0]      # no second sweep
    s2 = uvs_b200.Solver(0)
    try:
        for i in range(3):
            s2.upload([batch[i].copy()], opts)                   # the solved state, alone
            m = s2.marginalize(0, 0)
            g = priors[i]
            assert g["n"] == m["n"] and np.array_equal(g["kinds"], m["kinds"]) and np.array_equal(g["ids"], m["ids"])
            assert np.allclose(g["x0"], m["x0"], atol=1e-12)
            sA, sb = np.abs(m["A"]).max(), max(1.0, np.abs(m["b"]).max())
            assert np.abs(g["A"] - m["A"]).max() < 5e-6 * sA and np.abs(g["b"] - m["b"]).max() < 5e-6 * sb   # both within 1e-6 of the exact rule
    finally:
        s2.close()
    solver.solve()                                               # anything that moves the state invalidates the records
    l0 = solver.launch_count()
    solver.marginalize(0, 0)
    assert solver.launch_count() - l0 == counts[0]


def test_marginalization_on_the_repacked_state(solver, opts):
    """The reference marginalizes on the state AFTER double2vector() / vector2double() (yaw / position gauge fix,
    estimator.cpp:999-1006 and :1168), not on the raw solver output.  Shift the gauge on the host between solve and
    marginalize, hand the state back with uvs_upload_state (what GpuWindowProblem::marginalize does): linearisation
    points, A' and b' follow the shifted state (oracle on the same state) and differ from the un-shifted ones."""
    w = gw.make_window("C1")
    solver.upload([w], opts)
    solver.solve()
    solver.download()
    g_raw = solver.marginalize(0, 0)
    yaw, t = 0.02, np.array([0.3, -0.2, 0.05])
    Rz = np.array([[np.cos(yaw), -np.sin(yaw), 0.0], [np.sin(yaw), np.cos(yaw), 0.0], [0.0, 0.0, 1.0]])
    w.pose[:, :3] = w.pose[:, :3] @ Rz.T + t
    x, y, z, s_ = (w.pose[:, 3 + k].copy() for k in range(4))
    a, b = np.sin(yaw / 2), np.cos(yaw / 2)          # q' = (0, 0, a, b) * q   (x, y, z, w)
    w.pose[:, 3] = b * x - a * y
    w.pose[:, 4] = b * y + a * x
    w.pose[:, 5] = b * z + a * s_
    w.pose[:, 6] = b * s_ - a * z
    w.speed_bias[:, :3] = w.speed_bias[:, :3] @ Rz.T
    solver.upload_state()
    g = solver.marginalize(0, 0)
    m = orc.marginalize(w, opts, 0)
    assert g is not None and m is not None and g["n"] == m["n"]
    assert np.array_equal(g["kinds"], m["kinds"]) and np.array_equal(g["ids"], m["ids"])
    assert np.allclose(g["x0"], m["x0"], atol=1e-12)
    assert np.abs(g["x0"] - g_raw["x0"]).max() > 1e-2          # the prior really moved with the state
    sA, sb = np.abs(m["A"]).max(), max(1.0, np.abs(m["b"]).max())
    assert np.abs(g["A"] - m["A"]).max() < 2e-3 * sA            # same bar as test_marginalization_parity (MARGIN_OLD vs the oracle)
    assert np.abs(g["b"] - m["b"]).max() < 2e-3 * sb


@pytest.mark.parametrize("cfg,flag", [("tiny", 0), ("tiny", 1), ("C1", 0), ("C2", 0)])
def test_marginalization_against_exact_rule(solver, opts, cfg, flag):
    """GPU A', b' at the fixture's solved state against the reference's rule evaluated with mpmath at 40 digits
    (tests/golden/marg_*.npz).  This settles the loosened MARGIN_OLD tolerance of test_marginalization_parity (2e-3 between
    GPU and CPU oracle): the difference is the CPU restatement's, not the device's."""
    from tests.test_oracle import _marg_fixture
    w, z = _marg_fixture(cfg, flag)
    solver.upload([w], opts)
    g = solver.marginalize(0, flag)
    sA, sb = np.abs(z["A_exact"]).max(), max(1.0, np.abs(z["b_exact"]).max())
    As = np.tril(g["A"]) + np.tril(g["A"], -1).T
    eg, eo = np.abs(As - z["A_exact"]).max() / sA, np.abs(z["A_oracle"] - z["A_exact"]).max() / sA
    bg, bo = np.abs(g["b"] - z["b_exact"]).max() / sb, np.abs(z["b_oracle"] - z["b_exact"]).max() / sb
    print("marg %s flag %d: GPU vs exact A %.2e b %.2e | oracle vs exact A %.2e b %.2e" % (cfg, flag, eg, bg, eo, bo))
    # measured on B200: the device's cyclic Jacobi rotations resolve the small eigenvalues of Amm to full relative accuracy -
    # 3e-7 of max|A'| at worst against the exact rule, where the QL-based CPU restatement is at 2e-5 .. 2e-4.  So the GPU is
    # held to the north_star's 1e-6 against the EXACT value, and never behind the CPU restatement
    assert eg <= 1e-6 and bg <= 2e-6, (eg, bg)
    assert eg <= max(4 * eo, 1e-9) and bg <= max(4 * bo, 1e-9), (eg, eo, bg, bo)


def test_preintegration_on_device(solver):
    """IntegrationBase mid-point preintegration (integration_base.h:30-158) for many intervals in one launch,
    against the oracle's restatement and the independent numpy one of the generator"""
    rng = np.random.default_rng(11)
    n, per = 7, [20, 20, 13, 1, 40, 20, 5]
    off = np.concatenate([[0], np.cumsum(per)]).astype(np.int32)
    S = int(off[-1])
    dt = rng.uniform(0.004, 0.006, S)
    acc = rng.normal(0, 1.0, (S, 3)) + [0, 0, 9.8]
    gyr = rng.normal(0, 0.3, (S, 3))
    acc0 = rng.normal(0, 1.0, (n, 3)) + [0, 0, 9.8]; gyr0 = rng.normal(0, 0.3, (n, 3))
    ba = rng.normal(0, 0.02, (n, 3)); bg = rng.normal(0, 0.002, (n, 3))
    noise = [gw.ACC_N, gw.GYR_N, gw.ACC_W, gw.GYR_W]
    out = solver.preintegrate(off, dt, acc, gyr, acc0, gyr0, ba, bg, noise)
    for k in range(n):
        sl = slice(off[k], off[k + 1])
        ref = orc.preintegrate(dt[sl], acc[sl], gyr[sl], acc0[k], gyr0[k], ba[k], bg[k], noise)
        for name in ("delta_p", "delta_q", "delta_v"):
            assert np.allclose(out[name][k], ref[name], rtol=1e-12, atol=1e-14), (k, name)
        assert abs(out["sum_dt"][k] - ref["sum_dt"]) < 1e-14
        assert np.allclose(out["jacobian"][k].reshape(15, 15), ref["jacobian"], rtol=1e-10, atol=1e-16), k
        assert np.allclose(out["covariance"][k].reshape(15, 15), ref["covariance"], rtol=1e-10, atol=1e-24), k
    py = gw.preintegrate(dt[:20], acc[:20], gyr[:20], acc0[0], gyr0[0], ba[0], bg[0])
    assert np.allclose(out["covariance"][0].reshape(15, 15), py["covariance"], rtol=1e-9, atol=1e-24)


def test_batch_solve_pipelined_matches_plain(solver, windows):
    """uvs_batch_solve_pipelined cuts the batch into sub-batches on their own streams: same answer per window"""
    base = [windows[k] for k in ("C1", "tiny", "C2")]
    ws_a = [base[i % 3].copy() for i in range(9)]
    ws_b = [w.copy() for w in ws_a]
    opts = uvs_b200.default_options(max_num_iterations=6)
    sa = solver.batch_solve(ws_a, opts)
    sb = solver.batch_solve(ws_b, opts, groups=3)
    for i, (a, b) in enumerate(zip(ws_a, ws_b)):
        assert abs(sa[i].final_cost - sb[i].final_cost) <= 1e-9 * max(1.0, abs(sa[i].final_cost)), i
        assert sa[i].num_iterations == sb[i].num_iterations
        assert np.allclose(a.pose, b.pose, rtol=0, atol=1e-8), i
    # one-shot: the handle keeps no batch afterwards
    with pytest.raises(uvs_b200.UvsError):
        solver.cost()


def test_large_batch_concurrent_path_matches_oracle_and_single(solver, windows, opts):
    """Batches of >= 32 windows run the independent kernels of a stage side by side on auxiliary streams (fork / join
    events, all reduced-system updates as FP64 reductions).  A mixed batch of 48 windows (C2 / C1 / tiny, exact replicas)
    must reproduce the oracle's solve of each base window (1e-6 on the cost, 1e-4 on the poses) and agree with the plain
    in-order single-window solve; replicas must agree with each other (reductions are order-dependent: 1e-7)."""
    names = ("C2", "C1", "tiny")
    ref, single = {}, {}
    for k in names:
        r = windows[k].copy()
        ref[k] = (orc.solve(r, opts), r)
        c = windows[k].copy()
        solver.upload([c], opts)
        single[k] = (solver.solve()[0], c)
        solver.download()
    batch = [windows[names[i % 3]].copy() for i in range(48)]
    sums = solver.batch_solve(batch, opts)
    for i, w in enumerate(batch):
        k = names[i % 3]
        sm0, r = ref[k]
        sm1, c = single[k]
        assert sums[i].num_iterations == sm0.num_iterations == sm1.num_iterations, (i, k)
        assert abs(sums[i].final_cost - sm0.final_cost) <= 1e-6 * abs(sm0.final_cost), (i, k)
        assert np.abs(w.pose - r.pose).max() < STEP_TOL and np.abs(w.speed_bias - r.speed_bias).max() < STEP_TOL, (i, k)
        assert abs(sums[i].final_cost - sm1.final_cost) <= 1e-7 * abs(sm1.final_cost), (i, k)
        assert np.abs(w.pose - c.pose).max() < 1e-6, (i, k)
        if i >= 3:
            assert abs(sums[i].final_cost - sums[i - 3].final_cost) <= 1e-7 * abs(sums[i].final_cost), (i, k)


def test_rejected_steps_keep_the_system_complete(solver, windows):
    """A huge initial trust region (Gauss-Newton-like steps) makes the solver overshoot and reject steps (the oracle
    rejects five in a row here): the records of the last linearisation are reused while the reduced system is rebuilt
    every iteration (IMU / prior blocks included) - the iteration log must match the oracle."""
    w = windows["C1"].copy()
    r = w.copy()
    o = uvs_b200.default_options(max_num_iterations=12)
    o.initial_radius = 1e9
    sm0 = orc.solve(r, o)
    assert 0 in list(sm0.step_accepted[:sm0.num_iterations])
    big = [w.copy() for _ in range(33)]   # >= 32: concurrent path
    sums = solver.batch_solve(big, o)
    for i in (0, 16, 32):
        assert sums[i].num_iterations == sm0.num_iterations
        assert list(sums[i].step_accepted[:sm0.num_iterations]) == list(sm0.step_accepted[:sm0.num_iterations])
        # with next to no damping the weakly observable line parameters move by O(1) per step and differ by 1e-3 between
        # any two solvers (measured: poses agree to 3e-8, the cost to 4e-6): the cost bar is 2e-5 here, the pose bar stays
        assert abs(sums[i].final_cost - sm0.final_cost) <= 2e-5 * abs(sm0.final_cost)
        assert np.abs(big[i].pose - r.pose).max() < STEP_TOL


def test_10k_window_against_the_oracle(solver):
    """The north_star's 10 k-factor window (11 frames / 1500 points / 500 lines; tests/golden/window_10k.uvsw) on ONE GPU
    against the oracle: same accept sequence, per-iteration cost to 1e-6 while the iteration is well conditioned
    (through iteration 7: the cost falls from 2.5e10 to 1975), final pose delta to 1e-4.  The last iterations move the
    weakly observable line parameters by O(1): two GPU solves of the same upload differ there by 1e-4 in the cost
    (FP64 reductions are order-dependent), so the final cost is held to 5e-4 and checked under the oracle's own
    cost function."""
    import os
    w0 = uvs_b200.Window.load(os.path.join(os.path.dirname(__file__), "golden", "window_10k.uvsw"))
    opts = uvs_b200.default_options(max_num_iterations=10)
    ref = w0.copy()
    sm0 = orc.solve(ref, opts)
    w = w0.copy()
    solver.upload([w], opts)
    sm = solver.solve()[0]
    solver.download()
    n = sm.num_iterations
    assert n == sm0.num_iterations
    assert [sm.step_accepted[i] for i in range(n)] == [sm0.step_accepted[i] for i in range(n)]
    for i in range(min(n, 8)):
        assert abs(sm.cost[i] - sm0.cost[i]) <= 1e-6 * abs(sm0.cost[i]), (i, sm.cost[i], sm0.cost[i])
    assert abs(sm.final_cost - sm0.final_cost) <= 5e-4 * abs(sm0.final_cost)
    assert abs(orc.total_cost(w, opts) - sm.final_cost) <= 1e-9 * abs(sm.final_cost)   # the GPU's cost of its own solution is the oracle's
    dp, dq = _tangent_delta(ref, w)
    assert dp < STEP_TOL and dq < STEP_TOL
    assert np.abs(w.speed_bias - ref.speed_bias).max() < STEP_TOL


def test_fused_path_equals_record_path(windows, opts, monkeypatch):
    """The fused linearisation (uvs_lin.cu: factors evaluated inside the landmark elimination, no Jacobian records in HBM)
    against the record path (k_proj / k_line_vp -> k_core_* -> k_direct_fused) on the same mixed batch: same iteration
    log, cost to 1e-7 (FP64 reductions are order-dependent; measured 5e-9 after eight iterations), states to 1e-6.  UVS_NO_FUSE is read at every upload."""
    s = uvs_b200.Solver(0)
    try:
        names = ("C2", "C1", "tiny")
        a = [windows[names[i % 3]].copy() for i in range(9)]
        b = [w.copy() for w in a]
        monkeypatch.delenv("UVS_NO_FUSE", raising=False)
        s.upload(a, opts)
        l0 = s.launch_count()
        sa = s.solve()
        fused_launches = s.launch_count() - l0
        s.download()
        monkeypatch.setenv("UVS_NO_FUSE", "1")
        s.upload(b, opts)
        l0 = s.launch_count()
        sb = s.solve()
        record_launches = s.launch_count() - l0
        s.download()
        assert fused_launches < record_launches   # the Jacobian-mode k_proj / k_line_vp / k_core_* / k_direct launches are gone
        for i, (x, y) in enumerate(zip(a, b)):
            n = sa[i].num_iterations
            assert n == sb[i].num_iterations, i
            assert [sa[i].step_accepted[k] for k in range(n)] == [sb[i].step_accepted[k] for k in range(n)], i
            for k in range(n):
                assert abs(sa[i].cost[k] - sb[i].cost[k]) <= 1e-7 * abs(sb[i].cost[k]) + 1e-12, (i, k)
            assert np.abs(x.pose - y.pose).max() < 1e-6 and np.abs(x.speed_bias - y.speed_bias).max() < 1e-6, i
            assert np.abs(x.inv_depth - y.inv_depth).max() < 1e-6, i
    finally:
        s.close()


def test_materialised_sweep_timer(solver, windows, opts):
    """uvs_jacobian_sweep (the measurement entry point of bench.py) runs the four Jacobian kernels into the record arrays"""
    solver.upload([windows["C2"].copy() for _ in range(4)], opts)
    g, each = solver.jacobian_sweep(repeats=3)
    assert g > 0 and all(e > 0 for e in each)


def test_graph_replay_gives_the_same_solve(windows, opts):
    """uvs_set_graph_replay: the LM iteration replayed from a CUDA graph (captured at the first solve after an upload,
    reused by later solves of the same upload) must give what the plain launches give."""
    s = uvs_b200.Solver(0)
    try:
        batch = [windows["C2"].copy(), windows["C1"].copy(), windows["tiny"].copy()]
        s.upload(batch, opts)
        plain = s.solve()
        s.set_graph_replay(True)
        for _ in range(2):   # first solve captures, second one replays from iteration 0 on
            s.reset_state()
            g = s.solve()
            for a, b in zip(plain, g):
                assert a.num_iterations == b.num_iterations
                assert abs(a.final_cost - b.final_cost) <= 1e-7 * abs(a.final_cost)   # FP64 reductions are order-dependent
    finally:
        s.close()


def test_relocalisation_factors(solver, opts):
    """Relocalisation factors and the relo_Pose block (estimator.cpp:944-978; UvsWindow.n_relo / relo_*) against the oracle:
    same accept sequence, per-iteration cost to 1e-6, poses and relo_Pose to 1e-4; the marginalization built afterwards
    ignores them, exactly like the reference's (estimator.cpp:1003-1228)."""
    w0, truth = gw.make_window("C1", return_truth=True)
    w0 = gw.add_relocalisation(w0, truth, frame=3, n_match=15)
    ref, w = w0.copy(), w0.copy()
    sm0 = orc.solve(ref, opts)
    solver.upload([w], opts)
    assert abs(solver.cost()[0] - orc.total_cost(w0, opts)) <= 1e-9 * orc.total_cost(w0, opts)
    r, _ = solver.eval("proj", local=True)
    assert r.shape[0] == w0.n_proj + w0.n_relo           # the relocalisation factors are evaluated with their points' groups
    sm = solver.solve()[0]
    solver.download()
    n = sm.num_iterations
    assert n == sm0.num_iterations
    assert [sm.step_accepted[i] for i in range(n)] == [sm0.step_accepted[i] for i in range(n)]
    for i in range(n):
        assert abs(sm.cost[i] - sm0.cost[i]) <= 1e-6 * abs(sm0.cost[i]), (i, sm.cost[i], sm0.cost[i])
    dp, dq = _tangent_delta(ref, w)
    assert dp < STEP_TOL and dq < STEP_TOL
    assert np.abs(w.relo_pose - ref.relo_pose).max() < STEP_TOL
    assert np.abs(w.relo_pose - w0.relo_pose).max() > 1e-3        # and it moved
    assert np.abs(w.speed_bias - ref.speed_bias).max() < STEP_TOL and np.abs(w.inv_depth - ref.inv_depth).max() < STEP_TOL
    # marginalization: neither the block nor its factors take part
    g = solver.marginalize(0, 0)
    m = orc.marginalize(w, opts, 0)
    assert g["n"] == m["n"] and np.array_equal(g["kinds"], m["kinds"]) and np.array_equal(g["ids"], m["ids"])
    assert np.abs(g["A"] - m["A"]).max() < 2e-3 * np.abs(m["A"]).max()
    # a batch mixing windows with and without relocalisation
    a, b = w0.copy(), gw.make_window("C1")
    sums = solver.batch_solve([a, b], opts)
    assert abs(sums[0].final_cost - sm0.final_cost) <= 1e-6 * abs(sm0.final_cost)
    assert np.abs(a.relo_pose - ref.relo_pose).max() < STEP_TOL
    # argument checks: a matched point needs a projection factor; no td
    bad = w0.copy()
    bad.relo_point = np.array([w0.n_points - 1, w0.n_points], np.int32); bad.relo_pts_j = np.zeros((2, 3))
    with pytest.raises(uvs_b200.UvsError):
        solver.upload([bad], opts)

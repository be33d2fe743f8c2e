"""Triangulation: the numpy oracle (oracle/triangulate.py, a restatement of
feature_manager.cpp:427-589, 827-902) against exact synthetic geometry on the CPU, and the CUDA kernels
(uvs_triangulate_points / uvs_triangulate_lines / uvs_validate_lines) against the oracle on the GPU."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import triangulate as tri  # noqa: E402


def rot(axis, ang):
    axis = np.asarray(axis, float) / np.linalg.norm(axis)
    K = tri.skew(axis)
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K


def make_scene(seed, n_frames=11, n_tracks=200, n_lines=80, noise=0.0):
    rng = np.random.default_rng(seed)
    ric = rot([0.01, -0.02, 1.0], 1.55)          # EuRoC-like camera-IMU rotation
    tic = np.array([-0.02, -0.06, 0.01])
    Rs, Ps = [], []
    for f in range(n_frames):
        Rs.append(rot([0.2, 1.0, 0.1], 0.04 * f) @ rot([1, 0, 0], 0.02 * np.sin(f)))
        Ps.append(np.array([0.25 * f, 0.05 * np.sin(0.7 * f), 0.03 * f]))
    Rs, Ps = np.array(Rs), np.array(Ps)

    def cam(f):
        return Rs[f] @ ric, Ps[f] + Rs[f] @ tic

    def project(f, Xw):
        R, t = cam(f)
        x = R.T @ (Xw - t)
        return x / x[2]

    start, off, pts, depth = [], [0], [], []
    for _ in range(n_tracks):
        s = int(rng.integers(0, n_frames - 3))
        n = int(rng.integers(2, n_frames - s + 1))
        d = rng.uniform(2.0, 10.0)
        uv = np.array([rng.uniform(-0.4, 0.4), rng.uniform(-0.3, 0.3), 1.0])
        R, t = cam(s)
        Xw = R @ (uv * d) + t
        for k in range(n):
            p = project(s + k, Xw)
            p[:2] += rng.normal(0, noise, 2)
            pts.append(p)
        start.append(s); off.append(off[-1] + n); depth.append(d)
    lines = dict(first=[], last=[], sp0=[], ep0=[], sp1=[], ep1=[], nw=[], dw=[], a=[], b=[])
    for _ in range(n_lines):
        s = int(rng.integers(0, n_frames - 5))
        e = int(rng.integers(s + 2, n_frames))
        R, t = cam(s)
        a = R @ (np.array([rng.uniform(-0.3, 0.3), rng.uniform(-0.2, 0.2), 1.0]) * rng.uniform(3, 8)) + t
        dvec = rng.normal(size=3); dvec /= np.linalg.norm(dvec)
        b = a + dvec * rng.uniform(0.5, 2.0)
        # end points slide along the line from frame to frame (the detector's end points are not stable)
        a1, b1 = a + dvec * rng.uniform(-0.1, 0.1), b + dvec * rng.uniform(-0.1, 0.1)
        lines["first"].append(s); lines["last"].append(e)
        lines["sp0"].append(project(s, a)); lines["ep0"].append(project(s, b))
        lines["sp1"].append(project(e, a1)); lines["ep1"].append(project(e, b1))
        lines["nw"].append(np.cross(a, b)); lines["dw"].append(b - a)
        lines["a"].append(a); lines["b"].append(b)
    return dict(Rs=Rs, Ps=Ps, ric=ric, tic=tic, start=np.array(start, np.int32), off=np.array(off, np.int32), pts=np.array(pts),
                depth=np.array(depth), **{k: np.array(v) for k, v in lines.items()})


def oracle_points(sc):
    return np.array([tri.triangulate_point(sc["Rs"], sc["Ps"], sc["ric"], sc["tic"], int(sc["start"][t]),
                                           sc["pts"][sc["off"][t]:sc["off"][t + 1]]) for t in range(len(sc["start"]))])


def oracle_lines(sc):
    return np.array([tri.triangulate_line(sc["Rs"], sc["Ps"], sc["ric"], sc["tic"], int(sc["first"][t]), int(sc["last"][t]),
                                          sc["sp0"][t], sc["ep0"][t], sc["sp1"][t], sc["ep1"][t]) for t in range(len(sc["first"]))])


def ortho_to_plucker(o):
    """n_w ~ cos(phi) U[:,0], d_w ~ sin(phi) U[:,1], U = Rx Ry Rz (line_projection_factor.h:23-39)"""
    U = rot([1, 0, 0], o[0]) @ rot([0, 1, 0], o[1]) @ rot([0, 0, 1], o[2])
    return np.cos(o[3]) * U[:, 0], np.sin(o[3]) * U[:, 1]


# ---- CPU: the oracle on exact geometry ------------------------------------------------------------------------
def test_oracle_point_depth_is_exact_on_noise_free_tracks():
    sc = make_scene(1)
    d = oracle_points(sc)
    assert np.max(np.abs(d - sc["depth"]) / sc["depth"]) < 1e-9


def test_oracle_depth_below_threshold_falls_back_to_init_depth():
    sc = make_scene(2, n_tracks=4)
    pts = sc["pts"][sc["off"][0]:sc["off"][1]].copy()
    # a point 5 cm in front of the camera triangulates to < 0.1 -> INIT_DEPTH (feature_manager.cpp:474-477)
    R0, t0 = sc["Rs"][sc["start"][0]] @ sc["ric"], sc["Ps"][sc["start"][0]] + sc["Rs"][sc["start"][0]] @ sc["tic"]
    Xw = R0 @ np.array([0.0, 0.0, 0.05]) + t0
    for k in range(len(pts)):
        f = int(sc["start"][0]) + k
        R, t = sc["Rs"][f] @ sc["ric"], sc["Ps"][f] + sc["Rs"][f] @ sc["tic"]
        x = R.T @ (Xw - t)
        pts[k] = x / x[2]
    assert tri.triangulate_point(sc["Rs"], sc["Ps"], sc["ric"], sc["tic"], int(sc["start"][0]), pts, init_depth=5.0) == 5.0


def test_oracle_euler_angles_rebuild_the_rotation():
    rng = np.random.default_rng(3)
    for _ in range(50):
        m = rot(rng.normal(size=3), rng.uniform(-3, 3))
        a = tri.euler_angles_012(m)
        assert np.allclose(rot([1, 0, 0], a[0]) @ rot([0, 1, 0], a[1]) @ rot([0, 0, 1], a[2]), m, atol=1e-12)
        assert 0.0 <= a[0] <= np.pi + 1e-12   # Eigen's range for the first angle


def test_oracle_line_recovers_the_world_pluecker_line():
    sc = make_scene(4)
    o = oracle_lines(sc)
    for t in range(len(o)):
        n, d = ortho_to_plucker(o[t])
        nw, dw = sc["nw"][t], sc["dw"][t]
        s = np.linalg.norm(np.concatenate([nw, dw]))
        # the same line up to the common scale and sign of the homogeneous Pluecker coordinates
        v, w = np.concatenate([n, d]), np.concatenate([nw, dw]) / s
        assert min(np.linalg.norm(v - w), np.linalg.norm(v + w)) < 1e-8


# ---- GPU: kernels against the oracle -----------------------------------------------------------------------------
def behind_camera_copy(sc):
    """the same scene seen by cameras turned by 180 degrees about their y axis (R_wc' = R_wc diag(-1, 1, -1)): every line
    lies behind its first camera; the 'observations' are the central projections x / z of the end points (z < 0)"""
    flip = sc["ric"] @ np.diag([-1.0, 1.0, -1.0]) @ sc["ric"].T
    out = dict(sc)
    out["Rs"] = np.array([R @ flip for R in sc["Rs"]])
    out["Ps"] = np.array([P + R @ sc["tic"] - R2 @ sc["tic"] for P, R, R2 in zip(sc["Ps"], sc["Rs"], out["Rs"])])   # same camera centres
    sp0, ep0 = [], []
    for t in range(len(sc["first"])):
        f = int(sc["first"][t])
        R, c = out["Rs"][f] @ sc["ric"], out["Ps"][f] + out["Rs"][f] @ sc["tic"]
        xa, xb = R.T @ (sc["a"][t] - c), R.T @ (sc["b"][t] - c)
        sp0.append(xa / xa[2]); ep0.append(xb / xb[2])
    out["sp0"], out["ep0"] = np.array(sp0), np.array(ep0)
    return out


def oracle_flags(sc, ortho):
    res = [tri.line_solve_flag(sc["Rs"], sc["Ps"], sc["ric"], sc["tic"], int(sc["first"][t]), ortho[t], sc["sp0"][t], sc["ep0"][t])
           for t in range(len(sc["first"]))]
    return np.array([r[0] for r in res], np.int32), np.array([np.concatenate([r[1], r[2]]) for r in res])


def test_oracle_line_validity_recovers_the_end_points():
    """setLineOrtho's test (feature_manager.cpp:333-423): with the exact line the 3-D end points it computes are the world
    points that project onto the first observation's end points - in front of the camera flag 1, behind it flag 2"""
    sc = make_scene(31, n_lines=40)
    ortho = oracle_lines(sc)
    flag, ends = oracle_flags(sc, ortho)
    assert (flag == 1).all()
    assert np.abs(ends[:, :3] - sc["a"]).max() < 1e-6 and np.abs(ends[:, 3:] - sc["b"]).max() < 1e-6
    back = behind_camera_copy(sc)
    flag, ends = oracle_flags(back, ortho)      # the world line is the same
    assert (flag == 2).all()
    assert np.abs(ends[:, :3] - sc["a"]).max() < 1e-6 and np.abs(ends[:, 3:] - sc["b"]).max() < 1e-6


@pytest.fixture(scope="module")
def solver():
    import uvs_b200
    s = uvs_b200.Solver(0)
    yield s
    s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("seed,noise", [(11, 0.0), (12, 1.0 / 460.0), (13, 3.0 / 460.0)])
def test_gpu_points_match_oracle(solver, seed, noise):
    sc = make_scene(seed, n_tracks=500, noise=noise)
    ref = oracle_points(sc)
    got = solver.triangulate_points(sc["Rs"].reshape(-1, 9), sc["Ps"], sc["ric"].reshape(9), sc["tic"], sc["start"], sc["off"], sc["pts"])
    # same singular vector up to the conditioning of the 2n x 4 system: 1e-6 relative (north_star tolerance on residuals)
    assert np.max(np.abs(got - ref) / np.abs(ref)) < 1e-6
    assert np.median(np.abs(got - ref) / np.abs(ref)) < 1e-11


@pytest.mark.gpu
def test_gpu_points_fallback_and_single_observation(solver):
    sc = make_scene(14, n_tracks=8)
    # track 0 behind the camera (negative depth) -> init depth, as in the reference
    pts = sc["pts"].copy()
    pts[sc["off"][0]:sc["off"][1], :2] *= -1.0
    ref = np.array([tri.triangulate_point(sc["Rs"], sc["Ps"], sc["ric"], sc["tic"], int(sc["start"][t]), pts[sc["off"][t]:sc["off"][t + 1]], 7.5)
                    for t in range(8)])
    got = solver.triangulate_points(sc["Rs"].reshape(-1, 9), sc["Ps"], sc["ric"].reshape(9), sc["tic"], sc["start"], sc["off"], pts, init_depth=7.5)
    ok = ref != 7.5
    assert np.allclose(got[ok], ref[ok], rtol=1e-6)
    assert np.all(got[~ok] == 7.5)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [21, 22])
def test_gpu_lines_match_oracle(solver, seed):
    sc = make_scene(seed, n_lines=300)
    ref = oracle_lines(sc)
    got = solver.triangulate_lines(sc["Rs"].reshape(-1, 9), sc["Ps"], sc["ric"].reshape(9), sc["tic"], sc["first"], sc["last"],
                                   sc["sp0"], sc["ep0"], sc["sp1"], sc["ep1"])
    # angles are compared through what they parameterise (a branch flip of 2 pi would be the same rotation)
    for t in range(len(ref)):
        n0, d0 = ortho_to_plucker(ref[t])
        n1, d1 = ortho_to_plucker(got[t])
        assert np.linalg.norm(n0 - n1) < 1e-9 and np.linalg.norm(d0 - d1) < 1e-9
    assert np.max(np.abs(got - ref)) < 1e-8


@pytest.mark.gpu
def test_gpu_line_validity_matches_oracle(solver):
    """uvs_validate_lines against the oracle: lines in front of their first camera, the same lines behind it, and random
    orthonormal parameters (both outcomes); flags equal, world end points to 1e-9 relative"""
    sc = make_scene(41, n_lines=120)
    ortho = oracle_lines(sc)
    cases = [(sc, ortho), (behind_camera_copy(sc), ortho)]
    rng = np.random.default_rng(7)
    cases.append((sc, np.column_stack([rng.uniform(-np.pi, np.pi, (120, 3)), rng.uniform(0.05, 1.5, 120)])))
    seen = set()
    for scn, o in cases:
        f0, e0 = oracle_flags(scn, o)
        f1, e1 = solver.validate_lines(scn["Rs"], scn["Ps"], scn["ric"], scn["tic"], scn["first"], o, scn["sp0"], scn["ep0"], want_end_points=True)
        # camera-frame depths of the two end points: leave out lines whose depth is zero to rounding (the sign is then noise)
        R = np.array([scn["Rs"][int(f)] @ scn["ric"] for f in scn["first"]])
        c = np.array([scn["Ps"][int(f)] + scn["Rs"][int(f)] @ scn["tic"] for f in scn["first"]])
        zs = np.einsum("nij,ni->nj", R, e0[:, :3] - c)[:, 2]
        ze = np.einsum("nij,ni->nj", R, e0[:, 3:] - c)[:, 2]
        clear = (np.abs(zs) > 1e-9) & (np.abs(ze) > 1e-9) & np.isfinite(e0).all(axis=1)
        assert clear.sum() > 100
        assert np.array_equal(f0[clear], f1[clear])
        assert (np.abs(e1[clear] - e0[clear]) <= 1e-9 * np.maximum(1.0, np.abs(e0[clear]))).all()
        seen |= set(f0[clear].tolist())
    assert seen == {1, 2}
    f = solver.validate_lines(sc["Rs"], sc["Ps"], sc["ric"], sc["tic"], sc["first"], ortho, sc["sp0"], sc["ep0"])
    assert (f == 1).all()

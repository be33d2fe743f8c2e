"""Writes the committed window fixtures under tests/golden/ (uvs_window v1 blobs).

TOOLING.  The priors of these windows come from the CPU oracle's marginalisation of a solved
(F+1)-frame window (tools/gen_window.py), which is why bench.py reads the fixtures instead of
generating windows itself: bench.py's GPU arm must not execute the oracle.
    python tools/make_fixtures.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tools import gen_window as gw  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
FIXTURES = [("tiny", None), ("C1", None)] + [("C2", s) for s in (1002, 2002, 3002, 4002)] + [("10k", None)]


def main():
    os.makedirs(OUT, exist_ok=True)
    for cfg, seed in FIXTURES:
        w = gw.make_window(cfg, seed=seed)
        name = "window_%s%s.uvsw" % (cfg, "" if seed is None else "_s%d" % seed)
        w.save(os.path.join(OUT, name))
        print(name, "frames %d points %d lines %d proj %d line_obs %d vp_obs %d imu %d prior_n %d" % (
            w.n_frames, w.n_points, w.n_lines, w.n_proj, w.n_line_obs, w.n_vp_obs, w.n_imu, w.prior_n))


if __name__ == "__main__":
    main()


def oracle_goldens():
    """oracle outputs on the committed C1 fixture (regression pin used by tests/test_oracle.py)"""
    import numpy as np
    import uvs_b200
    from tests import orc
    w = uvs_b200.Window.load(os.path.join(OUT, "window_C1.uvsw"))
    o = uvs_b200.default_options()
    pr, _, _ = orc.eval_factors(w, o, orc.F_PROJ)
    ir, _, _ = orc.eval_factors(w, o, orc.F_IMU)
    ref = w.copy()
    sm = orc.solve(ref, o)
    np.savez(os.path.join(OUT, "oracle_C1.npz"), proj_r=pr, imu_r=ir, final_cost=sm.final_cost, pose=ref.pose)


if __name__ == "__main__":
    oracle_goldens()

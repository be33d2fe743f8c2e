"""Golden vectors for the marginalization parity test: for every configuration the window is solved by the CPU oracle,
the system (A, b) of MarginalizationInfo::marginalize is taken at that state, and A', b' of the reference's rule
(eigenvalues of Amm <= 1e-8 zeroed, marginalization_factor.cpp:263-281) are evaluated with mpmath at 40 digits
(tests/margref.py).  Stored per case: the solved state, exact A' / b', the FP64 oracle's A' / b', and the change of the
exact A' under a one-ulp perturbation of A (the conditioning of the rule).

    python tools/make_marg_fixtures.py        # writes tests/golden/marg_<cfg>_f<flag>.npz
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import uvs_b200  # noqa: E402
from tests import margref, orc  # noqa: E402
from tools import gen_window as gw  # noqa: E402

CASES = (("tiny", 0), ("tiny", 1), ("C1", 0), ("C2", 0))


def main():
    opts = uvs_b200.default_options()
    for cfg, flag in CASES:
        w = gw.make_window(cfg)
        orc.solve(w, opts)
        sysm = orc.marginalize_system(w, opts, flag)
        mo = orc.marginalize(w, opts, flag)
        A, b, m, n = sysm
        Ap, bp, ev = margref.exact_schur(A, b, m)
        cond = margref.conditioning(A, b, m)
        out = os.path.join(ROOT, "tests", "golden", "marg_%s_f%d.npz" % (cfg, flag))
        np.savez_compressed(out, pose=w.pose, speed_bias=w.speed_bias, ex_pose=w.ex_pose, inv_depth=w.inv_depth, ortho=w.ortho,
                            A_exact=Ap, b_exact=bp, A_oracle=mo["A"], b_oracle=mo["b"], cond=cond, m=m, n=n, evals_mm=ev)
        sA = np.abs(Ap).max()
        print("%s flag %d: m %d n %d | oracle vs exact %.2e | one-ulp conditioning %.2e | smallest / largest eigenvalue of Amm %.3g / %.3g"
              % (cfg, flag, m, n, np.abs(mo["A"] - Ap).max() / sA, cond, ev.min(), ev.max()))


if __name__ == "__main__":
    main()

"""Developer aid: per-stage device time of a single-window solve (the reference's own use case)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import glob
import numpy as np
import uvs_b200

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
w = uvs_b200.Window.load(sorted(glob.glob(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "window_C2_s*.uvsw")))[0])
o = uvs_b200.default_options(max_num_iterations=10, fixed_iterations=1)
s = uvs_b200.Solver(0)
s.upload([w.copy() for _ in range(B)], o)
for prof in (0, 1):
    s.set_profiling(prof)
    for _ in range(5):
        s.reset_state(); s.solve()
    ms = []
    for _ in range(20):
        s.reset_state(); s.solve(); ms.append(s.last_solve_ms())
    print("B=%d profiling=%d solve ms median %.3f min %.3f -> %.0f iter/s" % (B, prof, np.median(ms), np.min(ms), B * 10 / (np.median(ms) * 1e-3)))
st, n = s.last_stage_ms()
tot = sum(st.values())
print("stages (us per iteration):", {k: round(1e3 * v / n, 1) for k, v in st.items()}, "sum %.1f" % (1e3 * tot / n))

"""A/B of the fused point-kernel variants on the bench batch (developer tool): value and stage times per setting."""
import itertools
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402
import uvs_b200  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
ws = bench.load_workload(B)
opts = uvs_b200.default_options(max_num_iterations=10, fixed_iterations=1)
s = uvs_b200.Solver(0)
for llocc, occ in itertools.product((3, 4), (512, 640, 768)):
    dense, cta, pre = 1, 32, 0
    os.environ["UVS_PT_DENSE"], os.environ["UVS_PT_CTA"], os.environ["UVS_PT_PREFETCH"] = str(dense), str(cta), str(pre)
    os.environ["UVS_PT_OCC"], os.environ["UVS_LL_OCC"] = str(occ), str(llocc)
    s.upload(ws, opts)
    s.set_profiling(1)
    for _ in range(3):
        s.reset_state(); s.solve()
    ms, build = [], []
    for _ in range(6):
        s.reset_state(); s.solve(); ms.append(s.last_solve_ms())
        st, n = s.last_stage_ms(); build.append(st["build"] / n)
    print("lines occ %d  points threads/SM %d : dense %d cta %3d prefetch %d : %.3f ms/step  %.0f it/s  build %.3f ms/iter" % (llocc, occ, dense, cta, pre, np.median(ms), B * 10 / np.median(ms) * 1e3, np.median(build)), flush=True)
os.environ["UVS_NO_FUSE"] = "1"
s.upload(ws, opts)
s.set_profiling(1)
for _ in range(3):
    s.reset_state(); s.solve()
ms = []
for _ in range(6):
    s.reset_state(); s.solve(); ms.append(s.last_solve_ms())
print("record path (UVS_NO_FUSE) : %.3f ms/step  %.0f it/s" % (np.median(ms), B * 10 / np.median(ms) * 1e3))

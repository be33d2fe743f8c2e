"""Host-side phase timings of the reference-facing call (UVS_TRACE=1): pack / H2D / prep / solve / download."""
import os, sys, time
import numpy as np
if len(sys.argv) > 2 and sys.argv[2] == "trace":
    os.environ["UVS_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import uvs_b200 as uvs

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ws = bench.load_workload(B)
opts = uvs.default_options(max_num_iterations=bench.K_LM, fixed_iterations=1)
s = uvs.Solver(0)
sets = [[w.copy() for w in ws] for _ in range(4)]
views = [uvs.window_array(x) for x in sets]
for k in range(4):
    t0 = time.perf_counter()
    s.batch_solve(sets[k], opts, prepared=views[k])
    print("== call %d: %.2f ms" % (k, (time.perf_counter() - t0) * 1e3), file=sys.stderr, flush=True)

if "UVS_TRACE" in os.environ:
    sys.exit(0)
import numpy as np
ref = [w.pose.copy() for w in sets[0]]
noise = max(float(np.max(np.abs(a.pose - b))) for a, b in zip(sets[1], ref))
print("run-to-run pose noise (unpipelined): %.3g" % noise, file=sys.stderr)
for G in (None, 2, 3, 4, 6, 8):
    fresh = [[w.copy() for w in ws] for _ in range(4)]
    vs = [uvs.window_array(x) for x in fresh]
    s.batch_solve(fresh[0], opts, prepared=vs[0], groups=G)
    t0 = time.perf_counter()
    for k in range(1, 4):
        s.batch_solve(fresh[k], opts, prepared=vs[k], groups=G)
    dt = (time.perf_counter() - t0) / 3
    err = max(float(np.max(np.abs(a.pose - b))) for a, b in zip(fresh[1], ref))
    print("groups=%s: %.2f ms/step  %.0f it/s   max |pose - unpipelined| = %.3g" % (G, dt * 1e3, B * 10 / dt, err), file=sys.stderr, flush=True)

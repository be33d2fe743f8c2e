"""Host-side phase timings of the reference-facing call (UVS_TRACE=1): pack / H2D / prep / solve / download."""
import os, sys, time
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

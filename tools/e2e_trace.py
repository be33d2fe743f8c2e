import os, sys, time
os.environ["UVS_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import uvs_b200 as uvs
B = 1184
ws = bench.load_workload(B)
opts = uvs.default_options(max_num_iterations=bench.K_LM, fixed_iterations=1)
s = uvs.Solver(0)
sets = [[w.copy() for w in ws] for _ in range(4)]
views = [uvs.window_array(x) for x in sets]
for G in (4, 4, 4, 2, 3):
    for k in range(2):
        t0 = time.perf_counter()
        s.batch_solve(sets[k], opts, prepared=views[k], groups=G)
        print("== G=%d call %d: %.2f ms" % (G, k, (time.perf_counter() - t0) * 1e3), file=sys.stderr, flush=True)

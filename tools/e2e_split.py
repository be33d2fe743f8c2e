"""Developer aid: end-to-end time of the pipelined service call for several group counts (run under different UVS_PIPE_FIRST)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import uvs_b200 as uvs
B = 1184
ws = bench.load_workload(B)
opts = uvs.default_options(max_num_iterations=bench.K_LM, fixed_iterations=1)
s = uvs.Solver(0)
for G in (2, 3, 4):
    fresh = [[w.copy() for w in ws] for _ in range(5)]
    vs = [uvs.window_array(x) for x in fresh]
    s.batch_solve(fresh[0], opts, prepared=vs[0], groups=G)
    t0 = time.perf_counter()
    for k in range(1, 5):
        s.batch_solve(fresh[k], opts, prepared=vs[k], groups=G)
    dt = (time.perf_counter() - t0) / 4
    print("UVS_PIPE_FIRST=%s groups=%d: %.2f ms/step  %.0f it/s" % (os.environ.get("UVS_PIPE_FIRST", "1"), G, dt * 1e3, B * 10 / dt), flush=True)

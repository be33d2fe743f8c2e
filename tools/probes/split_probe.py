"""Does solving two half batches side by side (two handles, two host threads) beat one full batch?  The latency-bound
reduced solve of one half could hide under the FP64-bound linearisation of the other."""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import bench
import uvs_b200 as uvs

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
ws = bench.load_workload(B)
opts = uvs.default_options(max_num_iterations=bench.K_LM, fixed_iterations=1)
one = uvs.Solver(0)
one.upload(ws, opts)
for _ in range(3):
    one.reset_state(); one.solve()
t = []
for _ in range(10):
    one.reset_state(); t0 = time.perf_counter(); one.solve(); t.append(time.perf_counter() - t0)
print("one handle, %d windows: %.2f ms wall (device %.2f ms)" % (B, 1e3 * np.median(t), one.last_solve_ms()))
for G in (2, 3, 4):
    hs = [uvs.Solver(0) for _ in range(G)]
    cut = [B * g // G for g in range(G + 1)]
    for g in range(G):
        hs[g].upload(ws[cut[g]:cut[g + 1]], opts)
    def run(s):
        s.reset_state(); s.solve()
    for _ in range(3):
        th = [threading.Thread(target=run, args=(s,)) for s in hs]
        [x.start() for x in th]; [x.join() for x in th]
    t = []
    for _ in range(10):
        th = [threading.Thread(target=run, args=(s,)) for s in hs]
        t0 = time.perf_counter()
        [x.start() for x in th]; [x.join() for x in th]
        t.append(time.perf_counter() - t0)
    print("%d handles side by side: %.2f ms wall (device times %s)" % (G, 1e3 * np.median(t), [round(s.last_solve_ms(), 2) for s in hs]))
    for s in hs:
        s.close()

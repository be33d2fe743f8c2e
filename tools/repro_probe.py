"""Run-to-run reproducibility of the batched solve: the same 1024 windows solved as one batch and in three sub-batches;
windows whose solved poses differ by more than 1e-6 are listed with their iteration logs (accept sequence, costs)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import uvs_b200 as uvs

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ws = bench.load_workload(B)
opts = uvs.default_options(max_num_iterations=bench.K_LM, fixed_iterations=1)
s = uvs.Solver(0)
a = [w.copy() for w in ws]
sa = s.batch_solve(a, opts)
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
for rep in range(reps):
  b = [w.copy() for w in ws]
  sb = s.batch_solve(b, opts, groups=(None, 2, 3)[rep % 3])
  d = np.array([np.abs(x.pose - y.pose).max() for x, y in zip(a, b)])
  bad = np.nonzero(d > 1e-6)[0]
  print("run %d (groups %s): windows with |pose difference| > 1e-6: %d of %d (max %.3g)" % (rep, (None, 2, 3)[rep % 3], len(bad), B, d.max()))
  for i in bad[:3]:
    n = sa[i].num_iterations
    print("window %d (fixture %d): diff %.3g" % (i, i % 4, d[i]))
    print("   first run : accepted", [sa[i].step_accepted[k] for k in range(n)], "cost", ["%.9g" % sa[i].cost[k] for k in range(n)])
    print("   this run  : accepted", [sb[i].step_accepted[k] for k in range(n)], "cost", ["%.9g" % sb[i].cost[k] for k in range(n)])
    print("   rho       :", ["%.3g" % sa[i].relative_decrease[k] for k in range(n)], "|", ["%.3g" % sb[i].relative_decrease[k] for k in range(n)])
    print("   radius    :", ["%.3g" % sa[i].radius[k] for k in range(n)], "|", ["%.3g" % sb[i].radius[k] for k in range(n)])

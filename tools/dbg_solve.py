"""Developer aid: side-by-side iteration log of the GPU solve and the CPU oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, uvs_b200
from tools import gen_window as gw
from tests import orc
o = uvs_b200.default_options()
s = uvs_b200.Solver(0)
for cfg in sys.argv[1:]:
    w = gw.make_window(cfg); ref = w.copy()
    sm0 = orc.solve(ref, o)
    s.upload([w], o); sm = s.solve()[0]; s.download()
    n = sm.num_iterations
    print(cfg, "iters", n, sm0.num_iterations, "solve ms", s.last_solve_ms(), "sweep", s.last_sweep_ms())
    for i in range(n):
        print("  %2d cost %.10g %.10g rel %.2e | radius %.4g %.4g | acc %d %d | step %.3e %.3e | g %.3e %.3e | rd %.4f %.4f" % (i, sm.cost[i], sm0.cost[i], abs(sm.cost[i]-sm0.cost[i])/abs(sm0.cost[i]), sm.radius[i], sm0.radius[i], sm.step_accepted[i], sm0.step_accepted[i], sm.step_norm[i], sm0.step_norm[i], sm.gradient_max_norm[i], sm0.gradient_max_norm[i], sm.relative_decrease[i], sm0.relative_decrease[i]))
    print("  pose", np.abs(w.pose-ref.pose).max(), "sb", np.abs(w.speed_bias-ref.speed_bias).max(), "inv", np.abs(w.inv_depth-ref.inv_depth).max(), "ortho", np.abs(w.ortho-ref.ortho).max())

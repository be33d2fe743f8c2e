import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, uvs_b200
from tests import orc
w0 = uvs_b200.Window.load("/root/repo/tests/golden/window_10k.uvsw")
for fixed in (1, 0):
    opts = uvs_b200.default_options(max_num_iterations=10, fixed_iterations=fixed)
    ref = w0.copy(); sm0 = orc.solve(ref, opts)
    print("fixed", fixed, "oracle", sm0.final_cost, sm0.num_iterations, [sm0.step_accepted[i] for i in range(sm0.num_iterations)], [round(sm0.cost[i],4) for i in range(sm0.num_iterations)])
    s = uvs_b200.Solver(0)
    for env in (None, "1"):
        if env: os.environ["UVS_NO_FUSE"] = env
        else: os.environ.pop("UVS_NO_FUSE", None)
        for rep in range(3):
            w = w0.copy(); s.upload([w], opts); sm = s.solve()[0]; s.download()
            print("  nofuse", env, "gpu", sm.final_cost, sm.num_iterations, [sm.step_accepted[i] for i in range(sm.num_iterations)], "pose diff", np.abs(w.pose-ref.pose).max(), "cost under oracle", orc.total_cost(w, opts))
    print("  gpu costs", [round(sm.cost[i],4) for i in range(sm.num_iterations)])

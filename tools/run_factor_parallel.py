"""Factor-parallel solve of ONE window over all ranks (launch with torchrun); rank 0 checks the result against a
single-GPU solve of the same window and prints one JSON line.  Used by tests/test_multi_gpu.py and for the
break-even measurement of DESIGN.md 7."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import uvs_b200  # noqa: E402
from uvs_b200.parallel import make_allreduce  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "window_10k.uvsw"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = uvs_b200.Window.load(os.path.join(ROOT, "tests", "golden", name))
    opts = uvs_b200.default_options(max_num_iterations=10, fixed_iterations=1)
    s = uvs_b200.Solver(local)
    s.comm_init(rank, world, make_allreduce(dist, "cuda"))
    par = w.copy()
    s.upload([par], opts)
    for _ in range(2):
        s.reset_state(); s.solve()
    dist.barrier(); torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        s.reset_state()
        sm = s.solve()[0]
        ms.append(s.last_solve_ms())
    s.download()
    t = torch.tensor([float(np.median(ms))], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        one = uvs_b200.Solver(local)
        ref = w.copy()
        one.upload([ref], opts)
        for _ in range(2):
            one.reset_state(); one.solve()
        ms1 = []
        for _ in range(reps):
            one.reset_state(); sm1 = one.solve()[0]; ms1.append(one.last_solve_ms())
        one.download()
        out = {"window": name, "n_gpus": world, "ms_per_solve": float(t.item()), "ms_per_solve_1gpu": float(np.median(ms1)),
               "final_cost": sm.final_cost, "final_cost_1gpu": sm1.final_cost,
               "pose_diff": float(np.abs(par.pose - ref.pose).max()), "inv_depth_diff": float(np.abs(par.inv_depth - ref.inv_depth).max()),
               "iterations": sm.num_iterations}
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

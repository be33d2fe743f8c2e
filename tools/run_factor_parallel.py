"""Factor-parallel solve of ONE window over all ranks (launch with torchrun); rank 0 checks the result against a
single-GPU solve of the same window and prints one JSON line.  Used by tests/test_multi_gpu.py and bench.py --mode factor.

  torchrun --nproc-per-node N tools/run_factor_parallel.py WINDOW.uvsw [REPS] [nccl|callback]

nccl (default): the library's own communicator (uvs_comm_init_nccl: ncclCommInitRank from a unique id that rank 0 creates
and torch.distributed hands round); callback: the reduction as a callback into torch.distributed (uvs_comm_init)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import uvs_b200  # noqa: E402
from uvs_b200.parallel import init_factor_parallel  # noqa: E402


def run(name, reps, how="nccl", k_lm=10):
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    path = name if os.path.exists(name) else os.path.join(ROOT, "tests", "golden", name)
    w = uvs_b200.Window.load(path)
    opts = uvs_b200.default_options(max_num_iterations=k_lm, fixed_iterations=1)
    s = uvs_b200.Solver(local)
    init_factor_parallel(s, dist, rank, world, how)
    par = w.copy()
    s.upload([par], opts)
    for _ in range(3):
        s.reset_state(); s.solve()
    dist.barrier(); torch.cuda.synchronize()
    ms, c0 = [], s.collective_count()
    for _ in range(reps):
        s.reset_state()
        sm = s.solve()[0]
        ms.append(s.last_solve_ms())
    ncoll = (s.collective_count() - c0) / max(1, reps)
    s.download()
    t = torch.tensor([float(np.median(ms))], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out = None
    if rank == 0:
        one = uvs_b200.Solver(local)
        ref = w.copy()
        one.upload([ref], opts)
        for _ in range(3):
            one.reset_state(); one.solve()
        ms1 = []
        for _ in range(reps):
            one.reset_state(); sm1 = one.solve()[0]; ms1.append(one.last_solve_ms())
        one.download()
        one.close()
        out = {"window": os.path.basename(name), "n_gpus": world, "comm": how, "ms_per_solve": float(t.item()), "ms_per_solve_1gpu": float(np.median(ms1)),
               "speedup_vs_1gpu": float(np.median(ms1)) / float(t.item()), "collectives_per_solve": ncoll,
               "final_cost": sm.final_cost, "final_cost_1gpu": sm1.final_cost,
               "pose_diff": float(np.abs(par.pose - ref.pose).max()), "inv_depth_diff": float(np.abs(par.inv_depth - ref.inv_depth).max()),
               "iterations": sm.num_iterations, "n_proj": int(w.n_proj), "n_line_obs": int(w.n_line_obs), "n_vp_obs": int(w.n_vp_obs),
               "cam_dim": int(w.cam_dim)}
    dist.barrier()
    s.close()
    return out


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "window_10k.uvsw"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    how = sys.argv[3] if len(sys.argv) > 3 else "nccl"
    out = run(name, reps, how)
    if out is not None:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

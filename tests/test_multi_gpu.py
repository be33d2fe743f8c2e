"""Factor-parallel solve over 2 GPUs (NCCL all-reduce of the reduced camera system) equals the 1-GPU solve."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_factor_parallel_two_gpus_match_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29581", os.path.join(ROOT, "tools", "run_factor_parallel.py"), "window_C2_s1002.uvsw", "2"],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    line = [l for l in p.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert abs(out["final_cost"] - out["final_cost_1gpu"]) <= 1e-6 * abs(out["final_cost_1gpu"])
    assert out["pose_diff"] < 1e-4 and out["inv_depth_diff"] < 1e-4

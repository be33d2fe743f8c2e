"""Factor-parallel solve over 2 GPUs (NCCL all-reduce of the reduced camera system) equals the 1-GPU solve."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("window,how", [("window_C2_s1002.uvsw", "nccl"), ("window_10k.uvsw", "nccl"), ("window_C2_s1002.uvsw", "callback")])
def test_factor_parallel_two_gpus_match_single_gpu(window, how):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29581", os.path.join(ROOT, "tools", "run_factor_parallel.py"), window, "2", how],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    line = [l for l in p.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    # The 10 k-factor window starts at cost 2.5e10 and its weakly observable line parameters make the last two of the ten
    # iterations chaotic at the 1e-4 level for ANY summation order (measured: repeated single-GPU solves of the same
    # upload differ by 1e-4 in the final cost and agree to 2e-6 in the poses; tests/test_gpu_parity.py::test_10k_window);
    # the C2 window is held to the 1e-6 / 1e-4 bars of the north_star.
    tol = 5e-4 if "10k" in window else 1e-6
    assert abs(out["final_cost"] - out["final_cost_1gpu"]) <= tol * abs(out["final_cost_1gpu"])
    assert out["pose_diff"] < 1e-4 and out["inv_depth_diff"] < (1e-2 if "10k" in window else 1e-4)
    assert out["collectives_per_solve"] == 20   # ten LM iterations: one all-reduce of the reduced system + accumulators, one of the accumulators each

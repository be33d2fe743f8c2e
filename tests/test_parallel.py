"""N > 1 host logic on CPU: world_size-2 gloo processes exercise the window sharding and the all-reduce callback the
library calls in the factor-parallel mode (on host buffers here; the device path is tests/test_multi_gpu.py)."""
import ctypes as C
import os
import subprocess
import sys
import textwrap

import numpy as np

from uvs_b200.parallel import landmark_owner, shard_windows

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_windows_partitions_exactly():
    items = list(range(37))
    for world in (1, 2, 3, 8):
        parts = [shard_windows(items, r, world) for r in range(world)]
        assert sum(parts, []) == items
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_landmark_ownership_covers_every_landmark_once():
    idx = np.arange(1000)
    for world in (2, 4, 8):
        own = landmark_owner(idx, world)
        assert set(own) == set(range(world))
        assert all(np.count_nonzero(own == r) in (1000 // world, 1000 // world + 1) for r in range(world))


WORKER = textwrap.dedent("""
    import ctypes as C, os, sys
    sys.path.insert(0, %r)
    import numpy as np, torch, torch.distributed as dist
    import uvs_b200
    from uvs_b200.parallel import make_allreduce, shard_windows
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    fn = make_allreduce(dist, "cpu")
    # the callback contract of uvs_comm_init: sum `count` doubles in place over the ranks
    buf = (C.c_double * 1000)(*[float(rank + 1) * (i + 1) for i in range(1000)])
    assert fn(C.addressof(buf), 1000, 0) == 0
    exp = sum(r + 1 for r in range(world))
    assert all(abs(buf[i] - exp * (i + 1)) < 1e-9 for i in range(1000))
    # through the ctypes trampoline type the library receives
    cb = uvs_b200.binding.ALLREDUCE_FN(lambda user, p, n, st: fn(p, n, st))
    buf2 = (C.c_double * 8)(*([1.0] * 8))
    assert cb(None, C.addressof(buf2), 8, None) == 0 and abs(buf2[3] - world) < 1e-12
    # window-parallel: every rank takes its shard; the global iteration count is the all-reduced sum
    shard = shard_windows(list(range(10)), rank, world)
    t = torch.tensor([float(len(shard))], dtype=torch.float64)
    dist.all_reduce(t)
    assert t.item() == 10.0
    dist.barrier()
    print("rank", rank, "ok")
""")


def test_gloo_world_size_2_allreduce_callback(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29571", str(script)], capture_output=True, text=True, timeout=300, env=env)
    assert p.returncode == 0, p.stdout + p.stderr
    assert p.stdout.count("ok") == 2

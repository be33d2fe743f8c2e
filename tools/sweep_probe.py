"""The materialised Jacobian sweep alone (the roofline kernel group of bench.py) on the C2 x 1184 batch: for ncu captures
and A/B runs of k_proj<1> / k_line_vp<1> / k_imu_geom<1> + k_imu_weight / k_prior.

  python tools/sweep_probe.py [repeats] [windows]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import uvs_b200

rep = int(sys.argv[1]) if len(sys.argv) > 1 else 10
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1184
ws = bench.load_workload(B, 0, "C2")
opts = uvs_b200.default_options(max_num_iterations=10, fixed_iterations=1)
s = uvs_b200.Solver(0)
s.upload(ws, opts)
jac_bytes, _ = s.sweep_bytes()
ms, each = s.jacobian_sweep(repeats=rep)
print("group %.4f ms  %.1f GB/s  each %s" % (ms, jac_bytes / ms / 1e6, [round(v, 4) for v in each]))

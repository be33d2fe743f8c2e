#!/usr/bin/env python
"""bench.py — solver iterations/sec of the sliding-window backend solve (BASELINE.json's metric).

One "step" = one pass of the hot path over one batch: uvs_solve on B independent 11-frame / 200-point /
80-line / 3-VP windows ("C2", BASELINE.json configs[1]/[2]) per GPU, K_LM = 10 Levenberg-Marquardt
iterations each (config/euroc/euroc_config.yaml:56), convergence exits disabled so that every step
does exactly B x 10 iterations.  value = LM iterations per second over all windows and GPUs.

  python bench.py [--gpus N --steps K --warmup W]          our arm (CUDA, through the C ABI), C2 x 1184 windows per GPU
  python bench.py --config C1|C5|10k [...]                 the other BASELINE.json shapes, same measurement
  python bench.py --mode factor --window 10k|C5 [...]      ONE window sharded by landmark over the GPUs (NCCL all-reduce)
  python bench.py --impl reference [...]                   the CPU path on the host cores
The reference (ROS + Ceres + Eigen) cannot be built in this image, so the reference arm times the
Ceres-semantics CPU restatement in oracle/ (cpu_baseline.kind = "port"), all host threads.
"""
import argparse
import ctypes as C
import glob
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

K_LM = 10
METRIC = "solver iterations/sec (10-KF window)"
UNIT = "LM iterations/s"


DEFAULT_WINDOWS = {"C1": 2368, "C2": 1184, "C5": 256, "10k": 296}   # per GPU and step (multiples of 148 SMs where it matters)
SHAPES = {"C1": "C1 window (11 frames / 50 points / 20 lines / 1 VP; BASELINE.json configs[0])",
          "C2": "C2 window (11 frames / 200 points / 80 lines / 3 VP; BASELINE.json configs[1-2])",
          "C5": "C5 stress window (31 frames / 2000 points / 500 lines / 3 VP; BASELINE.json configs[4])",
          "10k": "10 k-factor window (11 frames / 1500 points / 500 lines / 3 VP; north_star)"}


def load_workload(n_windows, rank=0, config="C2"):
    """B windows of one shape: the committed fixtures (C5: the seeded generator), replicated with a seeded
    perturbation of the initial guess so that every window is a different problem."""
    from uvs_b200 import Window
    if config == "C2":
        paths = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "window_C2_s*.uvsw")))
        if not paths:
            raise SystemExit("fixtures missing: run python tools/make_fixtures.py")
        base = [Window.load(p) for p in paths]
    else:
        base = [named_window(config)]
    rng = np.random.default_rng(77 + 1000 * rank)
    out = []
    for i in range(n_windows):
        w = base[i % len(base)].copy()
        if i >= len(base):
            w.pose[:, :3] += rng.normal(0, 0.01, w.pose[:, :3].shape)
            dq = rng.normal(0, 0.002, (w.n_frames, 3))
            q = w.pose[:, 3:]
            # q <- q * (dq/2, 1), normalised
            x, y, z, s = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
            a, b, c = dq[:, 0] / 2, dq[:, 1] / 2, dq[:, 2] / 2
            qn = np.stack([s * a + x + y * c - z * b, s * b + y + z * a - x * c, s * c + z + x * b - y * a,
                           s - x * a - y * b - z * c], axis=1)
            w.pose[:, 3:] = qn / np.linalg.norm(qn, axis=1, keepdims=True)
            w.speed_bias[:, :3] += rng.normal(0, 0.01, (w.n_frames, 3))
            w.inv_depth *= 1.0 + rng.normal(0, 0.02, w.n_points)
            w.ortho += rng.normal(0, 0.005, w.ortho.shape)
        out.append(w)
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region"""

    def __init__(self, gpu=0):
        self.rows, self.proc, self.gpu = [], None, gpu

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for k, n in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


SWEEP_SOURCES = ("uvs_sweep.cu", "uvs_factors.cuh", "uvs_imu.cuh", "uvs_linefast.cuh", "uvs_math.cuh", "uvs_device.cuh")


def sources_sha():
    """sha256 over the sources of the materialised sweep kernels: ties the committed ncu profile (profiles/r2_sweep_ncu.json,
    written by tools/ncu_summary.py --sweep) to the source state it was taken at"""
    import hashlib
    h = hashlib.sha256()
    for name in SWEEP_SOURCES:
        h.update(open(os.path.join(ROOT, "uv-slam_b200", "csrc", name), "rb").read())
    return h.hexdigest()[:16]


def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_mod
    return rank, world, local, dist


def native_oracle():
    """The CPU baseline is timed with a -march=native build of the oracle made on THIS machine (oracle/Makefile `native`);
    the portable liborc.so of the tests is the fallback.  Must run before tests.orc is imported."""
    try:
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "native"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=300)
        path = os.path.join(ROOT, "oracle", "liborc_native.so")
        if os.path.exists(path):
            os.environ["UVS_ORC_LIB"] = path
            return "-O3 -march=native -ffp-contract=off (built on this host)"
    except Exception:
        pass
    return "-O3 -march=x86-64-v3 -ffp-contract=off (portable build; the native build failed)"


def run_cpu_sample(n_windows, threads, budget_s, k_lm=K_LM, config="C2"):
    """times the CPU oracle (test infrastructure) on a bounded sample of the same workload"""
    import uvs_b200
    from tests import orc
    ws = load_workload(n_windows, 0, config)
    o = uvs_b200.default_options(max_num_iterations=k_lm, fixed_iterations=1)
    t0 = time.perf_counter()
    orc.solve_batch([w.copy() for w in ws[:max(1, threads)]], o, threads)   # warm-up
    warm = time.perf_counter() - t0
    reps = max(1, int(budget_s / max(warm * n_windows / max(1, threads), 1e-3)))
    reps = min(reps, 400)
    t0 = time.perf_counter()
    for _ in range(reps):
        orc.solve_batch([w.copy() for w in ws], o, threads)
    dt = time.perf_counter() - t0
    return n_windows * k_lm * reps / dt, reps, dt


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    build = native_oracle()
    cfg = getattr(args, "config", "C2")
    n = max(threads * 2, 16) if cfg in ("C1", "C2") else max(threads, 4)
    # each "step" = one bounded sample: n windows x 10 LM iterations on all host threads
    import uvs_b200
    from tests import orc
    ws = load_workload(n, 0, cfg)
    o = uvs_b200.default_options(max_num_iterations=K_LM, fixed_iterations=1)
    for _ in range(args.warmup):
        orc.solve_batch([w.copy() for w in ws], o, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.solve_batch([w.copy() for w in ws], o, threads)
    dt = time.perf_counter() - t0
    value = n * K_LM * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s, %d LM iterations per window" % (SHAPES[cfg], K_LM),
                   "windows_per_step": n, "note": "reference = CPU restatement of the Ceres path (oracle/); Ceres itself cannot be built here"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d %s windows x %d LM iterations per step, %d host threads (window-parallel)" % (n, cfg, K_LM, threads),
                         "build": build},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def named_window(name):
    """-> Window for a configuration name (committed fixture, or generated with the seeded generator) or a .uvsw path"""
    from uvs_b200 import Window
    fixtures = {"10k": "window_10k.uvsw", "C2": "window_C2_s1002.uvsw", "C1": "window_C1.uvsw"}
    if name in fixtures:
        return Window.load(os.path.join(ROOT, "tests", "golden", fixtures[name]))
    if os.path.exists(name):
        return Window.load(name)
    from tools import gen_window as gw
    return gw.make_window(name)


def factor_arm(args):
    """ONE window, landmarks sharded over the ranks (SURVEY.md 8e factor-parallel): value = LM iterations/s of that window"""
    import uvs_b200
    rank, world, local, dist = dist_setup(args.gpus)
    w = named_window(args.window)
    opts = uvs_b200.default_options(max_num_iterations=K_LM, fixed_iterations=1)
    s = uvs_b200.Solver(local)
    if dist is not None:
        from uvs_b200.parallel import init_factor_parallel
        init_factor_parallel(s, dist, rank, world, "nccl")
    s.upload([w.copy()], opts)
    for _ in range(max(3, args.warmup)):
        s.reset_state(); s.solve()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if dist is not None:
        dist.barrier()
    ms, c0, l0 = [], s.collective_count(), s.launch_count()
    for _ in range(args.steps):
        s.reset_state()
        sm = s.solve()[0]
        ms.append(s.last_solve_ms())
    ncoll = (s.collective_count() - c0) / args.steps
    launches = s.launch_count() - l0
    if dist is not None:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    step_ms = float(np.mean(ms))
    if dist is not None:
        import torch
        t = torch.tensor([step_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_ms = float(t.item())
    # end to end: host window in, solved state out (upload + solve + download), all ranks
    e2e = []
    for _ in range(max(2, min(args.steps, 5))):
        c = w.copy()
        t0 = time.perf_counter()
        s.upload([c], opts); s.solve(); s.download()
        e2e.append(time.perf_counter() - t0)
    e2e_s = float(np.median(e2e))
    if dist is not None:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    jac_bytes, res_bytes = w.sweep_bytes()
    if rank == 0:
        d = w.cam_dim
        line = {
            "metric": METRIC, "value": K_LM / (step_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "ONE %s window (%d frames / %d points / %d lines: %d proj + %d line + %d VP factors), %d LM iterations per step"
                                   % (args.window, w.n_frames, w.n_points, w.n_lines, w.n_proj, w.n_line_obs, w.n_vp_obs, K_LM),
                       "parallelism": "factor-parallel x%d (landmark k on rank k %% N; IMU + prior on rank 0)" % world,
                       "collective": "ncclAllReduce (sum, f64) of [S | gS | g | column norms | accumulators] = %d doubles per LM iteration + %d doubles for the "
                                     "candidate cost; %g collectives per solve" % (d * d + 3 * d + 16, 16, ncoll),
                       "l2": "one window: inputs fit L2 (latency-bound regime, SURVEY.md 7)"},
            "e2e": {"value": K_LM / e2e_s, "unit": UNIT, "h2d_bytes_per_step": len(w.to_bytes()), "d2h_bytes_per_step": int(w.state_vector().nbytes),
                    "ms_per_step": 1e3 * e2e_s, "call": "uvs_upload_windows + uvs_solve + uvs_download_state on every rank"},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": (jac_bytes + res_bytes) * K_LM / (step_ms * 1e-3) / 1e9, "peak": 6550.1, "unit": "GB/s",
                         "frac": (jac_bytes + res_bytes) * K_LM / (step_ms * 1e-3) / 1e9 / 6550.1, "traffic": None,
                         "note": "whole solve of one window against the SURVEY 8d sweep bytes: launch- and latency-bound, not a bandwidth measurement"},
            "cpu_baseline": None, "final_cost": sm.final_cost,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    s.close()


def frame_path(local, k_lm=K_LM, n_slides=12):
    """The reference's own use - ONE window per image frame - end to end, two ways: (a) every window packed on the host and
    uploaded from scratch (uvs_upload_windows), (b) the device-resident window (uvs_window_*: only the new frame, the packed
    state and a plan cross the bus).  Host wall-clock of upload + solve + state download per frame, median over the slides."""
    import uvs_b200
    from tools import fm_ref, gen_sequence as gs
    W = 10
    seq = gs.Sequence(n_frames=W + 1 + n_slides, n_points=150, n_lines=50, seed=3)
    opts = uvs_b200.default_options(max_num_iterations=k_lm, fixed_iterations=1)
    dev, ref = uvs_b200.Solver(local), uvs_b200.Solver(local)
    dev.window_create(W, 5, 1024, 512)
    host = fm_ref.HostWindow(W, 5)
    nrng = np.random.default_rng(5)
    prior, t_res, t_scr, b_res, b_scr, slides = None, [], [], [], [], 0
    for k, fr in enumerate(seq.frames):
        host.push(k, fr)
        h0 = dev.h2d_bytes()
        t0 = time.perf_counter()
        dev.window_push_frame(fr.point_id, fr.point_xyz, fr.line_id, fr.line_sp, fr.line_ep, fr.line_vp, imu=fr.imu)
        t_push = time.perf_counter() - t0
        if len(host.frame_ids) < W + 1:
            continue
        pose, sb = seq.noisy_pose_sb(host.frame_ids, nrng)
        ep, el = host.eligible_points(), host.eligible_lines()
        inv = np.array([seq.inv_depth_of(t.id, host.frame_ids[t.start]) for t in ep]) * (1 + nrng.normal(0, 0.05, len(ep)))
        ortho = np.array([seq.ortho_of(t.id) for t in el]).reshape(-1, 4) + nrng.normal(0, 0.01, (len(el), 4))
        w_host = host.pack(pose, sb, seq.ex_pose(), inv, ortho, seq.ric, seq.tic, prior)
        w_dev = uvs_b200.Window(pose=pose.copy(), speed_bias=sb.copy(), ex_pose=seq.ex_pose(), inv_depth=inv.copy(), ortho=ortho.copy(),
                                line_ric=seq.ric.copy(), line_tic=seq.tic.copy())
        t0 = time.perf_counter()
        dev.window_upload(w_dev, opts); dev.solve(); dev.download()
        t_res.append(t_push + time.perf_counter() - t0)
        b_res.append(dev.h2d_bytes() - h0)
        h0 = ref.h2d_bytes()
        t0 = time.perf_counter()
        ref.upload([w_host], opts); ref.solve(); ref.download()
        t_scr.append(time.perf_counter() - t0)
        b_scr.append(ref.h2d_bytes() - h0)
        flag = 1 if slides % 3 == 2 else 0
        pd = dev.window_marginalize(flag)
        if pd is not None:
            prior = pd
        elif flag == 0:
            prior = None
        merged = ms = None
        if flag == 1 and len(host.imu) >= 2:
            ms = gs.merge_samples(host.imu_samples[-2], host.imu_samples[-1])
            merged = gs.imu_record(ms)
        host.slide(flag, merged, ms)
        dev.window_slide(flag, merged)
        slides += 1
    dev.close(); ref.close()
    med = lambda a: float(np.median(a[2:]))   # the first frames pay allocations
    return {"workload": "one 11-frame window per image frame, ~150 tracked points + ~50 lines per frame, %d LM iterations, %d frames" % (k_lm, slides),
            "from_scratch": {"ms_per_frame": 1e3 * med(t_scr), "h2d_bytes_per_frame": int(med(b_scr)), "call": "uvs_upload_windows + uvs_solve + uvs_download_state"},
            "device_resident": {"ms_per_frame": 1e3 * med(t_res), "h2d_bytes_per_frame": int(med(b_res)),
                                "call": "uvs_window_push_frame + uvs_window_upload + uvs_solve + uvs_download_state"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--windows", type=int, default=0, help="windows per GPU and step (default per --config; C2: 1184 = 8 x 148 SMs)")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work for the cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--mode", default="window", choices=["window", "factor"],
                    help="window: independent windows per GPU (default, weak scaling); factor: ONE window sharded by landmark over the GPUs "
                         "with an NCCL all-reduce of the reduced camera system per LM iteration (strong scaling)")
    ap.add_argument("--window", default="10k", help="--mode factor: 10k (11 frames / 1500 points / 500 lines), C5 (31 / 2000 / 500), C2, or a .uvsw path")
    ap.add_argument("--config", default="C2", choices=["C1", "C2", "C5", "10k"], help="--mode window: shape of the replicated window")
    ap.add_argument("--check", type=int, default=0, help="verify this many sampled windows of the timed batch against the CPU oracle (rank 0)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    if args.mode == "factor":
        return factor_arm(args)

    import uvs_b200
    rank, world, local, dist = dist_setup(args.gpus)
    B = args.windows or DEFAULT_WINDOWS[args.config]
    ws = load_workload(B, rank, args.config)
    opts = uvs_b200.default_options(max_num_iterations=K_LM, fixed_iterations=1)
    s = uvs_b200.Solver(local)
    s.upload(ws, opts)
    jac_bytes, res_bytes = s.sweep_bytes()
    s.set_profiling(1)

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(v):
        if dist is None:
            return v
        import torch
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing: inputs already in HBM, state rewound on the device between steps
    for _ in range(max(3, args.warmup)):
        s.reset_state(); s.solve()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    l0 = s.launch_count()
    dev_ms, stage_tot, iters_tot = 0.0, {}, 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        s.reset_state()
        s.solve()
        dev_ms += s.last_solve_ms()
        st, n_it = s.last_stage_ms()
        iters_tot += n_it
        for k, v in st.items():
            stage_tot[k] = stage_tot.get(k, 0.0) + v
    wall = time.perf_counter() - t0
    barrier()
    launches = s.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    step_ms = max_over_ranks(dev_ms / args.steps)
    value = world * B * K_LM / (step_ms * 1e-3)

    # ---- sampled parity of the TIMED batch: a few of its (perturbed) windows against the CPU oracle (test infrastructure)
    check = None
    if args.check > 0 and rank == 0:
        from tests import orc
        s.download()
        idx = sorted(set(np.linspace(0, B - 1, args.check).astype(int).tolist()))
        worst_cost, worst_pose = 0.0, 0.0
        fresh = load_workload(B, rank, args.config)
        for i in idx:
            ref = fresh[i]
            sm0 = orc.solve(ref, opts)
            worst_cost = max(worst_cost, abs(orc.total_cost(ws[i], opts) - sm0.final_cost) / abs(sm0.final_cost))
            worst_pose = max(worst_pose, float(np.abs(ws[i].pose - ref.pose).max()))
        check = {"windows": idx, "max_rel_cost_diff": worst_cost, "max_pose_diff": worst_pose, "ok": bool(worst_cost < 1e-6 and worst_pose < 1e-4)}

    # ---- materialised Jacobian sweep (the roofline kernel group): the four factor-type kernels write every residual and
    # tangent Jacobian block of the batch to HBM; CUDA events on the handle's stream (uvs_jacobian_sweep).  uvs_solve itself
    # takes the fused path (factors evaluated inside the landmark elimination, no records), so this is timed on its own.
    sweep_ms, sweep_each = s.jacobian_sweep(repeats=10)

    # ---- end to end through the reference-facing call: host buffers, H2D + solve + D2H per step
    host_sets = [[w.copy() for w in ws] for _ in range(2)]
    for k in range(2):
        s.batch_solve(host_sets[k % 2], opts, groups=0)
    n_e2e = max(2, min(args.steps, 5))
    fresh_sets = [[w.copy() for w in ws] for _ in range(n_e2e)]   # host copies made outside the timer
    views = [uvs_b200.window_array(fs) for fs in fresh_sets]     # ctypes structs of pointers to those host arrays
    barrier()
    t0 = time.perf_counter()
    for k in range(n_e2e):
        # uvs_batch_solve_pipelined: pack into pinned staging + H2D + solve + D2H, sub-batch k+1 uploading while k iterates
        s.batch_solve(fresh_sets[k], opts, prepared=views[k], groups=0)
    e2e_s = max_over_ranks((time.perf_counter() - t0) / n_e2e)
    e2e_value = world * B * K_LM / e2e_s
    h2d = sum(len(w.to_bytes()) for w in ws[:4]) // min(4, len(ws)) * B
    d2h = sum(w.state_vector().nbytes for w in ws) + B * C.sizeof(uvs_b200.UvsSummaryStruct)

    # ---- single-window latency (the reference's own use: one window per frame)
    s1 = uvs_b200.Solver(local)
    s1.upload([ws[0]], opts)
    for _ in range(5):
        s1.reset_state(); s1.solve()
    lat = []
    for _ in range(20):
        s1.reset_state(); s1.solve(); lat.append(s1.last_solve_ms())
    lat_ms = float(np.median(lat))
    # the same with the LM iteration replayed from a CUDA graph (pays only when one upload is solved repeatedly, as here)
    s1.set_graph_replay(True)
    for _ in range(3):
        s1.reset_state(); s1.solve()
    lat = []
    for _ in range(20):
        s1.reset_state(); s1.solve(); lat.append(s1.last_solve_ms())
    lat_graph_ms = float(np.median(lat))
    s1.close()
    frames = frame_path(local) if rank == 0 and args.config == "C2" else None

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    # ---- roofline of the Jacobian sweep (SURVEY.md 8d bytes) against the measured HBM peak
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "fallback of B200_PROFILING.md (of fallback)"
    achieved = jac_bytes / (sweep_ms * 1e-3) / 1e9
    total_stage = sum(stage_tot.values())
    shares = {k: round(v / total_stage, 4) for k, v in stage_tot.items() if v > 0}
    nproj, nline, nvp = sum(w.n_proj for w in ws), sum(w.n_line_obs for w in ws), sum(w.n_vp_obs for w in ws)
    nimu = sum(w.n_imu for w in ws)
    per_kernel = {}
    kbytes = (("k_proj", 384 * nproj), ("k_line_vp", 232 * nline + 120 * nvp), ("k_imu_geom+k_imu_weight", 6024 * nimu),
              ("k_prior", jac_bytes - 384 * nproj - 232 * nline - 120 * nvp - 6024 * nimu))
    for (k, nb), ms in zip(kbytes, sweep_each):
        if ms > 0:
            per_kernel[k] = {"ms_alone": round(ms, 4), "bytes": int(nb), "GB/s": round(nb / (ms * 1e-3) / 1e9, 1),
                             "frac": round(nb / (ms * 1e-3) / 1e9 / peak, 4)}
    sweep_serial_ms = float(sum(sweep_each))
    # fused linearisation stage of the solver (IMU + prior sweeps, point / line linearisation, tail, rank update): the same
    # algorithmic sweep bytes against its time (it never writes the point / line / VP records, so this can exceed the
    # materialised figure)
    lin_ms = stage_tot.get("build", 0.0) / max(1, iters_tot)
    iter_ms = total_stage / max(1, iters_tot)

    cpu = None
    if not args.no_cpu:
        build = native_oracle() if args.check <= 0 else "-O3 -march=x86-64-v3 -ffp-contract=off (portable build: --check loaded the checker first)"
        threads = os.cpu_count() or 1
        n_cpu = max(2 * threads, 16) if args.config in ("C1", "C2") else max(threads, 4)
        v, reps, dt = run_cpu_sample(n_cpu, threads, args.cpu_budget, config=args.config)
        v1, reps1, dt1 = run_cpu_sample(2 if args.config in ("C5", "10k") else 4, 1, min(4.0, args.cpu_budget / 3), config=args.config)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "%d %s windows x %d LM iterations x %d repetitions in %.1f s, %d host threads (window-parallel oracle)" % (
                   n_cpu, args.config, K_LM, reps, dt, threads),
               "single_thread_value": v1, "build": build}

    # DRAM traffic of the sweep launches from an ncu --set full capture of THIS source state: the profile names the kernel
    # sources it was taken at (sha256 over uv-slam_b200/csrc); a stale file is refused
    traffic, traffic_note = None, "no ncu capture of this source state"
    tp = os.path.join(ROOT, "profiles", "r2_sweep_ncu.json")
    if os.path.exists(tp) and args.config == "C2":
        prof = json.load(open(tp))
        if prof.get("windows") == B and prof.get("kernel_sources_sha") == sources_sha():
            traffic = prof["jacobian_sweep_dram_bytes"]
            traffic_note = "dram read+write of the sweep launches, " + prof.get("note", "")
        else:
            traffic_note = "profiles/r2_sweep_ncu.json was taken at another source state (%s, now %s) or batch size" % (prof.get("kernel_sources_sha"), sources_sha())

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "%s x %d independent windows per GPU, %d LM iterations per window per step" % (SHAPES[args.config], B, K_LM),
                   "windows_per_gpu": B, "lm_iterations": K_LM, "parallelism": "window-parallel x%d (no data-path collective)" % world,
                   "l2": "inputs larger than L2: %.0f MB of factor records per sweep" % (jac_bytes / 1e6)},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_s,
                "call": "uvs_batch_solve_pipelined (host UvsWindow arrays in, solved states + summaries out)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_note": traffic_note,
                     "kernel": "materialised Jacobian sweep = k_proj | k_line_vp | k_imu_geom + k_imu_weight | k_prior (Jacobian mode, "
                               "records to HBM), launched side by side on four streams (uvs_jacobian_sweep); achieved = SURVEY 8d bytes "
                               "of all four / CUDA-event time of the group",
                     "bytes_per_launch": int(jac_bytes), "ms_per_launch": sweep_ms, "ms_one_after_the_other": sweep_serial_ms,
                     "peak_source": peak_src, "per_kernel": per_kernel,
                     "fused_linearisation": {"ms_per_iteration": lin_ms, "GB/s_algorithmic": jac_bytes / (lin_ms * 1e-3) / 1e9 if lin_ms > 0 else None,
                                             "frac": jac_bytes / (lin_ms * 1e-3) / 1e9 / peak if lin_ms > 0 else None,
                                             "note": "solver path: factors evaluated inside the landmark elimination (no point / line / VP records); "
                                                     "the time also covers elimination, direct terms, IMU / prior blocks and the Schur rank update"},
                     "whole_iteration": {"ms": iter_ms, "bytes_algorithmic": int(jac_bytes + res_bytes),
                                         "frac": (jac_bytes + res_bytes) / (iter_ms * 1e-3) / 1e9 / peak if iter_ms > 0 else None}},
        "cpu_baseline": cpu,
        "stage_share": shares,
        "latency": {"single_window_ms_per_solve": lat_ms, "single_window_iterations_per_s": K_LM / (lat_ms * 1e-3),
                    "single_window_ms_per_solve_graph_replay": lat_graph_ms},
        "wall_ms_per_step": 1e3 * wall / args.steps,
    }
    if frames is not None:
        line["e2e"]["single_window_per_frame"] = frames
    if check is not None:
        line["parity_check"] = check
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

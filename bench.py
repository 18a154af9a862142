#!/usr/bin/env python
"""bench.py -- particle-updates/s of the fused MidasTouch filter step on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path (codebook query + motion + exact SE3_NN + weighting +
drift pruning + systematic resampling) over all particles.  Workload (BASELINE.json configs[2]):
"035_power_drill log 3, N=1e6 particles, 1xB200", synthetic stand-in assets,
codebook M=50 000, D=256 float64 (the reference's shipped width and dtype).
N>1: weak scaling, 1e6 particles per GPU, one all-gather of 8-byte weight sums per step.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_PER_GPU = 1_000_000
M, D = 50_000, 256
OBJ = "035_power_drill"
ALGO_BYTES_PER_UPDATE = 208  # SURVEY 8d / DESIGN.md
T_TRAJ = 64


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    def __init__(self, index):
        self.rows, self.stop = [], False
        self.index = index
        self.th = threading.Thread(target=self.run, daemon=True)

    def run(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            while not self.stop:
                self.rows.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                                  pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
                time.sleep(0.02)
        except Exception as e:  # noqa: BLE001
            self.err = str(e)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.th.join(timeout=2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        names = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
                 0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
                 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}
        bits = 0
        for _, r in self.rows:
            bits |= r
        return {"sm_mhz": statistics.median(c for c, _ in self.rows), "sm_max_mhz": getattr(self, "max", None),
                "reasons": [v for k, v in names.items() if bits & k and v != "gpu_idle"]}


def make_assets(seed=3):
    from midastouch_b200 import synth

    obj = synth.make_object(OBJ)
    cbs = synth.make_codebook(obj, M=M, D=D, seed=seed)
    gt, meas = synth.make_trajectory(obj, T=T_TRAJ, seed=seed)
    return obj, cbs, gt, meas


def run_ours(args):
    import torch.distributed as dist

    from midastouch_b200 import synth
    from midastouch_b200.engine import FilterEngine
    from midastouch_b200.tactile_tree import tactile_tree

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    obj, cbs, gt, meas = make_assets()
    cb = tactile_tree(cbs.poses, cbs.cam_poses, cbs.embeddings)
    cb.to_device(dev)
    n = N_PER_GPU
    cap = n + (n // 8 if world > 1 else 0)
    eng = FilterEngine(cb, capacity=cap, sig_t=2e-4, sig_r=0.5, seed=1234, rank=rank, world=world, n_global=n * world,
                       mesh_vertices=obj.vertices, pen_max=0.002)
    # particles start on codebook poses around the trajectory start (what init + snap produce)
    g = torch.Generator().manual_seed(100 + rank)
    sel = torch.randint(0, M, (n,), generator=g)
    eng.load_particles(cbs.poses.to(dev)[sel.to(dev)], nn_hint=sel.int().to(dev), spatial_sort=not args.no_sort)
    from midastouch_b200.engine import prepare_odom

    odoms = [prepare_odom(torch.inverse(meas[t - 1]) @ meas[t]) for t in range(1, T_TRAJ)]
    gts = [gt[t].float().contiguous() for t in range(T_TRAJ)]
    # tactile codes: pinned host buffers, one per frame (the step's only per-frame input)
    codes_h = [synth.make_query(cbs, int(sel[t]) if rank == 0 else 0, seed=t).pin_memory() for t in range(T_TRAJ - 1)]
    codes_d = [c.to(dev) for c in codes_h]
    us = torch.rand(4096, generator=torch.Generator().manual_seed(7)).tolist()
    l2flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def one(t, host_inputs):
        k = t % (T_TRAJ - 1)
        eng.step(codes_h[k] if host_inputs else codes_d[k], odoms[k], u=us[t % 4096], gt=gts[k + 1] if host_inputs else None)
        if host_inputs:
            return eng.rmse.cpu()  # D2H of the step's result (8 bytes)

    for t in range(args.warmup):
        one(t, False)
    sync()
    # ---- device-resident timing: per-step events, L2 flushed between steps (untimed)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with ClockSampler(local) as clk:
        sync()
        for t in range(args.steps):
            l2flush.zero_()
            evs[t][0].record()
            one(args.warmup + t, False)
            evs[t][1].record()
        sync()
    ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = torch.tensor([sum(ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    value = n * world * args.steps / (total_ms * 1e-3)

    stats_loop = eng.ctx.stats(reset=True)
    # ---- kernel-level timing of the two sweep kernels (same stream, CUDA events)
    a_ms, b_ms, q_ms = kernel_times(eng, codes_d, odoms, us, l2flush, min(args.steps, 20))

    # ---- end to end through the public API: host code + odom in, rmse out, every step
    sync()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for t in range(args.steps):
        one(args.warmup + args.steps + t, True)
    t1.record()
    sync()
    e2e_ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e = n * world * args.steps / (float(e2e_ms.item()) * 1e-3)

    if rank == 0:
        peak, how = peaks()
        sweep_ms = a_ms + b_ms
        achieved = ALGO_BYTES_PER_UPDATE * n / (sweep_ms * 1e-3) / 1e9
        out = {
            "metric": "particle-updates/sec", "value": value, "unit": "particle-updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 poses / f64 weights+prefix",
            "data": "synthetic (seeded stand-ins for YCB-Slide assets; no datasets offline)",
            "config": {"workload": f"{OBJ} log 3, N=1e6 particles per GPU, fused motion+SE3_NN+weight+systematic-resample step",
                       "particles_per_gpu": n, "codebook_M": M, "embedding_D": D, "embedding_dtype": "f64",
                       "noise": "in-kernel Philox4x32-10", "particle_order": "random" if args.no_sort else "sorted by codebook cell at load",
                       "drift_pruning": "pen_max 2 mm against the 1 mm surface vertex set (density of nontextured.stl[::10])", "l2": "flushed (256 MiB write) before every timed step",
                       "parallelism": f"particles sharded x{world}"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": how, "kernel": "k_step_a + k_step_b (particle sweep)",
                         "algorithmic_bytes_per_launch_pair": ALGO_BYTES_PER_UPDATE * n,
                         "k_step_a_ms": a_ms, "k_step_b_ms": b_ms, "k_cosine_rows_ms": q_ms,
                         "codebook_query_gbs": (M * D * 8 + D * 8 + M * 16) / (q_ms * 1e-3) / 1e9},
            "e2e": {"value": e2e, "unit": "particle-updates/s", "h2d_bytes_per_step": D * 8 + 64 + 64 + 4, "d2h_bytes_per_step": 8},
            "gpu_launches": 6 * args.steps, "clocks": clk.summary(),
            "engine_stats": {"nn_grid_fallbacks_per_step": stats_loop["nn_fallbacks"] / (args.steps + args.warmup),
                             "on_surface_last_step": stats_loop["on_surface"], "overflow": stats_loop["overflow"]},
        }
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline(budget_s=15.0)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def kernel_times(eng, codes_d, odoms, us, l2flush, reps):
    """average device time of k_cosine_rows, k_step_a, k_step_b inside a live step: L2 flushed once
    before the step (as in the timed loop), CUDA events on the launching stream between the kernels."""
    import ctypes as C

    from midastouch_b200._lib import call, ptr, stream_ptr
    from midastouch_b200.context import dtype_code

    acc = [0.0, 0.0, 0.0]
    for r in range(reps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        k = r % len(odoms)
        a = eng._fill(odoms[k], us[r], None, None, None, True)
        q = codes_d[k].reshape(-1).contiguous()
        s = stream_ptr()
        l2flush.zero_()
        ev[0].record()
        call("mt_codebook_query", eng.ctx.h, ptr(q), dtype_code(q), 0, s)
        ev[1].record()
        call("mt_step_a", eng.ctx.h, C.byref(a), s)
        ev[2].record()
        if eng.world > 1:
            eng._allgather_sums()
        ev[3].record()
        call("mt_step_b", eng.ctx.h, C.byref(a), s)
        ev[4].record()
        eng.cur = 1 - eng.cur
        if eng.world > 1:
            eng.use_n_dev = True
        eng.t += 1
        torch.cuda.synchronize()
        acc[0] += ev[0].elapsed_time(ev[1])
        acc[1] += ev[1].elapsed_time(ev[2])
        acc[2] += ev[3].elapsed_time(ev[4])
    return acc[1] / reps, acc[2] / reps, acc[0] / reps


def cpu_baseline(budget_s=15.0, n=65536, steps=None):
    """the reference's algorithm (oracle port: same torch-CPU ops as particle_filter.py /
    tactile_tree.py, cKDTree standing in for pynanoflann) on the host cores, on a bounded
    sample of the same workload."""
    from oracle import oracle as O

    obj, cbs, gt, meas = make_assets()
    keys = O.r3_se3(cbs.poses)
    from scipy.spatial import cKDTree

    tree = cKDTree(keys.numpy().astype("float64"))
    vds = obj.vertices
    g = torch.Generator().manual_seed(100)
    sel = torch.randint(0, M, (n,), generator=g)
    poses = cbs.poses[sel].clone()
    from midastouch_b200 import synth

    done, t_total = 0, 0.0
    t = 0
    while (t_total < budget_s and (steps is None)) or (steps is not None and done < steps):
        odom = torch.inverse(meas[t % (T_TRAJ - 1)]) @ meas[t % (T_TRAJ - 1) + 1]
        q = synth.make_query(cbs, int(sel[t % n]), seed=t)
        t0 = time.perf_counter()
        tn, rot = O.draw_motion_noise(n, 2e-4, 0.5)
        moved, _ = O.motion_model(poses, odom, tn, rot)                      # motionModel
        qk = O.r3_se3(moved).numpy().astype("float64")
        _, idx = tree.query(qk, k=1, workers=-1)                               # SE3_NN (16-thread k-d tree)
        w = O.get_similarity(q, cbs.embeddings[torch.from_numpy(idx)], True)  # gather N x D f64 + cosine + softmax
        w, _ = O.remove_invalid(moved, w, vds, 0.002)                          # remove_invalid_particles (k-d tree)
        anc = O.low_var_indices(w, float(torch.rand(1)))                      # resampler("low_var") (vectorised form)
        poses = moved[anc.clamp(min=0)]
        t_total += time.perf_counter() - t0
        done += 1
        t += 1
    return {"value": n * done / t_total, "unit": "particle-updates/s", "cores": os.cpu_count(),
            "torch_threads": torch.get_num_threads(), "kind": "port",
            "sample": f"{done} steps of N={n} particles (same codebook M={M}, D={D} f64); SE3_NN via scipy cKDTree (workers=all) "
                      "standing in for pynanoflann; resampler low_var in its vectorised searchsorted form"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.manual_seed(0)
    n = 65536
    # warm-up
    cpu_baseline(steps=max(1, min(args.warmup, 2)), n=n)
    t0 = time.perf_counter()
    cb = cpu_baseline(steps=args.steps, n=n)
    wall = time.perf_counter() - t0
    out = {"impl": "reference", "metric": "particle-updates/sec", "value": cb["value"], "unit": "particle-updates/s",
           "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * n / cb["value"],
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 poses / f64 weights",
           "data": "synthetic (seeded stand-ins)",
           "config": {"workload": f"{OBJ} log 3, reference algorithm on host CPU, bounded sample N={n} per step",
                      "codebook_M": M, "embedding_D": D, "embedding_dtype": "f64"},
           "cpu_baseline": {**cb, "value": cb["value"]},
           "e2e": {"value": cb["value"], "unit": "particle-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "wall_s": wall}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-sort", action="store_true", help="keep the particles in random order (no spatial sort at load)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

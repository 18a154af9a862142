#!/usr/bin/env python
"""bench.py -- particle-updates/s of the fused MidasTouch filter step on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path (codebook query + motion + exact SE3_NN + weighting +
drift pruning + systematic resampling) over all particles.  Workload (BASELINE.json configs[2]):
"035_power_drill log 3, N=1e6 particles, 1xB200", synthetic stand-in assets,
codebook M=50 000, D=256 float64 (the reference's shipped width and dtype).
N>1: weak scaling, 1e6 particles per GPU, one all-gather of 8-byte weight sums per step.
Every measured phase is a fresh filter run from global initialisation (the cloud converges onto the
true pose within ~20 steps under the smooth synthetic embedding).
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
ALGO_BYTES_PER_UPDATE = 208  # whole sweep, SURVEY 8d / DESIGN.md
A_BYTES_PER_UPDATE = 104     # k_step_a: pose 48 R + 48 W, match index 4 R + 4 W
T_TRAJ = 256  # frames of the synthetic slide; every measured phase restarts the filter at frame 0


def _find_peak(obj, words, lo, hi):
    """first number in [lo, hi] stored under a key containing one of `words` (the file's layout is the driver's)"""
    if isinstance(obj, dict):
        for k, v in obj.items():
            if isinstance(v, (int, float)) and lo <= float(v) <= hi and any(w in str(k).lower() for w in words):
                return float(v)
        for k, v in obj.items():
            r = _find_peak(v, words, lo, hi)
            if r is not None:
                return r
    elif isinstance(obj, list):
        for v in obj:
            r = _find_peak(v, words, lo, hi)
            if r is not None:
                return r
    return None


def peaks():
    """(HBM GB/s, source): MEASURED_PEAKS.json when the driver wrote one, else the profiling recipe's fallback"""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            v = _find_peak(json.load(open(p)), ("hbm", "copy", "bandwidth", "gbs", "gb_s", "gb/s"), 1000.0, 12000.0)
            if v is not None:
                return v, "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def tensor_peak():
    """(dense bf16 TFLOP/s, source)"""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            v = _find_peak(json.load(open(p)), ("bf16", "tflop", "tensor", "tf_s", "tfs"), 200.0, 3000.0)
            if v is not None:
                return v, "measured"
        except Exception:
            pass
    return 1590.0, "fallback"


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region by a separate `nvidia-smi -lms` process (nothing of it
    runs under this process' GIL); parsed after the region."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.active"

    def __init__(self, index):
        self.index, self.proc, self.rows, self.max, self.err = index, None, [], None, None
        self.path = os.path.join("/tmp", f"mt_bench_clocks_{os.getpid()}_{index}.csv")

    def __enter__(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=self.fh, stderr=subprocess.DEVNULL)
            time.sleep(0.15)  # first samples land before the region starts
        except Exception as e:  # noqa: BLE001
            self.err = str(e)
        return self

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.05)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:  # noqa: BLE001
                self.proc.kill()
            self.fh.close()
            try:
                for line in open(self.path):
                    f = [x.strip() for x in line.split(",")]
                    if len(f) >= 3 and f[0].isdigit():
                        self.rows.append((int(f[0]), int(f[2], 16)))
                        self.max = int(f[1])
                os.remove(self.path)
            except Exception as e:  # noqa: BLE001
                self.err = str(e)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable: " + str(self.err)]}
        names = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
                 0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
                 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}
        bits = 0
        for _, r in self.rows:
            bits |= r
        return {"sm_mhz": statistics.median(c for c, _ in self.rows), "sm_max_mhz": self.max, "samples": len(self.rows),
                "sampler": "nvidia-smi -lms 20 (separate process)", "reasons": [v for k, v in names.items() if bits & k and v != "gpu_idle"]}


def make_assets(seed=3):
    from midastouch_b200 import synth

    obj = synth.make_object(OBJ)
    cbs = synth.make_codebook(obj, M=M, D=D, seed=seed, embedding="smooth")
    gt, meas = synth.make_trajectory(obj, T=T_TRAJ, seed=seed)
    return obj, cbs, gt, meas


def workload_config(world, warmup, steps, no_sort=False):
    """the `config` object of the JSON line -- identical in both arms (--impl ours / reference)"""
    return {"workload": f"{OBJ} log 3, N=1e6 particles per GPU, one filter step = motion + SE3_NN + weighting + drift pruning + resampling",
            "embeddings": "smooth synthetic pose embedding (random Fourier features), query = embedding of the true pose + noise",
            "particles_per_gpu": N_PER_GPU, "codebook_M": M, "embedding_D": D, "embedding_dtype": "f64",
            "drift_pruning": "pen_max 2 mm against the 1 mm surface vertex set (density of nontextured.stl[::10])",
            "l2": "GPU arm: flushed (256 MiB write + read-back, untimed) before every device-timed step; codebook-side tables (keys, drift-test "
                  "vertices, weight tables) are loaded with an L2 evict_last policy and the particle arrays with evict_first, by design; "
                  "the e2e leg does not flush",
            "parallelism": f"particles sharded x{world}"}


def run_ours(args):
    import ctypes as C

    import torch.distributed as dist

    from midastouch_b200 import synth
    from midastouch_b200._lib import call
    from midastouch_b200.engine import FilterEngine, prepare_odom
    from midastouch_b200.tactile_tree import tactile_tree

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # NCCL prints its version banner on stdout: keep fd 1 clean for the one JSON line
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    obj, cbs, gt, meas = make_assets()
    cb = tactile_tree(cbs.poses, cbs.cam_poses, cbs.embeddings)
    n = N_PER_GPU
    # sharded: children follow their parents, so the shards drift apart while the cloud converges (the share of a GPU
    # settles at the share of the eventual survivors it happened to hold: +-20 % at 8 GPUs); 50 % headroom, evened out
    # every 64 steps (FilterEngine.rebalance); an overflow raises
    cap = n + (n // 2 if world > 1 else 0)
    cb.to_device(dev)
    # one context for everything this process will run on the codebook (the sharded check's single-GPU engine holds all
    # world x 1e6 particles on rank 0): sized once, up front
    cb.ctx.ensure_capacity(max(cap, n * world if (world > 1 and rank == 0 and not args.no_shard_check) else cap))
    eng = FilterEngine(cb, capacity=cap, sig_t=2e-4, sig_r=0.5, seed=1234, rank=rank, world=world, n_global=n * world,
                       mesh_vertices=obj.vertices, pen_max=0.002)
    eng.use_graph = not args.no_graph
    # particles start on codebook poses (what init_filter + the SE3_NN snap of filter.py:159-160 produce)
    g = torch.Generator().manual_seed(100 + (0 if os.environ.get("MT_BENCH_SAME_SEED") else rank))
    sel = torch.randint(0, M, (n,), generator=g)
    poses0, hint0 = cbs.poses.to(dev)[sel.to(dev)], sel.int().to(dev)

    def restart():
        """every measured phase is a fresh filter run: particles spread over the whole codebook"""
        eng.load_particles(poses0, nn_hint=hint0, spatial_sort=not args.no_sort)
        eng.t = 0
        eng.use_n_dev = False
        eng.ctx.stats(reset=True)

    restart()

    odoms = [prepare_odom(torch.inverse(meas[t - 1]) @ meas[t]) for t in range(1, T_TRAJ)]
    gts = [gt[t].float().contiguous() for t in range(T_TRAJ)]
    # tactile codes: pinned host buffers, one per frame (the step's only per-frame tensor input)
    # frame t+1's code = smooth embedding of the true pose + noise (what a trained TCN would return)
    codes_h = [synth.make_pose_query(gt[t + 1], D, seed=3, frame=t).pin_memory() for t in range(T_TRAJ - 1)]
    codes_d = [c.to(dev) for c in codes_h]
    us = torch.rand(4096, generator=torch.Generator().manual_seed(7)).tolist()
    l2buf = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def l2flush():
        """L2 flush between timed steps (untimed): write a 256 MiB buffer (> 126 MB of L2), then read it so that the cache
        is left holding clean lines -- otherwise the first timed kernel also pays for writing back the flush's dirty lines."""
        l2buf.zero_()
        l2buf.sum()

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    RUN = T_TRAJ - 6  # frames of one filter run; longer benchmarks restart the filter (untimed) and slide again

    def timed_pass(first, count, per_kernel):
        """`count` steps starting at frame `first` of the current filter run, L2 flushed before every step (untimed), one
        CUDA-event pair per step on the launching stream.  per_kernel: stream launches with the library's events between
        the kernels (roofline of k_step_a); else the step is ONE graph replay."""
        eng.use_graph = (not args.no_graph) and not per_kernel
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(6 if per_kernel else 2)] for _ in range(count)]
        for e in evs:
            for x in e:
                x.record()  # materialises the cudaEvent_t handles
        sync()
        for t in range(count):
            k = (first + t) % RUN
            if k == 0 and t:
                restart()
            l2flush()
            e = evs[t]
            if per_kernel:
                call("mt_ctx_set_timing_events", eng.ctx.h, (C.c_void_p * 4)(*[x.cuda_event for x in e[1:5]]))
            e[0].record()
            eng.step(codes_d[k], odoms[k], u=us[(first + t) % 4096])
            e[-1].record()
        sync()
        if per_kernel:
            call("mt_ctx_set_timing_events", eng.ctx.h, None)
        eng.use_graph = not args.no_graph
        return evs

    # ---- warm-up (graph instantiation for both buffer parities, allocator, flush)
    for t in range(args.warmup):
        l2flush()
        eng.step(codes_d[t], odoms[t], u=us[t])
    sync()
    # ---- (1) the headline: K steps, each ONE CUDA-graph replay (query | motion+SE3_NN -> queue consumers -> resampling)
    with ClockSampler(local) as clk:
        evs = timed_pass(args.warmup, args.steps, per_kernel=False)
    ms = [e[0].elapsed_time(e[1]) for e in evs]
    total_ms = torch.tensor([sum(ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    value = n * world * args.steps / (total_ms * 1e-3)
    replays = C.c_longlong(0)
    ncached = C.c_int(0)
    call("mt_step_graph_info", eng.ctx.h, C.byref(replays), C.byref(ncached))
    xdbg = (C.c_ulonglong * 3)()
    call("mt_dist_debug", eng.ctx.h, xdbg)
    stats_loop = eng.ctx.stats(reset=True)
    # ---- (2) the same steps again as stream launches with events between the kernels: roofline of k_step_a
    restart()
    for t in range(args.warmup):
        l2flush()
        eng.step(codes_d[t], odoms[t], u=us[t])
    kb = min(args.steps, 40)
    evk = timed_pass(args.warmup, kb, per_kernel=True)
    k_a = sum(e[1].elapsed_time(e[2]) for e in evk) / kb
    k_q = sum(e[2].elapsed_time(e[3]) for e in evk) / kb
    k_b = sum(e[3].elapsed_time(e[5]) for e in evk) / kb   # includes any wait for the side-stream query
    k_step = sum(e[0].elapsed_time(e[5]) for e in evk) / kb
    k_max = {"k_step_a": max(e[1].elapsed_time(e[2]) for e in evk), "queue consumers": max(e[2].elapsed_time(e[3]) for e in evk),
             "k_step_bw": max(e[3].elapsed_time(e[5]) for e in evk)}
    # ---- (3) converged cloud: frames 120.. of the same run (the cloud sits on a few hundred codebook poses)
    conv = None
    if not args.no_converged:
        for t in range(args.warmup + kb, 120):
            eng.step(codes_d[t], odoms[t], u=us[t])
        evc = timed_pass(120, 30, per_kernel=True)
        evg = timed_pass(150, 30, per_kernel=False)
        conv_ms = torch.tensor([sum(e[0].elapsed_time(e[1]) for e in evg)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(conv_ms, op=dist.ReduceOp.MAX)
        conv = {"frames": "150..180 (graph replays) / 120..150 (per-kernel events)", "value": n * world * 30 / (float(conv_ms.item()) * 1e-3),
                "ms_per_step": float(conv_ms.item()) / 30, "step_ms_every_3rd": [round(e[0].elapsed_time(e[1]), 4) for e in evg[::3]],
                "stream_form_step_ms": sum(e[0].elapsed_time(e[5]) for e in evc) / 30,
                "k_step_a_ms": sum(e[1].elapsed_time(e[2]) for e in evc) / 30,
                "k_step_a_frac": A_BYTES_PER_UPDATE * n / (sum(e[1].elapsed_time(e[2]) for e in evc) / 30 * 1e-3) / 1e9 / peaks()[0]}
    per_rank = None
    if world > 1:
        mine = torch.tensor([k_a, k_q, k_b, sum(ms) / args.steps, float(eng.count()), float(xdbg[2] - xdbg[1]), float(xdbg[2] - xdbg[0])],
                            dtype=torch.float64, device=dev)
        allr = torch.zeros(world * 7, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allr, mine)
        per_rank = [{"k_step_a_ms": r[0], "queue_consumers_ms": r[1], "k_step_bw_ms (stream form, incl. waiting for query and peers)": r[2],
                     "graph_step_ms": r[3], "particles_at_end": int(r[4]), "last_step_peer_wait_ns (sums sent -> all sums received)": int(r[5]),
                     "last_step_exchange_ns (barrier exit -> all sums received)": int(r[6])} for r in allr.reshape(world, 7).tolist()]

    # ---- codebook query kernel alone (it overlaps the motion/NN kernel inside a step)
    q_ms = query_time(eng, codes_d, l2flush, min(args.steps, 20))

    # ---- end to end through the public API: every step the tactile code (pinned host) + odometry +
    # ground truth go host->device and the step's result (rmse, 8 bytes) comes back to the host.  The
    # read-back is asynchronous into pinned memory and consumed one step later, so the host prepares
    # step t+1 while the GPU runs step t (the reference's loop reads .item() synchronously instead).
    sync()
    restart()
    for t in range(args.warmup):
        eng.step(codes_h[t], odoms[t], u=us[t], gt=gts[t + 1])
    sync()
    res_pinned = [torch.zeros(2, dtype=torch.float32).pin_memory() for _ in range(2)]
    res_ev = [None, None]
    results = []
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for t in range(args.steps):
        k = (args.warmup + t) % RUN
        if k == 0 and t:
            restart()  # (timed here: ~1 ms once per 250 steps)
        eng.step(codes_h[k], odoms[k], u=us[k % 4096], gt=gts[k + 1])  # the code travels host -> device on the engine's copy stream
        res_ev[t & 1] = eng.read_rmse_async(res_pinned[t & 1])        # and the result device -> host, also beside the kernels
        if t:
            res_ev[(t - 1) & 1].synchronize()
            results.append(float(res_pinned[(t - 1) & 1][0]))
    res_ev[(args.steps - 1) & 1].synchronize()
    results.append(float(res_pinned[(args.steps - 1) & 1][0]))
    t1.record()
    torch.cuda.current_stream().wait_stream(eng._io)
    sync()
    assert len(results) == args.steps and all(r == r for r in results)
    e2e_ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e = n * world * args.steps / (float(e2e_ms.item()) * 1e-3)

    # ---- sharded runs: one teacher-forced check that the concatenated shards equal the single-GPU engine bit for bit
    shard_check = sharded_check(eng, cb, obj, cbs, gt, meas, dev, rank, world) if (world > 1 and not args.no_shard_check) else None

    # ---- tactile code network (one forward per frame; random weights, 4096-point contact patch)
    tcn_ms = tcn_time(dev) if rank == 0 else None
    # ---- tactile depth network (image -> height map -> contact mask; cuDNN, random weights)
    tdn_ms = tdn_time(dev) if rank == 0 else None
    # ---- batched codebook query on the tensor cores (tcgen05, 3xTF32): 1024 codes against the codebook
    gemm = gemm_time(cb, dev) if rank == 0 else None
    # ---- a whole frame: tactile code network on the frame's 4096-point cloud, then the filter step, in stream order
    frame = None
    if rank == 0 and world == 1:
        tcn_net, tcn_cloud = tcn_time.last
        restart()
        for t in range(args.warmup):
            tcn_net.embed_clouds(tcn_cloud)
            eng.step(codes_d[t], odoms[t], u=us[t], gt=gts[t + 1])
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync()
        nf = min(args.steps, T_TRAJ - 1 - args.warmup)
        f0.record()
        for t in range(args.warmup, args.warmup + nf):
            tcn_net.embed_clouds(tcn_cloud)  # (random weights: the code it returns is discarded, the filter is fed the synthetic code of the frame)
            eng.step(codes_d[t], odoms[t], u=us[t], gt=gts[t + 1])
        f1.record()
        sync()
        fms = f0.elapsed_time(f1) / nf
        frame = {"ms_per_frame": fms, "particle_updates_per_s": n / (fms * 1e-3), "frames": nf,
                 "what": "mt_tcn_embed (MinkLoc3D on one 4096-point cloud) + one filter step per frame, same stream, L2 not flushed"}

    if rank == 0:
        peak, how = peaks()
        a_gbs = A_BYTES_PER_UPDATE * n / (k_a * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("k_step_a", {}).get("dram_bytes_per_launch")
        cfg = workload_config(world, args.warmup, args.steps, args.no_sort)
        run_info = {"frames": f"{args.warmup}..{args.warmup + args.steps} of a filter run from global initialisation",
                    "noise": "in-kernel Philox4x32-10", "resampling": "systematic (low_var), fused with the weighting",
                    "particle_order": "random" if args.no_sort else "sorted by the rank of the matched codebook key in the search index (k-d tree leaf order) at load",
                    "launch": "one CUDA-graph replay per step" if not args.no_graph else "stream launches"}
        out = {
            "metric": "particle-updates/sec", "value": value, "unit": "particle-updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 poses / f64 weights+prefix",
            "data": "synthetic (seeded stand-ins for YCB-Slide assets; no datasets offline)",
            "config": cfg, "run": run_info,
            "roofline": {"bound": "hbm", "achieved": a_gbs, "peak": peak, "unit": "GB/s", "frac": a_gbs / peak,
                         "traffic": traffic, "traffic_source": "profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel (round 2), not measured in this run",
                         "peak_source": how, "kernel": "k_step_a (motion + drift voxel test + hint-graph SE3_NN)",
                         "algorithmic_bytes_per_launch": A_BYTES_PER_UPDATE * n, "avg_launch_ms": k_a,
                         "timed": f"CUDA events between the kernels, stream-launch form, frames {args.warmup}..{args.warmup + kb}",
                         "sweep": {"kernels_ms": {"k_step_a": k_a, "queue consumers (k_step_meshq + k_step_meshq2 + k_step_nnq)": k_q,
                                                  "k_step_bw (sums + resample; incl. wait for the query)": k_b},
                                   "kernels_ms_max": k_max, "stream_form_step_ms": k_step, "graph_form_step_ms": total_ms / args.steps,
                                   "algorithmic_bytes_per_step": ALGO_BYTES_PER_UPDATE * n,
                                   "achieved": ALGO_BYTES_PER_UPDATE * n * world / (total_ms / args.steps * 1e-3) / 1e9 / world,
                                   "frac": ALGO_BYTES_PER_UPDATE * n / (total_ms / args.steps * 1e-3) / 1e9 / peak},
                         "codebook_query": {"kernel": "k_codebook_query<double> (graph branch parallel to k_step_a)", "ms": q_ms, "algorithmic_bytes": M * D * 8 + D * 8 + M * 16,
                                            "achieved": (M * D * 8 + D * 8 + M * 16) / (q_ms * 1e-3) / 1e9,
                                            "frac": (M * D * 8 + D * 8 + M * 16) / (q_ms * 1e-3) / 1e9 / peak}},
            "e2e": {"value": e2e, "unit": "particle-updates/s", "h2d_bytes_per_step": D * 8 + 64 + 64 + 4, "d2h_bytes_per_step": 8,
                    "readback": "rmse of every step, asynchronous into pinned memory on the engine's copy stream (FilterEngine.read_rmse_async), consumed one step later; the code is uploaded on the same copy stream, double-buffered", "l2": "not flushed"},
            "converged_cloud": conv,
            "tcn_forward_ms": tcn_ms, "tdn_forward_ms": tdn_ms, "frame_with_tcn": frame, "codebook_gemm": gemm,
            "gpu_launches": int(replays.value) * (6 if eng.prune else 4) if not args.no_graph else None,
            "graph": {"replays_so_far": int(replays.value), "instantiated": int(ncached.value), "kernel_nodes_per_replay": 6 if eng.prune else 4},
            "clocks": clk.summary(),
            "filter": {"rmse_t_mm_last_e2e_step": 1e3 * results[-1], "rmse_t_mm_first_e2e_step": 1e3 * results[0],
                       "step_ms_every_5th": [round(x, 4) for x in ms[::5]]},
            "per_rank": per_rank, "sharded_bit_exact": None if shard_check is None else shard_check["ok"], "sharded_check": shard_check,
            "parity": {"oracle": "oracle/oracle.py restates the path; pinned to the unmodified reference module by tests/golden/*.npz",
                       "unpinned": "theseus SO3 log/quaternion thresholds, pynanoflann accumulation order, MinkowskiEngine (third-party, absent): see DESIGN.md"},
            "engine_stats": {"nn_box_searches_per_step": stats_loop["nn_fallbacks"] / (args.steps + args.warmup),
                             "drift_tests_deferred_per_step": stats_loop["mesh_deferred"] / (args.steps + args.warmup),
                             "on_surface_last_step": stats_loop["on_surface"], "overflow": stats_loop["overflow"],
                             "leaves_max_one_search": stats_loop["grid_rows_max"]},
        }
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline(budget_s=20.0)
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def sharded_check(eng_unused, cb, obj, cbs, gt, meas, dev, rank, world, steps=2):
    """N>1: `steps` teacher-forced filter steps (same noise arrays, same offsets) on (a) the sharded engines, 1e6
    particles per GPU, strongly skewed weights so that the shard sizes drift, never reading the counts back between
    the steps, and (b) ONE engine holding all world x 1e6 particles on rank 0.  The concatenated shards must equal (b)
    bit for bit (poses and matches)."""
    import torch.distributed as dist

    from midastouch_b200 import synth
    from midastouch_b200.engine import FilterEngine

    n = N_PER_GPU
    N = n * world
    gen = torch.Generator(device=dev).manual_seed(4242)  # the same stream on every GPU
    sel = torch.randint(0, M, (N,), generator=gen, device=dev)
    poses = cbs.poses.to(dev)[sel]
    noise = [(2e-4 * torch.randn(N, 3, generator=gen, device=dev), 0.5 * torch.randn(N, 3, generator=gen, device=dev)) for _ in range(steps)]
    us = [0.37, 0.81, 0.11, 0.5][:steps]
    # a sharp query: weights differ by the full factor e between the two ends of the object
    q = synth.make_pose_query(gt[40], D, seed=3, frame=40).to(dev)
    odom = torch.inverse(meas[0]) @ meas[1]
    lo = rank * n
    eng = FilterEngine(cb, capacity=n + n // 2, sig_t=2e-4, sig_r=0.5, seed=1, rank=rank, world=world, n_global=N,
                       mesh_vertices=obj.vertices, pen_max=0.002)
    eng.rebalance_every = 0
    eng.load_particles(poses[lo:lo + n], nn_hint=sel[lo:lo + n].int())
    # teacher forcing needs every rank's noise slice to follow its (drifting) shard: the children of step k are ordered
    # globally, so shard r of step k+1 starts at the sum of the previous shards' counts -- taken from the device counts
    # by an all-gather on the device (no host read of the local count before the step is enqueued)
    off = torch.tensor([lo], dtype=torch.int64, device=dev)
    cnt_hist = []
    for k in range(steps):
        o = int(off.item()) if k else lo   # (k = 0: known; later: one host read per step, after the previous step)
        c = n if k == 0 else int(eng.n_dev[eng.cur].item())
        eng.step(q, odom, u=us[k], tn=noise[k][0][o:o + c].contiguous(), rot=noise[k][1][o:o + c].contiguous())
        cnts = torch.zeros(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(cnts, eng.n_dev[eng.cur].reshape(1))
        off = cnts[:rank].sum().reshape(1)
        cnt_hist.append(cnts.tolist())
    n_loc = eng.count()
    mine = torch.cat([eng.poses().reshape(n_loc, 16), eng.nn_idx().float().reshape(n_loc, 1)], 1)
    parts = [torch.zeros((int(c), 17), device=dev) for c in cnt_hist[-1]] if rank == 0 else None
    # gather to rank 0 (variable sizes): send / recv
    if rank == 0:
        parts[0] = mine
        for r in range(1, world):
            dist.recv(parts[r], src=r)
    else:
        dist.send(mine, dst=0)
    ok, detail = True, None
    if rank == 0:
        full = torch.cat(parts)
        del parts
        ref = FilterEngine(cb, capacity=N, sig_t=2e-4, sig_r=0.5, seed=1, mesh_vertices=obj.vertices, pen_max=0.002)
        ref.load_particles(poses, nn_hint=sel.int())
        for k in range(steps):
            ref.step(q, odom, u=us[k], tn=noise[k][0], rot=noise[k][1])
        want = torch.cat([ref.poses().reshape(N, 16), ref.nn_idx().float().reshape(N, 1)], 1)
        same = full.shape == want.shape and bool(torch.equal(full, want))
        ok = same
        detail = {"ok": same, "steps": steps, "particles": N, "children_per_rank": cnt_hist,
                  "what": "poses + codebook matches of the concatenated shards == one engine holding all particles (teacher-forced noise, skewed weights)"}
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    return detail if rank == 0 else {"ok": bool(flag.item())}


def tcn_time(dev, reps=20):
    """device time of one TCN forward (mt_tcn_forward: MinkLoc3D on one 4096-point cloud), random weights"""
    import numpy as np

    from midastouch_b200.tcn import TCN
    import types

    rng = np.random.default_rng(0)
    m = types.SimpleNamespace(tcn_weights="", model="MinkFPN", num_points=4096, batch_size=100, mink_quantization_size=0.001,
                              planes="32,64,64", layers="1,1,1", num_top_down=1, conv0_kernel_size=5, feature_size=256, output_dim=256)
    P = {}

    def conv(name, kvol, cin, cout):
        w = torch.from_numpy(rng.normal(size=(kvol, cin, cout)).astype("float32") * (2.0 / (kvol * cin)) ** 0.5)
        P[name] = w if kvol > 1 else w[0]

    def bn(name, c):
        P[f"{name}.bn.weight"], P[f"{name}.bn.bias"] = torch.ones(c), torch.zeros(c)
        P[f"{name}.bn.running_mean"], P[f"{name}.bn.running_var"] = torch.zeros(c), torch.ones(c)

    conv("backbone.conv0.kernel", 125, 1, 32), bn("backbone.bn0", 32)
    inpl = 32
    for s_, pl in enumerate((32, 64, 64)):
        conv(f"backbone.convs.{s_}.kernel", 8, inpl, inpl), bn(f"backbone.bn.{s_}", inpl)
        b = f"backbone.blocks.{s_}.0"
        conv(f"{b}.conv1.kernel", 27, inpl, pl), bn(f"{b}.norm1", pl), conv(f"{b}.conv2.kernel", 27, pl, pl), bn(f"{b}.norm2", pl)
        if inpl != pl:
            conv(f"{b}.downsample.0.kernel", 1, inpl, pl), bn(f"{b}.downsample.1", pl)
        inpl = pl
    conv("backbone.conv1x1.0.kernel", 1, 64, 256), conv("backbone.tconvs.0.kernel", 8, 256, 256), conv("backbone.conv1x1.1.kernel", 1, 64, 256)
    P["pooling.p"] = torch.tensor([3.0])
    tcn = TCN(types.SimpleNamespace(model=m, train=types.SimpleNamespace(normalize_embeddings=True)), device=dev, weights=P)
    xy = rng.uniform(-1, 1, size=(4096, 2))
    z = 0.3 * (xy[:, 0] ** 2 + xy[:, 1] ** 2) - 0.2
    cloud = torch.from_numpy(np.concatenate([xy, z[:, None]], 1).astype("float32")).to(dev)[None]
    tcn_time.last = (tcn, cloud)  # (scripts/tcn_prof.py)
    for _ in range(3):
        tcn.embed_clouds(cloud)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        tcn.embed_clouds(cloud)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def tdn_time(dev, reps=20):
    """device time of TDN.image2heightmap + heightmap2mask on one 320 x 240 tactile frame (FCRN-ResNet50-UpProj through
    cuDNN, BatchNorms folded, fused up-projections; random weights), host image in: normalisation and upload included"""
    import numpy as np
    import types

    from midastouch_b200.tdn import TDN, fcrn_parameter_shapes

    g = torch.Generator().manual_seed(0)
    S = {}
    for name, shape in fcrn_parameter_shapes().items():
        if len(shape) == 4:
            S[name] = torch.randn(shape, generator=g) * (2.0 / (shape[2] * shape[3] * shape[0])) ** 0.5
        elif name.endswith("running_var") or name.endswith(".weight"):
            S[name] = torch.ones(shape)
        else:
            S[name] = torch.zeros(shape, dtype=torch.long if name.endswith("tracked") else torch.float32)
    fc = types.SimpleNamespace(blend_sz=0, border=1, ratio=0.2, clip=5, batch_size=1)
    rng = np.random.default_rng(0)
    img = rng.integers(0, 255, (320, 240, 3)).astype(np.uint8)
    t = TDN(types.SimpleNamespace(tdn_weights="", fcrn=types.SimpleNamespace(sim=fc, real=fc)), bg=np.zeros((320, 240), np.float32), device=dev, weights=S)
    for _ in range(3):
        t.heightmap2mask(t.image2heightmap(img))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        t.heightmap2mask(t.image2heightmap(img))
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def gemm_time(cb, dev, nq=1024, reps=20):
    """k_codebook_gemm_tma (TMA-fed tcgen05 / TMEM, persistent, warp-specialised): nq codes x M rows; the float64 codebook's float32
    split planes are built once per upload, the queries' planes per call (inside the timed region)."""
    M_, D_ = cb.embeddings.shape
    Qb = torch.rand(nq, D_, generator=torch.Generator().manual_seed(5)).to(dev)
    for _ in range(3):
        cb.query_batched(Qb)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        cb.query_batched(Qb)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flops = 2.0 * M_ * D_ * nq
    peak_bf16, src = tensor_peak()
    tensor_tflops = 3.0 * flops / (ms * 1e-3) / 1e12
    return {"kernel": "k_split_planes(queries) + k_codebook_gemm_tma (cp.async.bulk.tensor -> 3-stage smem ring -> tcgen05.mma kind::tf32, 3 MMAs per product: 3xTF32)", "nq": nq, "ms": ms,
            "algorithmic_tflops": flops / (ms * 1e-3) / 1e12,
            "roofline": {"bound": "tensor", "achieved": tensor_tflops, "peak": peak_bf16 / 2, "unit": "TFLOP/s",
                         "frac": tensor_tflops / (peak_bf16 / 2),
                         "peak_source": f"{src} bf16 dense peak / 2 (TF32 runs at half the bf16 rate)"}}


def query_time(eng, codes_d, l2flush, reps):
    """average device time of the codebook query (k_to_f64 + k_cosine_rows) with the L2 flushed first."""
    from midastouch_b200._lib import call, ptr, stream_ptr
    from midastouch_b200.context import dtype_code

    acc = 0.0
    for r in range(reps):
        q = codes_d[r % len(codes_d)].reshape(-1).contiguous()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l2flush()
        e0.record()
        call("mt_codebook_query", eng.ctx.h, ptr(q), dtype_code(q), 0, stream_ptr())
        e1.record()
        torch.cuda.synchronize()
        acc += e0.elapsed_time(e1)
    return acc / reps


class _CpuArm:
    """The reference's loop body (filter.py:154-190) on the host cores, on the bench workload.

    kind "reference": the UNMODIFIED reference functions -- particle_filter.motionModel, get_similarity,
    remove_invalid_particles, resampler("weighted_random", the loop's default) and pose.get_logmap_from_matrix --
    from the verbatim copies under oracle/_ref (oracle/build_ref.py), loaded through oracle/ref_shim.py (stub modules
    for the absent trimesh / theseus / omegaconf; theseus SO3 log / quaternion delegate to the oracle's closed forms).
    tactile_tree.SE3_NN needs pynanoflann (absent): its k-d tree query is served by scipy.cKDTree over the same
    R3_SE3 keys, with all host threads, followed by the reference's gather of the matched embeddings.
    kind "port": the same sequence on the oracle's restatement, when oracle/_ref did not travel."""

    def __init__(self, n):
        import numpy as np
        from scipy.spatial import cKDTree

        from midastouch_b200 import synth
        from oracle import oracle as O
        from oracle import ref_shim

        torch.set_num_threads(os.cpu_count() or 1)  # (torchrun exports OMP_NUM_THREADS=1)
        self.n, self.O, self.synth = n, O, synth
        self.obj, self.cbs, self.gt, self.meas = make_assets()
        self.kind = "port"
        self.pf = None
        if ref_shim.reference_available():
            try:
                pfm, posem = ref_shim.load_reference()
                vpath = os.path.join("/tmp", f"mt_bench_vertices_{os.getpid()}.npy")
                np.save(vpath, self.obj.vertices)
                # downsample=1: obj.vertices already has the density of nontextured.stl[::10] (the drift test's vertex set)
                self.pf = pfm.particle_filter(ref_shim.default_cfg(num_particles=n), vpath, 1.0, False, 1)
                os.remove(vpath)
                self.pfm, self.posem = pfm, posem
                self.kind = "reference"
            except Exception as e:  # noqa: BLE001
                print(f"[bench] reference modules not usable ({e}); timing the oracle port", file=sys.stderr)
        self.keys = O.r3_se3(self.cbs.poses)
        self.tree = cKDTree(self.keys.numpy().astype("float64"))
        g = torch.Generator().manual_seed(100)
        sel = torch.randint(0, M, (n,), generator=g)
        self.poses = self.cbs.poses[sel].clone()
        self.t = 0

    def step(self):
        """one filter step; returns its wall time"""
        O, n = self.O, self.n
        k = self.t % (T_TRAJ - 2)
        odom = torch.inverse(self.meas[k]) @ self.meas[k + 1]
        q = self.synth.make_pose_query(self.gt[k + 1], D, seed=3, frame=k)
        t0 = time.perf_counter()
        if self.kind == "reference":
            pfm, pf = self.pfm, self.pf
            parts = pf.motionModel(pfm.Particles(self.poses), odom)                                   # filter.py:154-155
            logm = self.posem.get_logmap_from_matrix(parts.poses[:, :3, :3])                                  # R3_SE3, tactile_tree.py:73-77
            qk = torch.cat(((1.0 - 0.01) * parts.poses[:, :3, 3], 0.01 * logm), dim=1).numpy().astype("float64")
            _, idx = self.tree.query(qk, k=1, workers=-1)                                            # SE3_NN (stand-in for nanoflann)
            nn_codes = self.cbs.embeddings[torch.from_numpy(idx)]                                    # tactile_tree.py:55-57
            parts.weights = pf.get_similarity(q, nn_codes, softmax=True)                             # filter.py:170-173
            parts, _ = pf.remove_invalid_particles(parts)                                            # filter.py:176
            parts = pf.resampler(parts, resample="weighted_random")                                  # filter.py:190
            self.poses = parts.poses
        else:
            tn, rot = O.draw_motion_noise(n, 2e-4, 0.5)
            moved, _ = O.motion_model(self.poses, odom, tn, rot)
            qk = O.r3_se3(moved).numpy().astype("float64")
            _, idx = self.tree.query(qk, k=1, workers=-1)
            w = O.get_similarity(q, self.cbs.embeddings[torch.from_numpy(idx)], True)
            w, _ = O.remove_invalid(moved, w, self.obj.vertices, 0.002)
            anc = torch.multinomial(w / w.sum(), n, replacement=True)
            self.poses = moved[anc]
        self.t += 1
        return time.perf_counter() - t0

    def describe(self, done):
        what = ("the unmodified reference functions (particle_filter.motionModel / get_similarity / remove_invalid_particles / "
                "resampler('weighted_random'), pose.get_logmap_from_matrix) from oracle/_ref" if self.kind == "reference"
                else "the oracle port of the reference functions (oracle/_ref did not travel)")
        return (f"{done} steps of N={self.n} particles (codebook M={M}, D={D} f64): {what}; SE3_NN's nanoflann query served by "
                f"scipy.cKDTree (workers=all) over the same keys; theseus SO3 log by the oracle's closed form")


def cpu_baseline(budget_s=20.0, n=N_PER_GPU):
    """cpu_baseline of the N=1 line: a bounded sample (about budget_s seconds) of the same workload on the host cores"""
    arm = _CpuArm(n)
    arm.step()  # warm-up (allocations, thread pools)
    done, t_total = 0, 0.0
    while t_total < budget_s:
        t_total += arm.step()
        done += 1
    return {"value": n * done / t_total, "unit": "particle-updates/s", "cores": os.cpu_count(),
            "torch_threads": torch.get_num_threads(), "kind": arm.kind, "sample": arm.describe(done)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.manual_seed(0)
    # a step of the reference at N = 1e6 takes seconds (an N x D float64 gather of 2 GB among other things): the run is a
    # bounded sample -- at most REF_BUDGET_S seconds of timed steps, at least 2 -- of the same workload
    n = N_PER_GPU
    budget = float(os.environ.get("MT_REF_BUDGET_S", "150"))
    arm = _CpuArm(n)
    t_start = time.perf_counter()
    for _ in range(max(1, min(args.warmup, 1))):
        arm.step()
    done, t_total = 0, 0.0
    while done < args.steps and (done < 2 or t_total + t_total / done < budget):
        t_total += arm.step()
        done += 1
    wall = time.perf_counter() - t_start
    value = n * done / t_total
    cb = {"value": value, "unit": "particle-updates/s", "cores": os.cpu_count(), "torch_threads": torch.get_num_threads(),
          "kind": arm.kind, "sample": arm.describe(done) + f"; {done} of the requested {args.steps} steps timed (budget {budget:.0f} s)"}
    out = {"impl": "reference", "metric": "particle-updates/sec", "value": value, "unit": "particle-updates/s",
           "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * n / value,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 poses / f64 weights",
           "data": "synthetic (seeded stand-ins for YCB-Slide assets; no datasets offline)",
           "config": workload_config(world, args.warmup, args.steps),
           "cpu_baseline": cb,
           "e2e": {"value": value, "unit": "particle-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "steps_timed": done, "wall_s": wall}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-sort", action="store_true", help="keep the particles in random order (no spatial sort at load)")
    ap.add_argument("--no-graph", action="store_true", help="stream launches instead of CUDA-graph replays")
    ap.add_argument("--no-converged", action="store_true", help="skip the converged-cloud pass")
    ap.add_argument("--no-shard-check", action="store_true", help="N>1: skip the sharded-vs-single-GPU bit-exactness check")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

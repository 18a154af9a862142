"""BASELINE.json configs 2 and 5 on one GPU (diagnostic sweeps, written to gpurun_out/sweeps.json):
  * codebook query GB/s, M = 50 000, D in {256, 512}, embeddings stored float32 / float64, L2 cold and warm;
  * per-step latency of the engine over N per GPU in 2^16 .. 2^21 (cotter-pin stand-in, sigma_t = 1e-4).
Timing: CUDA events on the launching stream, L2 flushed (256 MiB write + read-back) before every cold sample."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from midastouch_b200 import synth
from midastouch_b200._lib import call, ptr, stream_ptr
from midastouch_b200.context import dtype_code
from midastouch_b200.engine import FilterEngine, prepare_odom
from midastouch_b200.tactile_tree import tactile_tree

dev = torch.device("cuda:0")
peak, peak_src = bench.peaks()
l2buf = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)


def flush():
    l2buf.zero_()
    l2buf.sum()


def ev_time(fn, reps, cold):
    acc = 0.0
    for _ in range(reps):
        if cold:
            flush()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        acc += e0.elapsed_time(e1)
    return acc / reps


out = {"peak_GBps": peak, "peak_source": peak_src, "codebook_query": [], "latency_sweep": []}
# ---- config 2: codebook query
box = synth.make_object("004_sugar_box")
for D in (256, 512):
    cbs = synth.make_codebook(box, M=50000, D=D, seed=3, embedding="smooth")
    for dt in (torch.float32, torch.float64):
        cb = tactile_tree(cbs.poses, cbs.cam_poses, cbs.embeddings.to(dt))
        cb.to_device(dev)
        q = synth.make_pose_query(cbs.poses[7], D, seed=3).to(dev).to(dt).reshape(-1).contiguous()
        fn = lambda: call("mt_codebook_query", cb.ctx.h, ptr(q), dtype_code(q), 0, stream_ptr())
        for _ in range(5):
            fn()
        esz = 4 if dt == torch.float32 else 8
        nbytes = 50000 * D * esz + D * esz + 50000 * 8 * 3  # rows + query + (cached norm in, cos + exp out) per row
        for cold in (True, False):
            ms = ev_time(fn, 30, cold)
            out["codebook_query"].append({"M": 50000, "D": D, "stored": str(dt).split(".")[-1], "l2": "cold" if cold else "warm",
                                          "us": 1e3 * ms, "algorithmic_bytes": nbytes, "GBps": nbytes / ms / 1e6,
                                          "frac_of_peak": nbytes / ms / 1e6 / peak})
        del cb
# ---- config 5: per-step latency over N per GPU
pin = synth.make_object("cotter-pin")
cbs = synth.make_codebook(pin, M=50000, D=256, seed=5, embedding="smooth")
cb = tactile_tree(cbs.poses, cbs.cam_poses, cbs.embeddings)
cb.to_device(dev)
gt, meas = synth.make_trajectory(pin, T=64, seed=5, step=1e-4)
odoms = [prepare_odom(torch.inverse(meas[t - 1]) @ meas[t]) for t in range(1, 64)]
codes = [synth.make_pose_query(gt[t + 1], 256, seed=5, frame=t).to(dev) for t in range(63)]
for lg in range(16, 22):
    n = 1 << lg
    eng = FilterEngine(cb, capacity=n, sig_t=1e-4, sig_r=0.5, seed=1, mesh_vertices=pin.vertices)
    g = torch.Generator().manual_seed(lg)
    sel = torch.randint(0, 50000, (n,), generator=g)
    eng.load_particles(cbs.poses.to(dev)[sel.to(dev)], nn_hint=sel.int().to(dev), spatial_sort=True)
    for t in range(10):
        eng.step(codes[t], odoms[t], u=0.3)
    import ctypes as C
    ts, ka, kq, kb = [], [], [], []
    for t in range(10, 60):
        flush()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        for x in e:
            x.record()
        call("mt_ctx_set_timing_events", eng.ctx.h, (C.c_void_p * 4)(*[x.cuda_event for x in e[1:5]]))
        e[0].record()
        eng.step(codes[t], odoms[t], u=0.37)
        e[5].record()
        torch.cuda.synchronize()
        ts.append(1e3 * e[0].elapsed_time(e[5]))
        ka.append(1e3 * e[1].elapsed_time(e[2])), kq.append(1e3 * e[2].elapsed_time(e[3])), kb.append(1e3 * e[4].elapsed_time(e[5]))
    call("mt_ctx_set_timing_events", eng.ctx.h, None)
    med = lambda a: sorted(a)[len(a) // 2]
    ts.sort()
    out["latency_sweep"].append({"n_per_gpu": n, "us_median": ts[len(ts) // 2], "us_p10": ts[5], "us_p90": ts[45],
                                 "us_k_step_a": med(ka), "us_k_step_nnq": med(kq), "us_k_step_bw": med(kb),
                                 "updates_per_s": n / (ts[len(ts) // 2] * 1e-6), "stats": eng.ctx.stats(reset=True)})
    del eng
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/sweeps.json", "w"), indent=1)
for r in out["codebook_query"]:
    print("query D=%d %s %s: %.1f us  %.0f GB/s (%.2f of peak)" % (r["D"], r["stored"], r["l2"], r["us"], r["GBps"], r["frac_of_peak"]))
for r in out["latency_sweep"]:
    print("N=%8d  median %.1f us  p10 %.1f p90 %.1f  (A %.1f, nnq %.1f, bw %.1f)  %.3g updates/s  fallbacks/step %.0f" % (
        r["n_per_gpu"], r["us_median"], r["us_p10"], r["us_p90"], r["us_k_step_a"], r["us_k_step_nnq"], r["us_k_step_bw"], r["updates_per_s"],
        r["stats"]["nn_fallbacks"] / 60))

"""A/B micro-benchmark of the step kernels for one library build (MIDAS_B200_LIB=...): per-kernel CUDA-event
times on the bench workload, steps 5..55 (global initialisation) and 120..170 (converged cloud), L2 flushed."""
import ctypes as C, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from midastouch_b200 import synth
from midastouch_b200._lib import call
from midastouch_b200.engine import FilterEngine, prepare_odom
from midastouch_b200.tactile_tree import tactile_tree

dev = torch.device("cuda:0")
obj, cbs, gt, meas = bench.make_assets()
cb = tactile_tree(cbs.poses, cbs.cam_poses, cbs.embeddings)
cb.to_device(dev)
n = int(os.environ.get("AB_N", bench.N_PER_GPU))
eng = FilterEngine(cb, capacity=n, sig_t=2e-4, sig_r=0.5, seed=1234, mesh_vertices=obj.vertices, pen_max=0.002)
g = torch.Generator().manual_seed(100)
sel = torch.randint(0, bench.M, (n,), generator=g)
eng.use_graph = bool(os.environ.get("AB_GRAPH"))
eng.fuse_sums = not bool(os.environ.get("AB_UNFUSED"))
eng.load_particles(cbs.poses.to(dev)[sel.to(dev)], nn_hint=sel.int().to(dev), spatial_sort=True)
odoms = [prepare_odom(torch.inverse(meas[t - 1]) @ meas[t]) for t in range(1, bench.T_TRAJ)]
codes = [synth.make_pose_query(gt[t + 1], bench.D, seed=3, frame=t).to(dev) for t in range(bench.T_TRAJ - 1)]
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
us = torch.rand(4096, generator=torch.Generator().manual_seed(7)).tolist()
T = int(os.environ.get("AB_STEPS", 170))
noflush = bool(os.environ.get("AB_NOFLUSH"))
rows = []
for t in range(T):
    if not noflush:
        flush.zero_(); flush.sum()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    for e in ev: e.record()
    call("mt_ctx_set_timing_events", eng.ctx.h, (C.c_void_p * 4)(*[x.cuda_event for x in ev[1:5]]))
    ev[0].record()
    eng.step(codes[t], odoms[t], u=us[t])
    ev[5].record()
    rows.append(ev)
torch.cuda.synchronize()
call("mt_ctx_set_timing_events", eng.ctx.h, None)
def avg(lo, hi, a, b):
    return 1e3 * sum(r[a].elapsed_time(r[b]) for r in rows[lo:hi]) / (hi - lo)
st = eng.ctx.stats(reset=True)
out = {"lib": os.path.basename(os.environ.get("MIDAS_B200_LIB", "default")), "graph": eng.use_graph}
for name, (lo, hi) in {"init": (5, 55), "conv": (120, min(170, T))}.items():
    if hi <= lo: continue
    out[name] = {"step": round(avg(lo, hi, 0, 5), 1), "a": round(avg(lo, hi, 1, 2), 1), "nnq": round(avg(lo, hi, 2, 3), 1), "bw": round(avg(lo, hi, 3, 5), 1)}
out["fallbacks/step"] = st["nn_fallbacks"] / T
out["on_surface"] = st["on_surface"]
out["mesh_deferred/step"] = st["mesh_deferred"] / T
out["scan_deferred/step"] = st["scan_deferred"] / T
print(json.dumps(out))

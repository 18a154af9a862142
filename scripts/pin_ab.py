"""cotter-pin stand-in (BASELINE config 5): per-step latency of the engine at N = 2^20 (and AB_N) on one GPU"""
import ctypes as C, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from midastouch_b200 import synth
from midastouch_b200._lib import call
from midastouch_b200.engine import FilterEngine, prepare_odom
from midastouch_b200.tactile_tree import tactile_tree
dev = torch.device("cuda:0")
pin = synth.make_object("cotter-pin")
cbs = synth.make_codebook(pin, M=50000, D=256, seed=5, embedding="smooth")
cb = tactile_tree(cbs.poses, cbs.cam_poses, cbs.embeddings)
cb.to_device(dev)
gt, meas = synth.make_trajectory(pin, T=64, seed=5, step=1e-4)
odoms = [prepare_odom(torch.inverse(meas[t - 1]) @ meas[t]) for t in range(1, 64)]
codes = [synth.make_pose_query(gt[t + 1], 256, seed=5, frame=t).to(dev) for t in range(63)]
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
n = int(os.environ.get("AB_N", 1 << 20))
eng = FilterEngine(cb, capacity=n, sig_t=1e-4, sig_r=0.5, seed=1, mesh_vertices=pin.vertices)
eng.use_graph = bool(os.environ.get("AB_GRAPH"))
g = torch.Generator().manual_seed(20)
sel = torch.randint(0, 50000, (n,), generator=g)
eng.load_particles(cbs.poses.to(dev)[sel.to(dev)], nn_hint=sel.int().to(dev), spatial_sort=True)
for t in range(10):
    eng.step(codes[t], odoms[t], u=0.3)
eng.ctx.stats(reset=True)
rows = []
for t in range(10, 50):
    flush.zero_(); flush.sum()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    for e in ev: e.record()
    call("mt_ctx_set_timing_events", eng.ctx.h, (C.c_void_p * 4)(*[x.cuda_event for x in ev[1:5]]))
    ev[0].record()
    eng.step(codes[t], odoms[t], u=0.37)
    ev[5].record()
    rows.append(ev)
torch.cuda.synchronize()
call("mt_ctx_set_timing_events", eng.ctx.h, None)
avg = lambda a, b: 1e3 * sum(r[a].elapsed_time(r[b]) for r in rows) / len(rows)
st = eng.ctx.stats(reset=True)
print(json.dumps({"lib": os.path.basename(os.environ.get("MIDAS_B200_LIB", "default")), "n": n, "step_us": round(avg(0, 5), 1), "a": round(avg(1, 2), 1),
                  "consumers": round(avg(2, 3), 1), "bw": round(avg(3, 5), 1), "updates_per_s": n / (avg(0, 5) * 1e-6),
                  "fallbacks/step": st["nn_fallbacks"] / 40, "leaves/search": st["grid_rows"] / max(st["nn_fallbacks"], 1), "leaves_max": st["grid_rows_max"],
                  "mesh_deferred/step": st["mesh_deferred"] / 40, "on_surface": st["on_surface"]}))

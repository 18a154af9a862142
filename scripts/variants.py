"""GPU micro-benchmark of the step kernels under engine options (diagnostic, not bench.py)."""
import sys, os, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from midastouch_b200 import synth
from midastouch_b200.engine import FilterEngine, prepare_odom
from midastouch_b200.tactile_tree import tactile_tree

dev = torch.device("cuda:0")
obj, cbs, gt, meas = bench.make_assets()
cb = tactile_tree(cbs.poses, cbs.cam_poses, cbs.embeddings)
cb.to_device(dev)
n = bench.N_PER_GPU
odoms = [prepare_odom(torch.inverse(meas[t - 1]) @ meas[t]) for t in range(1, bench.T_TRAJ)]
codes = [synth.make_query(cbs, t, seed=t).to(dev) for t in range(bench.T_TRAJ - 1)]
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
res = {}
for name, kw in (("prune+sort", dict(prune=True, sort=True)), ("noprune+sort", dict(prune=False, sort=True)),
                 ("prune+nosort", dict(prune=True, sort=False))):
    eng = FilterEngine(cb, capacity=n, seed=1, mesh_vertices=obj.vertices if kw["prune"] else None)
    g = torch.Generator().manual_seed(100)
    sel = torch.randint(0, bench.M, (n,), generator=g)
    eng.load_particles(cbs.poses.to(dev)[sel.to(dev)], nn_hint=sel.int().to(dev), spatial_sort=kw["sort"])
    eng.ctx.stats(reset=True)
    ts, fbs = [], []
    for t in range(40):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.step(codes[t % len(codes)], odoms[t % len(odoms)], u=0.3)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
        st = eng.ctx.stats(reset=True)
        fbs.append((st["nn_fallbacks"], st["grid_rows"], st["grid_rows_max"]))
    res[name] = {"us_step_0_4": ts[:5], "us_step_35_39": ts[35:], "fallbacks_per_step": sum(f[0] for f in fbs) / 40, "on_surface": st["on_surface"],
                 "us_all": [round(x) for x in ts], "fallbacks_all": fbs}
    del eng
print(json.dumps(res, indent=1))

"""host-side cost of FilterEngine.step (diagnostic)"""
import sys, os, time
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
codes_h = [synth.make_query(cbs, t, seed=t).pin_memory() for t in range(bench.T_TRAJ - 1)]
codes_d = [c.to(dev) for c in codes_h]
gts = [gt[t].float().contiguous() for t in range(bench.T_TRAJ)]
eng = FilterEngine(cb, capacity=n, seed=1, mesh_vertices=obj.vertices)
g = torch.Generator().manual_seed(100)
sel = torch.randint(0, bench.M, (n,), generator=g)
eng.load_particles(cbs.poses.to(dev)[sel.to(dev)], nn_hint=sel.int().to(dev), spatial_sort=True)
for t in range(10):
    eng.step(codes_d[t], odoms[t], u=0.3)
torch.cuda.synchronize()
for name, kw in (("device code, no gt", dict(h=False, g=False)), ("host code, no gt", dict(h=True, g=False)), ("host code + gt", dict(h=True, g=True))):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(50):
        k = t % 60
        eng.step(codes_h[k] if kw["h"] else codes_d[k], odoms[k], u=0.3, gt=gts[k + 1] if kw["g"] else None)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{name}: host {1e6*(t1-t0)/50:.0f} us/step issue, {1e6*(t2-t0)/50:.0f} us/step incl. drain")
# sync readback variants
for name in ("sync .cpu()", "pipelined pinned"):
    torch.cuda.synchronize()
    pin = [torch.zeros(2).pin_memory() for _ in range(2)]
    ev = [torch.cuda.Event() for _ in range(2)]
    t0 = time.perf_counter()
    for t in range(50):
        k = t % 60
        eng.step(codes_h[k], odoms[k], u=0.3, gt=gts[k + 1])
        if name.startswith("sync"):
            r = eng.rmse.cpu()
        else:
            pin[t & 1].copy_(eng.rmse, non_blocking=True)
            ev[t & 1].record()
            if t:
                ev[(t - 1) & 1].synchronize()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{name}: {1e6*(t2-t0)/50:.0f} us/step")

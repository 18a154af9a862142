"""diagnostic: distribution of the hint-graph scan length in k_step_a (library built with -DMT_SCAN_HIST)."""
import os, sys, ctypes as C, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from midastouch_b200 import synth, _lib
from midastouch_b200.engine import FilterEngine, prepare_odom
from midastouch_b200.tactile_tree import tactile_tree

dev = torch.device("cuda:0")
obj, cbs, gt, meas = bench.make_assets()
cb = tactile_tree(cbs.poses, cbs.cam_poses, cbs.embeddings); cb.to_device(dev)
n = bench.N_PER_GPU
eng = FilterEngine(cb, capacity=n, seed=1, mesh_vertices=obj.vertices)
g = torch.Generator().manual_seed(100)
sel = torch.randint(0, bench.M, (n,), generator=g)
eng.load_particles(cbs.poses.to(dev)[sel.to(dev)], nn_hint=sel.int().to(dev), spatial_sort=True)
odoms = [prepare_odom(torch.inverse(meas[t - 1]) @ meas[t]) for t in range(1, bench.T_TRAJ)]
codes = [synth.make_pose_query(gt[t + 1], bench.D, seed=3, frame=t).to(dev) for t in range(bench.T_TRAJ - 1)]
lib = _lib.lib()
f = lib.mt_debug_scan_hist
f.argtypes = [C.c_void_p, C.c_int]
out = {}
for t in range(60):
    eng.step(codes[t], odoms[t], u=0.3)
    if t in (0, 5, 10, 20, 40, 59):
        torch.cuda.synchronize()
        h = (C.c_ulonglong * 132)()
        f(h, 1)
        out[t] = {"particle": list(h)[:66], "warp_max": list(h)[66:]}
    else:
        torch.cuda.synchronize()
        h = (C.c_ulonglong * 132)(); f(h, 1)
for t, v in out.items():
    p, w = v["particle"], v["warp_max"]
    tp, tw = sum(p), sum(w)
    def cum(a, tot): 
        c, o = 0, []
        for k in (0, 2, 4, 6, 8, 12, 16, 24, 32, 48, 64, 65):
            o.append("%d:%.3f" % (k, sum(a[:k + 1]) / tot))
        return " ".join(o)
    print("step", t, "mean entries/particle %.2f" % (sum(i * x for i, x in enumerate(p)) / tp), "| mean warp max %.2f" % (sum(i * x for i, x in enumerate(w)) / tw))
    print("  particle cdf", cum(p, tp)); print("  warpmax  cdf", cum(w, tw))

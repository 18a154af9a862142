import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from midastouch_b200 import synth, particle_filter as PF, tactile_tree as TT
from midastouch_b200.engine import FilterEngine
dev = torch.device("cuda:0")
box = synth.make_object("004_sugar_box")
cbs = synth.make_codebook(box, M=20000, D=64, seed=4, embedding="smooth")
cb = TT.tactile_tree(cbs.poses, cbs.cam_poses, cbs.embeddings); cb.to_device(dev)
gt, meas = synth.make_trajectory(box, T=40, seed=4)
N = 20000
eng = FilterEngine(cb, capacity=N, mesh_vertices=box.vertices, pen_max=0.002, seed=0)
g = torch.Generator().manual_seed(0)
sel = torch.randint(0, 20000, (N,), generator=g)
eng.load_particles(cbs.poses[sel].to(dev), nn_hint=sel.int().to(dev))
keys = cb.logmap_pose
for t in range(1, 30):
    odom = torch.inverse(meas[t-1]) @ meas[t]
    eng.ctx.stats(reset=True)
    eng.step(synth.make_pose_query(gt[t], 64, seed=4, frame=t), odom, resample=False)
    st = eng.ctx.stats(reset=True)
    k = TT.R3_SE3(eng.poses())
    nn = eng.nn_idx().long()
    d = (k - keys[nn]).norm(dim=1)
    dt = (k[:, :3] - keys[nn][:, :3]).norm(dim=1); dr = (k[:, 3:] - keys[nn][:, 3:]).norm(dim=1)
    if t % 4 == 1: print(t, "fallbacks", st["nn_fallbacks"], "on_surface", st["on_surface"], "d* med %.2f p90 %.2f max %.2f mm | transl med %.2f rot med %.2f" % (1e3*d.median(), 1e3*d.quantile(0.9), 1e3*d.max(), 1e3*dt.median(), 1e3*dr.median()))
    # resample by running the full step on the same state? emulate: second step call would move again; instead do resampling step
    eng.step(synth.make_pose_query(gt[t], 64, seed=4, frame=t), torch.eye(4), tn=torch.zeros(N,3,device=dev), rot=torch.zeros(N,3,device=dev))

import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from midastouch_b200 import synth, particle_filter as PF, tactile_tree as TT
from midastouch_b200.config import compose
from midastouch_b200.filter_loop import run_filter, run_filter_engine
dev = torch.device("cuda:0")
box = synth.make_object("004_sugar_box")
cfg = compose(overrides=["expt.params.num_particles=20000", "expt.params.resample=low_var"])
cbs = synth.make_codebook(box, M=20000, D=64, seed=4, embedding="smooth")
cb = TT.tactile_tree(cbs.poses, cbs.cam_poses, cbs.embeddings); cb.to_device(dev)
gt, meas = synth.make_trajectory(box, T=80, seed=4)
code_fn = lambda idx: synth.make_pose_query(gt[idx], 64, seed=4, frame=idx)
for seed in (0, 1):
    torch.manual_seed(seed)
    pf = PF.particle_filter(cfg, box.vertices, downsample=1)
    st = run_filter(cfg, pf, cb, lambda i: code_fn(i).to(dev), gt.to(dev), meas.to(dev), floor=5000)
    print("dropin seed", seed, [round(1e3*x,1) for x in st["rmse_t"][::8]], st["num_particles"][::16])
    torch.manual_seed(seed)
    pf2 = PF.particle_filter(cfg, box.vertices, downsample=1)
    se = run_filter_engine(cfg, pf2, cb, code_fn, gt.to(dev), meas.to(dev), seed=seed)
    print("engine seed", seed, [round(1e3*x,1) for x in se["rmse_t"][::8]], se["engine"].ctx.stats())

"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz by running the UNMODIFIED
reference module (``/root/reference/midastouch/modules/particle_filter.py``, loaded by
``oracle/ref_shim.py``) on seeded synthetic inputs.  Runs only in the build container
(the reference tree does not travel); the .npz files it writes are committed.

    python -m oracle.gen_golden
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from midastouch_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    os.makedirs(OUT, exist_ok=True)
    pfm, posem = ref_shim.load_reference()
    torch.set_num_threads(1)  # fixed summation order in the reference's torch reductions

    obj = synth.make_object("004_sugar_box")
    vpath = os.path.join(OUT, "_vertices_tmp.npy")
    np.save(vpath, obj.vertices)
    cfg = ref_shim.default_cfg()
    pf = pfm.particle_filter(cfg, vpath, 1.0)
    os.remove(vpath)

    N, M, D = 1024, 4096, 256
    cb = synth.make_codebook(obj, M=M, D=D, seed=0)
    gt, meas = synth.make_trajectory(obj, T=8, seed=0)

    # ---- particles: codebook poses near the gt start + jitter
    g = torch.Generator().manual_seed(7)
    sel = torch.randint(0, M, (N,), generator=g)
    poses = cb.poses[sel].clone()

    # ---- euler_angles_to_matrix (pose.py:215-269)
    rot = torch.randn(64, 3, generator=g) * 30.0
    Rn = posem.euler_angles_to_matrix(torch.deg2rad(rot), "ZYX")
    np.savez(os.path.join(OUT, "euler_zyx.npz"), rot_deg=rot.numpy(), Rn=Rn.numpy())

    # ---- motionModel (particle_filter.py:359-377) with the reference's own RNG draw
    odom = torch.inverse(meas[0]) @ meas[1]
    torch.manual_seed(11)
    tn = torch.normal(mean=0.0, std=1.0 * pf.motion_noise["sig_t"], size=(N, 3))
    rn = torch.normal(mean=0.0, std=1.0 * pf.motion_noise["sig_r"], size=(N, 3))
    torch.manual_seed(11)
    moved = pf.motionModel(pfm.Particles(poses.clone()), odom, multiplier=1.0)
    assert len(moved) == N
    np.savez(os.path.join(OUT, "motion.npz"), poses=poses.numpy(), odom=odom.numpy(), tn=tn.numpy(), rot_deg=rn.numpy(),
             moved=moved.poses.numpy(), sig_t=pf.motion_noise["sig_t"], sig_r=pf.motion_noise["sig_r"], seed=11)

    # ---- get_similarity (particle_filter.py:449-469)
    q = synth.make_query(cb, int(sel[0]), seed=0)
    targets = cb.embeddings[sel]
    w_soft = pf.get_similarity(q, targets, softmax=True)
    w_raw = pf.get_similarity(q, targets, softmax=False)
    w_const = pf.get_similarity(q, targets[:1].repeat(16, 1), softmax=True)
    heat = pf.get_similarity(q, cb.embeddings, softmax=False)
    np.savez(os.path.join(OUT, "similarity.npz"), q=q.numpy(), sel=sel.numpy(), emb_head=cb.embeddings[:64].numpy(),  # full table = synth.make_codebook(box, 4096, 256, seed=0)
            
             w_soft=w_soft.numpy(), w_raw=w_raw.numpy(), w_const=w_const.numpy(), heat=heat.numpy())

    # ---- resampler low_var (the real Python loop) and low_var_batch (230-307)
    res = {}
    for name, w in (("soft", w_soft), ("raw", w_raw), ("masked", w_soft * (torch.arange(N) % 3 != 0)),
                    ("peaked", torch.softmax(80.0 * w_raw, 0))):
        for seed in (3, 4):
            torch.manual_seed(seed)
            u = torch.rand(1)
            torch.manual_seed(seed)
            labels = torch.arange(N, dtype=torch.float32)  # labels carry the ancestor index
            out = pf.resampler(pfm.Particles(moved.poses.clone(), w.clone(), labels), resample="low_var")
            torch.manual_seed(seed)
            outb = pf.resampler(pfm.Particles(moved.poses.clone(), w.clone(), labels.clone()), resample="low_var_batch")
            res[f"{name}_{seed}_w"] = w.numpy()
            res[f"{name}_{seed}_u"] = u.numpy()
            res[f"{name}_{seed}_anc"] = out.labels.numpy().astype(np.int64)
            res[f"{name}_{seed}_filled"] = (out.poses[:, 3, 3] == 1).numpy()
            res[f"{name}_{seed}_anc_batch"] = outb.labels.numpy().astype(np.int64)
            res[f"{name}_{seed}_poses0"] = out.poses[:8].numpy()
    res["in_poses"] = moved.poses.numpy()
    np.savez(os.path.join(OUT, "resample_low_var.npz"), **res)

    # ---- particle_rmse (472-496)
    rt, rr = pfm.particle_rmse(pfm.Particles(moved.poses.clone()), gt[1])
    np.savez(os.path.join(OUT, "rmse.npz"), poses=moved.poses.numpy(), gt=gt[1].numpy(), rmse_t=rt.numpy(), rmse_r=rr.numpy())

    # ---- remove_invalid_particles (379-403): push a third of the particles off the surface
    drift = moved.poses.clone()
    drift[::3, :3, 3] += 0.004 * drift[::3, :3, 2]
    out, drifted = pf.remove_invalid_particles(pfm.Particles(drift.clone(), w_soft.clone()))
    np.savez(os.path.join(OUT, "prune.npz"), poses=drift.numpy(), w_in=w_soft.numpy(), w_out=out.weights.numpy(),
             drifted=bool(drifted), vertices_ds=obj.vertices[::10], pen_max=pf.pen_max)

    # ---- annealing (405-447)
    ann = {}
    pf.particle_var = torch.tensor([float("inf")])
    p0 = pf.annealing(pfm.Particles(moved.poses.clone(), w_soft.clone()), torch.tensor(1e-3), floor=100)
    p1 = pf.annealing(p0, torch.tensor(0.8e-3), floor=100)  # ratio 0.8 -> remove
    ann["remove_n"] = len(p1)
    ann["remove_w"] = p1.weights.numpy()
    pf.particle_var = torch.tensor(1e-3)
    pf.init_particles = 2 * N
    p2 = pf.annealing(pfm.Particles(moved.poses.clone(), w_soft.clone()), torch.tensor(1.2e-3), floor=100)  # add
    ann["add_n"] = len(p2)
    ann["add_w"] = p2.weights.numpy()
    ann["w_in"] = w_soft.numpy()
    np.savez(os.path.join(OUT, "annealing.npz"), **ann)
    print("golden vectors written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print(" ", f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()

"""-m gpu: parity at BASELINE.json's own sizes (configs 2 and 3), teacher-forced against the oracle:
   config 3  035_power_drill, N = 1e6 particles, M = 50 000, D = 256, fused step (drift pruning on), CUDA-graph form
   config 2  004_sugar_box,   N = 65 536,        M = 50 000, D = 512 (float64 and float32 codebook)
Every particle is checked: poses 1e-5, SE3_NN index == the oracle's exact nearest key of the same pose, weights 1e-10,
ancestors == oracle.low_var_indices on the same weights (at most one flip, only where a sample location sits within
1e-13 of a CDF boundary: the reference's CPU cumsum is strictly sequential, no parallel scan reproduces its rounding)."""
import numpy as np
import pytest
import torch

from midastouch_b200 import synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _check_ancestors(a, w, u):
    n = w.shape[0]
    ref = O.low_var_indices(w, u)
    ok = ref >= 0
    bad = ((a != ref) & ok).nonzero().flatten()
    assert bad.numel() <= 1, bad.numel()
    if bad.numel():
        _, Cd = O.systematic_cdf(w)
        locs = O.systematic_locs(n, u)
        for j in bad.tolist():
            i0, i1 = sorted((int(a[j]), int(ref[j])))
            assert i1 - i0 == 1 or float(w[i0 + 1: i1].sum()) == 0.0
            assert abs(float(Cd[i0]) - float(locs[j])) < 1e-13
    return ref


def _headline(obj_name, N, M, D, emb_dtype, use_graph, seed):
    from midastouch_b200.engine import FilterEngine
    from midastouch_b200.tactile_tree import R3_SE3, tactile_tree

    dev = torch.device("cuda:0")
    obj = synth.make_object(obj_name)
    cbs = synth.make_codebook(obj, M=M, D=D, seed=seed, embedding="smooth")
    emb = cbs.embeddings.to(emb_dtype)
    cb = tactile_tree(cbs.poses, cbs.cam_poses, emb)
    cb.to_device(dev)
    gt, meas = synth.make_trajectory(obj, T=6, seed=seed)
    g = torch.Generator().manual_seed(seed)
    sel = torch.randint(0, M, (N,), generator=g)
    poses = cbs.poses[sel]
    # start the cloud between codebook poses (one untracked motion step), some of it off the surface
    poses = poses @ O.noisy_odom(torch.eye(4), 4e-4 * torch.randn(N, 3, generator=g), 1.0 * torch.randn(N, 3, generator=g))
    poses[::7, :3, 3] += 0.004 * poses[::7, :3, 2]
    tn = 2e-4 * torch.randn(N, 3, generator=g)
    rot = 0.5 * torch.randn(N, 3, generator=g)
    q = synth.make_pose_query(gt[1], D, seed=seed, frame=1)
    odom = torch.inverse(meas[0]) @ meas[1]
    u = 0.4142
    # stage 1: weighting only
    eng = FilterEngine(cb, capacity=N, mesh_vertices=obj.vertices, pen_max=0.002)
    eng.use_graph = use_graph
    eng.load_particles(poses.to(dev), nn_hint=sel.int().to(dev))
    eng.step(q, odom, u=u, tn=tn.to(dev), rot=rot.to(dev), gt=gt[1], resample=False)
    got_moved = eng.poses().cpu()
    moved, keep = O.motion_model(poses, odom, tn, rot)
    assert keep.all() and torch.allclose(got_moved, moved, rtol=1e-5, atol=1e-6)
    nn = eng.nn_idx().cpu().long()
    keys_cb = cb.logmap_pose.cpu().numpy()
    gkeys = R3_SE3(eng.poses()).cpu().numpy()
    want_nn = O.nn_exact(keys_cb, gkeys)
    assert np.array_equal(nn.numpy(), want_nn), int((nn.numpy() != want_nn).sum())
    sim = O.codebook_similarity(q, emb.double() if emb_dtype == torch.float64 else emb)
    w_soft = torch.exp(sim.double()[nn])
    w_ref, drifted = O.remove_invalid(got_moved, w_soft, obj.vertices, 0.002)
    assert not drifted and 0.05 < float((w_ref == 0).double().mean()) < 0.6
    w = eng.weights().cpu()
    assert torch.equal(w == 0, w_ref == 0)
    assert torch.allclose(w, w_ref / w_ref.sum(), rtol=1e-10 if emb_dtype == torch.float64 else 1e-6, atol=0)
    rt, rr = O.particle_rmse(moved, gt[1])
    r2 = eng.rmse.cpu()
    assert abs(float(r2[0]) - float(rt)) <= 1e-5 * float(rt)
    # rotation: this cloud covers all orientations, including faces exactly opposite to the true one.  There
    # acos((tr - 1) / 2) is ill-conditioned (a 1-ulp difference in the float32 trace moves the angle by up to
    # sqrt(eps) ~ 0.02 deg) and an argument that rounds below -1 turns into NaN -> 0 deg (nan_to_num,
    # particle_filter.py:486): each such particle moves the mean square by 180^2.  The 1e-5 bar for rmse_r is held by
    # the golden-vector tests; here only the size of that effect is bounded.
    assert abs(float(r2[1]) - float(rr)) <= 1e-3 * float(rr)
    # stage 2: the full step (two of them, so that both buffer parities / graph instantiations run)
    eng2 = FilterEngine(cb, capacity=N, mesh_vertices=obj.vertices, pen_max=0.002)
    eng2.use_graph = use_graph
    eng2.load_particles(poses.to(dev), nn_hint=sel.int().to(dev))
    eng2.step(q, odom, u=u, tn=tn.to(dev), rot=rot.to(dev))
    a = eng2.ancestors().cpu().long()
    ref = _check_ancestors(a, w, u)
    assert torch.equal(eng2.poses().cpu(), got_moved[a])
    assert torch.equal(eng2.nn_idx().cpu().long(), nn[a])
    assert bool((w[a] > 0).all())
    assert cb.ctx.stats()["overflow"] == 0
    return int((a != ref).sum())


def test_config3_drill_1e6_fused_step_graph():
    _headline("035_power_drill", 1_000_000, 50_000, 256, torch.float64, True, seed=3)


def test_config3_drill_1e6_fused_step_stream():
    _headline("035_power_drill", 1_000_000, 50_000, 256, torch.float64, False, seed=33)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_config2_sugar_box_65536_d512(dtype):
    _headline("004_sugar_box", 65_536, 50_000, 512, dtype, True, seed=2)

"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI
(midastouch_b200._lib -> libmidas_b200.so), against the oracle and the golden vectors of the
unmodified reference.  Bars: ancestor / NN indices bit-exact; float32 poses rtol 1e-5
(atol 1e-6 for entries near zero); float64 weights rtol 1e-5 (typically 1e-14)."""
import ctypes as C
import math

import numpy as np
import pytest
import torch

from oracle import oracle as O
from midastouch_b200 import synth

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-5, 1e-6


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def mt():
    import midastouch_b200 as m
    from midastouch_b200 import _lib, context, engine, particle_filter, tactile_tree

    class NS:
        pass

    ns = NS()
    ns.lib, ns.ctxm, ns.eng, ns.pf, ns.tt = _lib, context, engine, particle_filter, tactile_tree
    return ns


@pytest.fixture(scope="module")
def box():
    return synth.make_object("004_sugar_box")


@pytest.fixture(scope="module")
def cb_small(box, mt, dev):
    cbs = synth.make_codebook(box, M=4096, D=256, seed=0)
    cb = mt.tt.tactile_tree(cbs.poses, cbs.cam_poses, cbs.embeddings)
    cb.to_device(dev)
    return cbs, cb


@pytest.fixture(scope="module")
def cb_big(box, mt, dev):
    cbs = synth.make_codebook(box, M=50000, D=256, seed=0)
    cb = mt.tt.tactile_tree(cbs.poses, cbs.cam_poses, cbs.embeddings)
    cb.to_device(dev)
    return cbs, cb


def T(a):
    return torch.from_numpy(np.asarray(a))


def philox_u01_53(seed: int, stream_id: int, n: int) -> np.ndarray:
    """numpy model of the library's per-draw uniform: Philox4x32-10 with counter (j, j >> 32, stream lo, stream hi)
    and key (seed lo, seed hi); u = ((x >> 5) << 26 | (y >> 6)) * 2^-53 (mt_math.cuh mt_philox / mt_u01_53; pinned
    against the C source in tests/test_host_math.py::test_categorical_draw_from_cdf)."""
    j = np.arange(n, dtype=np.uint64)
    c = [j & np.uint64(0xFFFFFFFF), j >> np.uint64(32), np.full(n, stream_id & 0xFFFFFFFF, dtype=np.uint64),
         np.full(n, (stream_id >> 32) & 0xFFFFFFFF, dtype=np.uint64)]
    k0, k1 = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
    M0, M1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> np.uint64(32)) ^ c[1] ^ k0) & MASK, p1 & MASK, ((p0 >> np.uint64(32)) ^ c[3] ^ k1) & MASK, p0 & MASK]
        k0, k1 = (k0 + np.uint64(0x9E3779B9)) & MASK, (k1 + np.uint64(0xBB67AE85)) & MASK
    bits = ((c[0] >> np.uint64(5)) << np.uint64(26)) | (c[1] >> np.uint64(6))
    return bits.astype(np.float64) * (1.0 / 9007199254740992.0)


# ----------------------------------------------------------------------------- layout
def test_aos_soa_roundtrip(mt, dev):
    p = torch.randn(1000, 4, 4, device=dev)
    p[:, 3, :] = torch.tensor([0, 0, 0, 1.0], device=dev)
    soa = mt.ctxm.aos_to_soa(p, stride=1024)
    assert torch.equal(mt.ctxm.soa_to_aos(soa, 1000), p)


# ----------------------------------------------------------------------------- keys + NN
def test_keys_match_oracle(mt, dev, cb_big):
    cbs, cb = cb_big
    ref = O.r3_se3(cbs.poses)
    assert torch.allclose(cb.logmap_pose.cpu(), ref, rtol=1e-5, atol=2e-7)


@pytest.mark.parametrize("exhaustive", [False, True])
def test_nn_bit_exact_on_given_keys(mt, dev, cb_big, exhaustive):
    """given identical float32 keys the index search is bit-exact vs the oracle (ties -> lowest)."""
    cbs, cb = cb_big
    keys_cb = cb.logmap_pose.cpu().numpy()
    rng = np.random.default_rng(3)
    n = 4096 if exhaustive else 65536
    base = keys_cb[rng.integers(0, 50000, n)]
    q = (base + rng.normal(size=(n, 6)).astype(np.float32) * np.float32(6e-4)).astype(np.float32)
    q[:64] = keys_cb[:64]  # exact hits
    qd = torch.from_numpy(q).to(dev)
    idx = torch.empty(n, dtype=torch.int32, device=dev)
    mt.lib.call("mt_nn_assign", cb.ctx.h, qd.data_ptr(), n, 0, 1 if exhaustive else 0, idx.data_ptr(), mt.lib.stream_ptr())
    ref = O.nn_exact(keys_cb, q)
    assert np.array_equal(idx.cpu().numpy().astype(np.int64), ref)
    if not exhaustive:  # any hint, good or bad, gives the same answer
        for hint in (torch.from_numpy(ref.astype(np.int32)), torch.randint(0, 50000, (n,), dtype=torch.int32)):
            hd = hint.to(dev)
            idx2 = torch.empty_like(idx)
            mt.lib.call("mt_nn_assign", cb.ctx.h, qd.data_ptr(), n, hd.data_ptr(), 0, idx2.data_ptr(), mt.lib.stream_ptr())
            assert torch.equal(idx2, idx)


def test_nn_far_queries_and_duplicates(mt, dev, box):
    cbs = synth.make_codebook(box, M=3000, D=8, seed=5)
    poses = torch.cat([cbs.poses, cbs.poses[:200]])  # duplicate keys -> ties
    cb = mt.tt.tactile_tree(poses, poses, torch.cat([cbs.embeddings, cbs.embeddings[:200]]))
    cb.to_device(dev)
    keys_cb = cb.logmap_pose.cpu().numpy()
    rng = np.random.default_rng(0)
    q = np.concatenate([keys_cb[:200], (rng.normal(size=(500, 6)) * 0.2).astype(np.float32)]).astype(np.float32)
    qd = torch.from_numpy(q).to(dev)
    n = q.shape[0]
    for mode in (0, 1):
        idx = torch.empty(n, dtype=torch.int32, device=dev)
        mt.lib.call("mt_nn_assign", cb.ctx.h, qd.data_ptr(), n, 0, mode, idx.data_ptr(), mt.lib.stream_ptr())
        assert np.array_equal(idx.cpu().numpy().astype(np.int64), O.nn_brute(keys_cb, q)), mode


def _nbr_view(mt, cb, dev):
    p, k = C.c_void_p(), C.c_int()
    mt.lib.call("mt_codebook_nbr_info", cb.ctx.h, C.byref(p), C.byref(k))
    M = len(cb)
    holder = type("H", (), {})()
    holder.__cuda_array_interface__ = {"shape": (M, k.value, 8), "typestr": "<f4", "data": (p.value, False), "version": 3, "strides": None}
    return torch.as_tensor(holder, device=dev).clone().cpu().numpy()


def test_neighbour_graph_matches_host_model(mt, dev, box):
    """k_build_nbr: every key's 64 nearest other keys, ascending (distance, index), duplicates included."""
    cbs = synth.make_codebook(box, M=3000, D=8, seed=5)
    poses = torch.cat([cbs.poses, cbs.poses[:50]])
    cb = mt.tt.tactile_tree(poses, poses, torch.cat([cbs.embeddings, cbs.embeddings[:50]]))
    cb.to_device(dev)
    keys = cb.logmap_pose.cpu().numpy()
    nbr = _nbr_view(mt, cb, dev)
    M, K = keys.shape[0], nbr.shape[1]
    assert K == 64
    for h in list(range(0, M, 97)) + [0, 10, 3049, M - 1]:
        d = O.l2_sq_f32(keys, keys[h])
        d[h] = np.inf
        order = np.lexsort((np.arange(M), d))[:K]
        assert np.array_equal(nbr[h, :, 7].view(np.int32), order.astype(np.int32)), h
        assert np.array_equal(nbr[h, :, :6], keys[order])
        assert np.allclose(nbr[h, :, 6], np.sqrt(d[order]), rtol=1e-6, atol=0)
    # tiny codebook: lists are padded with (inf, -1)
    cb2 = mt.tt.tactile_tree(cbs.poses[:5], cbs.poses[:5], cbs.embeddings[:5])
    cb2.to_device(dev)
    nb2 = _nbr_view(mt, cb2, dev)
    assert (nb2[:, 4:, 7].view(np.int32) == -1).all() and np.isinf(nb2[:, 4:, 6]).all()
    assert (nb2[:, :4, 7].view(np.int32) >= 0).all()
    q = torch.from_numpy(cb2.logmap_pose.cpu().numpy()[[3, 1, 4]] + np.float32(1e-5)).to(dev)
    idx = torch.empty(3, dtype=torch.int32, device=dev)
    hint = torch.tensor([0, 0, 0], dtype=torch.int32, device=dev)
    mt.lib.call("mt_nn_assign", cb2.ctx.h, q.data_ptr(), 3, hint.data_ptr(), 0, idx.data_ptr(), mt.lib.stream_ptr())
    assert idx.cpu().tolist() == [3, 1, 4]


def test_nn_stale_hints_near_pi(mt, dev, cb_big):
    """keys whose rotation vector flips sign near angle pi make the previous match a useless
    hint (distance ~ 2*pi*w in key space): the search must stay exact and bounded."""
    cbs, cb = cb_big
    keys_cb = cb.logmap_pose.cpu().numpy()
    rng = np.random.default_rng(11)
    n = 20000
    base = rng.integers(0, 50000, n)
    q = keys_cb[base].copy()
    q[:, 3:] *= -1.0  # antipodal rotation vector, same translation
    q += rng.normal(size=(n, 6)).astype(np.float32) * np.float32(2e-4)
    qd = torch.from_numpy(q.astype(np.float32)).to(dev)
    hd = torch.from_numpy(base.astype(np.int32)).to(dev)
    idx = torch.empty(n, dtype=torch.int32, device=dev)
    cb.ctx.stats(reset=True)
    mt.lib.call("mt_nn_assign", cb.ctx.h, qd.data_ptr(), n, hd.data_ptr(), 0, idx.data_ptr(), mt.lib.stream_ptr())
    assert np.array_equal(idx.cpu().numpy().astype(np.int64), O.nn_exact(keys_cb, q.astype(np.float32), k=16))
    fb = cb.ctx.stats(reset=True)["nn_fallbacks"]
    assert 0 <= fb <= n
    print("near-pi stale hints: grid fallbacks", fb, "of", n)


def test_se3_nn_dropin(mt, dev, cb_big):
    cbs, cb = cb_big
    g = torch.Generator().manual_seed(0)
    sel = torch.randint(0, 50000, (2048,), generator=g)
    poses = cbs.poses[sel].clone()
    poses[:, :3, 3] += 3e-4 * torch.randn(2048, 3, generator=g)
    p, c, e = cb.SE3_NN(poses.to(dev))
    ref = O.se3_nn(O.r3_se3(cbs.poses), poses)
    # oracle keys and CUDA keys differ by float32 ulps: indices must agree unless the two
    # candidates are equidistant to rounding
    idx = cb.SE3_NN_idx(poses.to(dev)).cpu().long()
    diff = (idx != ref).nonzero().flatten()
    keys_cb = O.r3_se3(cbs.poses).numpy()
    qk = O.r3_se3(poses).numpy()
    for j in diff.tolist():
        da, db = O.l2_sq_f32(keys_cb[idx[j]], qk[j]), O.l2_sq_f32(keys_cb[ref[j]], qk[j])
        assert abs(float(da) - float(db)) <= 1e-5 * float(db) + 1e-12
    assert len(diff) <= 2
    assert torch.equal(p.cpu(), cbs.poses[idx]) and torch.equal(c.cpu(), cbs.cam_poses[idx]) and torch.equal(e.cpu(), cbs.embeddings[idx])


def test_r3_se3_other_weights(mt, dev, cb_small):
    """R3_SE3(poses, w) for w != 0.01 (tactile_tree.py:73-77) against the oracle."""
    cbs, _ = cb_small
    for w in (0.01, 0.05, 0.5):
        got = mt.tt.R3_SE3(cbs.poses.to(dev), w=w).cpu()
        assert torch.allclose(got, O.r3_se3(cbs.poses, w=w), rtol=1e-5, atol=2e-7)


@pytest.mark.parametrize("nn", [2, 5, 64])
def test_se3_nn_k_neighbours(mt, dev, cb_small, nn):
    """SE3_NN(query, nn > 1) (tactile_tree.py:43-58): the nn nearest codebook poses of every query, nearest first,
    against a float32 brute force in the library's accumulation order on the library's keys."""
    cbs, cb = cb_small
    g = torch.Generator().manual_seed(1)
    sel = torch.randint(0, 4096, (37,), generator=g)
    poses = cbs.poses[sel].clone()
    poses[:, :3, 3] += 5e-4 * torch.randn(37, 3, generator=g)
    p, c, e = cb.SE3_NN(poses.to(dev), nn=nn)
    assert p.shape == (37, nn, 4, 4) and c.shape == (37, nn, 4, 4) and e.shape == (37, nn, cbs.embeddings.shape[1])
    keys_cb = cb.logmap_pose.cpu().numpy()
    qk = mt.tt.R3_SE3(poses.to(dev)).cpu().numpy()
    for j in range(37):
        d = O.l2_sq_f32(keys_cb, qk[j]).astype(np.float32)
        order = np.lexsort((np.arange(4096), d))[:nn]  # ascending (distance, index)
        assert torch.equal(p[j].cpu(), cbs.poses[torch.from_numpy(order)])
        assert torch.equal(e[j].cpu(), cbs.embeddings[torch.from_numpy(order)])
    # a single query is squeezed like the reference's indices_p.squeeze()
    p1, _, _ = cb.SE3_NN(poses[:1].to(dev), nn=nn)
    assert p1.shape == (nn, 4, 4) and torch.equal(p1.cpu(), p[0].cpu())


# ----------------------------------------------------------------------------- motion
def test_motion_vs_reference_golden(mt, dev, golden, box):
    g = golden("motion")
    cfg = synth_cfg()
    pf = mt.pf.particle_filter(cfg, box.vertices)
    torch.manual_seed(int(g["seed"]))
    out = pf.motionModel(mt.pf.Particles(T(g["poses"]).to(dev)), T(g["odom"]))
    assert len(out) == g["poses"].shape[0]
    assert torch.allclose(out.poses.cpu(), T(g["moved"]), rtol=RTOL, atol=ATOL)
    assert float((out.poses.cpu() - T(g["moved"])).abs().max()) < 1e-6


def synth_cfg(n=1024):
    from oracle.ref_shim import default_cfg

    return default_cfg(num_particles=n)


def test_motion_philox_statistics(mt, dev, cb_small):
    cbs, cb = cb_small
    n = 200000
    poses = torch.eye(4)[None].repeat(n, 1, 1).to(dev)
    soa = mt.ctxm.aos_to_soa(poses)
    eye = torch.eye(4).contiguous()
    for step in (0, 1):
        out = torch.empty_like(soa)
        mt.lib.call("mt_motion", soa.data_ptr(), out.data_ptr(), n, n, eye.data_ptr(), 0, 0, 2e-4, 0.5, 7, step, 0, 0, 0, mt.lib.stream_ptr())
        p = mt.ctxm.soa_to_aos(out, n).cpu()
        t = p[:, :3, 3].double()
        assert abs(t.mean()) < 5e-6 and abs(t.std() / 2e-4 - 1) < 0.01
        ang = torch.rad2deg(torch.acos(((p[:, 0, 0] + p[:, 1, 1] + p[:, 2, 2] - 1) / 2).clamp(-1, 1))).double()
        # |rotation| of three independent 0.5 deg Euler angles ~ chi(3) * 0.5
        assert abs(ang.pow(2).mean().sqrt() / (0.5 * math.sqrt(3)) - 1) < 0.02
        if step == 0:
            first = p
    assert not torch.equal(first, p)
    out2 = torch.empty_like(soa)
    mt.lib.call("mt_motion", soa.data_ptr(), out2.data_ptr(), n, n, eye.data_ptr(), 0, 0, 2e-4, 0.5, 7, 1, 0, 0, 0, mt.lib.stream_ptr())
    assert torch.equal(out2, out)  # deterministic in (seed, step, particle)


def test_init_filter_vs_oracle(mt, dev, box):
    pf = mt.pf.particle_filter(synth_cfg(), box.vertices)
    gt, _ = synth.make_trajectory(box, T=2)
    torch.manual_seed(3)
    parts = pf.init_filter(gt[0].to(dev), 4096)
    torch.manual_seed(3)
    tn = torch.normal(mean=0.0, std=pf.init_noise[0], size=(4096, 3))
    rot = torch.normal(mean=0.0, std=pf.init_noise[1], size=(4096, 3))
    ref = O.init_filter(gt[0], tn, rot)
    assert torch.allclose(parts.poses.cpu(), ref, rtol=RTOL, atol=2e-6)


# ----------------------------------------------------------------------------- similarity
def test_similarity_vs_reference_golden(mt, dev, golden, box, cb_small):
    g = golden("similarity")
    cbs, cb = cb_small
    pf = mt.pf.particle_filter(synth_cfg(), box.vertices)
    q, sel = T(g["q"]).to(dev), T(g["sel"])
    targets = cbs.embeddings[sel].to(dev)
    w = pf.get_similarity(q, targets, softmax=True).cpu()
    assert torch.allclose(w, T(g["w_soft"]), rtol=1e-12, atol=0)
    w = pf.get_similarity(q, targets, softmax=False).cpu()
    assert torch.allclose(w, T(g["w_raw"]), rtol=1e-12, atol=0)
    w = pf.get_similarity(q, targets[:1].repeat(16, 1), softmax=True).cpu()  # constant -> softmax skipped
    assert torch.allclose(w, T(g["w_const"]), rtol=1e-12, atol=0)
    heat = cb.query(q).cpu()
    assert torch.allclose(heat, T(g["heat"]), rtol=1e-12, atol=0)
    # float32 storage of the codebook: still inside the 1e-5 bar
    cb32 = mt.tt.tactile_tree(cbs.poses, cbs.cam_poses, cbs.embeddings.float())
    cb32.to_device(dev)
    assert torch.allclose(cb32.query(q).cpu(), T(g["heat"]), rtol=RTOL, atol=0)


def test_cosine_ragged_and_nan(mt, dev):
    pf_ctx = mt.pf._ctx_for(dev, 10)
    for D, dt in ((6, torch.float64), (12, torch.float32), (514, torch.float64)):
        q = torch.rand(D, dtype=torch.float64)
        t = torch.rand(37, D, dtype=dt)
        t[5] = 0  # zero row: clamp at eps
        out = torch.empty(37, dtype=torch.float64, device=dev)
        qd, td = q.to(dev), t.to(dev)
        mt.lib.call("mt_cosine_rows", pf_ctx.h, qd.data_ptr(), 1, td.data_ptr(), 1 if dt == torch.float64 else 0, 37, D,
                    out.data_ptr(), mt.lib.stream_ptr())
        ref = torch.nn.functional.cosine_similarity(q[None], t.double())
        assert torch.allclose(out.cpu(), ref, rtol=1e-6 if dt == torch.float32 else 1e-12, atol=1e-300)


def test_cosine_batched(mt, dev):
    Q = torch.rand(70, 256)
    Tm = torch.rand(1000, 256)
    out = torch.empty(70, 1000, device=dev)
    ctx = mt.pf._ctx_for(dev, 10)
    Qd, Td = Q.to(dev), Tm.to(dev)
    mt.lib.call("mt_cosine_batched", ctx.h, Qd.data_ptr(), 70, Td.data_ptr(), 1000, 256, out.data_ptr(), mt.lib.stream_ptr())
    ref = torch.nn.functional.cosine_similarity(Q.double()[:, None, :], Tm.double()[None], dim=2)
    assert torch.allclose(out.cpu().double(), ref, rtol=1e-5, atol=0)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_codebook_query_batched_tensor_core(mt, dev, box, dtype):
    """tcgen05 / TMEM GEMM (3xTF32): nq codes against the whole codebook, within 1e-5 of the float64
    cosine; ragged sizes (M, nq not multiples of the 128 x 128 tile, D not a multiple of 32)."""
    g = torch.Generator().manual_seed(5)
    for M, D, nq in ((1000, 256, 70), (4133, 200, 300)):
        poses = torch.eye(4)[None].repeat(M, 1, 1)
        poses[:, :3, 3] = torch.rand(M, 3, generator=g) * 0.05
        emb = torch.rand(M, D, dtype=torch.float64, generator=g)
        emb[3] *= 1e-3  # a short row: the norm matters
        cb = mt.tt.tactile_tree(poses, poses, emb.to(dtype))
        cb.to_device(dev)
        Q = torch.rand(nq, D, generator=g)
        Q[1] = emb[7].float()
        out = cb.query_batched(Q.to(dev)).cpu().double()
        ref = torch.nn.functional.cosine_similarity(Q.double()[:, None, :], emb.to(dtype).double()[None], dim=2)
        assert out.shape == (nq, M)
        err = ((out - ref).abs() / ref.abs().clamp_min(1e-12)).max()
        assert float(err) < 1e-5, float(err)
        # and the single-query kernel agrees with row 0
        one = cb.query(Q[0].double().to(dev)).cpu()
        assert torch.allclose(out[0], one, rtol=1e-5, atol=0)


# ----------------------------------------------------------------------------- resampling
@pytest.mark.parametrize("name", ["soft", "raw", "masked", "peaked"])
@pytest.mark.parametrize("seq", [0, 1])
def test_low_var_vs_reference_loop_golden(mt, dev, golden, name, seq):
    g = golden("resample_low_var")
    ctx = mt.pf._ctx_for(dev, 1024)
    for seed in (3, 4):
        w = T(g[f"{name}_{seed}_w"]).to(dev)
        u = float(g[f"{name}_{seed}_u"][0])
        n = w.shape[0]
        anc = torch.empty(n, dtype=torch.int32, device=dev)
        mt.lib.call("mt_resample_systematic", ctx.h, w.data_ptr(), n, C.c_float(u), seq, anc.data_ptr(), 0, mt.lib.stream_ptr())
        filled = g[f"{name}_{seed}_filled"]
        a = anc.cpu().numpy().astype(np.int64)
        assert np.array_equal(a[filled], g[f"{name}_{seed}_anc"][filled])
        if seq:
            assert np.array_equal(a >= 0, filled)


def test_low_var_dropin_vs_reference_golden(mt, dev, golden, box):
    g = golden("resample_low_var")
    pf = mt.pf.particle_filter(synth_cfg(), box.vertices)
    poses = T(g["in_poses"]).to(dev)
    w = T(g["soft_3_w"]).to(dev)
    n = w.shape[0]
    labels = torch.arange(n, dtype=torch.float32, device=dev)
    torch.manual_seed(3)
    out = pf.resampler(mt.pf.Particles(poses, w, labels), resample="low_var", u=float(g["soft_3_u"][0]))
    assert np.array_equal(out.labels.cpu().numpy().astype(np.int64), g["soft_3_anc"])
    assert torch.equal(out.poses[:8].cpu(), T(g["soft_3_poses0"]))
    assert torch.equal(out.weights.cpu(), T(g["soft_3_w"])[T(g["soft_3_anc"])])
    # guards (particle_filter.py:237-241): all-zero and NaN weights return the input
    z = pf.resampler(mt.pf.Particles(poses, torch.zeros_like(w), labels), resample="low_var")
    assert torch.equal(z.poses, poses)
    wn = w.clone()
    wn[5] = float("nan")
    z = pf.resampler(mt.pf.Particles(poses, wn, labels), resample="low_var")
    assert torch.equal(z.poses, poses)


@pytest.mark.parametrize("n", [1, 2, 255, 256, 257, 65536, 1000003])
def test_low_var_sizes_vs_oracle(mt, dev, n):
    ctx = mt.pf._ctx_for(dev, n)
    g = torch.Generator().manual_seed(n)
    for kind in ("flat", "peaked", "sparse"):
        w = torch.rand(n, dtype=torch.float64, generator=g) + 0.5
        if kind == "peaked":
            w = torch.softmax(20 * w, 0)
        if kind == "sparse":
            w = w * (torch.rand(n, generator=g) < 0.1)
            if w.sum() == 0:
                w[0] = 1.0
        for u in (0.0, 0.73, float(np.nextafter(np.float32(1), np.float32(0)))):
            anc = torch.empty(n, dtype=torch.int32, device=dev)
            wd = w.to(dev)
            mt.lib.call("mt_resample_systematic", ctx.h, wd.data_ptr(), n, C.c_float(u), 0, anc.data_ptr(), 0, mt.lib.stream_ptr())
            a = anc.cpu().long()
            ref = O.low_var_indices(w, u)
            ok = ref >= 0
            nd = int((a[ok] != ref[ok]).sum())
            if nd:
                # perf mode uses a parallel float64 prefix: an ancestor may flip only where a
                # sample location sits within rounding distance of a CDF boundary
                _, Cd = O.systematic_cdf(w)
                locs = O.systematic_locs(n, u)
                bad = (a != ref).nonzero().flatten()
                for j in bad.tolist():
                    i0, i1 = sorted((int(a[j]), int(ref[j])))
                    assert i1 - i0 == 1 or float(w[i0 + 1 : i1].sum()) == 0.0
                    assert abs(float(Cd[i0]) - float(locs[j])) < 1e-13
            assert nd <= 1, (n, kind, u, nd)
            assert (a[1:] >= a[:-1]).all()
            if kind == "sparse":
                assert (w[a] > 0).all()  # zero-weight particles never get children


def test_low_var_heavy_parent(mt, dev):
    """one particle owning ~all slots exercises the warp-cooperative child writes."""
    n = 50000
    w = torch.full((n,), 1e-9, dtype=torch.float64)
    w[12345] = 1.0
    ctx = mt.pf._ctx_for(dev, n)
    anc = torch.empty(n, dtype=torch.int32, device=dev)
    wd = w.to(dev)
    mt.lib.call("mt_resample_systematic", ctx.h, wd.data_ptr(), n, C.c_float(0.5), 0, anc.data_ptr(), 0, mt.lib.stream_ptr())
    assert torch.equal(anc.cpu().long(), O.low_var_indices(w, 0.5))


# ----------------------------------------------------------------------------- rmse
def test_rmse_vs_reference_golden(mt, dev, golden):
    g = golden("rmse")
    rt, rr = mt.pf.particle_rmse(T(g["poses"]).to(dev), T(g["gt"]))
    assert abs(float(rt) - float(g["rmse_t"])) <= RTOL * float(g["rmse_t"])
    assert abs(float(rr) - float(g["rmse_r"])) <= RTOL * float(g["rmse_r"])


# ----------------------------------------------------------------------------- drift pruning
def test_prune_vs_reference_golden(mt, dev, golden, box):
    g = golden("prune")
    pf = mt.pf.particle_filter(synth_cfg(), box.vertices)
    assert np.array_equal(pf.mesh_vertices_ds, g["vertices_ds"]) and pf.pen_max == float(g["pen_max"])
    parts = mt.pf.Particles(T(g["poses"]).to(dev), T(g["w_in"]).to(dev))
    out, drifted = pf.remove_invalid_particles(parts)
    assert torch.equal(out.weights.cpu(), T(g["w_out"]))
    assert bool(drifted) == bool(g["drifted"])
    assert out.weights.data_ptr() == parts.weights.data_ptr()  # in-place like the reference (401)
    # everything far away -> drifted
    far = T(g["poses"]).clone()
    far[:, :3, 3] += 1.0
    out, drifted = pf.remove_invalid_particles(mt.pf.Particles(far.to(dev), T(g["w_in"]).to(dev)))
    assert bool(drifted) and float(out.weights.abs().sum()) == 0.0
    # explicit invalid_dist and float32 default weights
    out, drifted = pf.remove_invalid_particles(mt.pf.Particles(T(g["poses"]).to(dev)), invalid_dist=10.0)
    assert not bool(drifted) and out.weights.dtype == torch.float32 and bool((out.weights == 1).all())


def test_prune_boundary_vs_oracle(mt, dev, box):
    """points at ~pen_max from the nearest vertex: the float64 distance test must agree with the
    k-d tree oracle point by point (n = 200k, incl. points outside the bounding box)."""
    pf = mt.pf.particle_filter(synth_cfg(), box.vertices)
    rng = np.random.default_rng(5)
    n = 200000
    v = pf.mesh_vertices_ds[rng.integers(0, pf.mesh_vertices_ds.shape[0], n)]
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    pts = (v + d * rng.uniform(0.0, 2.0 * pf.pen_max, (n, 1))).astype(np.float32)
    pts[:1000] += 0.3
    poses = torch.eye(4)[None].repeat(n, 1, 1)
    poses[:, :3, 3] = torch.from_numpy(pts)
    w = torch.rand(n, dtype=torch.float64)
    out, drifted = pf.remove_invalid_particles(mt.pf.Particles(poses.to(dev), w.clone().to(dev)))
    w_ref, dr = O.remove_invalid(poses, w.clone(), pf.mesh_vertices_ds, pf.pen_max)
    assert torch.equal(out.weights.cpu(), w_ref) and bool(drifted) == dr
    assert 0.2 < float((w_ref == 0).double().mean()) < 0.8


# ----------------------------------------------------------------------------- annealing / cluster centres
def test_annealing_vs_reference_golden(mt, dev, golden, box):
    g = golden("annealing")
    gm = golden("motion")
    pf = mt.pf.particle_filter(synth_cfg(), box.vertices)
    N = g["w_in"].shape[0]
    poses = T(gm["moved"])[:N].to(dev)
    w = T(g["w_in"]).to(dev)
    labels = torch.arange(N, dtype=torch.float32, device=dev)
    p0 = pf.annealing(mt.pf.Particles(poses, w.clone(), labels), torch.tensor(1e-3), floor=100)  # first call: records var
    assert len(p0) == N and pf.init_particles == N
    p1 = pf.annealing(p0, torch.tensor(0.8e-3), floor=100)  # ratio 0.8 -> remove the lowest 20 %
    assert len(p1) == int(g["remove_n"])
    assert torch.equal(p1.weights.cpu(), T(g["remove_w"]))  # survivors keep their order (torch_delete)
    assert torch.equal(p1.poses, poses[p1.labels.long()])
    pf.particle_var = torch.tensor(1e-3)
    pf.init_particles = 2 * N
    p2 = pf.annealing(mt.pf.Particles(poses, w.clone(), labels), torch.tensor(1.2e-3), floor=100)  # add the top 20 %
    assert len(p2) == int(g["add_n"])
    assert torch.equal(p2.weights[:N].cpu(), T(g["w_in"]))
    assert torch.equal(torch.sort(p2.weights[N:].cpu()).values, torch.sort(T(g["add_w"])[N:]).values)
    assert torch.equal(p2.poses[N:], poses[p2.labels[N:].long()])
    # var == 0 and growth beyond init_particles are no-ops (particle_filter.py:417-419, 438-439)
    assert len(pf.annealing(p2, torch.tensor(0.0))) == len(p2)
    pf.init_particles = N
    pf.particle_var = torch.tensor(1e-3)
    assert len(pf.annealing(mt.pf.Particles(poses, w.clone(), labels), torch.tensor(2e-3))) == N


@pytest.mark.parametrize("n,k", [(1, 1), (1000, 1), (1000, 999), (65536, 20000), (1000003, 333334)])
def test_select_k_vs_sort(mt, dev, n, k):
    g = torch.Generator().manual_seed(n + k)
    w = torch.rand(n, dtype=torch.float64, generator=g)
    w[::5] = w[0]          # heavy ties
    w[1::97] = 0.0
    if n > 10:
        w[7] = -1.5        # negative and special values order correctly
        w[9] = float("inf")
    ctx = mt.pf._ctx_for(dev, n)
    wd = w.to(dev)
    for largest in (0, 1):
        sel = torch.empty(k, dtype=torch.int32, device=dev)
        keep = torch.empty(n - k, dtype=torch.int32, device=dev)
        mt.lib.call("mt_select_k", ctx.h, wd.data_ptr(), n, k, largest, sel.data_ptr(), keep.data_ptr(), mt.lib.stream_ptr())
        s, kp = sel.cpu().long(), keep.cpu().long()
        assert bool((s[1:] > s[:-1]).all()) and bool((kp[1:] > kp[:-1]).all())
        assert torch.equal(torch.sort(torch.cat([s, kp])).values, torch.arange(n))
        ref = torch.sort(w, descending=bool(largest)).values[:k]
        assert torch.equal(torch.sort(w[s], descending=bool(largest)).values, ref)
        # ties at the threshold go to the lowest indices
        thr = ref[-1]
        eq_sel = s[w[s] == thr]
        eq_all = (w == thr).nonzero().flatten()
        assert torch.equal(eq_sel, eq_all[: len(eq_sel)])


def test_cluster_centers_vs_oracle(mt, dev, box):
    pf = mt.pf.particle_filter(synth_cfg(), box.vertices)
    cbs = synth.make_codebook(box, M=3000, D=8, seed=9)
    g = torch.Generator().manual_seed(2)
    # three tight clusters around codebook poses + noise, labels -1/0/1 like DBSCAN output
    centres = cbs.poses[[10, 500, 2000]]
    n = 6000
    lab = torch.randint(0, 3, (n,), generator=g)
    tn = 1e-3 * torch.randn(n, 3, generator=g)
    rot = 3.0 * torch.randn(n, 3, generator=g)
    poses = centres[lab] @ O.noisy_odom(torch.eye(4), tn, rot)
    w = torch.rand(n, dtype=torch.float64, generator=g) + 0.1
    labels = (lab - 1).float()
    cp, cs = pf.get_cluster_centers(mt.pf.Particles(poses.to(dev), w.to(dev), labels.to(dev)), method="quat_avg")
    rp, rs = O.cluster_centers(poses, w, labels)
    assert cp.shape == (3, 4, 4) and cs.shape == (3, 3)
    assert torch.allclose(cp.cpu(), rp, rtol=1e-4, atol=2e-6), (cp.cpu() - rp).abs().max()
    assert torch.allclose(cs.cpu(), rs, rtol=1e-3, atol=1e-7), (cs.cpu() - rs).abs().max()
    # constant weights -> uniform (particle_filter.py:178-184); single default label (zeros)
    cp1, cs1 = pf.get_cluster_centers(mt.pf.Particles(poses[lab == 0].to(dev)), method="quat_avg")
    rp1, rs1 = O.cluster_centers(poses[lab == 0], torch.ones(int((lab == 0).sum())), torch.zeros(int((lab == 0).sum())))
    assert torch.allclose(cp1.cpu(), rp1, rtol=1e-4, atol=2e-6) and torch.allclose(cs1.cpu(), rs1, rtol=1e-3, atol=1e-7)
    RtR = cp[:, :3, :3].transpose(1, 2) @ cp[:, :3, :3]
    assert float((RtR.cpu() - torch.eye(3)).abs().max()) < 1e-5


# ----------------------------------------------------------------------------- fused step
def _engine_case(mt, dev, cbs, cb, N, seed, softmax=True):
    g = torch.Generator().manual_seed(seed)
    M = cbs.poses.shape[0]
    sel = torch.randint(0, M, (N,), generator=g)
    poses = cbs.poses[sel].clone()
    obj = synth.make_object("004_sugar_box")
    gt, meas = synth.make_trajectory(obj, T=4, seed=seed)
    odom = torch.inverse(meas[0]) @ meas[1]
    torch.manual_seed(seed)
    tn, rot = O.draw_motion_noise(N, 2e-4, 0.5)
    q = synth.make_query(cbs, int(sel[0]), seed=seed)
    return poses, sel, odom, tn, rot, q, gt


@pytest.mark.parametrize("N,big", [(1024, False), (4099, False), (65536, True)])
def test_fused_step_vs_oracle(mt, dev, cb_small, cb_big, N, big):
    cbs, cb = cb_big if big else cb_small
    poses, sel, odom, tn, rot, q, gt = _engine_case(mt, dev, cbs, cb, N, seed=N)
    u = 0.618
    eng = mt.eng.FilterEngine(cb, capacity=N + 100)
    eng.load_particles(poses.to(dev), nn_hint=sel.int().to(dev))
    # stage 1: weighting only (kernel A)
    eng.step(q, odom, u=u, tn=tn.to(dev), rot=rot.to(dev), gt=gt[1], resample=False)
    keys_cb = cb.logmap_pose.cpu()
    moved, keep = O.motion_model(poses, odom, tn, rot)
    assert keep.all()
    got_moved = eng.poses().cpu()
    assert torch.allclose(got_moved, moved, rtol=RTOL, atol=ATOL)
    # NN on the CUDA-moved poses must be exactly the oracle's NN of those same poses/keys
    nn = eng.nn_idx().cpu().long()
    gkeys = mt.tt.R3_SE3(got_moved.to(dev)).cpu().numpy()
    assert np.array_equal(nn.numpy(), O.nn_exact(keys_cb.numpy(), gkeys))
    sim = O.codebook_similarity(q, cbs.embeddings)
    w_ref = torch.softmax(sim[nn], 0)
    w = eng.weights().cpu()
    assert torch.allclose(w, w_ref, rtol=1e-10, atol=0)
    assert abs(float(w.sum()) - 1.0) < 1e-12
    rt, rr = O.particle_rmse(moved, gt[1])
    r2 = eng.rmse.cpu()
    assert abs(float(r2[0]) - float(rt)) <= RTOL * float(rt) and abs(float(r2[1]) - float(rr)) <= RTOL * float(rr)
    # stage 2: a full step from the same start; ancestors vs the oracle's low_var on the CUDA weights
    eng2 = mt.eng.FilterEngine(cb, capacity=N + 100)
    eng2.load_particles(poses.to(dev))  # no hint: exercises the seeded search
    eng2.step(q, odom, u=u, tn=tn.to(dev), rot=rot.to(dev))
    anc = eng2.ancestors().cpu().long()
    ref_anc = O.low_var_indices(w, u)
    assert torch.equal(anc, ref_anc)
    assert torch.equal(eng2.poses().cpu(), got_moved[anc])
    assert torch.equal(eng2.nn_idx().cpu().long(), nn[anc])
    # and against the reference formulation end to end (gather N x D float64, then cosine)
    if N <= 4099:
        full = O.filter_step(poses, odom, tn, rot, keys_cb, cbs.embeddings, q, u)
        assert torch.equal(full["nn_idx"], nn) and torch.equal(full["anc"], anc)
        assert torch.allclose(w, full["weights"], rtol=1e-10)


def test_fused_step_with_prune_vs_oracle(mt, dev, cb_small, box):
    """loop order of filter.py:170-190 with remove_invalid_particles between weighting and
    resampling: drifted particles get weight 0 and never become ancestors."""
    cbs, cb = cb_small
    N = 6000
    poses, sel, odom, tn, rot, q, gt = _engine_case(mt, dev, cbs, cb, N, seed=21)
    poses[::4, :3, 3] += 0.005 * poses[::4, :3, 2]  # a quarter of the cloud floats 5 mm off the surface
    vds = box.vertices[::10]
    eng = mt.eng.FilterEngine(cb, capacity=N, mesh_vertices=vds, pen_max=0.002)
    eng.load_particles(poses.to(dev))
    eng.step(q, odom, u=0.37, tn=tn.to(dev), rot=rot.to(dev), resample=False)
    moved = eng.poses().cpu()
    nn = eng.nn_idx().cpu().long()
    w = eng.weights().cpu()
    w_soft = torch.softmax(O.codebook_similarity(q, cbs.embeddings)[nn], 0)
    w_ref, drifted = O.remove_invalid(moved, w_soft, vds, 0.002)
    assert not drifted and 0.15 < float((w_ref == 0).double().mean()) < 0.5
    assert torch.equal(w == 0, w_ref == 0)
    assert torch.allclose(w, w_ref / w_ref.sum(), rtol=1e-10, atol=0)
    st = cb.ctx.stats()
    assert st["drifted"] == 0 and st["on_surface"] == int((w_ref != 0).sum())
    eng2 = mt.eng.FilterEngine(cb, capacity=N, mesh_vertices=vds, pen_max=0.002)
    eng2.load_particles(poses.to(dev))
    eng2.step(q, odom, u=0.37, tn=tn.to(dev), rot=rot.to(dev))
    anc = eng2.ancestors().cpu().long()
    assert torch.equal(anc, O.low_var_indices(w_ref, 0.37))
    assert bool((w_ref[anc] > 0).all())
    assert torch.equal(eng2.poses().cpu(), moved[anc])


def test_fused_step_all_drifted_reprojects(mt, dev, cb_small, box):
    """filter.py:176-179: when every particle has drifted the poses are re-projected onto the
    codebook (SE3_NN) and the resampler keeps them (all-zero weights, particle_filter.py:240)."""
    cbs, cb = cb_small
    N = 3000
    poses, sel, odom, tn, rot, q, gt = _engine_case(mt, dev, cbs, cb, N, seed=22)
    poses[:, :3, 3] += 0.02 * poses[:, :3, 2]
    eng = mt.eng.FilterEngine(cb, capacity=N, mesh_vertices=box.vertices[::10], pen_max=0.002)
    eng.load_particles(poses.to(dev))
    eng.step(q, odom, u=0.5, tn=tn.to(dev), rot=rot.to(dev))
    moved, _ = O.motion_model(poses, odom, tn, rot)
    st = cb.ctx.stats()
    assert st["drifted"] == 1 and st["on_surface"] == 0 and st["resample_skipped"] == 1
    nn = eng.nn_idx().cpu().long()
    got_keys = mt.tt.R3_SE3(eng.poses()).cpu()
    assert torch.equal(eng.poses().cpu(), cbs.poses[nn])
    assert torch.equal(eng.ancestors().cpu().long(), torch.arange(N))
    ref_nn = O.se3_nn(O.r3_se3(cbs.poses), moved)
    assert float((nn != ref_nn).double().mean()) < 2e-3  # float32 key ulps only
    cb.ctx.stats(reset=True)


def test_engine_spatial_sort_is_a_permutation(mt, dev, cb_big):
    """load_particles(spatial_sort=True) only reorders: every per-particle result equals the
    unsorted engine's result at perm[i]."""
    cbs, cb = cb_big
    N = 20000
    poses, sel, odom, tn, rot, q, gt = _engine_case(mt, dev, cbs, cb, N, seed=31)
    e0 = mt.eng.FilterEngine(cb, capacity=N)
    e0.load_particles(poses.to(dev))
    e0.step(q, odom, u=0.1, tn=tn.to(dev), rot=rot.to(dev), resample=False)
    e1 = mt.eng.FilterEngine(cb, capacity=N)
    perm = e1.load_particles(poses.to(dev), spatial_sort=True)
    assert torch.equal(torch.sort(perm).values.cpu(), torch.arange(N))
    pc = perm.cpu()
    e1.step(q, odom, u=0.1, tn=tn[pc].to(dev), rot=rot[pc].to(dev), resample=False)
    assert torch.equal(e1.poses().cpu(), e0.poses().cpu()[pc])
    assert torch.equal(e1.nn_idx().cpu(), e0.nn_idx().cpu()[pc])
    assert torch.allclose(e1.weights().cpu(), e0.weights().cpu()[pc], rtol=1e-12, atol=0)
    rank = torch.empty(50000, dtype=torch.int32, device=dev)
    mt.lib.call("mt_codebook_rank", cb.ctx.h, rank.data_ptr(), mt.lib.stream_ptr())
    assert torch.equal(torch.sort(rank).values.cpu(), torch.arange(50000, dtype=torch.int32))
    r = rank[e1.nn_idx().long()].cpu()
    assert float((r[1:] >= r[:-1]).double().mean()) > 0.8  # still (nearly) in rank order after one step


def test_fused_step_philox_teacher_forced(mt, dev, cb_small):
    """perf mode (in-kernel Philox): recover the drawn noise from the standalone motion kernel
    with the same (seed, step, gid) and teacher-force the oracle with it."""
    cbs, cb = cb_small
    N = 8192
    poses, sel, odom, _, _, q, gt = _engine_case(mt, dev, cbs, cb, N, seed=5)
    eng = mt.eng.FilterEngine(cb, capacity=N, seed=99)
    eng.load_particles(poses.to(dev))
    eng.t = 3
    eng.step(q, odom, u=0.25)
    soa = mt.ctxm.aos_to_soa(poses.to(dev))
    out = torch.empty_like(soa)
    od = odom.contiguous()
    mt.lib.call("mt_motion", soa.data_ptr(), out.data_ptr(), N, N, od.data_ptr(), 0, 0, 2e-4, 0.5, 99, 3, 0, 0, 0, mt.lib.stream_ptr())
    moved = mt.ctxm.soa_to_aos(out, N).cpu()
    nn = torch.from_numpy(O.nn_exact(cb.logmap_pose.cpu().numpy(), mt.tt.R3_SE3(moved.to(dev)).cpu().numpy()))
    w = torch.softmax(O.codebook_similarity(q, cbs.embeddings)[nn], 0)
    anc = O.low_var_indices(w, 0.25)
    assert torch.equal(eng.ancestors().cpu().long(), anc)
    assert torch.equal(eng.poses().cpu(), moved[anc])


def test_fused_multi_step_invariants(mt, dev, cb_big):
    """N = 1e6 for 5 steps (BASELINE config 3 scale): size-independent properties."""
    cbs, cb = cb_big
    N = 1_000_000
    g = torch.Generator().manual_seed(0)
    sel = torch.randint(0, 50000, (N,), generator=g)
    obj = synth.make_object("004_sugar_box")
    gt, meas = synth.make_trajectory(obj, T=8)
    eng = mt.eng.FilterEngine(cb, capacity=N, seed=1)
    eng.load_particles(cbs.poses.to(dev)[sel.to(dev)], nn_hint=sel.int().to(dev))
    for t in range(1, 6):
        odom = torch.inverse(meas[t - 1]) @ meas[t]
        q = synth.make_query(cbs, int(sel[t]))
        eng.step(q, odom, gt=gt[t])
        anc = eng.ancestors()
        assert int(anc.min()) >= 0 and int(anc.max()) < N
        assert bool((anc[1:] >= anc[:-1]).all())  # systematic draw is sorted
        cnt = torch.bincount(anc.long(), minlength=N)
        assert int(cnt.sum()) == N and int(cnt.max()) <= 3  # weights within a factor e -> <= ceil(e) children
        nn = eng.nn_idx()
        assert int(nn.min()) >= 0 and int(nn.max()) < 50000
        p = eng.poses()
        RtR = p[:, :3, :3].transpose(1, 2) @ p[:, :3, :3]
        assert float((RtR - torch.eye(3, device=dev)).abs().max()) < 1e-4
        assert torch.isfinite(eng.rmse).all()
    # exactness spot check on a 20k subsample of the final NN assignment
    sub = torch.randperm(N, generator=g)[:20000]
    keys = mt.tt.R3_SE3(p[sub.to(dev)])
    # nn_idx holds the NN of the parents' moved poses == children poses
    ref = O.nn_exact(cb.logmap_pose.cpu().numpy(), keys.cpu().numpy())
    assert np.array_equal(eng.nn_idx()[sub.to(dev)].cpu().numpy().astype(np.int64), ref)


def test_missing_library_fails_loudly(mt, monkeypatch):
    monkeypatch.setattr(mt.lib, "LIB_PATH", "/nonexistent/libmidas_b200.so")
    monkeypatch.setattr(mt.lib, "_lib", None)
    with pytest.raises(mt.lib.MidasError):
        mt.lib.lib()


def test_cpu_tensor_rejected(mt, box):
    pf = mt.pf.particle_filter(synth_cfg(), box.vertices)
    with pytest.raises(mt.lib.MidasError):
        pf.get_similarity(torch.rand(1, 8, dtype=torch.float64), torch.rand(4, 8, dtype=torch.float64))


# ----------------------------------------------------------------------------- the loop (filter.py:131-233)
def test_filter_loop_dropin_and_engine_converge(mt, dev, box):
    """both forms of the loop -- the reference's sequence of drop-in calls and the resident engine --
    localise a sliding touch on the sugar box: translation RMSE falls from the global initialisation
    (~ object scale) to millimetres; filter_stats carries the reference's keys."""
    from midastouch_b200.config import compose
    from midastouch_b200.filter_loop import run_filter, run_filter_engine

    cfg = compose(overrides=["expt.params.num_particles=20000", "expt.params.resample=low_var"])
    cbs = synth.make_codebook(box, M=20000, D=64, seed=4, embedding="smooth")
    cb = mt.tt.tactile_tree(cbs.poses, cbs.cam_poses, cbs.embeddings)
    cb.to_device(dev)
    pf = mt.pf.particle_filter(cfg, box.vertices, downsample=1)
    gt, meas = synth.make_trajectory(box, T=80, seed=4)
    code_fn = lambda idx: synth.make_pose_query(gt[idx], 64, seed=4, frame=idx)  # noqa: E731
    torch.manual_seed(0)
    st = run_filter(cfg, pf, cb, lambda i: code_fn(i).to(dev), gt.to(dev), meas.to(dev), floor=5000)
    for k in ("rmse_t", "rmse_r", "time", "traj_size", "avg_time", "total_time", "cluster_poses", "cluster_stds", "obj_name",
              "tree_size", "noise_ratio", "init_noise", "init_particles", "num_particles", "log_id", "trial_id"):
        assert k in st, k
    assert st["traj_size"] == 80 and st["tree_size"] == 20000 and st["log_id"] == "00"
    assert st["rmse_t"][0] > 0.03 and st["rmse_t"][-1] < 0.012, (st["rmse_t"][0], st["rmse_t"][-1])
    assert 5000 <= st["num_particles"][-1] <= 20000 and st["cluster_poses"][-1].shape == (1, 4, 4)
    # the resident engine running the same loop body (cluster centres + annealing, varying particle count), teacher-forced
    # with the drop-in's own random draws: same particle-count trajectory, same convergence
    torch.manual_seed(0)
    torch.cuda.manual_seed(0)
    st0 = run_filter(cfg, mt.pf.particle_filter(cfg, box.vertices, downsample=1), cb, lambda i: code_fn(i).to(dev), gt.to(dev), meas.to(dev),
                     floor=5000, resample="low_var")
    torch.manual_seed(0)
    torch.cuda.manual_seed(0)
    se = run_filter_engine(cfg, mt.pf.particle_filter(cfg, box.vertices, downsample=1), cb, code_fn, gt.to(dev), meas.to(dev), floor=5000)
    assert se["rmse_t"][0] > 0.03 and se["rmse_t"][-1] < 0.012, (se["rmse_t"][0], se["rmse_t"][-1])
    a, b = np.array(st0["num_particles"]), np.array(se["num_particles"])
    assert a.shape == b.shape and b.min() >= 5000 and b[-1] < 20000
    assert np.abs(a - b).max() <= 0.02 * 20000, (a.tolist(), b.tolist())  # identical up to rare float64 weight ties
    first = lambda r: next(i for i, x in enumerate(r) if x < 0.012)  # noqa: E731
    assert abs(first(st0["rmse_t"]) - first(se["rmse_t"])) <= 3, (first(st0["rmse_t"]), first(se["rmse_t"]))
    assert se["engine"].ctx.stats()["overflow"] == 0
    # fixed-N form (one CUDA-graph replay per frame, in-kernel noise)
    torch.manual_seed(0)
    sf = run_filter_engine(cfg, mt.pf.particle_filter(cfg, box.vertices, downsample=1), cb, code_fn, gt.to(dev), meas.to(dev), anneal=False,
                           teacher_forced=False)
    assert sf["rmse_t"][0] > 0.03 and sf["rmse_t"][-1] < 0.012 and set(sf["num_particles"]) == {20000}


def test_sharded_engine_matches_single_gpu():
    """2 GPUs (skipped on a 1-GPU box): the sharded step reproduces the single-GPU children exactly and
    rebalance() preserves the global particle sequence (scripts/multigpu_check.py under torchrun)."""
    import os
    import subprocess
    import sys

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(root, "scripts", "multigpu_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("name,M", [("cotter-pin", 20000), ("025_mug", 30000), ("035_power_drill", 50000)])
def test_engine_on_other_objects(mt, dev, name, M):
    """the other BASELINE.json objects (tiny McMaster part, curved mug, L-shaped drill): one engine step
    from codebook poses, exact SE3_NN of the moved poses and ancestors vs the oracle on a subsample."""
    obj = synth.make_object(name)
    cbs = synth.make_codebook(obj, M=M, D=32, seed=6, embedding="smooth")
    cb = mt.tt.tactile_tree(cbs.poses, cbs.cam_poses, cbs.embeddings)
    cb.to_device(dev)
    N = 60000
    g = torch.Generator().manual_seed(8)
    sel = torch.randint(0, M, (N,), generator=g)
    poses = cbs.poses[sel]
    gt, meas = synth.make_trajectory(obj, T=6, seed=6, step=1e-4 if name == "cotter-pin" else 2.5e-4)
    sig_t = 1e-4 if name == "cotter-pin" else 2e-4
    tn = sig_t * torch.randn(N, 3, generator=g)
    rot = 0.5 * torch.randn(N, 3, generator=g)
    q = synth.make_pose_query(gt[1], 32, seed=6, frame=1)
    odom = torch.inverse(meas[0]) @ meas[1]
    eng = mt.eng.FilterEngine(cb, capacity=N, sig_t=sig_t, mesh_vertices=obj.vertices, pen_max=0.002)
    eng.load_particles(poses.to(dev), nn_hint=sel.int().to(dev))
    eng.step(q, odom, u=0.77, tn=tn.to(dev), rot=rot.to(dev), resample=False)
    moved = eng.poses()
    nn = eng.nn_idx().cpu().long()
    keys_cb = cb.logmap_pose.cpu().numpy()
    sub = torch.randperm(N, generator=g)[:8000]
    gk = mt.tt.R3_SE3(moved[sub.to(dev)]).cpu().numpy()
    assert np.array_equal(nn[sub].numpy(), O.nn_exact(keys_cb, gk, k=16))
    w = eng.weights().cpu()
    assert abs(float(w.sum()) - 1.0) < 1e-12
    eng2 = mt.eng.FilterEngine(cb, capacity=N, sig_t=sig_t, mesh_vertices=obj.vertices, pen_max=0.002)
    eng2.load_particles(poses.to(dev))
    eng2.step(q, odom, u=0.77, tn=tn.to(dev), rot=rot.to(dev))
    assert torch.equal(eng2.ancestors().cpu().long(), O.low_var_indices(w, 0.77))
    assert cb.ctx.stats()["overflow"] == 0


_FALLBACKS_64 = {}


@pytest.mark.parametrize("nbr_k", [64, 0])
def test_engine_heavy_fallback_is_exact(mt, dev, monkeypatch, nbr_k):
    """thin rod (cotter-pin stand-in): after a few steps the particles' rotations have drifted off the key
    manifold.  With 64-entry neighbour lists (forced) a large share of the hint scans is inconclusive and the
    box-hierarchy search carries the load; the matches must still be the exact nearest keys, and the hierarchy must
    prune (a search that visited every leaf would also be exact).  With the list length left to the upload (0: a
    codebook this dense gets longer lists) the same matches must come out with far fewer box searches."""
    if nbr_k:
        monkeypatch.setenv("MIDAS_B200_NBR_K", str(nbr_k))
    else:
        monkeypatch.delenv("MIDAS_B200_NBR_K", raising=False)
    obj = synth.make_object("cotter-pin")
    M, N, D = 20000, 40000, 32
    cbs = synth.make_codebook(obj, M=M, D=D, seed=9, embedding="smooth")
    cb = mt.tt.tactile_tree(cbs.poses, cbs.cam_poses, cbs.embeddings)
    cb.to_device(dev)
    kk = C.c_int(0)
    mt.lib.call("mt_codebook_nbr_info", cb.ctx.h, None, C.byref(kk))
    assert kk.value == 64 if nbr_k else kk.value in (128, 256), kk.value
    g = torch.Generator().manual_seed(9)
    sel = torch.randint(0, M, (N,), generator=g)
    gt, meas = synth.make_trajectory(obj, T=16, seed=9, step=1e-4)
    eng = mt.eng.FilterEngine(cb, capacity=N, sig_t=1e-4, seed=3, mesh_vertices=obj.vertices, pen_max=0.002)
    eng.load_particles(cbs.poses[sel].to(dev), nn_hint=sel.int().to(dev), spatial_sort=True)
    cb.ctx.stats(reset=True)
    for t in range(12):
        q = synth.make_pose_query(gt[t + 1], D, seed=9, frame=t)
        eng.step(q, torch.inverse(meas[t]) @ meas[t + 1], u=0.1 + 0.07 * t)
    q = synth.make_pose_query(gt[13], D, seed=9, frame=12)
    eng.step(q, torch.inverse(meas[12]) @ meas[13], u=0.5, resample=False)
    st = cb.ctx.stats()
    moved, nn = eng.poses(), eng.nn_idx().cpu().long().numpy()
    sub = torch.randperm(N, generator=g)[:6000]
    gk = mt.tt.R3_SE3(moved[sub.to(dev)]).cpu().numpy()
    assert np.array_equal(nn[sub.numpy()], O.nn_exact(cb.logmap_pose.cpu().numpy(), gk, k=32))
    n_leaf = cb.ctx.grid_info()[1][0]
    if nbr_k:
        assert st["nn_fallbacks"] > N // 4, st   # the fallback path really was exercised (13 steps x N particles)
        _FALLBACKS_64["n"] = st["nn_fallbacks"]
    elif "n" in _FALLBACKS_64:
        assert st["nn_fallbacks"] < 0.8 * _FALLBACKS_64["n"], (st, _FALLBACKS_64)  # the longer lists settle a good part of them
    assert 0 < st["grid_rows_max"] < (3 * n_leaf) // 4, (st, n_leaf)
    assert st["overflow"] == 0


def test_weighted_random_resampler(mt, dev, box):
    """resampler(..., "weighted_random") (the reference's default, particle_filter.py:243-250: N draws with
    replacement ~ weights): same distribution (the random stream is the library's Philox, not torch's), zero-weight
    particles are never drawn, reproducible under torch.manual_seed, guards as in the reference, and the draws
    equal the host model of the same arithmetic for a fixed seed."""
    pf = mt.pf.particle_filter(synth_cfg(), box.vertices)
    n = 200000
    g = torch.Generator().manual_seed(12)
    w = torch.rand(n, generator=g, dtype=torch.float64)
    w[torch.rand(n, generator=g) < 0.3] = 0.0
    w[-5:] = 0.0
    poses = torch.eye(4).repeat(n, 1, 1)
    poses[:, 0, 3] = torch.arange(n, dtype=torch.float32)
    labels = torch.arange(n, dtype=torch.float32)
    P = mt.pf.Particles(poses.to(dev), w.to(dev), labels.to(dev))
    torch.manual_seed(5)
    out = pf.resampler(P)  # default mode
    idx = out.labels.long().cpu()
    assert len(out) == n and torch.equal(out.poses[:, 0, 3].cpu(), idx.float()) and torch.equal(out.weights.cpu(), w[idx])
    assert bool((w[idx] > 0).all())
    # counts per block of 400 consecutive particles against their expectation (~400 draws each -> normal z-scores)
    counts = torch.bincount(idx // 400, minlength=n // 400).double()
    exp = (n * w / w.sum()).reshape(-1, 400).sum(1)
    z = (counts - exp) / exp.sqrt()
    assert float(z.abs().max()) < 5.0 and abs(float(z.mean())) < 0.15 and abs(float(z.std()) - 1.0) < 0.15
    torch.manual_seed(5)
    again = pf.resampler(P)
    assert torch.equal(again.labels, out.labels)
    other = pf.resampler(P)
    assert not torch.equal(other.labels, out.labels)
    # guards (237-241)
    z0 = pf.resampler(mt.pf.Particles(P.poses, torch.zeros_like(P.weights), P.labels))
    assert torch.equal(z0.poses, P.poses)
    wn = P.weights.clone()
    wn[5] = float("nan")
    z1 = pf.resampler(mt.pf.Particles(P.poses, wn, P.labels))
    assert torch.equal(z1.poses, P.poses)
    one = pf.resampler(mt.pf.Particles(P.poses[:1], P.weights[:1] + 1.0, P.labels[:1]))
    assert len(one) == 1 and torch.equal(one.poses, P.poses[:1])
    # C ABI against the host model of the same draw: sequential float64 CDF, u = Philox(seed; j, stream), first C_i > u*S
    m = 70000
    ctx = mt.pf._ctx_for(dev, m)
    wd = w[:m].to(dev).contiguous()
    idx32 = torch.empty(m, dtype=torch.int32, device=dev)
    scratch = torch.empty(m, dtype=torch.float64, device=dev)
    mt.lib.call("mt_resample_multinomial", ctx.h, wd.data_ptr(), m, m, 99, 7, scratch.data_ptr(), idx32.data_ptr(), None, mt.lib.stream_ptr())
    cdf = scratch.cpu().numpy()
    assert np.all(np.diff(cdf) >= 0) and np.all((np.diff(cdf) == 0) == (w[1:m].numpy() == 0))
    assert np.allclose(cdf, np.cumsum(w[:m].numpy()), rtol=1e-12)
    sp = C.c_void_p()
    mt.lib.call("mt_step_local_sum_ptr", ctx.h, C.byref(sp))
    S = float(mt.eng._device_double_view(sp.value + 4 * 8, dev).item())  # the total the resamplers normalise by
    assert cdf[-1] <= S and abs(S - cdf[-1]) <= 1e-12 * S
    u = philox_u01_53(99, 7, m)
    want = np.searchsorted(cdf, u * S, side="right")
    assert want.max() < m and np.array_equal(idx32.cpu().numpy(), want.astype(np.int32))

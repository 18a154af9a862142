"""CPU: the per-particle arithmetic header shared by every CUDA kernel
(midastouch_b200/csrc/mt_math.cuh) compiled for the host and checked against the oracle
and the golden vectors.  This is a test of the kernel source, not a CPU product path."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import oracle as O
from midastouch_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def H():
    build = os.path.join(HERE, "_build")
    os.makedirs(build, exist_ok=True)
    so = os.path.join(build, "host_math.so")
    src = os.path.join(HERE, "host_math_harness.cpp")
    hdrs = [os.path.join(HERE, "..", "midastouch_b200", "csrc", h) for h in ("mt_math.cuh", "mt_nn.cuh", "mt_cluster.cuh")]
    if not os.path.exists(so) or os.path.getmtime(so) < max([os.path.getmtime(src)] + [os.path.getmtime(h) for h in hdrs]):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-x", "c++", "-shared", "-fPIC", "-o", so, src])
    return ctypes.CDLL(so)


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_keys_match_oracle(H):
    obj = synth.make_object("004_sugar_box")
    cb = synth.make_codebook(obj, M=5000, D=8)
    poses = cb.poses.numpy().copy()
    keys = np.zeros((5000, 6), np.float32)
    H.h_se3_keys(P(poses), ctypes.c_longlong(5000), P(keys))
    ref = O.r3_se3(cb.poses).numpy()
    assert np.abs(keys - ref).max() < 2e-7
    # rotations near identity / near pi exercise both branches
    from scipy.spatial.transform import Rotation as R

    rng = np.random.default_rng(0)
    ax = rng.normal(size=(300, 3))
    ax /= np.linalg.norm(ax, axis=1, keepdims=True)
    ang = np.concatenate([np.full(100, 1e-3), np.full(100, np.pi - 1e-3), rng.uniform(0, np.pi, 100)])
    T = np.zeros((300, 4, 4), np.float32)
    T[:, :3, :3] = R.from_rotvec(ax * ang[:, None]).as_matrix()
    T[:, 3, 3] = 1
    keys = np.zeros((300, 6), np.float32)
    H.h_se3_keys(P(T), ctypes.c_longlong(300), P(keys))
    ref = O.r3_se3(torch.from_numpy(T)).numpy()
    assert np.abs(keys - ref).max() < 1e-6


def test_motion_matches_reference_golden(H, golden):
    g = golden("motion")
    poses = g["poses"].astype(np.float32).copy()
    n = poses.shape[0]
    out = np.zeros_like(poses)
    H.h_motion(P(poses), ctypes.c_longlong(n), P(g["odom"].astype(np.float32).copy()), P(g["tn"].copy()), P(g["rot_deg"].copy()), P(out))
    assert np.allclose(out, g["moved"], rtol=1e-5, atol=1e-6)
    assert np.abs(out - g["moved"]).max() < 5e-7


def test_key_dist_bit_exact(H):
    rng = np.random.default_rng(1)
    keys = (rng.normal(size=(4096, 6)) * 0.05).astype(np.float32)
    q = (rng.normal(size=6) * 0.05).astype(np.float32)
    out = np.zeros(4096, np.float32)
    H.h_key_dist(P(keys), ctypes.c_longlong(4096), P(q), P(out))
    assert np.array_equal(out, O.l2_sq_f32(keys, q))


@pytest.mark.parametrize("name", ["soft", "raw", "masked", "peaked"])
def test_slot_logic_matches_reference_loop(H, golden, name):
    g = golden("resample_low_var")
    for seed in (3, 4):
        w = torch.from_numpy(g[f"{name}_{seed}_w"])
        u = float(g[f"{name}_{seed}_u"][0])
        _, C = O.systematic_cdf(w)
        C = C.numpy().copy()
        n = len(C)
        anc = np.zeros(n, np.int64)
        H.h_ancestors_from_cdf(P(C), ctypes.c_longlong(n), ctypes.c_float(u), P(anc))
        filled = g[f"{name}_{seed}_filled"]
        assert np.array_equal(anc >= 0, filled)
        assert np.array_equal(anc[filled], g[f"{name}_{seed}_anc"][filled])


@pytest.mark.parametrize("n", [1, 2, 7, 1024, 65536, 1000003])
def test_locs_and_counts(H, n):
    for u in (0.0, 0.37, float(np.nextafter(np.float32(1), np.float32(0)))):
        locs = np.zeros(n)
        H.h_locs(ctypes.c_longlong(n), ctypes.c_float(u), P(locs))
        assert np.array_equal(locs, O.systematic_locs(n, u).numpy())
        w = torch.rand(n, dtype=torch.float64, generator=torch.Generator().manual_seed(n)) + 1e-3
        _, C = O.systematic_cdf(w)
        anc = np.zeros(n, np.int64)
        H.h_ancestors_from_cdf(P(C.numpy().copy()), ctypes.c_longlong(n), ctypes.c_float(u), P(anc))
        assert np.array_equal(anc, O.low_var_indices(w, u).numpy())


@pytest.mark.parametrize("n", [1, 2, 3, 7, 1000, 1024, 65536, 999983, 1000000, 1000003, (1 << 24) - 1, 16000000, (1 << 31) - 1])
def test_division_free_quotient_is_correctly_rounded(H, n):
    """mt_div_rn(k, N, fl(1/N)) == fl(k / N) bit for bit (the systematic sample locations k/N of particle_filter.py:254-261
    are computed without a division in the resampling kernel)"""
    rng = np.random.default_rng(n)
    starts = [0] if n <= 70000 else [0, n - 50000] + [int(x) for x in rng.integers(0, n - 50000, 6)]
    for k0 in starts:
        cnt = min(n, 70000) if k0 == 0 else 50000
        out = np.empty(cnt, dtype=np.float64)
        H.h_div_rn(ctypes.c_longlong(k0), ctypes.c_longlong(cnt), ctypes.c_longlong(n), out.ctypes.data_as(ctypes.c_void_p))
        want = np.arange(k0, k0 + cnt, dtype=np.float64) / np.float64(n)
        assert np.array_equal(out, want), (n, k0, int((out != want).sum()))


def test_philox_normals_statistics(H):
    n = 200000
    out = np.zeros((n, 6), np.float32)
    H.h_normals(ctypes.c_ulonglong(123), ctypes.c_ulonglong(5), ctypes.c_longlong(n), P(out))
    assert np.isfinite(out).all()
    assert np.abs(out.mean(0)).max() < 0.01 and np.abs(out.std(0) - 1).max() < 0.01
    c = np.corrcoef(out.T)
    assert np.abs(c - np.eye(6)).max() < 0.01
    out2 = np.zeros((n, 6), np.float32)
    H.h_normals(ctypes.c_ulonglong(123), ctypes.c_ulonglong(6), ctypes.c_longlong(n), P(out2))
    assert np.abs(np.corrcoef(out[:, 0], out2[:, 0])[0, 1]) < 0.01
    from scipy import stats

    assert stats.kstest(out[:, 3].astype(np.float64), "norm").pvalue > 1e-3


def test_rot_err_matches_reference_golden(H, golden):
    g = golden("rmse")
    poses = g["poses"].astype(np.float32).copy()
    n = poses.shape[0]
    out = np.zeros(n, np.float32)
    H.h_rot_err(P(g["gt"].astype(np.float32).copy()), P(poses), ctypes.c_longlong(n), P(out))
    rr = np.sqrt(np.mean(out.astype(np.float64) ** 2))
    assert abs(rr - float(g["rmse_r"])) <= 1e-5 * float(g["rmse_r"])


def _nbr_table(keys, K):
    """host model of k_build_nbr: the K nearest other keys of every key, ascending (distance, index)."""
    M = keys.shape[0]
    out = np.zeros((M, K, 8), np.float32)
    out[:, :, 6] = np.inf
    out[:, :, 7] = np.int32(-1).view(np.float32)
    for h in range(M):
        d = O.l2_sq_f32(keys, keys[h])
        d[h] = np.inf
        order = np.lexsort((np.arange(M), d))[: min(K, M - 1)]
        out[h, : len(order), :6] = keys[order]
        out[h, : len(order), 6] = np.sqrt(d[order].astype(np.float32))
        out[h, : len(order), 7] = order.astype(np.int32).view(np.float32)
    return out


@pytest.mark.parametrize("M,K", [(3000, 64), (3000, 8), (20, 64), (2, 64)])
def test_hint_graph_search_is_exact(H, M, K):
    """whenever the neighbour-list scan claims a proven answer it equals the exhaustive argmin
    (ties -> lowest index), for good, stale and random hints; duplicates included."""
    obj = synth.make_object("004_sugar_box")
    cb = synth.make_codebook(obj, M=max(M, 8), D=8, seed=2)
    keys = O.r3_se3(cb.poses).numpy()[:M].copy()
    if M >= 100:
        keys[50:60] = keys[40:50]  # duplicate keys -> ties
    nbr = _nbr_table(keys, K)
    rng = np.random.default_rng(1)
    n = 4000
    base = rng.integers(0, M, n)
    q = (keys[base] + rng.normal(size=(n, 6)).astype(np.float32) * np.float32(8e-4)).astype(np.float32)
    q[:100] = keys[base[:100]]  # exact hits
    ref = O.nn_brute(keys, q)
    for kind in ("true", "base", "random"):
        hint = {"true": ref, "base": base, "random": rng.integers(0, M, n)}[kind].astype(np.int32)
        idx = np.zeros(n, np.int32)
        ok = np.zeros(n, np.int32)
        dist = np.zeros(n, np.float32)
        H.h_hint_scan(P(keys), ctypes.c_longlong(M), P(nbr), K, P(q), ctypes.c_longlong(n), P(hint), P(idx), P(ok), P(dist))
        proven = ok.astype(bool)
        assert np.array_equal(idx[proven].astype(np.int64), ref[proven]), kind
        if M - 1 <= K:
            assert proven.all()  # the list holds every other key
        elif kind != "random" and K >= 32:
            assert proven.mean() > 0.9, (kind, proven.mean())
        # unproven queries still carry a real candidate for the grid search
        d_ref = O.l2_sq_f32(keys[idx], q)
        assert np.array_equal(d_ref, dist)


def test_so3_to_quaternion_matches_oracle(H):
    from scipy.spatial.transform import Rotation as R

    rng = np.random.default_rng(4)
    ax = rng.normal(size=(600, 3))
    ax /= np.linalg.norm(ax, axis=1, keepdims=True)
    ang = np.concatenate([rng.uniform(0, np.pi, 400), np.full(100, np.pi - 1e-4), np.full(100, 1e-4)])
    T = np.zeros((600, 4, 4), np.float32)
    T[:, :3, :3] = R.from_rotvec(ax * ang[:, None]).as_matrix()
    T[:, 3, 3] = 1
    q = np.zeros((600, 4), np.float32)
    H.h_so3_to_quat(P(T), ctypes.c_longlong(600), P(q))
    ref = O.so3_to_quaternion(torch.from_numpy(T[:, :3, :3].copy())).numpy()
    assert np.abs(q - ref).max() < 2e-6


def test_box_hierarchy_search_is_exact(H):
    """the fallback search index (6-D Morton order, boxes of 32 keys, two 32-ary levels; mt_nn.cuh) built and
    searched on the host with the pruning rule the CUDA kernel uses: exact argmin (ties -> lowest index) for
    queries on the key manifold, far off it, outside the bounding box, with and without a seed candidate, with
    duplicate keys, and np.argmin semantics for NaN queries -- while visiting a small part of the leaves."""
    rng = np.random.default_rng(11)
    M = 6000
    # keys on a curved 3-D manifold embedded in 6-D (position on a thin rod, normal direction, yaw), like a codebook
    u, th, yaw = rng.uniform(0, 0.03, M), rng.uniform(0, 2 * np.pi, M), rng.uniform(-np.pi, np.pi, M)
    keys = np.stack([u, 0.0015 * np.cos(th), 0.0015 * np.sin(th), 0.01 * th, 0.01 * yaw, 0.005 * np.sin(yaw + th)], 1).astype(np.float32)
    keys[100] = keys[7]          # duplicates: the lower index must win
    keys[5000] = keys[7]
    near = keys[rng.integers(0, M, 600)] + rng.normal(0, 2e-4, (600, 6)).astype(np.float32)
    far = keys[rng.integers(0, M, 600)] + rng.normal(0, 8e-3, (600, 6)).astype(np.float32)
    outside = (rng.uniform(-0.3, 0.3, (100, 6))).astype(np.float32)
    exact = keys[[7, 100, 5000, 42]].copy()
    q = np.ascontiguousarray(np.concatenate([near, far, outside, exact]).astype(np.float32))
    n = q.shape[0]
    want = O.nn_brute(keys, q)
    assert want[-4] == 7 and want[-3] == 7 and want[-2] == 7
    dims = np.zeros(3, dtype=np.int32)
    for seeded in (False, True):
        seeds = rng.integers(0, M, n).astype(np.int32) if seeded else np.full(n, -1, dtype=np.int32)
        idx, vis = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
        rc = H.h_bvh_search(P(keys), ctypes.c_longlong(M), P(q), ctypes.c_longlong(n), P(seeds), P(idx), P(vis), P(dims))
        assert rc == 0 and tuple(dims) == ((M + 31) // 32, ((M + 31) // 32 + 31) // 32, 1)
        assert np.array_equal(idx, want)
        assert vis.mean() < dims[0] / 3  # (this host loop is not best-first: an upper bound of the kernel's work)
    # a good seed (the true neighbour) leaves only the verification
    idx, vis = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
    H.h_bvh_search(P(keys), ctypes.c_longlong(M), P(q), ctypes.c_longlong(n), P(want.astype(np.int32)), P(idx), P(vis), P(dims))
    assert vis[:600].mean() < 8 and vis[600:1200].mean() < 25  # verification only: a handful of leaves
    assert np.array_equal(idx, want)
    # NaN query -> index 0 (np.argmin of all-NaN distances); NaN key -> build refuses
    qn = q[:2].copy()
    qn[0, 3] = np.nan
    idx2, vis2 = np.zeros(2, dtype=np.int32), np.zeros(2, dtype=np.int32)
    H.h_bvh_search(P(keys), ctypes.c_longlong(M), P(qn), ctypes.c_longlong(2), P(np.full(2, -1, dtype=np.int32)), P(idx2), P(vis2), P(dims))
    assert idx2[0] == 0 and idx2[1] == want[1]
    bad = keys.copy()
    bad[3, 2] = np.nan
    assert H.h_bvh_search(P(bad), ctypes.c_longlong(M), P(q), ctypes.c_longlong(1), P(np.full(1, -1, dtype=np.int32)), P(idx2), P(vis2), P(dims)) == -1


def test_search_index_is_a_kd_tree_layout(H):
    """mt_bvh_build orders the keys as the leaves of a balanced k-d tree (median split of the widest coordinate, left part
    a multiple of the level's node size): the order is a permutation, leaves are stored in index order, level-1 nodes
    (1024 keys) and leaves (32 keys) are whole subtrees -- sibling boxes are disjoint along some coordinate -- and on a curved
    codebook-like manifold the leaf boxes are clearly smaller than those of runs of a 6-D Morton order."""
    rng = np.random.default_rng(5)
    M = 50000
    u, th, yaw = rng.uniform(0, 0.05, M), rng.uniform(0, 2 * np.pi, M), rng.uniform(-0.3, 0.3, M)
    keys = np.stack([u, 0.0015 * np.cos(th), 0.0015 * np.sin(th), 0.01 * np.cos(th) * 3, 0.01 * np.sin(th) * 3, 0.01 * yaw], 1).astype(np.float32)
    n_leaf, n_l1 = (M + 31) // 32, ((M + 31) // 32 + 31) // 32
    order, leaf, l1 = np.zeros(M, np.int32), np.zeros((n_leaf, 12), np.float32), np.zeros((n_l1, 12), np.float32)
    assert H.h_bvh_layout(P(keys), ctypes.c_longlong(M), P(order), P(leaf), P(l1)) == 0
    assert np.array_equal(np.sort(order), np.arange(M))
    full = order[: (M // 32) * 32].reshape(-1, 32)
    assert (np.diff(full, axis=1) > 0).all()  # leaves in index order
    ks = keys[order]
    for j in range(n_leaf):  # the stored boxes are the boxes of the leaves' keys
        blk = ks[32 * j:32 * j + 32]
        assert np.array_equal(leaf[j, :6], blk.min(0)) and np.array_equal(leaf[j, 6:], blk.max(0))
    # two leaves of the same level-1 node never overlap in all six coordinates at once (interiors): they were separated by a split
    for g in (0, 7, n_l1 - 2):
        L = leaf[32 * g:32 * g + 32]
        lo, hi = L[:, None, :6], L[None, :, 6:]
        overlap = (np.maximum(L[:, None, :6], L[None, :, :6]) < np.minimum(L[:, None, 6:], L[None, :, 6:])).all(2)
        np.fill_diagonal(overlap, False)
        assert not overlap.any()
    # against runs of 32 of the round-1 order (6-D Morton code): the mean box diagonal is a third smaller
    lo6, hi6 = keys.min(0), keys.max(0)
    cell = (hi6 - lo6).max() / 1023.0
    qk = np.clip(np.floor((keys - lo6) / cell), 0, 1023).astype(np.uint64)
    code = np.zeros(M, dtype=np.uint64)
    for bit in range(9, -1, -1):
        for k in range(6):
            code = (code << np.uint64(1)) | ((qk[:, k] >> np.uint64(bit)) & np.uint64(1))
    mo = np.argsort(code, kind="stable")
    km = keys[mo][: (M // 32) * 32].reshape(-1, 32, 6)
    diag_morton = np.linalg.norm(km.max(1) - km.min(1), axis=1).mean()
    diag_kd = np.linalg.norm(leaf[: M // 32, 6:] - leaf[: M // 32, :6], axis=1).mean()
    assert diag_kd < 0.75 * diag_morton, (diag_kd, diag_morton)
    # deterministic
    order2 = np.zeros(M, np.int32)
    H.h_bvh_layout(P(keys), ctypes.c_longlong(M), P(order2), P(leaf.copy()), P(l1.copy()))
    assert np.array_equal(order, order2)


def test_categorical_draw_from_cdf(H):
    """mt_cdf_draw / mt_u01_53 (the multinomial resampler's per-draw arithmetic): equals
    searchsorted(C, u*S, side='right'), never returns an item of zero weight, handles u*S landing on or
    beyond the last CDF value, and the 53-bit uniforms are in [0,1) and uniform."""
    rng = np.random.default_rng(2)
    n = 5000
    w = rng.random(n)
    w[rng.random(n) < 0.3] = 0.0
    w[-7:] = 0.0                       # trailing zero-weight items
    C = np.cumsum(w)
    S = float(C[-1])
    nd = 200000
    idx, u = np.zeros(nd, dtype=np.int32), np.zeros(nd)
    H.h_cdf_draws(P(C), ctypes.c_longlong(n), ctypes.c_double(S), ctypes.c_ulonglong(1234), ctypes.c_ulonglong(0),
                  ctypes.c_longlong(nd), P(idx), P(u))
    assert (u >= 0).all() and (u < 1).all() and abs(u.mean() - 0.5) < 5e-3 and abs(np.mean(u < 0.1) - 0.1) < 5e-3
    from test_gpu_parity import philox_u01_53  # the numpy model the GPU test uses: pinned to the C source here
    assert np.array_equal(u, philox_u01_53(1234, 0, nd))
    want = np.searchsorted(C, u * S, side="right")
    assert np.array_equal(idx, np.minimum(want, n - 8))  # (u*S == S cannot pick a trailing zero-weight item)
    assert (w[idx] > 0).all()
    counts = np.bincount(idx, minlength=n)
    z = (counts - nd * w / S) / np.sqrt(np.maximum(nd * w / S, 1e-9))
    assert np.abs(z[w > 0]).max() < 6.0
    # a CDF whose top is below S (chunk clamping): draws beyond it fall on the last item of positive weight
    idx2 = np.zeros(4, dtype=np.int32)
    H.h_cdf_draws(P(C), ctypes.c_longlong(n), ctypes.c_double(S * (1 + 1e-3)), ctypes.c_ulonglong(5), ctypes.c_ulonglong(1),
                  ctypes.c_longlong(4), P(idx2), None)
    assert (w[idx2] > 0).all()
    one = np.array([3.0])
    H.h_cdf_draws(P(one), ctypes.c_longlong(1), ctypes.c_double(3.0), ctypes.c_ulonglong(5), ctypes.c_ulonglong(1), ctypes.c_longlong(4), P(idx2), None)
    assert (idx2 == 0).all()

"""The TCN oracle (sparse float64 restatement of MinkLoc3D, oracle/tcn_oracle.py) against a dense
torch.nn.functional.conv3d evaluation of the same network with re-masking after every layer --
the pin of the sparse-convolution semantics in the absence of MinkowskiEngine."""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import tcn_oracle as T


def dense_weight(W, k, cin, cout):
    """(k^3, Cin, Cout) with x fastest -> conv3d weight (Cout, Cin, kz, ky, kx)"""
    return torch.from_numpy(np.asarray(W, np.float64).reshape(k, k, k, cin, cout)).permute(4, 3, 0, 1, 2).contiguous()


def dense_forward(coords, P, G):
    g = lambda n: np.asarray(P[n], np.float64)  # noqa: E731

    def bn(x, n):
        sh = (1, -1, 1, 1, 1)
        w, b, m, v = (torch.from_numpy(g(f"{n}.bn.{k}")).reshape(sh) for k in ("weight", "bias", "running_mean", "running_var"))
        return (x - m) / torch.sqrt(v + 1e-5) * w + b

    B = int(coords[:, 0].max()) + 1
    m0 = torch.zeros((B, 1, G, G, G), dtype=torch.float64)
    m0[coords[:, 0], 0, coords[:, 3], coords[:, 2], coords[:, 1]] = 1.0  # (b, c, z, y, x)
    x = F.conv3d(m0, dense_weight(g("backbone.conv0.kernel"), 5, 1, 32), padding=2)
    x = torch.relu(bn(x, "backbone.bn0")) * m0
    mask, inpl, fm = m0, 32, None
    for s, pl in enumerate((32, 64, 64)):
        x = F.conv3d(x, dense_weight(g(f"backbone.convs.{s}.kernel"), 2, inpl, inpl), stride=2)
        mask = F.max_pool3d(mask, 2)
        x = torch.relu(bn(x, f"backbone.bn.{s}")) * mask
        b = f"backbone.blocks.{s}.0"
        y = torch.relu(bn(F.conv3d(x, dense_weight(g(f"{b}.conv1.kernel"), 3, inpl, pl), padding=1), f"{b}.norm1")) * mask
        y = bn(F.conv3d(y, dense_weight(g(f"{b}.conv2.kernel"), 3, pl, pl), padding=1), f"{b}.norm2") * mask
        res = x
        if f"{b}.downsample.0.kernel" in P:
            res = bn(F.conv3d(x, dense_weight(g(f"{b}.downsample.0.kernel"), 1, inpl, pl)), f"{b}.downsample.1") * mask
        x = torch.relu(y + res) * mask
        inpl = pl
        if s == 1:
            fm = (x, mask)
    x = F.conv3d(x, dense_weight(g("backbone.conv1x1.0.kernel"), 1, 64, 256)) * mask
    wt = torch.from_numpy(g("backbone.tconvs.0.kernel").reshape(2, 2, 2, 256, 256)).permute(3, 4, 0, 1, 2).contiguous()
    x = F.conv_transpose3d(x, wt, stride=2) * fm[1]
    x = x + F.conv3d(fm[0], dense_weight(g("backbone.conv1x1.1.kernel"), 1, 64, 256)) * fm[1]
    p = float(P["pooling.p"][0])
    out = torch.zeros((B, 256), dtype=torch.float64)
    for bi in range(B):
        act = fm[1][bi, 0] > 0
        v = x[bi][:, act]  # (256, n_active)
        out[bi] = (v.clamp(min=1e-6) ** p).mean(dim=1) ** (1.0 / p)
    return out.numpy(), x, fm[1]


def test_sparse_oracle_equals_dense_conv3d():
    rng = np.random.default_rng(0)
    G = 16
    clouds = []
    for b in range(2):
        # a thin curved sheet (what a tactile contact patch looks like) + a few strays
        xy = rng.integers(0, G, size=(150, 2))
        z = np.clip((4 + 0.3 * xy[:, 0] + rng.integers(-1, 2, 150)), 0, G - 1).astype(np.int64)
        c = np.unique(np.concatenate([np.stack([xy[:, 0], xy[:, 1], z], 1), rng.integers(0, G, size=(10, 3))]), axis=0)
        clouds.append(c)
    coords = T.batched(clouds)
    P = T.random_state_dict(seed=3)
    sparse, trace = T.minkloc_forward(coords, P)
    dense, xd, maskd = dense_forward(coords, P, G)
    assert np.allclose(sparse, dense, rtol=1e-9, atol=1e-12)
    # the per-point FPN output too, not just the pooled descriptor
    fc, fx = trace["fpn"]
    got = xd[fc[:, 0], :, fc[:, 3] // 4, fc[:, 2] // 4, fc[:, 1] // 4].numpy()
    assert np.allclose(fx, got, rtol=1e-9, atol=1e-10)
    assert int(maskd.sum()) == fc.shape[0]


def test_quantize_and_front_end():
    rng = np.random.default_rng(1)
    cloud = rng.uniform(-0.01, 0.01, size=(500, 3)).astype(np.float32)
    sc = T.scale_cloud(cloud)
    assert sc.min() == -1.0 and sc.max() == 1.0
    q = T.quantize(sc, 0.001)
    assert q.min() >= -1000 and q.max() <= 1000 and len(np.unique(q, axis=0)) == len(q)
    # negative coordinates floor towards -inf
    assert (T.quantize(np.array([[-0.0005, 0.0005, -0.0015]], np.float32), 0.001) == np.array([[-1, 0, -2]])).all()
    d = T.down_coords(np.array([[0, -1, 0, 3], [0, -2, 1, 2]]), 1)
    assert (d == np.array([[0, -2, 0, 2]])).all()
    depth = np.full((6, 4), 0.02)
    mask = np.zeros((6, 4))
    mask[2:4, 1:3] = 1
    pts = T.heightmap_to_pointcloud(depth, mask, f=100.0, width=4, height=6)
    assert pts.shape == (4, 3) and np.allclose(pts[:, 2], -0.02)

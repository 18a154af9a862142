"""TEST INFRASTRUCTURE ONLY -- CPU oracle of the tactile code network (TCN = MinkLoc3D).

Restates, in float64 numpy, the forward pass the reference runs through MinkowskiEngine:
``TCN.cloud_to_tactile_code`` (contrib/tcn_minkloc/tcn.py:52-148), ``MinkLoc.forward`` /
``GeM`` (minkloc.py:45-95), ``MinkFPN.forward`` (minkfpn.py:110-138) and ME's ``BasicBlock``
(conv3-bn-relu-conv3-bn + residual, relu).

PARITY UNPINNED: MinkowskiEngine is neither vendored nor pinned by the reference
(README.md:85-87) and is not installable here, and the trained weights
(tcn_weights.pth.tar) are not in the container.  The sparse-tensor semantics restated here are
the published ones -- a convolution is evaluated at the active output coordinates only, with
inactive inputs contributing nothing; stride-2 convolutions create the coordinates
floor(c / 2s) * 2s; a transposed convolution writes onto the coordinate map that already
exists at the finer stride; kernel offset i <-> (ox, oy, oz) with x fastest, centred for odd
kernel sizes and 0..k-1 for even ones; kernels are stored (k^3, Cin, Cout).  They are pinned
against a dense ``torch.nn.functional.conv3d`` evaluation with re-masking after every layer
(tests/test_oracle_tcn.py); they are NOT pinned against MinkowskiEngine itself.
"""
from __future__ import annotations

import numpy as np


# ------------------------------------------------------------------ coordinates
def quantize(cloud: np.ndarray, q: float) -> np.ndarray:
    """ME.utils.sparse_quantize (tcn.py:124-130): floor(coords / q), unique rows (sorted)."""
    c = np.floor(np.asarray(cloud, dtype=np.float32) / np.float32(q)).astype(np.int64)
    return np.unique(c, axis=0)


def batched(coords_list) -> np.ndarray:
    """ME.utils.batched_coordinates: (n,4) [batch, x, y, z]."""
    return np.concatenate([np.concatenate([np.full((c.shape[0], 1), b, np.int64), c], 1) for b, c in enumerate(coords_list)], 0)


def _pack(c: np.ndarray) -> np.ndarray:
    off = 1 << 17  # 18 bits per coordinate (|c| < 131072), 9 bits of batch index
    return ((c[:, 0] << 54) | ((c[:, 1] + off) << 36) | ((c[:, 2] + off) << 18) | (c[:, 3] + off)).astype(np.int64)


def _lookup(keys_sorted, order, query):
    pos = np.searchsorted(keys_sorted, query)
    pos = np.clip(pos, 0, len(keys_sorted) - 1)
    hit = keys_sorted[pos] == query
    return np.where(hit, order[pos], -1)


def kernel_offsets(k: int, dil: int) -> np.ndarray:
    """offset of kernel index i, x fastest; centred for odd k, 0..k-1 for even k."""
    r = np.arange(k) - (k // 2 if k % 2 else 0)
    oz, oy, ox = np.meshgrid(r, r, r, indexing="ij")
    return np.stack([ox.ravel(), oy.ravel(), oz.ravel()], 1) * dil


def down_coords(coords: np.ndarray, stride: int) -> np.ndarray:
    c = coords.copy()
    c[:, 1:] = np.floor_divide(c[:, 1:], 2 * stride) * (2 * stride)
    return np.unique(c, axis=0)


# ------------------------------------------------------------------ layers
def conv(in_coords, in_feats, out_coords, W, k, dil):
    """out[p] = sum_i in[p + offset_i] @ W[i] over the active inputs (regular and strided conv)."""
    keys = _pack(in_coords)
    order = np.argsort(keys)
    ks = keys[order]
    W = np.asarray(W, np.float64).reshape(k**3, in_feats.shape[1], -1)
    out = np.zeros((out_coords.shape[0], W.shape[2]))
    for i, o in enumerate(kernel_offsets(k, dil)):
        q = out_coords.copy()
        q[:, 1:] += o
        idx = _lookup(ks, order, _pack(q))
        ok = idx >= 0
        out[ok] += in_feats[idx[ok]] @ W[i]
    return out


def conv_transpose(coarse_coords, coarse_feats, fine_coords, W, fine_stride):
    """k=2, stride=2 transposed conv onto the existing finer coordinate map."""
    keys = _pack(coarse_coords)
    order = np.argsort(keys)
    ks = keys[order]
    W = np.asarray(W, np.float64).reshape(8, coarse_feats.shape[1], -1)
    out = np.zeros((fine_coords.shape[0], W.shape[2]))
    for i, o in enumerate(kernel_offsets(2, fine_stride)):
        q = fine_coords.copy()
        q[:, 1:] -= o
        aligned = np.all(np.mod(q[:, 1:], 2 * fine_stride) == 0, axis=1)
        idx = _lookup(ks, order, _pack(q))
        ok = (idx >= 0) & aligned
        out[ok] += coarse_feats[idx[ok]] @ W[i]
    return out


def bn(x, p, eps=1e-5):
    return (x - p["running_mean"]) / np.sqrt(p["running_var"] + eps) * p["weight"] + p["bias"]


def relu(x):
    return np.maximum(x, 0.0)


# ------------------------------------------------------------------ the network
def minkloc_forward(coords: np.ndarray, P: dict, planes=(32, 64, 64), conv0_k=5):
    """coords (n,4) int64 [b,x,y,z] unique; P: float64 state dict with the reference's parameter
    names (``backbone.conv0.kernel`` ...).  Returns (B, feature_size) GeM descriptors and the
    per-level intermediate tensors (for tests)."""
    g = lambda n: np.asarray(P[n], np.float64)  # noqa: E731
    bnp = lambda n: {k: g(f"{n}.bn.{k}") for k in ("weight", "bias", "running_mean", "running_var")}  # noqa: E731
    x = np.ones((coords.shape[0], 1))
    x = relu(bn(conv(coords, x, coords, g("backbone.conv0.kernel"), conv0_k, 1), bnp("backbone.bn0")))
    c, stride, fmaps, trace = coords, 1, [], {"conv0": x}
    for s in range(3):
        cd = down_coords(c, stride)
        x = relu(bn(conv(c, x, cd, g(f"backbone.convs.{s}.kernel"), 2, stride), bnp(f"backbone.bn.{s}")))
        c, stride = cd, stride * 2
        b = f"backbone.blocks.{s}.0"
        y = relu(bn(conv(c, x, c, g(f"{b}.conv1.kernel"), 3, stride), bnp(f"{b}.norm1")))
        y = bn(conv(c, y, c, g(f"{b}.conv2.kernel"), 3, stride), bnp(f"{b}.norm2"))
        res = x
        if f"{b}.downsample.0.kernel" in P:
            res = bn(conv(c, x, c, g(f"{b}.downsample.0.kernel"), 1, stride), bnp(f"{b}.downsample.1"))
        x = relu(y + res)
        trace[f"stage{s}"] = (c, x)
        if s == 1:  # num_bottom_up - 1 - num_top_down <= ndx < len(convs) - 1  (minkfpn.py:124)
            fmaps.append((c, x, stride))
    x = conv(c, x, c, g("backbone.conv1x1.0.kernel"), 1, stride)
    fc, fx, fs = fmaps[-1]
    x = conv_transpose(c, x, fc, g("backbone.tconvs.0.kernel"), fs) + conv(fc, fx, fc, g("backbone.conv1x1.1.kernel"), 1, fs)
    trace["fpn"] = (fc, x)
    p, eps = float(np.asarray(P["pooling.p"]).reshape(-1)[0]), 1e-6
    B = int(coords[:, 0].max()) + 1
    out = np.zeros((B, x.shape[1]))
    for b in range(B):
        sel = fc[:, 0] == b
        out[b] = np.mean(np.maximum(x[sel], eps) ** p, axis=0) ** (1.0 / p)  # GeM (minkloc.py:84-95)
    return out, trace


def l2_normalize(x, eps=1e-12):
    return x / np.maximum(np.linalg.norm(x, axis=1, keepdims=True), eps)


def random_state_dict(seed=0, planes=(32, 64, 64), feature=256, conv0_k=5, scale=1.0):
    """random parameters with the reference's names and MinkowskiEngine's shapes."""
    rng = np.random.default_rng(seed)
    P = {}

    def convp(name, kvol, cin, cout):
        w = rng.normal(size=(kvol, cin, cout)) * scale * np.sqrt(2.0 / (kvol * cin) * 4)
        P[name] = w if kvol > 1 else w[0]

    def bnp(name, c):
        P[f"{name}.bn.weight"] = rng.uniform(0.5, 1.5, c)
        P[f"{name}.bn.bias"] = rng.normal(size=c) * 0.1
        P[f"{name}.bn.running_mean"] = rng.normal(size=c) * 0.1
        P[f"{name}.bn.running_var"] = rng.uniform(0.5, 1.5, c)

    convp("backbone.conv0.kernel", conv0_k**3, 1, planes[0])
    bnp("backbone.bn0", planes[0])
    inpl = planes[0]
    for s, pl in enumerate(planes):
        convp(f"backbone.convs.{s}.kernel", 8, inpl, inpl)
        bnp(f"backbone.bn.{s}", inpl)
        b = f"backbone.blocks.{s}.0"
        convp(f"{b}.conv1.kernel", 27, inpl, pl)
        bnp(f"{b}.norm1", pl)
        convp(f"{b}.conv2.kernel", 27, pl, pl)
        bnp(f"{b}.norm2", pl)
        if inpl != pl:
            convp(f"{b}.downsample.0.kernel", 1, inpl, pl)
            bnp(f"{b}.downsample.1", pl)
        inpl = pl
    convp("backbone.conv1x1.0.kernel", 1, planes[2], feature)
    convp("backbone.tconvs.0.kernel", 8, feature, feature)
    convp("backbone.conv1x1.1.kernel", 1, planes[1], feature)
    P["pooling.p"] = np.array([3.0])
    return P


# ------------------------------------------------------------------ point-cloud front end
def heightmap_to_pointcloud(depth: np.ndarray, mask, f: float, width: int, height: int) -> np.ndarray:
    """``heightmap2Pointcloud`` (digit_renderer.py:210-248) after ``correct_image_height_map``
    (the caller passes the corrected depth): pixel grid -> metres, masked-off points dropped."""
    hv = depth * mask if mask is not None else depth
    ys, xs = np.meshgrid(np.arange(depth.shape[0]), np.arange(depth.shape[1]), indexing="ij")
    x = (xs - width / 2.0) / f * depth
    y = -((ys - height / 2.0) / f) * depth
    pts = np.stack([x.reshape(-1), y.reshape(-1), -hv.reshape(-1)], 1)
    return pts[pts[:, 2] != 0]


def scale_cloud(cloud: np.ndarray) -> np.ndarray:
    """global min-max to [-1, 1] over all coordinates (tcn.py:111-116), float32 like the reference."""
    c = np.asarray(cloud, np.float32)
    lo, hi = c.min(), c.max()
    return (np.float32(2.0) * (c - lo) / (hi - lo) - np.float32(1.0)).astype(np.float32)

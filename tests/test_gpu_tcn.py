"""GPU parity of the tactile code network (mt_tcn_forward through the drop-in TCN class) against
the float64 sparse oracle (oracle/tcn_oracle.py, itself pinned to dense conv3d).  float32 kernels
vs a float64 oracle: descriptors agree to 2e-4 relative (the bar written here; the reference's own
float32 MinkowskiEngine path is not bit-defined either)."""
import types

import numpy as np
import pytest
import torch

from oracle import tcn_oracle as T

pytestmark = pytest.mark.gpu


def tcn_cfg(num_points=4096, batch_size=100):
    m = types.SimpleNamespace(tcn_weights="tcn_weights.pth.tar", model="MinkFPN", num_points=num_points, batch_size=batch_size,
                              mink_quantization_size=0.001, planes="32,64,64", layers="1,1,1", num_top_down=1,
                              conv0_kernel_size=5, feature_size=256, output_dim=256)
    return types.SimpleNamespace(model=m, train=types.SimpleNamespace(normalize_embeddings=True))


def torch_quantize(clouds):
    """voxel coordinates exactly as the reference gets them: ME.utils.sparse_quantize divides the CUDA
    tensor by the Python scalar with torch (tcn.py:124-130), which is a multiply by the reciprocal on
    CUDA -- so the oracle is teacher-forced with torch's result rather than numpy's true division."""
    out = []
    for c in clouds:
        ijk = torch.floor(torch.from_numpy(np.asarray(c, np.float32)).cuda() / 0.001).to(torch.int64).cpu().numpy()
        out.append(np.unique(ijk, axis=0))
    return out


def contact_cloud(rng, n=6000):
    """a curved contact patch like a DIGIT height map: (n,3) metres"""
    xy = rng.uniform(-0.008, 0.008, size=(n, 2))
    z = 0.02 - 0.5 * (xy[:, 0] ** 2 + 2 * xy[:, 1] ** 2) / 0.01 + rng.normal(size=n) * 2e-5
    return np.concatenate([xy, z[:, None]], 1).astype(np.float32)


@pytest.fixture(scope="module")
def net():
    from midastouch_b200.tcn import TCN

    P = T.random_state_dict(seed=5)
    tcn = TCN(tcn_cfg(), device="cuda:0", weights={k: torch.from_numpy(np.asarray(v)) for k, v in P.items()})
    return tcn, P


def test_embed_clouds_vs_oracle(net):
    tcn, P = net
    rng = np.random.default_rng(0)
    clouds = np.stack([T.scale_cloud(contact_cloud(rng)[:4096]) for _ in range(3)])
    out = tcn.embed_clouds(torch.from_numpy(clouds).cuda()).cpu().numpy()
    coords = T.batched(torch_quantize(clouds))
    # the oracle is parameterised by float32-rounded weights (what the library holds)
    P32 = {k: np.asarray(v, np.float32).astype(np.float64) for k, v in P.items()}
    ref, trace = T.minkloc_forward(coords, P32)
    ref = T.l2_normalize(ref)
    assert out.shape == (3, 256) and out.dtype == np.float64
    assert np.allclose(np.linalg.norm(out, axis=1), 1.0, atol=1e-12)
    assert np.allclose(out, ref, rtol=2e-4, atol=1e-6), np.abs(out - ref).max()
    # deterministic: same input, same bits
    out2 = tcn.embed_clouds(torch.from_numpy(clouds).cuda()).cpu().numpy()
    assert np.array_equal(out, out2)
    # batch elements are independent
    one = tcn.embed_clouds(torch.from_numpy(clouds[1:2]).cuda()).cpu().numpy()
    assert np.allclose(one[0], out[1], rtol=1e-6, atol=1e-9)


def test_level_counts_and_negative_coordinates(net):
    tcn, P = net
    from midastouch_b200 import _lib
    from midastouch_b200.tcn import pack_coordinates

    rng = np.random.default_rng(3)
    ijk = np.unique(rng.integers(-40, 40, size=(3000, 3)), axis=0)
    coords = T.batched([ijk])
    keys = torch.unique(pack_coordinates(torch.zeros(len(ijk), dtype=torch.int64), torch.from_numpy(ijk))).cuda()
    out = torch.empty((1, 256), dtype=torch.float64, device="cuda")
    counts = torch.zeros(4, dtype=torch.int32, device="cuda")
    tcn._ensure(len(ijk), 1)
    _lib.call("mt_tcn_forward", tcn._h, keys.data_ptr(), len(ijk), 1, 0, out.data_ptr(), counts.data_ptr(), _lib.stream_ptr())
    c = coords
    want = [len(c)]
    for s in (1, 2, 4):
        c = T.down_coords(c, s)
        want.append(len(c))
    assert counts.cpu().tolist() == want
    P32 = {k: np.asarray(v, np.float32).astype(np.float64) for k, v in P.items()}
    ref, _ = T.minkloc_forward(coords, P32)
    assert np.allclose(out.cpu().numpy(), ref, rtol=2e-4, atol=1e-6)


def test_cloud_to_tactile_code_api(net):
    """tcn.py:52-148 end to end: height map + mask -> code; the torch.multinomial draw is
    reproduced on the oracle side with the same generator state."""
    tcn, P = net
    from midastouch_b200.tcn import PointcloudRenderer

    rng = np.random.default_rng(7)
    H, W, f = 120, 160, 200.0
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    depth = (0.02 + 0.002 * np.exp(-((xs - 80) ** 2 + (ys - 60) ** 2) / 900.0) + rng.normal(size=(H, W)) * 1e-5).astype(np.float32)
    mask = (((xs - 80) ** 2 + (ys - 60) ** 2) < 45**2).astype(np.float32)
    rend = PointcloudRenderer(f, W, H)
    small = TCNsmall(tcn)
    torch.manual_seed(11)
    code = small.cloud_to_tactile_code(rend, torch.from_numpy(depth).cuda(), torch.from_numpy(mask).cuda())
    assert code.shape == (1, 256) and code.dtype == torch.float64
    # oracle front end + the same sampling draw
    pts = T.heightmap_to_pointcloud(depth.astype(np.float64), mask.astype(np.float64), f, W, H).astype(np.float32)
    got_pts = rend.heightmap2Pointcloud(torch.from_numpy(depth).cuda(), torch.from_numpy(mask).cuda()).cpu().numpy()
    assert np.allclose(got_pts, pts, rtol=1e-6, atol=1e-9)
    torch.manual_seed(11)
    idxs = torch.arange(pts.shape[0], device="cuda", dtype=torch.float)
    ids = torch.multinomial(idxs, num_samples=small.num_points, replacement=small.num_points > pts.shape[0]).cpu().numpy()
    cloud = T.scale_cloud(got_pts[ids])
    P32 = {k: np.asarray(v, np.float32).astype(np.float64) for k, v in P.items()}
    ref, _ = T.minkloc_forward(T.batched(torch_quantize([cloud])), P32)
    assert np.allclose(code.cpu().numpy(), T.l2_normalize(ref), rtol=2e-4, atol=1e-6)
    # empty contact -> num_points zeros -> a single voxel (tcn.py:89-94); must not crash
    z = small.cloud_to_tactile_code(rend, torch.from_numpy(depth).cuda(), torch.zeros(H, W).cuda())
    assert z.shape == (1, 256)


def TCNsmall(tcn):
    """same network, 2048 sample points (keeps the oracle fast)"""
    tcn.num_points = 2048
    return tcn


def test_cpu_input_rejected(net):
    from midastouch_b200._lib import MidasError

    with pytest.raises(MidasError):
        net[0].embed_clouds(torch.zeros(1, 16, 3))


def test_embed_entry_equals_keys_entry(net):
    """mt_tcn_embed (float clouds: quantisation + duplicate removal in the library, rows in order of first
    occurrence) against mt_tcn_forward on torch's sorted unique keys: the same voxel sets, the same per-point
    sums; only the GeM summation order differs (float64 partial sums)."""
    tcn, P = net
    from midastouch_b200 import _lib
    from midastouch_b200.tcn import pack_coordinates

    rng = np.random.default_rng(21)
    # dense enough for real neighbourhoods and duplicate voxels: coordinates span +-60 voxels
    clouds = (rng.uniform(-0.06, 0.06, size=(3, 3000, 3))).astype(np.float32)
    clouds[:, :, 2] = 0.3 * (clouds[:, :, 0] ** 2 + clouds[:, :, 1] ** 2) * 10
    cl = torch.from_numpy(clouds).cuda()
    got = tcn.embed_clouds(cl)
    B, Pn, _ = cl.shape
    ijk = torch.floor(cl.reshape(-1, 3) / tcn.quantization_size).to(torch.int64)
    keys = torch.unique(pack_coordinates(torch.arange(B, device="cuda").repeat_interleave(Pn), ijk))
    assert keys.numel() < B * Pn  # duplicates present
    out = torch.empty((B, 256), dtype=torch.float64, device="cuda")
    counts = torch.zeros(4, dtype=torch.int32, device="cuda")
    _lib.call("mt_tcn_forward", tcn._h, keys.data_ptr(), keys.numel(), B, 1, out.data_ptr(), counts.data_ptr(), _lib.stream_ptr())
    assert int(counts[0]) == keys.numel()
    assert torch.isfinite(got).all()
    assert torch.allclose(got, out, rtol=1e-9, atol=1e-12), (got - out).abs().max()
    # and against the oracle on the same voxels (dense neighbourhoods exercise every kernel offset)
    coords = T.batched(torch_quantize(clouds))
    P32 = {k: np.asarray(v, np.float32).astype(np.float64) for k, v in P.items()}
    ref, _ = T.minkloc_forward(coords, P32)
    assert np.allclose(got.cpu().numpy(), T.l2_normalize(ref), rtol=2e-4, atol=1e-6)


def test_voxel_out_of_range_poisons(net):
    tcn, _ = net
    cl = torch.zeros((1, 64, 3), device="cuda")
    cl[0, 5, 1] = 200.0  # 200 / 0.001 voxels: beyond the 18-bit key fields
    assert torch.isnan(tcn.embed_clouds(cl)).all()
    cl[0, 5, 1] = 0.5
    assert torch.isfinite(tcn.embed_clouds(cl)).all()


def test_many_clouds_ordered_compaction(net):
    """40 clouds x 1100 points: more than 32 compaction blocks (strided carry), batch ranges of every element,
    each cloud's code equal to the code it gets alone."""
    tcn, _ = net
    rng = np.random.default_rng(33)
    clouds = rng.uniform(-0.2, 0.2, size=(40, 1100, 3)).astype(np.float32)
    clouds[:, :, 2] = 0.5 * clouds[:, :, 0] * clouds[:, :, 1]
    clouds[7] = clouds[3]  # two identical clouds: identical codes
    cl = torch.from_numpy(clouds).cuda()
    out = tcn.embed_clouds(cl)
    assert out.shape == (40, 256) and torch.isfinite(out).all()
    assert torch.equal(out[7], out[3])
    for b in (0, 3, 19, 39):
        one = tcn.embed_clouds(cl[b:b + 1])
        assert torch.allclose(one[0], out[b], rtol=1e-9, atol=1e-12), (b, (one[0] - out[b]).abs().max())

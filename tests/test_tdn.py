"""Tactile depth network drop-in (midastouch_b200/tdn.py) against golden vectors of the UNMODIFIED reference
(contrib/tdn_fcrn/fcrn.py FCRN_net + tdn.py TDN methods, run by oracle/gen_golden_tdn.py with the seeded synthetic
parameters of oracle/tdn_oracle.py).  Plain torch: runs on the CPU here.  Tolerance: the drop-in folds 65 BatchNorms
into their convolutions and evaluates an up-projection's eight convolutions as one, so float32 sums are ordered
differently through ~60 layers: 2e-3 relative on the height map (measured: ~1e-4)."""
import os
import types

import numpy as np
import pytest
import torch

from oracle import tdn_oracle as TO

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tdn_fcrn.npz")


def cfg(real=False):
    fc = types.SimpleNamespace(real=types.SimpleNamespace(blend_sz=10, border=10, ratio=0.9, clip=5, batch_size=1),
                               sim=types.SimpleNamespace(blend_sz=0, border=1, ratio=0.2, clip=5, batch_size=1))
    return types.SimpleNamespace(tdn_weights="tdn_weights.pth.tar", fcrn=fc)


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.fixture(scope="module")
def net():
    from midastouch_b200.tdn import TDN

    torch.set_num_threads(max(torch.get_num_threads(), 4))
    return TDN(cfg(), device="cpu", weights={"state_dict": TO.synthetic_fcrn_state(seed=7)})


def test_parameter_table_matches_state_dict_layout():
    from midastouch_b200.tdn import fcrn_parameter_shapes

    S = fcrn_parameter_shapes()
    # ResNet-50 trunk (53 convolutions) + neck + 4 x (8 + 1) up-projection convolutions + head
    convs = [k for k, v in S.items() if len(v) == 4]
    assert len(convs) == 53 + 1 + 36 + 1
    assert S["layer3.5.conv2.weight"] == (256, 256, 3, 3) and S["up1.conv1_2.weight"] == (512, 1024, 2, 3)
    assert S["up4.conv2_3.weight"] == (64, 128, 3, 2) and S["conv3.weight"] == (1, 64, 3, 3)
    n = sum(int(np.prod(v)) for v in S.values())
    assert 63_000_000 < n < 64_500_000  # FCRN-ResNet50-UpProj


def test_image2heightmap_vs_reference(net, gold):
    for k in (0, 1):
        hm = net.image2heightmap(TO.synthetic_tactile_image(seed=k + 1))
        assert hm.shape == (320, 240) and hm.dtype == torch.float32
        got = hm.numpy()
        want = gold[f"heightmap{k}_sub"]
        scale = float(np.abs(want).max())
        assert np.allclose(got[::4, ::4], want, rtol=2e-3, atol=2e-4 * scale), np.abs(got[::4, ::4] - want).max() / scale
        st = gold[f"heightmap{k}_stats"]
        assert np.allclose([got.mean(), got.std(), got.min(), got.max()], st, rtol=2e-3, atol=1e-3 * scale)


def test_bottleneck_embedding_vs_reference(net, gold):
    from midastouch_b200.tdn import TDN

    t = TDN(cfg(), device="cpu", bottleneck=True)
    t.model = net.model
    z = t.model(t._image_tensor(TO.synthetic_tactile_image(seed=1)), bottleneck=True)
    assert tuple(z.shape) == tuple(gold["bottleneck_shape"]) == (1, 1024, 10, 8)
    want = gold["bottleneck_sub"]
    got = z[0, ::16, ::2, ::2].numpy()
    assert np.allclose(got, want, rtol=2e-3, atol=2e-4 * float(np.abs(want).max()))
    f = t.image2embedding(TO.synthetic_tactile_image(seed=1))
    assert f.shape == (1, 10 * 8 * 1024) and abs(float(f.norm()) - 1.0) < 1e-5
    with pytest.raises(AssertionError):
        t.image2heightmap(TO.synthetic_tactile_image(seed=1))  # bottleneck mode: tdn.py:105-107


def test_heightmap2mask_vs_reference(gold):
    from midastouch_b200.tdn import TDN

    hm, bg = torch.from_numpy(gold["mask_heightmap"]), gold["mask_bg"]
    for tag, real in (("sim", False), ("real", True)):
        t = TDN(cfg(), bg=bg, real=real, device="cpu")
        keep = hm.clone()
        m = t.heightmap2mask(hm)
        assert m.dtype == torch.bool and np.array_equal(m.numpy(), gold[f"mask_{tag}"])
        assert np.array_equal(t.heightmap2mask(hm, small_parts=True).numpy(), gold[f"mask_{tag}_small"])
        assert torch.equal(hm, keep)  # the caller's height map is not modified
    t = TDN(cfg(), bg=bg, device="cpu")
    assert np.array_equal(t.heightmap2mask(torch.from_numpy(bg) + 1.0).numpy(), gold["mask_empty"]) and not gold["mask_empty"].any()


def test_blend_heightmaps_vs_reference(gold):
    from midastouch_b200.tdn import TDN

    t = TDN(cfg(), device="cpu")
    t.blend_sz = 3
    outs = [t.blend_heightmaps(torch.from_numpy(f)).numpy() for f in gold["blend_in"]]
    assert np.allclose(np.stack(outs), gold["blend_out"], rtol=1e-6, atol=1e-7)
    t.blend_sz = 0  # sim: blending off (config/tdn/default.yaml:32)
    x = torch.ones(3, 3)
    assert t.blend_heightmaps(x) is x


def test_minmax_normalisation_matches_opencv():
    from midastouch_b200.tdn import normalize_minmax_255

    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    for img in (TO.synthetic_tactile_image(seed=4), rng.integers(17, 203, (32, 24, 3)).astype(np.uint8),
                rng.normal(size=(16, 12, 3)).astype(np.float32)):
        want = cv2.normalize(img, None, alpha=0, beta=255, norm_type=cv2.NORM_MINMAX)
        got = normalize_minmax_255(img)
        assert got.dtype == want.dtype
        if img.dtype == np.uint8:
            assert np.abs(got.astype(int) - want.astype(int)).max() <= 1 and (got != want).mean() < 0.01
        else:
            assert np.allclose(got, want, rtol=1e-5, atol=1e-4)


def test_missing_or_misshapen_weights_are_refused():
    from midastouch_b200.tdn import FCRN, TDN

    S = TO.synthetic_fcrn_state(seed=1)
    bad = dict(S)
    del bad["up2.conv2_4.weight"]
    with pytest.raises(KeyError):
        FCRN(bad, "cpu")
    bad = dict(S)
    bad["layer1.0.conv1.weight"] = torch.zeros(64, 64, 3, 3)
    with pytest.raises(ValueError):
        FCRN(bad, "cpu")
    with pytest.raises(RuntimeError):
        TDN(cfg(), device="cpu").image2heightmap(TO.synthetic_tactile_image(seed=1))


@pytest.mark.gpu
def test_image2heightmap_on_the_gpu(gold):
    """the same network through cuDNN: float32 convolutions against the reference's CPU golden at the CPU bar; with torch's
    default TF32 convolutions (what the reference itself runs with on an Ampere+ GPU) within 3 %."""
    from midastouch_b200.tdn import TDN

    t = TDN(cfg(), device="cuda:0", weights=TO.synthetic_fcrn_state(seed=7))
    img = TO.synthetic_tactile_image(seed=1)
    want = gold["heightmap0_sub"]
    scale = float(np.abs(want).max())
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        got = t.image2heightmap(img)
    assert got.is_cuda and got.shape == (320, 240)
    assert np.allclose(got.cpu().numpy()[::4, ::4], want, rtol=2e-3, atol=2e-4 * scale)
    got_tf32 = t.image2heightmap(img).cpu().numpy()[::4, ::4]
    assert np.abs(got_tf32 - want).max() < 3e-2 * scale
    bg = torch.from_numpy(gold["mask_bg"]).cuda()
    t2 = TDN(cfg(), bg=gold["mask_bg"], device="cuda:0")
    m = t2.heightmap2mask(torch.from_numpy(gold["mask_heightmap"]).cuda())
    assert m.is_cuda and np.array_equal(m.cpu().numpy(), gold["mask_sim"]) and bg.is_cuda


@pytest.mark.gpu
def test_frame_front_end_image_to_code(gold):
    """filter.py:140-147 on the drop-ins: image -> TDN height map -> mask -> TCN code (random weights: shapes, dtypes,
    determinism, and an all-background frame giving the empty-contact code)"""
    from midastouch_b200.filter_loop import tactile_code_fn
    from midastouch_b200.tcn import TCN, PointcloudRenderer
    from midastouch_b200.tdn import TDN
    from oracle import tcn_oracle as TC

    tdn = TDN(cfg(), bg=np.zeros((320, 240), np.float32), device="cuda:0", weights=TO.synthetic_fcrn_state(seed=7))
    m = types.SimpleNamespace(tcn_weights="", model="MinkFPN", num_points=2048, batch_size=100, mink_quantization_size=0.001,
                              planes="32,64,64", layers="1,1,1", num_top_down=1, conv0_kernel_size=5, feature_size=256, output_dim=256)
    tcn = TCN(types.SimpleNamespace(model=m, train=types.SimpleNamespace(normalize_embeddings=True)), device="cuda:0",
              weights={k: torch.from_numpy(np.asarray(v)) for k, v in TC.random_state_dict(seed=5).items()})
    images = [TO.synthetic_tactile_image(seed=s) for s in (1, 2)]
    fn = tactile_code_fn(tdn, tcn, PointcloudRenderer(200.0, 240, 320), images)
    torch.manual_seed(3)
    c0 = fn(0)
    assert c0.shape == (1, 256) and c0.dtype == torch.float64 and c0.is_cuda and torch.isfinite(c0).all()
    assert abs(float(c0.norm()) - 1.0) < 1e-9
    torch.manual_seed(3)
    assert torch.equal(fn(0), c0)  # same image, same sampling seed, same bits
    assert not torch.equal(fn(1), c0)


def test_device_side_normalisation_equals_the_numpy_restatement():
    from midastouch_b200.tdn import TDN, normalize_minmax_255

    t = TDN(cfg(), device="cpu")
    rng = np.random.default_rng(1)
    for img in (TO.synthetic_tactile_image(seed=5), rng.integers(3, 90, (40, 30, 3)).astype(np.uint8), np.full((8, 6, 3), 7, np.uint8),
                rng.normal(size=(16, 12, 3)).astype(np.float32)):
        want = torch.from_numpy(normalize_minmax_255(img)).permute(2, 0, 1).float()[None]
        got = t._image_tensor(img)
        assert got.shape == want.shape and got.dtype == torch.float32
        assert torch.equal(got, want) if img.dtype == np.uint8 else torch.allclose(got, want, rtol=1e-6, atol=1e-5)
        assert torch.equal(t._image_tensor(torch.from_numpy(img)), got)  # tensors are accepted as well

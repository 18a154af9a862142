"""TEST INFRASTRUCTURE ONLY.  Golden vectors of the tactile depth network from the UNMODIFIED reference
(contrib/tdn_fcrn/fcrn.py FCRN_net, tdn.py TDN methods), run here on the CPU with seeded synthetic parameters:

    python -m oracle.gen_golden_tdn      ->  tests/golden/tdn_fcrn.npz

Checks first that ``midastouch_b200.tdn.fcrn_parameter_shapes()`` equals the reference module's state dict (names and
shapes): that pins the parameter naming the drop-in loads ``tdn_weights.pth.tar`` by.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_shim, tdn_oracle  # noqa: E402
from midastouch_b200.tdn import fcrn_parameter_shapes  # noqa: E402


def main():
    torch.manual_seed(0)
    fcrn, tdn = ref_shim.load_reference_tdn()
    net = fcrn.FCRN_net(1, bottleneck=False)
    ref_shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    mine = {k: tuple(v) for k, v in fcrn_parameter_shapes().items()}
    assert ref_shapes == mine, (set(ref_shapes) ^ set(mine), [k for k in mine if k in ref_shapes and mine[k] != ref_shapes[k]][:5])
    assert list(net.state_dict().keys()) == list(fcrn_parameter_shapes().keys())
    state = tdn_oracle.synthetic_fcrn_state(seed=7)
    net.load_state_dict(state)
    net.eval()
    out = {}
    T = object.__new__(tdn.TDN)  # the reference class without its constructor (which needs weights on disk, tdn.py:38-50)
    T.model, T.device = net, torch.device("cpu")
    T.b, T.r, T.clip, T.batch_size, T.blend_sz = 1, 0.2, 5, 1, 0  # config/tdn/default.yaml:31-36 (sim)
    T.heightmap_window = __import__("collections").deque([])
    imgs = [tdn_oracle.synthetic_tactile_image(seed=s) for s in (1, 2)]
    hms = []
    for k, im in enumerate(imgs):
        hm = T.image2heightmap(im)  # tdn.py:94-115: cv2.normalize + FCRN_net.forward + blend (off in sim)
        hms.append(hm.numpy().copy())
        out[f"heightmap{k}_sub"] = hms[-1][::4, ::4].astype(np.float32)
        out[f"heightmap{k}_stats"] = np.array([hms[-1].mean(), hms[-1].std(), hms[-1].min(), hms[-1].max()], np.float64)
    net.bottleneck = True
    import cv2

    with torch.no_grad():  # the lines of image2embedding (tdn.py:131-134) up to the network output
        im0 = cv2.normalize(imgs[0], None, alpha=0, beta=255, norm_type=cv2.NORM_MINMAX)
        z = net(torch.from_numpy(im0).permute(2, 0, 1).float()[None])
    out["bottleneck_sub"] = z[0, ::16, ::2, ::2].numpy().astype(np.float32)
    out["bottleneck_shape"] = np.array(z.shape)
    net.bottleneck = False
    # heightmap2mask (tdn.py:139-165) on a scaled height map against a background, sim and real thresholds
    rng = np.random.default_rng(3)
    scale = 60.0 / max(float(hms[0].max()), 1e-9)
    hm_s = torch.from_numpy(np.ascontiguousarray(hms[0][::2, ::2]) * scale)  # (160, 120): keeps the fixture small
    bg = torch.from_numpy((hms[1][::2, ::2] * scale * 0.2 + rng.normal(0, 0.5, (160, 120))).astype(np.float32))
    out["mask_heightmap"], out["mask_bg"] = hm_s.numpy(), bg.numpy()
    for tag, (b, r, clip) in {"sim": (1, 0.2, 5), "real": (10, 0.9, 5)}.items():
        T.b, T.r, T.clip, T.bg = b, r, clip, bg
        out[f"mask_{tag}"] = T.heightmap2mask(hm_s.clone()).numpy()
        out[f"mask_{tag}_small"] = T.heightmap2mask(hm_s.clone(), small_parts=True).numpy()
    out["mask_empty"] = T.heightmap2mask(bg.clone() + 1.0).numpy()  # below the clip everywhere -> no contact
    # blend_heightmaps (tdn.py:60-92), window of 3 over 5 frames
    T.blend_sz, T.heightmap_window = 3, __import__("collections").deque([])
    frames = [torch.from_numpy(rng.normal(size=(6, 5)).astype(np.float32)) for _ in range(5)]
    out["blend_in"] = np.stack([f.numpy() for f in frames])
    out["blend_out"] = np.stack([T.blend_heightmaps(f).numpy() for f in frames])
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "tdn_fcrn.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})
    print("heightmap stats", out["heightmap0_stats"], "mask pixels", {k: int(v.sum()) for k, v in out.items() if k.startswith("mask_") and v.dtype == bool})


if __name__ == "__main__":
    main()

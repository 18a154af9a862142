"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Seeded synthetic parameters and inputs for the tactile depth network (there are no TDN weights in the container):
``synthetic_fcrn_state(seed)`` fills every entry of ``FCRN_net.state_dict()`` (fcrn.py:174-272; names and shapes from
``midastouch_b200.tdn.fcrn_parameter_shapes``, which ``oracle/gen_golden_tdn.py`` checks against the reference module)
with values that keep the activations of the 50-layer network in range: convolutions He-normal (fcrn.py:208-211),
BatchNorm scale in [0.8, 1.2], shift and running mean ~ N(0, 0.05), running variance in [0.8, 1.2] -- i.e. every
BatchNorm is non-trivial, so that folding errors would show.
"""
import numpy as np
import torch


def synthetic_fcrn_state(seed: int = 0) -> dict:
    from midastouch_b200.tdn import fcrn_parameter_shapes

    g = torch.Generator().manual_seed(seed)
    S = {}
    for name, shape in fcrn_parameter_shapes().items():
        if name.endswith("num_batches_tracked"):
            S[name] = torch.tensor(0, dtype=torch.long)
        elif len(shape) == 4:
            cout, cin, kh, kw = shape
            S[name] = torch.randn(shape, generator=g) * float(np.sqrt(2.0 / (kh * kw * cout)))
        elif name.endswith("running_var") or (name.endswith(".weight") and len(shape) == 1):
            S[name] = 0.8 + 0.4 * torch.rand(shape, generator=g)
        else:  # BatchNorm bias / running mean, convolution bias
            S[name] = 0.05 * torch.randn(shape, generator=g)
    return S


def synthetic_tactile_image(seed: int = 0, h: int = 320, w: int = 240) -> np.ndarray:
    """a DIGIT-like frame: smooth background + a bright blob, uint8 (h, w, 3)"""
    rng = np.random.default_rng(seed)
    ys, xs = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    base = 90 + 30 * np.sin(xs / 37.0)[..., None] + 20 * np.cos(ys / 51.0)[..., None] + np.array([10.0, -5.0, 20.0])
    blob = 80 * np.exp(-((xs - 0.6 * w) ** 2 + (ys - 0.4 * h) ** 2) / (2 * 28.0**2))[..., None] * np.array([1.0, 0.7, 0.4])
    img = base + blob + rng.normal(0, 2.0, (h, w, 3))
    return np.clip(img, 0, 255).astype(np.uint8)

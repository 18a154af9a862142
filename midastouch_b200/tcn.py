"""Drop-in for ``midastouch/contrib/tcn_minkloc/tcn.py`` (reference lines 18-148).

``TCN(cfg).cloud_to_tactile_code(tac_render, heightmaps, masks) -> (B, D) float64`` with the
MinkLoc3D forward pass (minkloc.py:45-95, minkfpn.py:110-138) running in libmidas_b200
(``mt_tcn_forward``, csrc/mt_tcn.cuh) instead of MinkowskiEngine.  Parameters keep the
reference's state-dict names (``backbone.conv0.kernel``, ``backbone.bn0.bn.weight`` ...,
``pooling.p``) so ``tcn_weights.pth.tar`` loads unchanged.  Point sampling stays
``torch.multinomial`` exactly as in the reference (tcn.py:96-108): it is the RNG contract.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from ._lib import MidasError, call, ptr, stream_ptr

_CONV_IDS = {"backbone.conv0.kernel": 0, "backbone.conv1x1.0.kernel": 13, "backbone.tconvs.0.kernel": 14,
             "backbone.conv1x1.1.kernel": 15}
_BN_IDS = {"backbone.bn0": 0}
for _s in range(3):
    _CONV_IDS[f"backbone.convs.{_s}.kernel"] = 1 + _s
    _CONV_IDS[f"backbone.blocks.{_s}.0.conv1.kernel"] = 4 + 2 * _s
    _CONV_IDS[f"backbone.blocks.{_s}.0.conv2.kernel"] = 5 + 2 * _s
    _CONV_IDS[f"backbone.blocks.{_s}.0.downsample.0.kernel"] = 10 + _s
    _BN_IDS[f"backbone.bn.{_s}"] = 1 + _s
    _BN_IDS[f"backbone.blocks.{_s}.0.norm1"] = 4 + 2 * _s
    _BN_IDS[f"backbone.blocks.{_s}.0.norm2"] = 5 + 2 * _s
    _BN_IDS[f"backbone.blocks.{_s}.0.downsample.1"] = 10 + _s


def pack_coordinates(batch_idx: torch.Tensor, ijk: torch.Tensor) -> torch.Tensor:
    """(n,) batch index + (n,3) integer voxel coordinates -> int64 keys of mt_tcn_forward."""
    o = 1 << 17
    ijk = ijk.to(torch.int64)
    if ijk.numel() and (int(ijk.abs().max()) >= o):
        raise MidasError("TCN: voxel coordinate outside +-131071")
    return (batch_idx.to(torch.int64) << 54) | ((ijk[:, 0] + o) << 36) | ((ijk[:, 1] + o) << 18) | (ijk[:, 2] + o)


class PointcloudRenderer:
    """the 39 lines of ``digit_renderer.heightmap2Pointcloud`` (digit_renderer.py:210-248) the TCN
    needs online; ``f``/``width``/``height`` are the TACTO camera constants (``renderer.f`` lives in
    the tacto fork, not in the reference tree).  The depth passed in is the corrected ("cam" frame)
    height map."""

    def __init__(self, f: float, width: int, height: int):
        self.f, self.width, self.height = float(f), int(width), int(height)

    def heightmap2Pointcloud(self, depth: torch.Tensor, contact_mask: torch.Tensor = None) -> torch.Tensor:
        hv = depth * contact_mask if contact_mask is not None else depth
        xv = torch.arange(hv.shape[1], device=hv.device)
        yv = torch.arange(hv.shape[0], device=hv.device)
        y, x = torch.meshgrid(yv, xv, indexing="ij")
        x = (x - self.width / 2.0) / self.f * depth
        y = -((y - self.height / 2.0) / self.f) * depth
        pts = torch.hstack((x.reshape(-1, 1), y.reshape(-1, 1), -hv.reshape(-1, 1)))
        return pts[pts[:, 2] != 0]


class TCN:
    def __init__(self, cfg, device=None, weights: dict | str | None = None):
        m = cfg.model
        if "MinkFPN" not in m.model:
            raise NotImplementedError("Model not implemented: {}".format(m.model))
        self.num_points = int(m.num_points)
        self.batch_size = int(m.batch_size)
        self.quantization_size = float(m.mink_quantization_size)
        self.planes = [int(e) for e in str(m.planes).split(",")]
        self.layers = [int(e) for e in str(m.layers).split(",")]
        if self.layers != [1, 1, 1] or int(m.num_top_down) != 1 or len(self.planes) != 3:
            raise NotImplementedError("TCN: libmidas_b200 implements the shipped topology (layers 1,1,1; num_top_down 1)")
        self.conv0_kernel_size = int(m.conv0_kernel_size)
        self.feature_size, self.output_dim = int(m.feature_size), int(m.output_dim)
        assert self.feature_size == self.output_dim, "output_dim must be the same as feature_size"
        train = getattr(cfg, "train", None)
        self.normalize_embeddings = bool(getattr(train, "normalize_embeddings", True)) if train is not None else True
        self.device = torch.device(device if device is not None else "cuda:0")
        if self.device.type != "cuda":
            raise MidasError("TCN: needs a CUDA device; there is no CPU path")
        self.index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self._h = C.c_void_p()
        self._cap = (0, 0)
        self._inv_q = None
        self.state = None
        if weights is not None:
            self.load_weights(weights)

    # ------------------------------------------------------------------ parameters
    def load_weights(self, weights):
        """path to tcn_weights.pth.tar / a state dict / {"state_dict": ...} (tcn.py:41-50)"""
        if isinstance(weights, (str, os.PathLike)):
            weights = torch.load(weights, map_location="cpu")
        if isinstance(weights, dict) and "state_dict" in weights:
            weights = weights["state_dict"]
        self.state = {k: torch.as_tensor(np.asarray(v)).detach().float().cpu().contiguous() if not torch.is_tensor(v)
                      else v.detach().float().cpu().contiguous() for k, v in weights.items()}
        self._upload()

    def _ensure(self, n: int, batch: int):
        if n <= self._cap[0] and batch <= self._cap[1]:
            return
        if self._h:
            torch.cuda.synchronize(self.index)
            call("mt_tcn_destroy", self._h)
        self._cap = (max(n, 2 * self._cap[0], 8192), max(batch, self._cap[1], 4))
        self._h = C.c_void_p()
        call("mt_tcn_create", self.index, self._cap[0], self._cap[1], C.byref(self._h))
        if self.state is not None:
            self._upload()

    def _upload(self):
        if not self._h:
            self._ensure(8192, 4)
            return
        st = self.state
        for name, cid in _CONV_IDS.items():
            if name not in st:
                if "downsample" in name:
                    continue
                raise MidasError(f"TCN: missing parameter {name}")
            w = st[name]
            w3 = w.reshape(1, *w.shape) if w.dim() == 2 else w
            call("mt_tcn_set_conv", self._h, cid, ptr(w3.contiguous()), w3.shape[0], w3.shape[1], w3.shape[2])
        for name, bid in _BN_IDS.items():
            if f"{name}.bn.weight" not in st:
                if "downsample" in name:
                    continue
                raise MidasError(f"TCN: missing parameter {name}.bn.weight")
            w, b, m, v = (st[f"{name}.bn.{k}"] for k in ("weight", "bias", "running_mean", "running_var"))
            call("mt_tcn_set_bn", self._h, bid, ptr(w), ptr(b), ptr(m), ptr(v), w.numel(), C.c_float(1e-5))
        call("mt_tcn_set_gem", self._h, C.c_float(float(st["pooling.p"].reshape(-1)[0])), C.c_float(1e-6))

    # ------------------------------------------------------------------ forward
    def embed_clouds(self, clouds: torch.Tensor) -> torch.Tensor:
        """clouds: (B, P, 3) float32 CUDA, already scaled to [-1, 1] -> (B, D) float64.
        Quantisation = ME.utils.sparse_quantize + batched_coordinates (tcn.py:124-131): floor(c / q),
        unique per cloud -- done inside the library (mt_tcn_embed): one enqueue, no sort, no host sync."""
        if self.state is None:
            raise MidasError("TCN: load_weights first")
        if not clouds.is_cuda:
            raise MidasError("TCN: clouds must be CUDA tensors; there is no CPU path")
        B, Pn, _ = clouds.shape
        if B > 511:
            raise MidasError("TCN: at most 511 clouds per call (batch_size)")
        pts = clouds.float().contiguous()
        self._ensure(B * Pn, B)
        out = torch.empty((B, self.output_dim), dtype=torch.float64, device=clouds.device)
        if self._inv_q is None:
            # `cloud / q` on a CUDA tensor is a multiplication by a float32 reciprocal of the Python scalar (torch 2.x
            # forms it in float64: float32(1 / q)); taken from torch itself, once, so the voxels are the reference's
            self._inv_q = float((torch.ones(1, dtype=torch.float32, device=clouds.device) / self.quantization_size).item())
        with torch.cuda.device(self.index):
            call("mt_tcn_embed", self._h, ptr(pts), B, Pn, C.c_float(self._inv_q), int(self.normalize_embeddings), ptr(out), 0,
                 stream_ptr())
        return out  # (a voxel coordinate outside +-131071 poisons the codes of the call with NaN)

    def cloud_to_tactile_code(self, tac_render, heightmaps, masks) -> torch.Tensor:
        """tcn.py:52-148: height maps + contact masks -> point clouds (``tac_render.heightmap2Pointcloud``)
        -> ``num_points`` samples (torch.multinomial with index-valued weights; an empty cloud becomes
        ``num_points`` zeros) -> global min-max scaling to [-1, 1] -> network -> float64 codes."""
        if type(heightmaps) is not list:
            heightmaps, masks = [heightmaps], [masks]
        n_points = self.num_points
        out = []
        for s in range(0, len(heightmaps), max(self.batch_size, 1)):
            clouds = []
            for h, c in zip(heightmaps[s:s + self.batch_size], masks[s:s + self.batch_size]):
                cloud = tac_render.heightmap2Pointcloud(h, c).to(self.device)
                if cloud.shape[0] == 0:
                    cloud = torch.zeros((n_points, 3), device=self.device)
                else:
                    idxs = torch.arange(cloud.shape[0], device=cloud.device, dtype=torch.float)
                    ids = torch.multinomial(idxs, num_samples=n_points, replacement=n_points > cloud.shape[0])
                    cloud = cloud[ids, :]
                clouds.append(2.0 * (cloud - torch.min(cloud)) / (torch.max(cloud) - torch.min(cloud)) - 1)
            out.append(self.embed_clouds(torch.stack(clouds, dim=0).float()))
        return torch.vstack(out).double()

    def __copy__(self):
        raise MidasError("TCN owns device state and cannot be shallow-copied")

    def __del__(self):
        try:
            if self._h:
                from . import _lib

                _lib.lib().mt_tcn_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

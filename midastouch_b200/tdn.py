"""Drop-in for ``midastouch/contrib/tdn_fcrn/tdn.py`` (reference lines 30-165): tactile image -> height map -> contact mask.

``TDN(cfg, bg, bottleneck, real).image2heightmap(image)`` / ``heightmap2mask(heightmap)`` / ``image2embedding(image)``
/ ``blend_heightmaps(heightmap)`` with the reference's semantics; ``tdn_weights.pth.tar`` (``{"state_dict": ...}`` of
``FCRN_net``, fcrn.py:174-272) loads unchanged.

The depth network is a dense batch-1 CNN (ResNet-50 encoder + four up-projections, Laina et al. 2016): library work
(cuDNN through ``torch.nn.functional``), not a hand-written kernel -- SURVEY.md section 8(f) rank 4.  What this file does
differently from the reference module, at load time, once:

* every BatchNorm (eval mode) is folded into the convolution in front of it (53 + 12 of them), so the forward pass
  is convolutions, ReLUs and one max-pool;
* an up-projection block (fcrn.py:62-169) evaluates eight convolutions (3x3, 2x3, 3x2, 2x2, twice) on differently
  padded copies of its input and interleaves the results with stack / permute / view.  With the paddings the
  reference uses, a 2x3 kernel is a 3x3 kernel with a zero last row, a 3x2 kernel one with a zero last column, and
  the interleave is a pixel shuffle: the block becomes ONE 3x3 convolution to 8 x C_out channels + ``pixel_shuffle(2)``
  (both branches, BatchNorms folded), then the 3x3 convolution of branch 1, the sum and the ReLU.  Same numbers
  (float32 summation order aside), a tenth of the launches.

No CUDA-extension dependency: this module is plain torch and runs wherever the tensors live.
"""
from __future__ import annotations

import collections
import os

import numpy as np
import torch
import torch.nn.functional as F

_RESNET50 = ((64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2))  # (planes, blocks, stride of the first block)
_UPS = ((1024, 512), (512, 256), (256, 128), (128, 64))
_UP_KERNELS = (("1", (3, 3)), ("2", (2, 3)), ("3", (3, 2)), ("4", (2, 2)))  # conv{b}_{1..4}: sub-pixel (0,0) (0,1) (1,0) (1,1)


def fcrn_parameter_shapes() -> "collections.OrderedDict[str, tuple]":
    """names and shapes of ``FCRN_net.state_dict()`` (fcrn.py:174-272), derived from the architecture"""
    S = collections.OrderedDict()

    def bn(name, c):
        S[f"{name}.weight"], S[f"{name}.bias"] = (c,), (c,)
        S[f"{name}.running_mean"], S[f"{name}.running_var"], S[f"{name}.num_batches_tracked"] = (c,), (c,), ()

    S["conv1.weight"] = (64, 3, 7, 7)
    bn("bn1", 64)
    inpl = 64
    for li, (planes, blocks, stride) in enumerate(_RESNET50, start=1):
        for b in range(blocks):
            p = f"layer{li}.{b}"
            S[f"{p}.conv1.weight"] = (planes, inpl, 1, 1)
            bn(f"{p}.bn1", planes)
            S[f"{p}.conv2.weight"] = (planes, planes, 3, 3)
            bn(f"{p}.bn2", planes)
            S[f"{p}.conv3.weight"] = (4 * planes, planes, 1, 1)
            bn(f"{p}.bn3", 4 * planes)
            if b == 0 and (stride != 1 or inpl != 4 * planes):
                S[f"{p}.downsample.0.weight"] = (4 * planes, inpl, 1, 1)
                bn(f"{p}.downsample.1", 4 * planes)
            inpl = 4 * planes
    S["conv2.weight"] = (1024, 2048, 1, 1)
    bn("bn2", 1024)
    for ui, (cin, cout) in enumerate(_UPS, start=1):
        for br in ("1", "2"):
            for k, (kh, kw) in _UP_KERNELS:
                S[f"up{ui}.conv{br}_{k}.weight"], S[f"up{ui}.conv{br}_{k}.bias"] = (cout, cin, kh, kw), (cout,)
        bn(f"up{ui}.bn1_1", cout)
        bn(f"up{ui}.bn1_2", cout)
        S[f"up{ui}.conv3.weight"], S[f"up{ui}.conv3.bias"] = (cout, cout, 3, 3), (cout,)
        bn(f"up{ui}.bn2", cout)
    S["conv3.weight"], S["conv3.bias"] = (1, 64, 3, 3), (1,)
    return S


def _fold(w, b, S, bn, eps=1e-5):
    """conv (w, b) followed by eval-mode BatchNorm ``bn`` -> one conv; float64 fold, float32 result"""
    g, beta = S[f"{bn}.weight"].double(), S[f"{bn}.bias"].double()
    mu, var = S[f"{bn}.running_mean"].double(), S[f"{bn}.running_var"].double()
    s = g / torch.sqrt(var + eps)
    w2 = w.double() * s.view(-1, 1, 1, 1)
    b2 = (torch.zeros_like(mu) if b is None else b.double()) * s + beta - mu * s
    return w2.float().contiguous(), b2.float().contiguous()


class FCRN:
    """functional, BatchNorm-folded form of ``FCRN_net`` (eval mode) built from its state dict"""

    def __init__(self, state: dict, device, out_size=(320, 240)):
        S = {k: v.detach().float().cpu() for k, v in state.items() if torch.is_tensor(v)}
        want = fcrn_parameter_shapes()
        missing = [k for k in want if k not in S and not k.endswith("num_batches_tracked")]
        if missing:
            raise KeyError(f"FCRN: state dict lacks {missing[:4]}{' ...' if len(missing) > 4 else ''}")
        for k, shp in want.items():
            if k in S and tuple(S[k].shape) != tuple(shp):
                raise ValueError(f"FCRN: parameter {k} has shape {tuple(S[k].shape)}, expected {shp}")
        dev = torch.device(device)
        P = {}

        def put(name, wb):
            P[name] = (wb[0].to(dev), wb[1].to(dev))

        put("stem", _fold(S["conv1.weight"], None, S, "bn1"))
        self.blocks = []
        inpl = 64
        for li, (planes, blocks, stride) in enumerate(_RESNET50, start=1):
            for b in range(blocks):
                p = f"layer{li}.{b}"
                for j in (1, 2, 3):
                    put(f"{p}.c{j}", _fold(S[f"{p}.conv{j}.weight"], None, S, f"{p}.bn{j}"))
                has_ds = f"{p}.downsample.0.weight" in S
                if has_ds:
                    put(f"{p}.ds", _fold(S[f"{p}.downsample.0.weight"], None, S, f"{p}.downsample.1"))
                self.blocks.append((p, stride if b == 0 else 1, has_ds))
                inpl = 4 * planes
        put("neck", _fold(S["conv2.weight"], None, S, "bn2"))
        for ui, (cin, cout) in enumerate(_UPS, start=1):
            # fused up-projection: channel (branch, c, sub-pixel) so that pixel_shuffle(2) interleaves the four kernels
            W = torch.zeros(2, cout, 4, cin, 3, 3, dtype=torch.float32)
            B = torch.zeros(2, cout, 4, dtype=torch.float32)
            for bi, br in enumerate(("1", "2")):
                for si, (k, (kh, kw)) in enumerate(_UP_KERNELS):
                    w, b = _fold(S[f"up{ui}.conv{br}_{k}.weight"], S[f"up{ui}.conv{br}_{k}.bias"], S, f"up{ui}.bn1_{br}")
                    W[bi, :, si, :, :kh, :kw] = w  # (the reference pads top / left only for the short side: zero last row / column)
                    B[bi, :, si] = b
            put(f"up{ui}.fused", (W.reshape(8 * cout, cin, 3, 3).contiguous(), B.reshape(8 * cout).contiguous()))
            put(f"up{ui}.c3", _fold(S[f"up{ui}.conv3.weight"], S[f"up{ui}.conv3.bias"], S, f"up{ui}.bn2"))
        put("head", (S["conv3.weight"].contiguous(), S["conv3.bias"].contiguous()))
        self.P, self.device, self.out_size = P, dev, tuple(out_size)
        self._graphs = {}  # (input shape, bottleneck, tf32) -> (CUDA graph, static input, static output)

    @torch.no_grad()
    def encode(self, x: torch.Tensor) -> torch.Tensor:
        """(B,3,H,W) float32 -> (B,1024,H/32,W/32): the bottleneck features (fcrn.py:243-260)"""
        P = self.P
        x = F.relu(F.conv2d(x, *P["stem"], stride=2, padding=3), inplace=True)
        x = F.max_pool2d(x, 3, 2, 1)
        for p, stride, has_ds in self.blocks:
            y = F.relu(F.conv2d(x, *P[f"{p}.c1"]), inplace=True)
            y = F.relu(F.conv2d(y, *P[f"{p}.c2"], stride=stride, padding=1), inplace=True)
            y = F.conv2d(y, *P[f"{p}.c3"])
            r = F.conv2d(x, *P[f"{p}.ds"], stride=stride) if has_ds else x
            x = F.relu(y.add_(r), inplace=True)
        return F.conv2d(x, *P["neck"])

    @torch.no_grad()
    def decode(self, x: torch.Tensor) -> torch.Tensor:
        """(B,1024,h,w) -> (B,1,320,240) height map (fcrn.py:262-272; dropout is the identity in eval mode)"""
        P = self.P
        for ui, (cin, cout) in enumerate(_UPS, start=1):
            y = F.pixel_shuffle(F.conv2d(x, *P[f"up{ui}.fused"], padding=1), 2)  # (B, 2*cout, 2h, 2w)
            b1 = F.conv2d(F.relu(y[:, :cout], inplace=False), *P[f"up{ui}.c3"], padding=1)
            x = F.relu(b1.add_(y[:, cout:]), inplace=True)
        x = F.relu(F.conv2d(x, *P["head"], padding=1), inplace=True)
        return F.interpolate(x, size=self.out_size, mode="bilinear", align_corners=False)

    def _forward(self, x: torch.Tensor, bottleneck: bool) -> torch.Tensor:
        z = self.encode(x)
        return z if bottleneck else self.decode(z)

    def __call__(self, x: torch.Tensor, bottleneck: bool = False) -> torch.Tensor:
        """On a CUDA device the ~170 small launches of a batch-1 forward are captured once per input shape into a CUDA
        graph and replayed (the pass is launch-bound from Python: 3.3 ms eager on a B200); MIDAS_B200_TDN_NO_GRAPH=1
        keeps it eager."""
        if not x.is_cuda or os.environ.get("MIDAS_B200_TDN_NO_GRAPH"):
            return self._forward(x, bottleneck)
        key = (tuple(x.shape), bool(bottleneck), bool(torch.backends.cudnn.allow_tf32))
        ent = self._graphs.get(key)
        with torch.cuda.device(x.device):
            if ent is None:
                static_in = x.clone()
                cur = torch.cuda.current_stream()
                side = torch.cuda.Stream()
                side.wait_stream(cur)
                with torch.cuda.stream(side):  # warm-up off the capturing stream: cuDNN picks its algorithms here
                    for _ in range(3):
                        self._forward(static_in, bottleneck)
                cur.wait_stream(side)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    static_out = self._forward(static_in, bottleneck)
                ent = self._graphs[key] = (graph, static_in, static_out)
            graph, static_in, static_out = ent
            static_in.copy_(x)
            graph.replay()
            return static_out.clone()


def normalize_minmax_255(image: np.ndarray) -> np.ndarray:
    """``cv2.normalize(image, None, alpha=0, beta=255, norm_type=cv2.NORM_MINMAX)`` (tdn.py:108): min-max over the whole
    array to [0, 255]; integer images are rounded half to even and saturated like OpenCV's ``saturate_cast``."""
    a = np.asarray(image)
    lo, hi = float(a.min()), float(a.max())
    scale = 255.0 / (hi - lo) if hi > lo else 0.0
    out = (a.astype(np.float64) - lo) * scale
    if np.issubdtype(a.dtype, np.integer):
        info = np.iinfo(a.dtype)
        return np.clip(np.rint(out), info.min, info.max).astype(a.dtype)
    return out.astype(a.dtype)


class TDN:
    def __init__(self, cfg, bg: np.ndarray = None, bottleneck: bool = False, real: bool = False, device=None, weights=None):
        fc = cfg.fcrn.real if real else cfg.fcrn.sim
        self.b, self.r, self.clip = int(fc.border), float(fc.ratio), float(fc.clip)
        self.batch_size = int(fc.batch_size)
        self.blend_sz = int(fc.blend_sz)
        self.bottleneck = bool(bottleneck)
        self.device = torch.device(device if device is not None else ("cuda:0" if torch.cuda.is_available() else "cpu"))
        self.heightmap_window = collections.deque([])
        self.model = None
        if bg is not None:
            self.bg = torch.as_tensor(np.asarray(bg), dtype=torch.float32).to(self.device)
        if weights is None:  # tdn.py:38: DIRS["weights"] / cfg.tdn_weights
            root = os.environ.get("MIDASTOUCH_WEIGHTS", "")
            cand = os.path.join(root, str(getattr(cfg, "tdn_weights", "tdn_weights.pth.tar")))
            weights = cand if os.path.isfile(cand) else None
        if weights is not None:
            self.load_weights(weights)

    def load_weights(self, weights):
        """path to tdn_weights.pth.tar / ``{"state_dict": ...}`` / a state dict (tdn.py:49-50)"""
        if isinstance(weights, (str, os.PathLike)):
            weights = torch.load(weights, map_location="cpu")
        if isinstance(weights, dict) and "state_dict" in weights:
            weights = weights["state_dict"]
        self.model = FCRN(weights, self.device)

    def _net(self) -> FCRN:
        if self.model is None:
            raise RuntimeError("TDN: no weights loaded (load_weights / weights= / $MIDASTOUCH_WEIGHTS)")
        return self.model

    # ------------------------------------------------------------------ tdn.py:60-92
    def blend_heightmaps(self, heightmap: torch.Tensor) -> torch.Tensor:
        if not self.blend_sz:
            return heightmap
        if len(self.heightmap_window) >= self.blend_sz:
            self.heightmap_window.popleft()
        self.heightmap_window.append(heightmap)
        n = len(self.heightmap_window)
        w = torch.tensor([x / n for x in range(1, n + 1)], device=heightmap.device)
        w = torch.exp(w) / torch.sum(torch.exp(w))
        stack = torch.stack(list(self.heightmap_window))
        return torch.sum((stack * w[:, None, None]) / w.sum(), dim=0)

    def _image_tensor(self, image) -> torch.Tensor:
        """(H,W,3) image -> (1,3,H,W) float32 on the device, min-max normalised to [0, 255] like ``normalize_minmax_255``
        (tdn.py:108-110) -- but evaluated on the device: the raw frame is uploaded as it is (230 KB for a uint8 DIGIT frame)
        and the normalisation is a handful of tiny kernels without a host synchronisation, instead of ~1 ms of numpy"""
        t = image if torch.is_tensor(image) else torch.from_numpy(np.ascontiguousarray(image))
        integer = not t.dtype.is_floating_point
        info = torch.iinfo(t.dtype) if integer else None
        x = t.to(self.device).to(torch.float64)
        lo, hi = x.min(), x.max()
        scale = torch.where(hi > lo, 255.0 / (hi - lo).clamp_min(1e-300), torch.zeros_like(hi))
        x = (x - lo) * scale
        if integer:  # OpenCV's saturate_cast: round half to even, clamp to the type's range
            x = torch.round(x).clamp_(info.min, info.max)
        elif t.dtype != torch.float64:
            x = x.to(t.dtype)
        return x.permute(2, 0, 1).float()[None].contiguous()

    # ------------------------------------------------------------------ tdn.py:94-115
    def image2heightmap(self, image: np.ndarray) -> torch.Tensor:
        assert self.bottleneck is False, "Bottleneck feature is enabled, can't carry out image2heightmap"
        out = self._net()(self._image_tensor(image))[0].squeeze()
        return self.blend_heightmaps(out)

    # ------------------------------------------------------------------ tdn.py:117-137
    def image2embedding(self, image: np.ndarray) -> torch.Tensor:
        if self.bottleneck is False:
            print("Bottleneck feature extraction not enabled, switching")
            self.bottleneck = True
        out = self._net()(self._image_tensor(image), bottleneck=True)[0].squeeze()
        feature = out.reshape((-1, 10 * 8 * 1024))
        return feature / torch.norm(feature, dim=1).reshape(-1, 1)

    # ------------------------------------------------------------------ tdn.py:139-165
    def heightmap2mask(self, heightmap: torch.Tensor, small_parts: bool = False) -> torch.Tensor:
        b = self.b
        hm = heightmap[b:-b, b:-b]
        diff = hm - self.bg[b:-b, b:-b]
        diff = torch.where(diff < self.clip, torch.zeros_like(diff), diff)  # (the reference zeroes a temporary in place)
        contact = diff > torch.quantile(diff, 0.8) * self.r
        padded = torch.zeros_like(self.bg, dtype=torch.bool)
        total = contact.shape[0] * contact.shape[1]
        atleast = 0.01 * total if small_parts else 0.1 * total
        if torch.count_nonzero(contact) < atleast:
            return padded
        padded[b:-b, b:-b] = contact
        return padded

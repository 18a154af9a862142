"""mt_ctx lifetime + small helpers shared by the drop-in classes and the engine."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import MidasError, call, ptr, stream_ptr

F32, F64 = 0, 1


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.float64:
        return F64
    raise MidasError(f"unsupported dtype {t.dtype}; libmidas_b200 takes float32 or float64")


def require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise MidasError(f"{what} must be a CUDA tensor: midastouch_b200 has no CPU path")


class Context:
    """owns one mt_ctx (scratch + optional codebook tables) on one GPU."""

    def __init__(self, device, capacity: int, M: int = 1, D: int = 4):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise MidasError("midastouch_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.capacity, self.M, self.D = int(capacity), int(M), int(D)
        self._h = C.c_void_p()
        call("mt_ctx_create", self.index, C.c_size_t(self.capacity), self.M, self.D, C.byref(self._h))
        self.generation = 0    # bumped whenever the mt_ctx is re-created (holders of addresses inside it must refresh)
        self._codebook = None  # (keys_host, emb) kept alive / for re-upload on growth
        self._mesh = None      # (vertices_host float64, cell)

    @property
    def h(self):
        return self._h

    def upload_codebook(self, keys_host: torch.Tensor, emb: torch.Tensor):
        assert keys_host.device.type == "cpu" and keys_host.dtype == torch.float32 and keys_host.shape == (self.M, 6)
        require_cuda(emb, "codebook embeddings")
        keys_host = keys_host.contiguous()
        emb = emb.contiguous()
        with torch.cuda.device(self.index):
            call("mt_codebook_upload", self._h, ptr(keys_host), ptr(emb), dtype_code(emb))
        self._codebook = (keys_host, emb)

    def ensure_capacity(self, n: int):
        if n <= self.capacity:
            return
        cb = self._codebook
        torch.cuda.synchronize(self.index)
        call("mt_ctx_destroy", self._h)
        self.capacity = int(max(n, 2 * self.capacity))
        self._h = C.c_void_p()
        call("mt_ctx_create", self.index, C.c_size_t(self.capacity), self.M, self.D, C.byref(self._h))
        self.generation += 1
        if cb is not None:
            self.upload_codebook(*cb)
        if self._mesh is not None:
            self.upload_mesh(*self._mesh)

    def upload_mesh(self, vertices, cell: float):
        """down-sampled mesh vertices (V,3) float64 host array for the drift test."""
        import numpy as np

        v = np.ascontiguousarray(np.asarray(vertices, dtype=np.float64).reshape(-1, 3))
        with torch.cuda.device(self.index):
            call("mt_mesh_upload", self._h, v.ctypes.data_as(C.c_void_p), v.shape[0], float(cell))
        self._mesh = (v, float(cell))

    def stats(self, reset: bool = False) -> dict:
        out = (C.c_longlong * 16)()
        with torch.cuda.device(self.index):
            call("mt_ctx_stats", self._h, out, int(reset))
        return {"overflow": out[0], "resample_skipped": out[1], "invalid_poses": out[2], "nn_fallbacks": out[3],
                "drifted": out[5], "on_surface": out[6], "grid_rows": out[4], "grid_rows_max": out[7], "mesh_deferred": out[8], "scan_deferred": out[9]}

    def grid_info(self):
        h = C.c_float()
        dims = (C.c_int * 3)()
        occ = C.c_int()
        call("mt_codebook_grid_info", self._h, C.byref(h), dims, C.byref(occ))
        return h.value, tuple(dims), occ.value

    def __del__(self):
        try:
            if self._h:
                _lib.lib().mt_ctx_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass


def aos_to_soa(poses: torch.Tensor, stride: int | None = None) -> torch.Tensor:
    """(N,4,4) float32 cuda -> (3, stride, 4) float32 SoA rows."""
    require_cuda(poses, "poses")
    poses = poses.reshape(-1, 4, 4).float().contiguous()
    n = poses.shape[0]
    stride = n if stride is None else stride
    soa = torch.empty((3, stride, 4), dtype=torch.float32, device=poses.device)
    with torch.cuda.device(poses.device):
        call("mt_aos_to_soa", ptr(poses), n, ptr(soa), stride, stream_ptr())
    return soa


def soa_to_aos(soa: torch.Tensor, n: int) -> torch.Tensor:
    out = torch.empty((n, 4, 4), dtype=torch.float32, device=soa.device)
    with torch.cuda.device(soa.device):
        call("mt_soa_to_aos", ptr(soa), soa.shape[1], n, ptr(out), stream_ptr())
    return out

"""Drop-in for ``midastouch/tactile_tree/tactile_tree.py`` (reference lines 13-77).

Same class name, constructor, attributes and methods; PyTorch tensors in and out.
The nanoflann k-d tree (tactile_tree.py:34-41) is replaced by a per-key neighbour graph
and a 6-D bounding-box hierarchy held inside the CUDA library, searched exactly
(``mt_nn_assign``).  Extra, additive API for the fast path: ``SE3_NN_idx`` (indices
only) and ``query`` (cos(q, E_m) for every row in one pass).
"""
from __future__ import annotations

import torch

from ._lib import MidasError, call, ptr, stream_ptr
from .context import Context, aos_to_soa, dtype_code, require_cuda


def R3_SE3(poses: torch.Tensor, w: float = 0.01) -> torch.Tensor:
    """tactile_tree.py:73-77 on the GPU: (N,4,4) -> (N,6) float32 keys."""
    require_cuda(poses, "poses")
    poses = poses.reshape(-1, 4, 4)
    n = poses.shape[0]
    soa = aos_to_soa(poses)
    keys = torch.empty((n, 6), dtype=torch.float32, device=poses.device)
    with torch.cuda.device(poses.device):
        call("mt_se3_keys_w", ptr(soa), n, n, float(w), ptr(keys), stream_ptr())
    return keys


class tactile_tree(torch.nn.Module):
    def __init__(self, poses, cam_poses, embeddings):
        super().__init__()
        self.poses = poses.float()
        self.cam_poses, self.embeddings = cam_poses.float(), embeddings
        self.logmap_pose = None
        self.tree, self.tree_size = None, 0
        self.ctx = None
        if self.poses.is_cuda:
            self.init_tree()
        else:
            self.tree_size = self.poses.shape[0]

    def __len__(self):
        return self.tree_size

    def __repr__(self):
        return "tactile Tree of size: {}".format(self.__len__())

    def to_device(self, device):
        self.poses = self.poses.to(device)
        self.cam_poses = self.cam_poses.to(device)
        self.embeddings = self.embeddings.to(device)
        self.init_tree()

    def init_tree(self, capacity: int = 65536):
        require_cuda(self.poses, "codebook poses")
        self.poses = self.poses.contiguous()
        self.cam_poses = self.cam_poses.contiguous()
        self.embeddings = self.embeddings.contiguous()
        self.logmap_pose = R3_SE3(self.poses)
        M, D = self.embeddings.shape
        self.ctx = Context(self.poses.device, capacity, M, D)
        self.ctx.upload_codebook(self.logmap_pose.cpu(), self.embeddings)
        self.tree = self.ctx
        self.tree_size = M

    # ---- additive fast-path API
    def SE3_NN_idx(self, query: torch.Tensor, hint: torch.Tensor | None = None, exhaustive: bool = False) -> torch.Tensor:
        """indices of the exact nearest codebook pose, int32 (N,)."""
        keys = R3_SE3(query.reshape(-1, 4, 4))
        n = keys.shape[0]
        idx = torch.empty(n, dtype=torch.int32, device=keys.device)
        with torch.cuda.device(keys.device):
            call("mt_nn_assign", self.ctx.h, ptr(keys), n, ptr(hint), 1 if exhaustive else 0, ptr(idx), stream_ptr())
        return idx

    def query(self, code: torch.Tensor) -> torch.Tensor:
        """cos(code, E_m) for all M rows, float64 (M,) -- one pass over the codebook."""
        code = code.reshape(-1).to(self.embeddings.device).contiguous()
        out = torch.empty(self.tree_size, dtype=torch.float64, device=self.embeddings.device)
        with torch.cuda.device(out.device):
            call("mt_codebook_query", self.ctx.h, ptr(code), dtype_code(code), ptr(out), stream_ptr())
        return out

    def query_batched(self, codes: torch.Tensor) -> torch.Tensor:
        """cos(codes_q, E_m) for a batch of codes: (nq, D) -> (nq, M) float32, on the tensor cores
        (tcgen05 / TMEM, 3xTF32; the eval script's M x M retrieval, single_touch_test.py:35-73)."""
        require_cuda(codes, "codes")
        q = torch.atleast_2d(codes).to(torch.float32).contiguous()
        out = torch.empty((q.shape[0], self.tree_size), dtype=torch.float32, device=q.device)
        with torch.cuda.device(q.device):
            call("mt_codebook_query_batched", self.ctx.h, ptr(q), q.shape[0], ptr(out), stream_ptr())
        return out

    def _gather(self, table: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        n = idx.shape[0]
        row = table[0].numel() * (2 if table.dtype == torch.float64 else 1)
        out = torch.empty((n,) + tuple(table.shape[1:]), dtype=table.dtype, device=table.device)
        with torch.cuda.device(table.device):
            call("mt_gather_rows_f32", ptr(table), ptr(idx), n, row, ptr(out), stream_ptr())
        return out

    # ---- reference API
    def SE3_NN(self, _query, nn=1):
        """tactile_tree.py:43-58: best SE(3) match of every query pose; returns the gathered
        (poses, cam_poses, embeddings) like the reference (squeezed for a single query)."""
        query = _query.reshape(-1, 4, 4)
        if nn != 1:  # kneighbors(n_neighbors=nn): (N, nn) indices, nearest first (exhaustive search in the library)
            if not 1 <= int(nn) <= min(64, self.tree_size):
                raise MidasError("SE3_NN: nn must be between 1 and min(64, codebook size)")
            keys = R3_SE3(query)
            n = keys.shape[0]
            idx = torch.empty((n, int(nn)), dtype=torch.int32, device=keys.device)
            with torch.cuda.device(keys.device):
                call("mt_nn_topk", self.ctx.h, ptr(keys), n, int(nn), ptr(idx), stream_ptr())
            ii = idx.long().squeeze()  # the reference's indices_p.squeeze()
            return self.poses[ii], self.cam_poses[ii], self.embeddings[ii]
        idx = self.SE3_NN_idx(query)
        p, c, e = self._gather(self.poses, idx), self._gather(self.cam_poses, idx), self._gather(self.embeddings, idx)
        if idx.shape[0] == 1:  # the reference's indices_p.squeeze()
            return p[0], c[0], e[0]
        return p, c, e

    def get_poses(self):
        return self.poses, self.cam_poses

    def get_pose(self, idx):
        return self.poses[idx, :]

    def get_embeddings(self):
        return self.embeddings

    def get_embedding(self, idx):
        return self.embeddings[idx, :]

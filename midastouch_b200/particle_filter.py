"""Drop-in for ``midastouch/modules/particle_filter.py`` (reference lines 33-496).

Same names (``Particles``, ``particle_filter``, ``particle_rmse``, ``torch_delete``),
argument meaning and error behaviour; PyTorch CUDA tensors in and out; the arithmetic runs
in libmidas_b200 (sm_100a).  Differences, all additive or documented in DESIGN.md:
  * ``resampler`` default stays "weighted_random" like the reference (``torch.multinomial`` there;
    here N categorical draws in the library, Philox-keyed -- same distribution, not the same random
    stream); "low_var" is the CUDA systematic resampler; "low_var_batch" is served by the same kernel
    (its N x N formulation has an off-by-one at particle_filter.py:279 that is not reproduced).
  * the mesh is given as a vertex array (or .npy path): trimesh is not a dependency.
"""
from __future__ import annotations

import copy
import ctypes as C
from typing import Tuple

import numpy as np
import torch

from ._lib import MidasError, call, ptr, stream_ptr
from .context import Context, aos_to_soa, dtype_code, require_cuda, soa_to_aos


class Particles:
    """particle_filter.py:33-79: [poses (N,4,4), weights (N,), labels (N,)]"""

    poses = None
    weights = None
    labels = None

    def __init__(self, poses: torch.Tensor, weights: torch.Tensor = None, labels: torch.Tensor = None):
        self.poses = poses
        self.weights = weights if weights is not None else torch.ones(self.poses.shape[0], device=poses.device)
        self.labels = labels if labels is not None else torch.zeros(self.poses.shape[0], device=poses.device)

    def __len__(self):
        return self.poses.shape[0]

    def remove(self, idxs: torch.Tensor) -> None:
        self.poses = torch_delete(self.poses, idxs, dim=0)
        self.weights = torch_delete(self.weights, idxs)
        self.labels = torch_delete(self.labels, idxs)

    def add(self, poses: torch.Tensor, weights: torch.Tensor, labels: torch.Tensor) -> None:
        self.poses = torch.cat((self.poses, poses), dim=0)
        self.weights = torch.cat((self.weights, weights))
        self.labels = torch.cat((self.labels, labels))


def torch_delete(arr: torch.Tensor, idxs: torch.Tensor, dim: int = 0) -> torch.Tensor:
    """np.delete equivalent (particle_filter.py:81-90) without the N x k boolean matrix:
    a keep-mask + stream compaction."""
    if idxs.nelement():
        keep = torch.ones(arr.size(dim), dtype=torch.bool, device=arr.device)
        keep[idxs.reshape(-1).to(arr.device)] = False
        return arr[keep]
    return arr


_CTX = {}


def _ctx_for(device, n: int) -> Context:
    device = torch.device(device)
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    c = _CTX.get(key)
    if c is None:
        c = _CTX[key] = Context(device, max(n, 65536))
    c.ensure_capacity(n)
    return c


def _load_vertices(mesh):
    if isinstance(mesh, str):
        return np.load(mesh)
    if hasattr(mesh, "vertices"):
        return np.asarray(mesh.vertices)
    return np.asarray(mesh)


class particle_filter:
    """particle filter class for update and propagation of SE(3) Particles on mesh"""

    def __init__(self, cfg, mesh_path, noise: float = 1.0, real: bool = False, downsample: int = 10):
        self.pen_max = cfg.tdn.render.pen.max
        verts = _load_vertices(mesh_path)
        self.mesh_vertices = np.asarray(verts, dtype=np.float64)
        self._scale = float(np.linalg.norm(self.mesh_vertices.max(0) - self.mesh_vertices.min(0)))
        self.mesh_vertices_ds = self.mesh_vertices[::downsample, :]
        nr, nt = cfg.expt.params.noise_r, cfg.expt.params.noise_t
        which = "real" if real else "sim"
        # mcmaster.yaml:19-20 gives scalars where particle_filter.py:114-121 reads .sim/.real
        sig_r = nr if isinstance(nr, (int, float)) else nr[which]
        sig_t = nt if isinstance(nt, (int, float)) else nt[which]
        self.motion_noise = {"mu": 0, "sig_r": float(sig_r), "sig_t": float(sig_t)}
        self.particle_var = torch.tensor([float("inf")])
        self.init_noise = [self.mesh_diagonal() / 3.0 * noise, 180.0 / 3.0 * noise]
        self._mesh_ctx = {}  # device index -> Context holding the vertex grid of THIS mesh

    def mesh_diagonal(self):
        return self._scale

    # ------------------------------------------------------------------ init (129-145)
    def init_filter(self, gt_pose: torch.Tensor = torch.eye(4), N: int = 10000) -> Particles:
        """gt @ [Rzyx(N(0,60 deg)) | N(0, scale/3)]: noise from the CPU generator exactly like the
        reference (two (N,3) float32 normals), composed on the GPU by the motion kernel with
        identity odometry."""
        require_cuda(gt_pose, "gt_pose")
        tn = torch.normal(mean=0.0, std=self.init_noise[0], size=(N, 3))
        rot = torch.normal(mean=0.0, std=self.init_noise[1], size=(N, 3))
        dev = gt_pose.device
        base = gt_pose.float()[None].expand(N, 4, 4).contiguous()
        soa = aos_to_soa(base)
        eye = torch.eye(4, dtype=torch.float32).contiguous()
        tn_d, rot_d = tn.to(dev).contiguous(), rot.to(dev).contiguous()
        with torch.cuda.device(dev):
            call("mt_motion", ptr(soa), ptr(soa), N, N, ptr(eye), ptr(tn_d), ptr(rot_d), 0.0, 0.0, 0, 0, 0, 0, 1, stream_ptr())
        return Particles(soa_to_aos(soa, N))

    # ------------------------------------------------------------------ motion (319-377)
    def motionModel(self, _particles: Particles, odom: torch.Tensor, multiplier: float = 1.0) -> Particles:
        if multiplier < 1.0:
            multiplier = 1.0
        particles = copy.copy(_particles)
        poses = particles.poses
        require_cuda(poses, "particle poses")
        N = poses.shape[0]
        # RNG contract of add_noise_to_odom (326-335): CPU default generator, translation first
        tn = torch.normal(mean=self.motion_noise["mu"], std=float(multiplier) * self.motion_noise["sig_t"], size=(N, 3))
        rot = torch.normal(mean=self.motion_noise["mu"], std=float(multiplier) * self.motion_noise["sig_r"], size=(N, 3))
        dev = poses.device
        soa = aos_to_soa(poses)
        odom_h = odom.detach().float().cpu().contiguous()
        tn_d, rot_d = tn.to(dev).contiguous(), rot.to(dev).contiguous()
        invalid = torch.zeros(1, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            call("mt_motion", ptr(soa), ptr(soa), N, N, ptr(odom_h), ptr(tn_d), ptr(rot_d), 0.0, 0.0, 0, 0, 0, ptr(invalid), 0, stream_ptr())
        particles.poses = soa_to_aos(soa, N)
        if int(invalid.item()):  # check_quats (347-357): rare path, prune non-finite poses
            bad = (~torch.isfinite(particles.poses.reshape(N, -1)).all(dim=1)).nonzero()
            particles.remove(bad)
        return particles

    # ------------------------------------------------------------------ measurement (449-469)
    def get_similarity(self, queries: torch.Tensor, targets: torch.Tensor, softmax=True) -> torch.Tensor:
        targets = torch.atleast_2d(targets)
        require_cuda(targets, "targets")
        targets = targets.contiguous()
        q = torch.atleast_2d(queries).to(targets.device)
        if q.shape[0] != 1:
            raise MidasError("get_similarity: one query row against N targets (every reference call site)")
        q = q.reshape(-1).contiguous()
        n, D = targets.shape
        ctx = _ctx_for(targets.device, n)
        w = torch.empty(n, dtype=torch.float64, device=targets.device)
        with torch.cuda.device(targets.device):
            call("mt_cosine_rows", ctx.h, ptr(q), dtype_code(q), ptr(targets), dtype_code(targets), n, D, ptr(w), stream_ptr())
            if softmax:
                call("mt_softmax_f64", ctx.h, ptr(w), n, ptr(w), stream_ptr())
        if targets.dtype != torch.float64:  # torch.cosine_similarity returns the operands' dtype
            w = w.to(targets.dtype)
        return w.squeeze()

    # ------------------------------------------------------------------ resampling (230-307)
    def resampler(self, _particles: Particles, resample: str = "weighted_random", u: float = None) -> Particles:
        particles = copy.copy(_particles)
        nSamples = len(particles)
        require_cuda(particles.poses, "particle poses")
        if resample == "weighted_random":
            # WeightedRandomSampler(weights, N, replacement=True) == N categorical draws (243-250); drawn in the
            # library (Philox4x32-10 keyed by a seed taken from torch's default generator, so torch.manual_seed
            # still makes a run reproducible)
            dev = particles.poses.device
            w = particles.weights.to(torch.float64).contiguous()
            seed = int(torch.randint(0, 2**62, (1,)).item())
            ctx = _ctx_for(dev, nSamples)
            idx32 = torch.empty(nSamples, dtype=torch.int32, device=dev)
            scratch = torch.empty(nSamples, dtype=torch.float64, device=dev)
            status = torch.zeros(1, dtype=torch.int32, device=dev)
            with torch.cuda.device(dev):
                call("mt_resample_multinomial", ctx.h, ptr(w), nSamples, nSamples, seed, 0, ptr(scratch), ptr(idx32), ptr(status),
                     stream_ptr())
            if int(status.item()):
                return particles  # all-zero / NaN weights: the reference returns the input (237-241)
            idxs = idx32.long()
            return Particles(particles.poses[idxs, :, :], particles.weights[idxs], particles.labels[idxs])
        if resample not in ("low_var", "low_var_batch"):
            raise MidasError(f"unknown resample mode {resample!r}")
        dev = particles.poses.device
        w = particles.weights.to(torch.float64).contiguous()
        if u is None:  # the reference draws torch.rand(1) on the weights' device (260)
            u = float(torch.rand(1, device=dev).item())
        ctx = _ctx_for(dev, nSamples)
        anc = torch.empty(nSamples, dtype=torch.int32, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            call("mt_resample_systematic", ctx.h, ptr(w), nSamples, C.c_float(u), 0, ptr(anc), ptr(status), stream_ptr())
        if int(status.item()):
            return particles  # all-zero / NaN weights: the reference returns the input (237-241)
        soa_in = aos_to_soa(particles.poses)
        soa_out = torch.empty_like(soa_in)
        w_in = particles.weights.contiguous()
        with torch.cuda.device(dev):
            call("mt_gather_soa", ptr(soa_in), nSamples, ptr(anc), nSamples, ptr(soa_out), nSamples, stream_ptr())
        idx = anc.long()
        return Particles(soa_to_aos(soa_out, nSamples), w_in[idx], particles.labels[idx])

    # ------------------------------------------------------------------ prune / anneal / clusters ("next" rows)
    def _mesh_context(self, device) -> Context:
        device = torch.device(device)
        idx = device.index if device.index is not None else torch.cuda.current_device()
        c = self._mesh_ctx.get(idx)
        if c is None:
            c = self._mesh_ctx[idx] = Context(device, 1024)
            c.upload_mesh(self.mesh_vertices_ds, self.pen_max)
        return c

    def remove_invalid_particles(self, _particles: Particles, invalid_dist: float = None) -> Tuple[Particles, bool]:
        """particle_filter.py:379-403: weights *= (nearest down-sampled mesh vertex within
        invalid_dist); returns (particles, drifted) with drifted a 0-dim bool CUDA tensor.
        Like the reference the multiply is in place on the (shallow-copied) weights tensor."""
        particles = copy.copy(_particles)
        poses = particles.poses
        require_cuda(poses, "particle poses")
        poses = poses.reshape(-1, 4, 4).float().contiguous()
        n = poses.shape[0]
        dist = self.pen_max if invalid_dist is None else float(invalid_dist)
        ctx = self._mesh_context(poses.device)
        w = particles.weights
        w64 = w if (w.dtype == torch.float64 and w.is_contiguous()) else w.to(torch.float64).contiguous()
        nvalid = torch.zeros(1, dtype=torch.int32, device=poses.device)
        with torch.cuda.device(poses.device):
            call("mt_prune_aos", ctx.h, ptr(poses), n, C.c_double(dist), ptr(w64), ptr(nvalid), stream_ptr())
        if w64 is not w:  # float32 default weights (Particles.__init__): keep the caller's dtype
            w.copy_(w64.to(w.dtype))
        return particles, (nvalid == 0).reshape(())

    def annealing(self, _particles: Particles, var: float, floor: int = 1000) -> Particles:
        """particle_filter.py:405-447: shrink / grow the particle set with the cluster variance.  The
        k lowest (highest) weights are found by the CUDA radix select (``mt_select_k``) instead of
        ``torch.topk``; survivors keep their order, duplicates are appended (in index order)."""
        particles = copy.copy(_particles)
        var = var if torch.is_tensor(var) else torch.tensor(float(var))
        if torch.isinf(self.particle_var).any():
            self.particle_var = var
            self.init_particles = len(particles.weights)
            return particles
        if float(var) == 0.0:
            return particles
        ratio = float(var) / float(self.particle_var)
        self.particle_var = var
        n_particles = len(particles.weights)
        N = particles.poses.shape[0]
        if ratio < 1:
            num = min(int((1.0 - ratio) * N), abs(n_particles - floor), n_particles // 3)
            if not num:
                return particles
            keep = self._select(particles.weights, num, largest=False)[1]
            particles.poses, particles.weights, particles.labels = particles.poses[keep], particles.weights[keep], particles.labels[keep]
        elif ratio > 1:
            num = min(int((ratio - 1.0) * N), n_particles // 3)
            if num + n_particles > self.init_particles or not num:
                return particles
            add = self._select(particles.weights, num, largest=True)[0]
            particles.add(particles.poses[add, :], particles.weights[add], particles.labels[add])
        return particles

    def _select(self, weights: torch.Tensor, k: int, largest: bool):
        require_cuda(weights, "particle weights")
        w = weights.to(torch.float64).contiguous()
        n = w.shape[0]
        ctx = _ctx_for(w.device, n)
        sel = torch.empty(k, dtype=torch.int32, device=w.device)
        keep = torch.empty(n - k, dtype=torch.int32, device=w.device)
        with torch.cuda.device(w.device):
            call("mt_select_k", ctx.h, ptr(w), n, k, int(largest), ptr(sel), ptr(keep), stream_ptr())
        return sel.long(), keep.long()

    def get_cluster_centers(self, _particles: Particles, method: str = "logmap"):
        """particle_filter.py:153-206 -> (cluster_poses (K,4,4), cluster_stds (K,3)) float32.  method "quat_avg"
        (what the loop passes, filter.py:184-186): Markley quaternion average; "logmap" (the reference's default):
        weighted mean of the SE(3) tangents, exponentiated (log_map_averaged, pose.py:101-109)."""
        if method not in ("quat_avg", "logmap"):
            raise MidasError(f"get_cluster_centers: unknown method {method!r}")
        particles = copy.copy(_particles)
        poses = particles.poses.reshape(-1, 4, 4)
        require_cuda(poses, "particle poses")
        poses = poses.float().contiguous()
        n = poses.shape[0]
        uniq, inv = torch.unique(particles.labels, return_inverse=True)  # the reference's torch.unique (164)
        K = int(uniq.shape[0])
        ctx = _ctx_for(poses.device, n)
        w = particles.weights.to(torch.float64).contiguous()
        centers = torch.empty((K, 4, 4), dtype=torch.float32, device=poses.device)
        stds = torch.empty((K, 3), dtype=torch.float32, device=poses.device)
        with torch.cuda.device(poses.device):
            for k0 in range(0, K, 16):  # the kernels reduce up to 16 clusters per pass
                kk = min(16, K - k0)
                lab = (inv - k0).to(torch.int32)
                lab = torch.where((lab >= 0) & (lab < kk), lab, torch.full_like(lab, -1)).contiguous()
                call("mt_cluster_centers", ctx.h, ptr(poses), ptr(w), ptr(lab), n, kk, int(method == "logmap"),
                     ptr(centers[k0:k0 + kk]), ptr(stds[k0:k0 + kk]), stream_ptr())
        return centers, stds

    def cluster_particles(self, _particles: Particles, method: str = "euclidean", eps: float = 1e-2) -> Particles:
        """particle_filter.py:208-228: DBSCAN(eps, min_samples = N / 5) on the particle translations, sklearn's
        semantics (labels int64, noise -1), run in the library (``mt_dbscan``).  method "logmap" (DBSCAN on the
        theseus SE(3) tangents; no call site in the reference uses it) is not implemented."""
        if method != "euclidean":
            raise MidasError("cluster_particles: only method='euclidean' (the value every reference call site uses) is implemented")
        particles = copy.copy(_particles)
        poses = particles.poses.reshape(-1, 4, 4)
        require_cuda(poses, "particle poses")
        poses = poses.float().contiguous()
        n = poses.shape[0]
        min_samples = max(int(n / 5), 1)
        ctx = _ctx_for(poses.device, n)
        labels = torch.empty(n, dtype=torch.int64, device=poses.device)
        with torch.cuda.device(poses.device):
            call("mt_dbscan", ctx.h, ptr(poses), n, C.c_double(float(eps)), min_samples, ptr(labels), None, stream_ptr())
        particles.labels = labels
        return particles


def particle_rmse(_particles, gt_pose: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """particle_filter.py:472-496 -> (rmse_t, rmse_r) 0-dim float32 CUDA tensors."""
    poses = _particles.poses if isinstance(_particles, Particles) else _particles
    poses = poses.reshape(-1, 4, 4)
    require_cuda(poses, "particle poses")
    n = poses.shape[0]
    ctx = _ctx_for(poses.device, n)
    soa = aos_to_soa(poses)
    gt_h = gt_pose.detach().float().cpu().contiguous()
    out = torch.empty(2, dtype=torch.float32, device=poses.device)
    with torch.cuda.device(poses.device):
        call("mt_rmse", ctx.h, ptr(soa), n, n, ptr(gt_h), ptr(out), stream_ptr())
    return out[0], out[1]

"""FilterEngine: the resident, fused form of the loop body ``filter.py:152-190``.

Particles live on the GPU as SoA 3 x float4 rows (48 B / particle) in two ping-pong
buffers; one ``step()`` is
    codebook query   sim[M] = cos(q, E_m)                      (get_similarity, 449-469)
    kernel A         motion + R3_SE3 key + exact 1-NN + weight lookup + chunk sums
                                                            (motionModel, SE3_NN, 170-173)
    [sharded only]   all-gather of the per-GPU weight sums     (one 8-byte value per GPU)
    kernel B         normalise + float64 prefix + systematic draw + child scatter
                                                            (resampler "low_var", 251-307)
with no host synchronisation.  Sharding: each GPU owns a contiguous block of particles and
the children of its own parents, so poses never cross NVLink; only the weight sums do.
"""
from __future__ import annotations

import ctypes as C

import torch

from ._lib import MidasError, StepArgs, call, lib, ptr, stream_ptr
from .context import aos_to_soa, dtype_code, require_cuda, soa_to_aos  # noqa: F401
from .tactile_tree import tactile_tree


Odom16 = C.c_float * 16


def prepare_odom(odom) -> "Odom16":
    """(4,4) tensor / 16 floats -> the by-value kernel argument (do this once per frame)."""
    vals = odom.reshape(-1).tolist() if isinstance(odom, torch.Tensor) else list(odom)
    return Odom16(*[float(x) for x in vals])


def rebalance_plan(counts, rank: int):
    """Particles are globally ordered rank-major.  Given the current per-rank counts, return
    (send_counts, recv_counts, targets) that move the shard boundaries to an even split while keeping
    the global order: rank r sends to rank s the overlap of its current range with s's target range."""
    G, total = len(counts), int(sum(counts))
    targets = [total // G + (1 if r < total % G else 0) for r in range(G)]
    cur_off = [sum(counts[:r]) for r in range(G)]
    tgt_off = [sum(targets[:r]) for r in range(G)]

    def overlap(a0, a1, b0, b1):
        return max(0, min(a1, b1) - max(a0, b0))

    send = [overlap(cur_off[rank], cur_off[rank] + counts[rank], tgt_off[s], tgt_off[s] + targets[s]) for s in range(G)]
    recv = [overlap(cur_off[s], cur_off[s] + counts[s], tgt_off[rank], tgt_off[rank] + targets[rank]) for s in range(G)]
    return send, recv, targets


class FilterEngine:
    def __init__(self, codebook: tactile_tree, capacity: int, sig_t: float = 2e-4, sig_r: float = 0.5,
                 seed: int = 0, rank: int = 0, world: int = 1, group=None, n_global: int | None = None,
                 mesh_vertices=None, pen_max: float = 0.002):
        if codebook.ctx is None:
            raise MidasError("FilterEngine: codebook.to_device(cuda) first")
        self.cb = codebook
        self.dev = codebook.poses.device
        self.capacity = int(capacity)
        codebook.ctx.ensure_capacity(self.capacity)
        self.ctx = codebook.ctx
        self._ctx_gen = self.ctx.generation  # the engine caches addresses inside the context (see _refresh_ctx)
        self.sig_t, self.sig_r, self.seed = float(sig_t), float(sig_r), int(seed)
        # drift pruning (remove_invalid_particles): needs the down-sampled mesh vertices
        self.pen_max = float(pen_max)
        self.prune = mesh_vertices is not None
        if self.prune:
            self.ctx.upload_mesh(mesh_vertices, self.pen_max)
        self.rank, self.world, self.group = int(rank), int(world), group
        self.n_global = n_global
        d = self.dev
        self.soa = [torch.zeros((3, self.capacity, 4), dtype=torch.float32, device=d) for _ in range(2)]
        self.nn = [torch.full((self.capacity,), -1, dtype=torch.int32, device=d) for _ in range(2)]
        self.anc = torch.zeros(self.capacity, dtype=torch.int32, device=d)
        # step outputs (rmse) and the tactile code are double-buffered by the parity of the particle buffers, so that the
        # copies of step t +- 1 (host -> device code, device -> host rmse) can run on a copy stream beside step t's kernels
        self._rm = [torch.zeros(2, dtype=torch.float32, device=d) for _ in range(2)]
        self._rm_last = 0
        self.n_dev = [torch.zeros(1, dtype=torch.int64, device=d) for _ in range(2)]
        self.shard_sums = torch.zeros(max(self.world, 1), dtype=torch.float64, device=d)
        self._q = [torch.zeros(codebook.embeddings.shape[1], dtype=torch.float64, device=d) for _ in range(2)]
        self._io = torch.cuda.Stream(device=d)  # copy stream
        self._ev_q = [torch.cuda.Event() for _ in range(2)]     # code of parity p has landed
        self._ev_step = [torch.cuda.Event() for _ in range(2)]  # the last step that used the buffers of parity p is done
        self._ev_read = torch.cuda.Event()
        self._ev_rd = [None, None]  # pending read_rmse_async of the result buffer of parity p
        self.cur = 0
        self._n_children = 0  # single GPU: children to draw in the next resampling (0 = as many as there are particles)
        self.n = 0  # host-side upper bound of the local particle count
        self.t = 0
        self.use_n_dev = False
        p = C.c_void_p()
        call("mt_step_local_sum_ptr", self.ctx.h, C.byref(p))
        self._local_sum_ptr = p.value
        self._a = StepArgs()
        # one library call per step; the step's kernels are replayed from a CUDA graph (query | motion + SE3_NN ->
        # queue consumers -> cooperative resampling kernel) whenever the step runs in its fused form
        self.use_graph = True
        self.overlap_query = True
        self.rebalance_every = 64  # sharded runs: even the shards out every this many steps (0 = never)
        self.fuse_sums = True  # single GPU: weight sums + resampling as one cooperative kernel
        self._side = torch.cuda.Stream(device=self.dev)
        self._ev_table = torch.cuda.Event()
        self._ev_free = torch.cuda.Event()
        self._rng = torch.Generator().manual_seed(self.seed)
        # sharded: map the peers' exchange buffers so that the weight sums travel inside the fused kernel
        self.peer_exchange = False
        if self.world > 1:
            self.connect_peers()

    def _refresh_ctx(self):
        """The shared context of a codebook is re-created when a later user needs more capacity
        (Context.ensure_capacity): addresses cached from the old one (local weight sum, peer mappings) are stale.
        Single GPU: re-query them.  Sharded: the peers hold mappings of the old exchange buffer, which only a collective
        connect_peers() on every rank can renew -- raise rather than read freed memory."""
        if self._ctx_gen == self.ctx.generation:
            return
        if self.world > 1:
            raise MidasError("the codebook's context was re-created (a later engine needed more capacity) while a sharded FilterEngine "
                             "was using it: create the largest engine first, or call connect_peers() on every rank")
        p = C.c_void_p()
        call("mt_step_local_sum_ptr", self.ctx.h, C.byref(p))
        self._local_sum_ptr = p.value
        if self.prune:
            pass  # Context.ensure_capacity re-uploads the mesh itself
        self._ctx_gen = self.ctx.generation

    def connect_peers(self) -> bool:
        """sharded runs: exchange the CUDA IPC handles of the contexts' exchange buffers (one 64-byte
        all-gather, once) and map them; afterwards a step is codebook query + kernel A + ONE cooperative
        kernel per GPU that stores its weight sum into the peers' buffers over NVLink -- no collective on
        the data path.  Returns False (and keeps the all-gather path) when peer mapping is unavailable."""
        import torch.distributed as dist
        import warnings

        if not (dist.is_available() and dist.is_initialized()):
            return False
        mine = (C.c_ubyte * 64)()
        call("mt_dist_export", self.ctx.h, mine)
        local = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(self.dev)
        allh = torch.zeros(self.world * 64, dtype=torch.uint8, device=self.dev)
        dist.all_gather_into_tensor(allh, local, group=self.group)
        buf = (C.c_ubyte * (64 * self.world)).from_buffer_copy(bytes(allh.cpu().numpy().tobytes()))
        ok = torch.ones(1, dtype=torch.int32, device=self.dev)
        try:
            with torch.cuda.device(self.dev):
                call("mt_dist_import", self.ctx.h, self.rank, self.world, buf)
        except MidasError as e:
            warnings.warn(f"peer exchange unavailable, using the all-gather path: {e}")
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)  # all ranks take the same path
        self.peer_exchange = bool(ok.item())
        # rebalance() moves particles with an all-to-all: NCCL sets its point-to-point channels up on first use
        # (hundreds of milliseconds); do that here, at set-up time, not in the middle of a run (step 64)
        try:
            for elems in (1, 1 << 18):  # a tiny and a 1 MB-per-peer message: both protocols' buffers get set up
                one = torch.zeros(self.world * elems, dtype=torch.float32, device=self.dev)
                dist.all_to_all_single(torch.empty_like(one), one, group=self.group)
            torch.cuda.synchronize(self.dev)
        except RuntimeError:  # a backend without all-to-all: rebalance() will say so when it is needed
            pass
        self._ctx_gen = self.ctx.generation
        return self.peer_exchange

    # ------------------------------------------------------------------ state in / out
    def load_particles(self, poses: torch.Tensor, nn_hint: torch.Tensor | None = None, spatial_sort: bool = False):
        """poses: (n,4,4) CUDA.  nn_hint: optional codebook index per particle (any valid index is
        correct; a good one makes the first search cheap).  spatial_sort=True reorders the
        particles by the rank of their codebook match in the search index (k-d tree leaf order) so that neighbouring threads walk
        neighbouring keys (systematic resampling preserves the order afterwards); the applied
        permutation is returned (poses[perm] is what the engine holds), else None."""
        require_cuda(poses, "poses")
        n = poses.shape[0]
        if n > self.capacity:
            raise MidasError("load_particles: more particles than capacity")
        perm = None
        if spatial_sort:
            poses = poses.reshape(-1, 4, 4).float().contiguous()
            idx = self.cb.SE3_NN_idx(poses, hint=None if nn_hint is None else nn_hint.to(torch.int32).contiguous())
            rank = torch.empty(len(self.cb), dtype=torch.int32, device=self.dev)
            with torch.cuda.device(self.dev):
                call("mt_codebook_rank", self.ctx.h, ptr(rank), stream_ptr())
            perm = torch.sort(rank[idx.long()], stable=True).indices  # one-off, at load time only
            poses, nn_hint = poses[perm], idx[perm]
        with torch.cuda.device(self.dev):
            call("mt_aos_to_soa", ptr(poses.reshape(-1, 4, 4).float().contiguous()), n, ptr(self.soa[self.cur]), self.capacity, stream_ptr())
        if nn_hint is None:
            self.nn[self.cur][:n] = -1
        else:
            self.nn[self.cur][:n] = nn_hint.to(torch.int32)
        self.n = n
        self.n_dev[self.cur].fill_(n)
        if self.n_global is None or self.world == 1:
            self.n_global = n if self.world == 1 else self.n_global
        return perm

    def snap_to_codebook(self):
        """filter.py:159-160: particles.poses = codebook.SE3_NN(particles.poses)[0] (exhaustive NN)."""
        n = self.n
        keys = torch.empty((n, 6), dtype=torch.float32, device=self.dev)
        idx = self.nn[self.cur]
        with torch.cuda.device(self.dev):
            s = stream_ptr()
            call("mt_se3_keys", ptr(self.soa[self.cur]), self.capacity, n, ptr(keys), s)
            call("mt_nn_assign", self.ctx.h, ptr(keys), n, 0, 1, ptr(idx), s)
            aos = torch.empty((n, 4, 4), dtype=torch.float32, device=self.dev)
            call("mt_gather_rows_f32", ptr(self.cb.poses), ptr(idx), n, 16, ptr(aos), s)
            call("mt_aos_to_soa", ptr(aos), n, ptr(self.soa[self.cur]), self.capacity, s)

    def poses(self) -> torch.Tensor:
        n = self.count()
        return soa_to_aos(self.soa[self.cur], n)

    def count(self) -> int:
        if self.use_n_dev:
            self.n = int(self.n_dev[self.cur].item())
            self.check()
        return self.n

    def nn_idx(self) -> torch.Tensor:
        """codebook match of every particle (particles zeroed by the drift test are stored as
        -(idx+2) inside the engine; decoded here)."""
        s = self.nn[self.cur][: self.count()]
        return torch.where(s < -1, -(s + 2), s)

    def pruned_mask(self) -> torch.Tensor:
        return self.nn[self.cur][: self.count()] < -1

    def ancestors(self) -> torch.Tensor:
        return self.anc[: self.count()]

    @property
    def rmse(self) -> torch.Tensor:
        """(2,) float32 CUDA tensor: translation / rotation RMSE of the last step that was given a ground truth."""
        return self._rm[self._rm_last]

    def read_rmse_async(self, out_pinned: torch.Tensor) -> torch.cuda.Event:
        """Copy the last step's rmse (8 bytes) into a pinned host tensor on the engine's copy stream, i.e. beside the next
        step's kernels instead of between two launches on the compute stream; returns the event to synchronise on.
        The result buffers alternate with the particle buffers, so the value stays valid for one more step."""
        self._ev_read.record(torch.cuda.current_stream(self.dev))
        self._io.wait_event(self._ev_read)
        ev = torch.cuda.Event()
        with torch.cuda.stream(self._io):
            out_pinned.copy_(self._rm[self._rm_last], non_blocking=True)
            ev.record(self._io)
        self._ev_rd[self._rm_last] = ev
        return ev

    # ------------------------------------------------------------------ one filter step
    def _fill(self, odom, u, tn, rot, gt, softmax, prune=True, resample=True):
        a = self._a
        a.d_soa_cur, a.d_soa_next = ptr(self.soa[self.cur]), ptr(self.soa[1 - self.cur])
        a.stride = self.capacity
        a.d_nn_cur, a.d_nn_next, a.d_anc = ptr(self.nn[self.cur]), ptr(self.nn[1 - self.cur]), ptr(self.anc)
        a.n = self.n
        a.odom = odom
        a.d_tn, a.d_rot = ptr(tn), ptr(rot)
        a.sig_t, a.sig_r = self.sig_t, self.sig_r
        a.seed, a.step, a.first_gid = self.seed, self.t, self.rank * (1 << 40)
        a.softmax, a.u, a.resample = int(bool(softmax)), float(u), int(bool(resample))
        a.gt = ptr(gt) if gt is not None else None
        a.d_rmse2 = ptr(self._rm[self.cur])
        a.rank, a.world = self.rank, self.world
        a.n_global = int(self.n_global if (self.world > 1 and self.n_global is not None) else self._n_children)
        a.d_shard_sums = ptr(self.shard_sums) if self.world > 1 else None
        a.d_n_out = ptr(self.n_dev[1 - self.cur])
        a.d_n_in = ptr(self.n_dev[self.cur]) if self.use_n_dev else None
        a.prune_dist = self.pen_max if (self.prune and prune) else 0.0
        a.d_cb_poses = ptr(self.cb.poses) if self.prune else None
        a.table_ready_event = None
        a.fuse_sums = int(self.fuse_sums and (self.world == 1 or (self.peer_exchange and resample)))
        return a

    def step(self, code: torch.Tensor, odom: torch.Tensor, u: float | None = None, tn: torch.Tensor | None = None,
             rot: torch.Tensor | None = None, gt: torch.Tensor | None = None, softmax: bool = True,
             resample: bool = True, prune: bool = True):
        """code: (1,D)/(D,) tactile code, host or device, float32/float64.
        odom: (4,4) host tensor / 16 floats.  u: systematic offset in [0,1) (drawn from the
        engine's CPU generator when None).  tn/rot: optional (n,3) float32 CUDA noise
        (parity mode); Philox in-kernel otherwise.  gt: optional (4,4) host pose -> self.rmse.
        prune: apply remove_invalid_particles (needs mesh_vertices at construction)."""
        if self.n == 0:
            raise MidasError("step: no particles loaded")
        self._refresh_ctx()
        if u is None:
            u = float(torch.rand(1, generator=self._rng).item())
        # the code goes into the engine's own buffer of this parity (the step graph is keyed on that address).  A host
        # tensor -- the step's only per-frame tensor input, D*8 bytes -- is copied on the copy stream: the transfer
        # runs beside the previous step's kernels instead of between two graph launches on the compute stream.
        par = self.cur
        q = self._q[par]
        src = code.reshape(-1)
        main = torch.cuda.current_stream(self.dev)
        if self._ev_rd[par] is not None:  # this parity's result buffer is about to be rewritten
            main.wait_event(self._ev_rd[par])
            self._ev_rd[par] = None
        if src.device.type == "cpu":
            self._io.wait_event(self._ev_step[par])  # the last step that read this buffer has finished
            with torch.cuda.stream(self._io):
                q.copy_(src, non_blocking=True)
                self._ev_q[par].record(self._io)
            main.wait_event(self._ev_q[par])
        else:
            q.copy_(src, non_blocking=True)
        odom16 = odom if isinstance(odom, Odom16) else prepare_odom(odom)
        gt_h = None
        if gt is not None:
            gt_h = gt if (gt.device.type == "cpu" and gt.dtype == torch.float32 and gt.is_contiguous()) else gt.detach().float().cpu().contiguous()
        if tn is not None:
            tn, rot = tn.contiguous(), rot.contiguous()
        a = self._fill(odom16, u, tn, rot, gt_h, softmax, prune, resample)
        with torch.cuda.device(self.dev):
            s = stream_ptr()
            fused = C.c_int(1)
            if self.world > 1:
                call("mt_step_is_fused", self.ctx.h, C.byref(a), C.byref(fused))
            if fused.value or not resample:
                # query (side stream / graph branch) + motion + SE3_NN + queue consumers + resampling: ONE call
                call("mt_step", self.ctx.h, C.byref(a), ptr(q), dtype_code(q), int(self.use_graph and resample), s)
            else:  # peers not mappable: sums -> NCCL all-gather (8 bytes / GPU) -> resampling
                call("mt_codebook_query", self.ctx.h, ptr(q), dtype_code(q), 0, s)
                call("mt_step_a", self.ctx.h, C.byref(a), s)
                self._allgather_sums()
                call("mt_step_b", self.ctx.h, C.byref(a), s)
        self._ev_step[par].record(main)
        self._rm_last = par
        if resample:
            self.cur = 1 - self.cur
            if self.world > 1:
                # the children of this GPU's parents: the count lives on the device from now on (the grids cover the
                # whole capacity, an overflow raises MT_STAT_OVERFLOW, checked in count() / rebalance())
                self.use_n_dev = True
                if self.rebalance_every and (self.t + 1) % self.rebalance_every == 0:
                    self.rebalance()
        self.t += 1

    # ------------------------------------------------------------------ the reference's loop body, particle count varying
    def step_loop(self, code, odom, u=None, tn=None, rot=None, gt=None, softmax=True, prune=True, count=0, floor=1000,
                  cluster_every=50, eps=1e-2):
        """One iteration of filter.py:154-190 INCLUDING cluster_particles (every `cluster_every`-th call), get_cluster_centers
        ("quat_avg") and annealing: the particle count shrinks / grows with the cluster variance like the reference's
        (particle_filter.py:405-447).  Single GPU.  Same kernels as step(); between the weighting and the resampling
        the cluster moments, the order statistic of the weights (radix select) and -- when growing -- the duplication of
        the heaviest particles run as library calls; the variance (4 bytes) is read by the host, which takes the
        annealing decision exactly like the reference (one synchronisation per frame, where the reference has several).
        Shrinking is done in place: the k lightest particles are marked weightless and the resampler draws n - k
        children, which is what removing them first gives.  Returns (cluster_poses, cluster_stds)."""
        if self.world != 1:
            raise MidasError("step_loop: single GPU")
        self.step(code, odom, u=0.0, tn=tn, rot=rot, gt=gt, softmax=softmax, resample=False, prune=prune)
        n = self.n
        dev = self.dev
        if not hasattr(self, "labels") or self.labels.shape[0] != self.capacity:
            self.labels = torch.zeros(self.capacity, dtype=torch.int64, device=dev)
            self._anneal_var, self._init_particles = None, n
        w = self.weights()  # normalised float64, 0 for pruned particles
        aos = soa_to_aos(self.soa[self.cur], n)
        with torch.cuda.device(dev):
            s = stream_ptr()
            if cluster_every and count % cluster_every == 0:  # filter.py:182-183
                call("mt_dbscan", self.ctx.h, ptr(aos), n, C.c_double(float(eps)), max(int(n / 5), 1), ptr(self.labels), None, s)
            uniq, inv = torch.unique(self.labels[:n], return_inverse=True)
            K = int(uniq.shape[0])
            centers = torch.empty((K, 4, 4), dtype=torch.float32, device=dev)
            stds = torch.empty((K, 3), dtype=torch.float32, device=dev)
            for k0 in range(0, K, 16):
                kk = min(16, K - k0)
                lab = (inv - k0).to(torch.int32)
                lab = torch.where((lab >= 0) & (lab < kk), lab, torch.full_like(lab, -1)).contiguous()
                call("mt_cluster_centers", self.ctx.h, ptr(aos), ptr(w), ptr(lab), n, kk, 0, ptr(centers[k0:k0 + kk]), ptr(stds[k0:k0 + kk]), s)
        var = float(torch.mean(stds).item())  # the frame's one host read
        # ---- annealing (particle_filter.py:405-447), decision on the host like the reference
        n_children = n
        if self._anneal_var is None:
            self._anneal_var, self._init_particles = var, n
        elif var != 0.0:
            ratio = var / self._anneal_var
            self._anneal_var = var
            if ratio < 1:
                num = min(int((1.0 - ratio) * n), abs(n - floor), n // 3)
                if num:
                    sel = torch.empty(num, dtype=torch.int32, device=dev)
                    with torch.cuda.device(dev):
                        call("mt_select_k", self.ctx.h, ptr(w), n, num, 0, ptr(sel), None, stream_ptr())
                    nn = self.nn[self.cur]
                    v = nn[sel.long()]
                    nn[sel.long()] = torch.where(v >= 0, -(v + 2), v)  # weightless: the resampler gives them no children
                    n_children = n - num
            elif ratio > 1:
                num = min(int((ratio - 1.0) * n), n // 3)
                if num and num + n <= self._init_particles and num + n <= self.capacity:
                    sel = torch.empty(num, dtype=torch.int32, device=dev)
                    with torch.cuda.device(dev):
                        call("mt_select_k", self.ctx.h, ptr(w), n, num, 1, ptr(sel), None, stream_ptr())
                    sl = sel.long()
                    soa, nn = self.soa[self.cur], self.nn[self.cur]
                    soa[:, n:n + num] = soa[:, sl]  # Particles.add: duplicates appended in index order
                    nn[n:n + num] = nn[sl]
                    self.labels[n:n + num] = self.labels[sl]
                    n = n + num
                    n_children = n
                    self.n = n
                    self.n_dev[self.cur].fill_(n)
        if u is None:
            u = float(torch.rand(1, generator=self._rng).item())
        # ---- resampling of the annealed set: n_children systematic draws
        self._n_children = n_children
        odom16 = odom if isinstance(odom, Odom16) else prepare_odom(odom)
        a = self._fill(odom16, u, tn, rot, None, softmax, prune, True)
        a.step = self.t - 1
        with torch.cuda.device(dev):
            call("mt_step_b", self.ctx.h, C.byref(a), stream_ptr())
        self._n_children = 0
        self.cur = 1 - self.cur
        anc = self.anc[:n_children].long()
        self.labels[:n_children] = self.labels[anc]
        self.n = n_children
        self.n_dev[self.cur].fill_(n_children)
        return centers, stds

    def check(self):
        """raise if a step overflowed the particle buffers or a peer's weight sum never arrived (synchronises)"""
        st = self.ctx.stats()
        if st["overflow"] == 1:
            raise MidasError("FilterEngine: a shard outgrew the engine capacity (children were dropped); rebalance more often or raise the capacity")
        if st["overflow"] == 2:
            raise MidasError("FilterEngine: a peer's weight sum never arrived in the fused exchange (10 s); that step kept its particles")

    def rebalance(self):
        """sharded runs: children follow their parents, so a GPU whose particles carry more weight
        accumulates particles (slowly: weights are within a factor e).  This evens the shards out again
        with one all-to-all of 52 B per moved particle (poses + matches); global order is preserved.
        Synchronises with the host (reads the device-side counts)."""
        if self.world == 1:
            return
        import torch.distributed as dist

        n = self.count()  # (also raises on overflow / exchange time-out)
        counts = torch.zeros(self.world, dtype=torch.int64, device=self.dev)
        dist.all_gather_into_tensor(counts, self.n_dev[self.cur].reshape(1), group=self.group)
        counts = counts.cpu().tolist()
        send, recv, targets = rebalance_plan(counts, self.rank)
        if counts == targets:
            return
        soa, nn = self.soa[self.cur], self.nn[self.cur]
        rows = torch.cat([soa[0, :n], soa[1, :n], soa[2, :n], nn[:n].view(torch.float32).reshape(n, 1)], dim=1).contiguous()  # (n,13)
        out = torch.empty((sum(recv), 13), dtype=torch.float32, device=self.dev)
        dist.all_to_all_single(out, rows, output_split_sizes=recv, input_split_sizes=send, group=self.group)
        m = out.shape[0]
        if m > self.capacity:
            raise MidasError("rebalance: shard exceeds the engine capacity")
        for r in range(3):
            soa[r, :m] = out[:, 4 * r:4 * r + 4]
        nn[:m] = out[:, 12].contiguous().view(torch.int32)
        self.n = m
        self.n_dev[self.cur].fill_(m)
        self.use_n_dev = True

    def _allgather_sums(self):
        import torch.distributed as dist

        local = torch.empty(0)  # placeholder for type checkers
        # view of the library's local weight sum (device double written by kernel A)
        local = _device_double_view(self._local_sum_ptr, self.dev)
        dist.all_gather_into_tensor(self.shard_sums, local, group=self.group)

    def weights(self) -> torch.Tensor:
        """normalised float64 weights of the current particles (valid after a step with
        resample=False, i.e. before the children replace them)."""
        n = self.count()
        w = torch.empty(n, dtype=torch.float64, device=self.dev)
        a = self._a
        with torch.cuda.device(self.dev):
            call("mt_step_weights", self.ctx.h, C.byref(a), ptr(w), stream_ptr())
        return w


def _device_double_view(address: int, device) -> torch.Tensor:
    """zero-copy (1,) float64 CUDA tensor over a raw device address (library scratch)."""

    class _Holder:
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {
        "shape": (1,), "typestr": "<f8", "data": (address, False), "version": 3, "strides": None,
    }
    return torch.as_tensor(h, device=device)

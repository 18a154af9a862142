"""The loop of ``midastouch/filter/filter.py:131-233`` without the GUI / image networks: tactile codes
come from a callable (the TCN, or precomputed codes), everything else is the reference's sequence

    t>0: odom = inv(meas[prev]) @ meas[idx]; particles = pf.motionModel(particles, odom)      (154-155)
    t=0: particles = pf.init_filter(gt[idx], N); poses = codebook.SE3_NN(poses)[0]             (159-160)
    rmse = particle_rmse(particles, gt[idx])                                                     (164)
    weights = pf.get_similarity(code, codebook.SE3_NN(poses)[2])                              (170-173)
    particles, drifted = pf.remove_invalid_particles(particles); drifted -> re-project        (176-179)
    cluster_poses, cluster_stds = pf.get_cluster_centers(particles, "quat_avg")               (184-186)
    particles = pf.annealing(particles, mean(cluster_stds)); particles = pf.resampler(...)    (189-190)

    count % 50 == 0: particles = pf.cluster_particles(particles)                                (182-183)

with one deliberate difference: the frame index follows a fixed schedule instead of wall-clock pacing (134-139;
reproducible).  ``filter_stats`` has the reference's keys (99-116).  ``run_filter_engine`` is the same loop on the
resident FilterEngine: ``step_loop`` (cluster centres + annealing, varying particle count) or, with anneal=False, the
fixed-N ``step`` replayed from a CUDA graph."""
from __future__ import annotations

import time

import torch

from .engine import FilterEngine
from .particle_filter import Particles, particle_filter, particle_rmse


def _new_stats(cfg, pf, codebook, traj_size, init_particles):
    return {"rmse_t": [], "rmse_r": [], "time": [], "traj_size": traj_size, "avg_time": None, "total_time": 0,
            "cluster_poses": [], "cluster_stds": [], "obj_name": cfg.expt.obj_model, "tree_size": len(codebook),
            "noise_ratio": cfg.expt.params.noise_ratio, "init_noise": pf.init_noise, "init_particles": init_particles,
            "num_particles": [], "log_id": str(cfg.expt.log_id).zfill(2), "trial_id": 0}


def tactile_code_fn(tdn, tcn, tac_render, tactile_images, small_parts: bool = False):
    """the front of a frame as filter.py:140-147 runs it -- image -> height map (``TDN.image2heightmap``) -> contact mask
    (``TDN.heightmap2mask``) -> tactile code (``TCN.cloud_to_tactile_code``) -- as the ``code_fn`` of the loops below"""

    def code_fn(idx):
        heightmap = tdn.image2heightmap(tactile_images[idx])
        mask = tdn.heightmap2mask(heightmap, small_parts=small_parts)
        return tcn.cloud_to_tactile_code(tac_render, heightmap, mask)

    return code_fn


def run_filter(cfg, pf: particle_filter, codebook, code_fn, gt_p: torch.Tensor, meas_p: torch.Tensor, schedule=None,
               softmax: bool = True, floor: int = 1000, resample: str | None = None) -> dict:
    """drop-in classes, reference order.  code_fn(idx) -> (1,D) tactile code.  gt_p / meas_p: (T,4,4) CUDA."""
    N = int(cfg.expt.params.num_particles)
    resample = resample or getattr(cfg.expt.params, "resample", "weighted_random")
    schedule = list(range(gt_p.shape[0])) if schedule is None else list(schedule)
    stats = _new_stats(cfg, pf, codebook, len(schedule), N)
    prev_idx, particles, count = 0, None, 0
    for idx in schedule:
        t0 = time.time()
        code = code_fn(idx)
        if prev_idx > 0:
            odom = torch.inverse(meas_p[prev_idx]) @ meas_p[idx]
            particles = pf.motionModel(particles, odom)
        else:  # the reference re-initialises until the previous frame index was > 0 (filter.py:152,230)
            particles = pf.init_filter(gt_p[idx], N)
            particles.poses, _, _ = codebook.SE3_NN(particles.poses)
        rmse_t, rmse_r = particle_rmse(particles, gt_p[idx])
        stats["rmse_t"].append(rmse_t.item())
        stats["rmse_r"].append(rmse_r.item())
        _, _, nn_codes = codebook.SE3_NN(particles.poses)
        particles.weights = pf.get_similarity(code, nn_codes, softmax=softmax)
        particles, drifted = pf.remove_invalid_particles(particles)
        if drifted:
            particles.poses, _, _ = codebook.SE3_NN(particles.poses)
        if count % 50 == 0:
            particles = pf.cluster_particles(particles)
        cluster_poses, cluster_stds = pf.get_cluster_centers(particles, method="quat_avg")
        particles = pf.annealing(particles, torch.mean(cluster_stds), floor=floor)
        particles = pf.resampler(particles, resample=resample)
        stats["cluster_poses"].append(cluster_poses)
        stats["cluster_stds"].append(cluster_stds)
        stats["num_particles"].append(len(particles))
        stats["time"].append(time.time() - t0)
        stats["total_time"] = sum(stats["time"])
        prev_idx = idx
        count += 1
    stats["avg_time"] = stats["total_time"] / max(len(schedule), 1)
    return stats


def run_filter_engine(cfg, pf: particle_filter, codebook, code_fn, gt_p: torch.Tensor, meas_p: torch.Tensor, schedule=None,
                      softmax: bool = True, seed: int = 0, anneal: bool = True, floor: int = 1000, teacher_forced: bool = True) -> dict:
    """same loop on the resident engine.  anneal=True (default): one ``FilterEngine.step_loop`` per frame = the reference's
    whole loop body, cluster_particles every 50th frame, cluster centres and annealing included, so the particle count
    follows the cluster variance like ``run_filter``'s.  teacher_forced: motion noise and the systematic offset are
    drawn exactly like the drop-in classes draw them (CPU default generator / CUDA generator), which makes the two
    loops comparable frame by frame; otherwise the engine's in-kernel Philox noise.  anneal=False: fixed particle
    count, ``FilterEngine.step`` (one CUDA-graph replay per frame, no host synchronisation inside the loop)."""
    N = int(cfg.expt.params.num_particles)
    schedule = list(range(gt_p.shape[0])) if schedule is None else list(schedule)
    stats = _new_stats(cfg, pf, codebook, len(schedule), N)
    eng = FilterEngine(codebook, capacity=N, sig_t=pf.motion_noise["sig_t"], sig_r=pf.motion_noise["sig_r"], seed=seed,
                       mesh_vertices=pf.mesh_vertices_ds, pen_max=pf.pen_max)
    dev = gt_p.device
    gt_h, meas_h = gt_p.detach().float().cpu(), meas_p.detach().float().cpu()
    rm = torch.zeros((len(schedule), 2), dtype=torch.float32, device=dev)
    prev_idx, count = 0, 0
    t_start = time.time()
    for k, idx in enumerate(schedule):
        code = code_fn(idx)
        tn = rot = None
        if prev_idx > 0:
            odom = torch.inverse(meas_h[prev_idx]) @ meas_h[idx]
            if teacher_forced:  # the draws of particle_filter.add_noise_to_odom (326-335): translation first, CPU generator
                n = eng.n
                tn = torch.normal(mean=0.0, std=pf.motion_noise["sig_t"], size=(n, 3)).to(dev)
                rot = torch.normal(mean=0.0, std=pf.motion_noise["sig_r"], size=(n, 3)).to(dev)
        else:
            parts = pf.init_filter(gt_p[idx], N)
            eng.load_particles(parts.poses)
            eng.snap_to_codebook()
            # frame without motion: weights + resampling only (identity odometry, no noise)
            odom = torch.eye(4)
            tn = rot = torch.zeros((N, 3), dtype=torch.float32, device=dev)
        u = float(torch.rand(1, device=dev).item()) if teacher_forced else None  # particle_filter.py:260
        if anneal:
            cluster_poses, cluster_stds = eng.step_loop(code, odom, u=u, tn=tn, rot=rot, gt=gt_h[idx], softmax=softmax, count=count, floor=floor)
            stats["cluster_poses"].append(cluster_poses)
            stats["cluster_stds"].append(cluster_stds)
        else:
            eng.step(code, odom, u=u, tn=tn, rot=rot, gt=gt_h[idx], softmax=softmax)
        rm[k].copy_(eng.rmse)
        stats["num_particles"].append(eng.n)
        prev_idx = idx
        count += 1
    rm = rm.cpu()
    stats["rmse_t"], stats["rmse_r"] = rm[:, 0].tolist(), rm[:, 1].tolist()
    stats["total_time"] = time.time() - t_start
    stats["avg_time"] = stats["total_time"] / max(len(schedule), 1)
    stats["time"] = [stats["avg_time"]] * len(schedule)
    stats["engine"] = eng
    return stats

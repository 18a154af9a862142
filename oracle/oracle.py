"""TEST INFRASTRUCTURE ONLY -- the CPU oracle for the MidasTouch per-step hot path.

A CPU restatement (torch-CPU / numpy, the reference's own arithmetic library) of
the functions SURVEY.md section 8a lists, each citing the reference file:line it
follows (paths relative to /root/reference/).  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import this module; the product (``midastouch_b200``) never
does and fails loudly without its CUDA library.

PARITY PIN STATUS (see DESIGN.md "Oracle"):
  * The reference ships no tests, golden vectors or fixtures for this path
    (SURVEY.md section 4), so the pins are outputs of the UNMODIFIED reference
    module itself, produced in the build container by ``oracle/gen_golden.py``
    through ``oracle/ref_shim.py`` and committed under ``tests/golden/``:
    get_similarity, resampler("low_var" -- the real Python loop -- and
    "low_var_batch"), add_noise_to_odom / motionModel, remove_invalid_particles,
    annealing, particle_rmse, euler_angles_to_matrix.  ``tests/test_oracle.py``
    checks every function below against those vectors.
  * theseus (SO3.log_map / to_quaternion), pynanoflann (exact L2 k-NN) and
    MinkowskiEngine are third-party, un-vendored and un-pinned in the reference
    (README.md:74,85-87; environment.yml:45).  Their published algorithms are
    restated here (so3_log_map, so3_to_quaternion, se3_nn) and pinned against
    scipy.spatial.transform.Rotation / brute force instead: for those three the
    status is "parity unpinned against the dependency itself".
"""
from __future__ import annotations

import math

import numpy as np
import torch

# --------------------------------------------------------------------------------------
# theseus restatements (third-party; call sites pose.py:19-23, 26-34)
# --------------------------------------------------------------------------------------
# theseus/constants.py: SO3 float32 thresholds (near-zero 5e-3, near-pi 1e-2); float64
# (5e-6, 1e-7).  The reference always calls these on float32 poses.
_SO3_NEAR_ZERO_EPS = {torch.float32: 5e-3, torch.float64: 5e-6}
_SO3_NEAR_PI_EPS = {torch.float32: 1e-2, torch.float64: 1e-7}


def _sine_axis(R: torch.Tensor) -> torch.Tensor:
    s = R.new_zeros(R.shape[0], 3)
    s[:, 0] = 0.5 * (R[:, 2, 1] - R[:, 1, 2])
    s[:, 1] = 0.5 * (R[:, 0, 2] - R[:, 2, 0])
    s[:, 2] = 0.5 * (R[:, 1, 0] - R[:, 0, 1])
    return s


def _major_axis_rows(R: torch.Tensor, cosine: torch.Tensor):
    """near-pi helper shared by log_map and to_quaternion (theseus SO3)."""
    n = R.shape[0]
    aux = torch.arange(n)
    dd = torch.diagonal(R, dim1=1, dim2=2)
    major = torch.logical_and(dd[:, 1] > dd[:, 0], dd[:, 1] > dd[:, 2]).long() + 2 * torch.logical_and(
        dd[:, 2] > dd[:, 0], dd[:, 2] > dd[:, 1]
    ).long()
    sel = 0.5 * (R[aux, major] + R[aux, :, major])
    sel[aux, major] -= cosine
    return aux, major, sel


def so3_log_map(R: torch.Tensor) -> torch.Tensor:
    """theseus ``SO3(tensor=R).log_map()`` restated (used by pose.py:19-23).

    R: (N,3,3) float32 -> (N,3) rotation vectors, same dtype.
    """
    R = R.reshape(-1, 3, 3)
    dt = R.dtype
    sa = _sine_axis(R)
    cosine = 0.5 * (torch.diagonal(R, dim1=1, dim2=2).sum(dim=1) - 1)
    sine = sa.norm(dim=1)
    theta = torch.atan2(sine, cosine)
    near_zero = theta < _SO3_NEAR_ZERO_EPS[dt]
    near_pi = 1 + cosine <= _SO3_NEAR_PI_EPS[dt]
    nz_or_pi = torch.logical_or(near_zero, near_pi)
    one = torch.ones(1, dtype=dt)
    sine_nz = torch.where(nz_or_pi, one, sine)
    scale = torch.where(nz_or_pi, 1 + sine**2 / 6, theta / sine_nz)
    ret = sa * scale.view(-1, 1)
    aux, major, sel = _major_axis_rows(R, cosine)
    axis = sel / torch.where(near_zero, one, sel.norm(dim=1)).view(-1, 1)
    sign_tmp = sa[aux, major].sign()
    sign = torch.where(sign_tmp != 0, sign_tmp, torch.ones_like(sign_tmp))
    return torch.where(near_pi.view(-1, 1), axis * (theta * sign).view(-1, 1), ret)


def so3_to_quaternion(R: torch.Tensor) -> torch.Tensor:
    """theseus ``SO3(tensor=R).to_quaternion()`` restated -> (N,4) as (w,x,y,z)
    (used by tf_to_xyzquat, pose.py:26-34)."""
    R = R.reshape(-1, 3, 3)
    dt = R.dtype
    sa = _sine_axis(R)
    w = 0.5 * (1 + torch.diagonal(R, dim1=1, dim2=2).sum(dim=1)).clamp(0, 4).sqrt()
    near_pi = w <= _SO3_NEAR_PI_EPS[dt]
    one = torch.ones(1, dtype=dt)
    ret = R.new_zeros(R.shape[0], 4)
    ret[:, 0] = w
    ret[:, 1:] = 0.5 * sa / torch.where(near_pi, one, w).view(-1, 1)
    cosine = 0.5 * (torch.diagonal(R, dim1=1, dim2=2).sum(dim=1) - 1)
    aux, major, sel = _major_axis_rows(R, cosine)
    nrm = sel.norm(dim=1)
    axis = sel / torch.where(nrm == 0, one, nrm).view(-1, 1)
    sign_tmp = sa[aux, major].sign()
    sign = torch.where(sign_tmp != 0, sign_tmp, torch.ones_like(sign_tmp))
    sin_half = (1 - w * w).clamp(0, 1).sqrt()
    ret[:, 1:] = torch.where(near_pi.view(-1, 1), axis * (sin_half * sign).view(-1, 1), ret[:, 1:])
    return ret


# --------------------------------------------------------------------------------------
# codebook: 6-D keys and exact 1-NN  (tactile_tree.py:43-58, 73-77)
# --------------------------------------------------------------------------------------
def r3_se3(poses: torch.Tensor, w: float = 0.01) -> torch.Tensor:
    """``R3_SE3`` (tactile_tree.py:73-77): key = [(1-w) t, w Log_SO3(R)] float32 (N,6)."""
    poses = poses.reshape(-1, 4, 4)
    return torch.cat(((1.0 - w) * poses[:, :3, 3], w * so3_log_map(poses[:, :3, :3])), dim=1)


def fma_f32(a: np.ndarray, b: np.ndarray, c: np.ndarray) -> np.ndarray:
    """correctly rounded float32 fma(a, b, c) = round(a*b + c) (what fmaf / FFMA compute).
    a*b is exact in float64; the float64 sum may round, and rounding that to float32 could then land on
    the wrong side of a float32 midpoint (double rounding), so the TwoSum residual decides such ties."""
    a64, b64, c64 = (np.asarray(x, dtype=np.float32).astype(np.float64) for x in (a, b, c))
    t = a64 * b64
    s = t + c64
    bb = s - t
    e = (t - (s - bb)) + (c64 - bb)  # s + e == t + c exactly
    r = s.astype(np.float32)
    diff = s - r.astype(np.float64)
    with np.errstate(invalid="ignore"):
        r2 = np.nextafter(r, np.where(diff > 0, np.float32(np.inf), np.float32(-np.inf)).astype(np.float32))
        mid = (diff != 0) & np.isfinite(r2) & (np.abs(diff) == np.abs(r2.astype(np.float64) - s))
        # only when s is exactly a float32 midpoint can the residual change the result: it breaks the tie
        toward_r2 = mid & (((e > 0) & (r2 > r)) | ((e < 0) & (r2 < r)))
    out = np.where(toward_r2, r2, r)
    return out.astype(np.float32)


def l2_sq_f32(keys: np.ndarray, q: np.ndarray) -> np.ndarray:
    """squared L2 in float32 with the fixed operation order the CUDA kernels use (mt_math.cuh mt_key_dist):
    lo = fma(d4, d4, fma(d2, d2, d0*d0)), hi = fma(d5, d5, fma(d3, d3, d1*d1)), result = lo + hi, with
    d_k = key_k - q_k; every operation correctly rounded to float32.
    keys (M,6) f32, q (6,) or (N,6) f32 broadcasting against keys."""
    d = (keys.astype(np.float32) - q.astype(np.float32)).astype(np.float32)
    if d.shape[-1] != 6:
        raise ValueError("l2_sq_f32: 6-D keys")
    lo = (d[..., 0] * d[..., 0]).astype(np.float32)
    hi = (d[..., 1] * d[..., 1]).astype(np.float32)
    lo = fma_f32(d[..., 2], d[..., 2], lo)
    hi = fma_f32(d[..., 3], d[..., 3], hi)
    lo = fma_f32(d[..., 4], d[..., 4], lo)
    hi = fma_f32(d[..., 5], d[..., 5], hi)
    return (lo + hi).astype(np.float32)


def l2_sq_f32_sequential(keys: np.ndarray, q: np.ndarray) -> np.ndarray:
    """squared L2 in float32 accumulated the way nanoflann's L2_Simple_Adaptor does for a 6-D point: a plain left-to-right
    sum of (a_k - b_k)^2, every operation rounded to float32, no fused multiply-add.  Only used to COUNT how many nearest
    neighbours would change under that order (the kernels and ``l2_sq_f32`` use two interleaved fma chains)."""
    d = (keys.astype(np.float32) - q.astype(np.float32)).astype(np.float32)
    acc = np.zeros(d.shape[:-1], dtype=np.float32)
    for k in range(6):
        acc = (acc + (d[..., k] * d[..., k]).astype(np.float32)).astype(np.float32)
    return acc


def nn_order_sensitivity(keys: np.ndarray, queries: np.ndarray, k: int = 4, workers: int = -1):
    """(#queries whose argmin differs between the two float32 accumulation orders, #queries with a float32 near-tie,
    indices under each order): the float64 k-d tree proposes the k nearest, both orders re-rank them."""
    from scipy.spatial import cKDTree

    keys = np.ascontiguousarray(keys, dtype=np.float32)
    queries = np.ascontiguousarray(queries, dtype=np.float32)
    tree = cKDTree(keys.astype(np.float64))
    _, cand = tree.query(queries.astype(np.float64), k=k, workers=workers)

    def pick(d):
        dmin = d.min(axis=1, keepdims=True)
        return np.where(d == dmin, cand, np.iinfo(np.int64).max).min(axis=1).astype(np.int64)

    a = pick(l2_sq_f32(keys[cand], queries[:, None, :]))
    b = pick(l2_sq_f32_sequential(keys[cand], queries[:, None, :]))
    d64 = ((keys[cand].astype(np.float64) - queries[:, None, :].astype(np.float64)) ** 2).sum(-1)
    d64.sort(axis=1)
    near = int(((d64[:, 1] - d64[:, 0]) <= 4e-7 * d64[:, 1]).sum())
    return int((a != b).sum()), near, a, b


def nn_brute(keys: np.ndarray, queries: np.ndarray, chunk: int = 512) -> np.ndarray:
    """exact 1-NN by exhaustive search, ties -> lowest codebook index.  This is the
    semantics of pynanoflann.KDTree(metric="L2").kneighbors(n_neighbors=1)
    (tactile_tree.py:34-41,50-53; nanoflann is an exact k-d tree)."""
    keys = np.ascontiguousarray(keys, dtype=np.float32)
    queries = np.ascontiguousarray(queries, dtype=np.float32)
    out = np.empty(queries.shape[0], dtype=np.int64)
    for s in range(0, queries.shape[0], chunk):
        q = queries[s : s + chunk]
        d = l2_sq_f32(keys[None, :, :], q[:, None, :])
        out[s : s + chunk] = np.argmin(d, axis=1)
    return out


def nn_exact(keys: np.ndarray, queries: np.ndarray, k: int = 8, workers: int = -1) -> np.ndarray:
    """exact 1-NN for large N: a float64 k-d tree (scipy cKDTree, exact) proposes the
    k nearest, the float32 fixed-order distance re-ranks them (ties -> lowest index),
    which equals ``nn_brute`` (checked in tests/test_oracle.py)."""
    from scipy.spatial import cKDTree

    keys = np.ascontiguousarray(keys, dtype=np.float32)
    queries = np.ascontiguousarray(queries, dtype=np.float32)
    k = min(k, keys.shape[0])
    tree = cKDTree(keys.astype(np.float64))
    _, cand = tree.query(queries.astype(np.float64), k=k, workers=workers)
    cand = cand.reshape(queries.shape[0], k)
    d = l2_sq_f32(keys[cand], queries[:, None, :])
    dmin = d.min(axis=1, keepdims=True)
    cand_masked = np.where(d == dmin, cand, np.iinfo(np.int64).max)
    return cand_masked.min(axis=1).astype(np.int64)


def se3_nn(cb_keys: torch.Tensor, query_poses: torch.Tensor, exact_large: bool = True) -> torch.Tensor:
    """index part of ``tactile_tree.SE3_NN`` (tactile_tree.py:43-53) -> (N,) int64."""
    q = r3_se3(query_poses.clone().float()).numpy()
    k = cb_keys.numpy()
    if exact_large and q.shape[0] * k.shape[0] > 5_000_000:
        return torch.from_numpy(nn_exact(k, q))
    return torch.from_numpy(nn_brute(k, q))


# --------------------------------------------------------------------------------------
# motion model  (particle_filter.py:319-377, pose.py:215-269)
# --------------------------------------------------------------------------------------
def euler_zyx_matrix(rot_deg: torch.Tensor) -> torch.Tensor:
    """``euler_angles_to_matrix(torch.deg2rad(rot), "ZYX")`` (pose.py:215-269):
    Rn = Rz(a0) @ Ry(a1) @ Rx(a2), float32."""
    a = torch.deg2rad(rot_deg)
    c, s = torch.cos(a), torch.sin(a)
    one, zero = torch.ones_like(a[:, 0]), torch.zeros_like(a[:, 0])

    def m(*flat):
        return torch.stack(flat, -1).reshape(-1, 3, 3)

    Rz = m(c[:, 0], -s[:, 0], zero, s[:, 0], c[:, 0], zero, zero, zero, one)
    Ry = m(c[:, 1], zero, s[:, 1], zero, one, zero, -s[:, 1], zero, c[:, 1])
    Rx = m(one, zero, zero, zero, c[:, 2], -s[:, 2], zero, s[:, 2], c[:, 2])
    return Rz @ Ry @ Rx


def draw_motion_noise(N: int, sig_t: float, sig_r: float, mul: float = 1.0):
    """the reference's RNG contract (particle_filter.py:326-335): two (N,3) float32
    normals from the CPU default generator, translation first."""
    tn = torch.normal(mean=0.0, std=float(mul) * sig_t, size=(N, 3))
    rot = torch.normal(mean=0.0, std=float(mul) * sig_r, size=(N, 3))
    return tn, rot


def noisy_odom(odom: torch.Tensor, tn: torch.Tensor, rot_deg: torch.Tensor) -> torch.Tensor:
    """``add_noise_to_odom`` (particle_filter.py:319-345) with the noise given."""
    N = tn.shape[0]
    Tn = torch.zeros((N, 4, 4), dtype=odom.dtype)
    Tn[:, :3, :3], Tn[:, :3, 3], Tn[:, 3, 3] = euler_zyx_matrix(rot_deg), tn, 1
    return odom[None].expand(N, 4, 4) @ Tn


def motion_model(poses: torch.Tensor, odom: torch.Tensor, tn: torch.Tensor, rot_deg: torch.Tensor):
    """``motionModel`` (particle_filter.py:359-377): poses @ (odom @ Tn), then the
    quaternion validity prune of ``check_quats`` (347-357).  Returns (poses, keep mask)."""
    out = poses @ noisy_odom(odom, tn, rot_deg)
    qn = torch.norm(so3_to_quaternion(out[:, :3, :3]), dim=1)
    keep = ~torch.logical_or(qn == 0, torch.isnan(qn))
    return out, keep


def init_filter(gt_pose: torch.Tensor, tn: torch.Tensor, rot_deg: torch.Tensor) -> torch.Tensor:
    """``init_filter`` (particle_filter.py:129-145): gt @ [R_zyx(rot) | tn] with scipy's
    float64 ``from_euler("zyx", degrees=True)`` cast into float32."""
    from scipy.spatial.transform import Rotation as R

    N = tn.shape[0]
    Rn = torch.tensor(R.from_euler("zyx", rot_deg.numpy(), degrees=True).as_matrix())
    Tn = torch.zeros((N, 4, 4), dtype=gt_pose.dtype)
    Tn[:, :3, :3], Tn[:, :3, 3], Tn[:, 3, 3] = Rn, tn, 1
    return gt_pose[None].expand(N, 4, 4) @ Tn


# --------------------------------------------------------------------------------------
# measurement: cosine similarity + softmax  (particle_filter.py:449-469)
# --------------------------------------------------------------------------------------
def get_similarity(queries: torch.Tensor, targets: torch.Tensor, softmax: bool = True) -> torch.Tensor:
    w = torch.nn.functional.cosine_similarity(torch.atleast_2d(queries), torch.atleast_2d(targets)).squeeze()
    if (not torch.isclose(w.max() - w.min(), torch.tensor([0.0], dtype=w.dtype))) and softmax:
        w = torch.softmax(w, dim=0)
    return w


def codebook_similarity(q: torch.Tensor, emb: torch.Tensor) -> torch.Tensor:
    """cos(q, E_m) for every codebook row (the heatmap form, filter.py:213-215);
    the engine's sim[M] table.  float64."""
    return torch.nn.functional.cosine_similarity(torch.atleast_2d(q).double(), emb.double())


# --------------------------------------------------------------------------------------
# resampling  (particle_filter.py:230-307)
# --------------------------------------------------------------------------------------
def systematic_cdf(weights: torch.Tensor):
    norm = weights / torch.sum(weights)
    return norm, torch.cumsum(norm, dim=0, dtype=torch.float64)


def systematic_locs(N: int, u: float) -> torch.Tensor:
    """sample locations exactly as particle_filter.py:254-261: float64 j/N plus a
    float32 offset u/N (u = torch.rand(1)), remainder 1."""
    locs = torch.tensor(range(0, N, 1), dtype=torch.float64) / N
    offset = torch.tensor([u], dtype=torch.float32) / N
    return torch.remainder(locs + offset, 1)


def low_var_indices(weights: torch.Tensor, u: float) -> torch.Tensor:
    """ancestor index of every output slot for ``resample="low_var"``
    (particle_filter.py:288-307), vectorised: the two-pointer loop assigns slot j the
    first i >= anc[j-1] with loc_j < C_i, i.e. cummax(searchsorted(C, loc, right=True));
    slots the loop never reaches (C_last <= loc) are returned as -1 (the reference
    leaves them as zero poses)."""
    N = weights.shape[0]
    _, C = systematic_cdf(weights)
    locs = systematic_locs(N, u)
    ss = torch.searchsorted(C, locs, right=True)
    anc = torch.cummax(ss, dim=0).values
    return torch.where(anc >= N, torch.full_like(anc, -1), anc)


def low_var_indices_loop(weights: torch.Tensor, u: float) -> torch.Tensor:
    """literal two-pointer loop (small N only)."""
    N = weights.shape[0]
    _, C = systematic_cdf(weights)
    locs = systematic_locs(N, u)
    out = torch.full((N,), -1, dtype=torch.int64)
    cur = 0
    for i in range(N):
        while cur < N and locs[cur] < C[i]:
            out[cur] = i
            cur += 1
    return out


def resample_gather(poses, weights, labels, anc):
    """gather with the reference's unfilled-slot behaviour (zeros)."""
    safe = anc.clamp(min=0)
    filled = anc >= 0
    p = torch.where(filled[:, None, None], poses[safe], torch.zeros_like(poses[safe]))
    w = torch.where(filled, weights[safe], torch.zeros_like(weights[safe]))
    l = torch.where(filled, labels[safe], torch.zeros_like(labels[safe]))
    return p, w, l


def resample_skip(weights: torch.Tensor) -> bool:
    """guard at particle_filter.py:237-241."""
    norm = weights / torch.sum(weights)
    return bool(torch.all(norm == 0)) or bool(torch.any(torch.isnan(norm)))


# --------------------------------------------------------------------------------------
# error metric  (particle_filter.py:472-496, pose.py:178-208)
# --------------------------------------------------------------------------------------
def particle_rmse(poses: torch.Tensor, gt_pose: torch.Tensor):
    poses = poses.reshape(-1, 4, 4)
    gt = gt_pose[None]
    R_diff = torch.matmul(gt[:, :3, :3], poses[:, :3, :3].permute(0, 2, 1))
    T_diff = gt[:, :3, 3] - poses[:, :3, 3]
    e_t = torch.norm(T_diff, dim=1)
    tr = R_diff[:, 0, 0] + R_diff[:, 1, 1] + R_diff[:, 2, 2]
    ang = torch.nan_to_num(torch.rad2deg(torch.acos((tr - 1.0) * 0.5)))
    ang = torch.where(ang > 180.0, ang - 360.0, ang)
    ang = torch.where(ang < -180.0, ang + 360.0, ang)
    return torch.sqrt(torch.mean(e_t**2)), torch.mean(torch.sqrt(torch.mean(ang**2, dim=0)))


# --------------------------------------------------------------------------------------
# "next" rows: drift prune, annealing, cluster centres
# --------------------------------------------------------------------------------------
def nearest_vertex_dist(vertices: np.ndarray, pts: np.ndarray) -> np.ndarray:
    """distance to the nearest mesh vertex (sklearn KDTree query at
    particle_filter.py:386-392; float64 like sklearn)."""
    from scipy.spatial import cKDTree

    d, _ = cKDTree(np.asarray(vertices, dtype=np.float64)).query(np.asarray(pts, dtype=np.float64), k=1)
    return d


def remove_invalid(poses: torch.Tensor, weights: torch.Tensor, vertices: np.ndarray, pen_max: float):
    """``remove_invalid_particles`` (particle_filter.py:379-403) -> (weights*m, drifted)."""
    dist = torch.tensor(nearest_vertex_dist(vertices, poses[:, :3, 3].numpy()))
    m = torch.ones(poses.shape[0])
    m[dist > pen_max] = 0.0
    return weights * m, bool(torch.sum(m) == 0)


def annealing_plan(n_particles: int, ratio: float, floor: int, init_particles: int):
    """the integer bookkeeping of ``annealing`` (particle_filter.py:421-446) ->
    ("remove"|"add"|"none", count)."""
    N = n_particles
    if ratio < 1:
        k = min(int((1.0 - ratio) * N), abs(n_particles - floor), n_particles // 3)
        return ("remove", k) if k else ("none", 0)
    if ratio > 1:
        k = min(int((ratio - 1.0) * N), n_particles // 3)
        if k + n_particles > init_particles:
            return ("none", 0)
        return ("add", k)
    return ("none", 0)


def quat_average(poses: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """``xyz_quat_averaged`` (pose.py:112-147) with ``torch.linalg.eigh`` standing in for
    the removed ``Tensor.eig`` (the 4x4 moment matrix is symmetric; the antipodal
    canonicalisation at pose.py:127,140 fixes the eigenvector sign).  -> (4,4) float32."""
    q = so3_to_quaternion(poses[:, :3, :3])  # w,x,y,z
    a = q[:, [1, 2, 3, 0]].clone()  # x,y,z,w
    a[a[:, 3] < 0] = -a[a[:, 3] < 0]
    Mm = (a[:, :, None] * w.view(-1, 1, 1) * a[:, None, :]).sum(0) / w.sum()
    evals, evecs = torch.linalg.eigh(Mm.double())
    v = evecs[:, -1].to(poses.dtype)
    if v[3] < 0:
        v = -v
    t = torch.sum(poses[:, :3, 3] * w[:, None] / w.sum(), dim=0)
    x, y, z, ww = (v / v.norm()).tolist()
    Rm = torch.tensor(
        [
            [1 - 2 * (y * y + z * z), 2 * (x * y - z * ww), 2 * (x * z + y * ww)],
            [2 * (x * y + z * ww), 1 - 2 * (x * x + z * z), 2 * (y * z - x * ww)],
            [2 * (x * z - y * ww), 2 * (y * z + x * ww), 1 - 2 * (x * x + y * y)],
        ],
        dtype=poses.dtype,
    )
    T = torch.eye(4, dtype=poses.dtype)
    T[:3, :3], T[:3, 3] = Rm, t
    return T


def cluster_centers(poses, weights, labels):
    """``get_cluster_centers(method="quat_avg")`` (particle_filter.py:153-206)."""
    weights = weights.float()
    uniq = torch.unique(labels)
    cp = torch.zeros((uniq.shape[0], 4, 4))
    cs = torch.zeros((uniq.shape[0], 3))
    for i, lab in enumerate(uniq):
        sel = labels == lab
        P, W = poses[sel], weights[sel]
        if torch.isclose(W.max() - W.min(), torch.tensor([0.0], dtype=W.dtype)):
            W = torch.ones_like(W)
        cp[i] = quat_average(P, W)
        cs[i] = torch.sqrt(torch.sum(((P[:, :3, 3] - cp[i, :3, 3]) ** 2 * W[:, None]) / W.sum(), dim=0))
    return cp, cs


# --------------------------------------------------------------------------------------
# one teacher-forced step of the loop body (filter.py:152-190), systematic resampling
# --------------------------------------------------------------------------------------
def filter_step(poses, odom, tn, rot_deg, cb_keys, cb_emb, q, u, softmax=True, gather_dot=True):
    """motion -> SE3_NN -> get_similarity -> low_var resample, exactly in the order of
    filter.py:154-190 (no prune/cluster/anneal).  Returns a dict of every intermediate."""
    moved, keep = motion_model(poses, odom, tn, rot_deg)
    nn_idx = se3_nn(cb_keys, moved)
    if gather_dot:  # reference form: gather (N,D) float64 then cosine (filter.py:170-173)
        w = get_similarity(q, cb_emb[nn_idx], softmax=softmax)
    else:
        w = codebook_similarity(q, cb_emb)[nn_idx]
        if softmax and not torch.isclose(w.max() - w.min(), torch.tensor([0.0], dtype=w.dtype)):
            w = torch.softmax(w, dim=0)
    anc = low_var_indices(w, u)
    return dict(moved=moved, keep=keep, nn_idx=nn_idx, weights=w, anc=anc, out_poses=moved[anc.clamp(min=0)])


# ------------------------------------------------------------------ SE(3) log / exp (theseus, unpinned)
def se3_log_map(poses: torch.Tensor) -> torch.Tensor:
    """th.SE3(tensor=T[:, :3, :]).log_map() (call site pose.py:101-105) restated in closed form, float64:
    tangent = [V^-1 t, Log_SO3(R)],  V^-1 = I - 1/2 [w]x + a [w]x^2,  a = (1 - th sin th / (2 (1 - cos th))) / th^2."""
    P = poses.reshape(-1, 4, 4).double()
    w = so3_log_map(P[:, :3, :3].float()).double()
    t = P[:, :3, 3]
    th2 = (w * w).sum(1)
    th = th2.sqrt()
    small = th < 5e-3
    ths = torch.where(small, torch.ones_like(th), th)
    a = torch.where(small, 1.0 / 12.0 + th2 / 720.0, (1.0 - 0.5 * ths * torch.sin(ths) / (1.0 - torch.cos(ths))) / (ths * ths))
    c1 = torch.cross(w, t, dim=1)
    c2 = torch.cross(w, c1, dim=1)
    return torch.cat([t - 0.5 * c1 + a[:, None] * c2, w], dim=1)


def se3_exp_map(x: torch.Tensor) -> torch.Tensor:
    """th.SE3.exp_map(tangent_vector=x).to_matrix() (pose.py:106-109), float64, one tangent (6,) -> (4,4)."""
    x = x.double().reshape(6)
    v, w = x[:3], x[3:]
    th2 = float((w * w).sum())
    th = th2 ** 0.5
    K = torch.tensor([[0.0, -w[2], w[1]], [w[2], 0.0, -w[0]], [-w[1], w[0], 0.0]], dtype=torch.float64)
    if th < 1e-6:
        A, B, Cc = 1.0 - th2 / 6.0, 0.5 - th2 / 24.0, 1.0 / 6.0 - th2 / 120.0
    else:
        import math

        A, B, Cc = math.sin(th) / th, (1.0 - math.cos(th)) / th2, (th - math.sin(th)) / (th2 * th)
    eye = torch.eye(3, dtype=torch.float64)
    T = torch.eye(4, dtype=torch.float64)
    T[:3, :3] = eye + A * K + B * (K @ K)
    T[:3, 3] = (eye + B * K + Cc * (K @ K)) @ v
    return T


def log_map_averaged(poses: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """pose.py:101-109: exp(sum_n w_n log(T_n) / sum w)."""
    lg = se3_log_map(poses)
    w = w.double()
    return se3_exp_map((lg * w[:, None]).sum(0) / w.sum()).float()


def cluster_centers_logmap(poses, weights, labels):
    """get_cluster_centers(method="logmap") (particle_filter.py:153-206) on the restated SE(3) maps."""
    uniq = torch.unique(labels)
    centers = torch.zeros((len(uniq), 4, 4))
    stds = torch.zeros((len(uniq), 3))
    wf = weights.float()
    for i, lab in enumerate(uniq):
        m = labels == lab
        tp, tw = poses[m], wf[m]
        if torch.isclose(tw.max() - tw.min(), torch.tensor([0.0])):
            tw = torch.ones_like(tw)
        centers[i] = log_map_averaged(tp, tw)
        stds[i] = torch.sqrt(torch.sum(((tp[:, :3, 3] - centers[i, :3, 3]) ** 2 * tw[:, None]) / tw.sum(), dim=0))
    return centers, stds


def dbscan_labels(poses: torch.Tensor, eps: float = 1e-2) -> torch.Tensor:
    """cluster_particles(method="euclidean") (particle_filter.py:208-228): sklearn itself is the oracle."""
    from sklearn.cluster import DBSCAN

    n = poses.shape[0]
    return torch.from_numpy(DBSCAN(eps=eps, min_samples=int(n / 5)).fit(poses[:, :3, 3].cpu().numpy()).labels_.astype("int64"))

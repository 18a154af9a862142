"""Seeded synthetic stand-ins for the YCB-Slide assets.

No meshes, codebooks, logs or network weights exist offline, so every BASELINE.json
config is realised with deterministic synthetic objects of the nominal extents:
surface samples + normals (the role of ``nontextured.stl``), a codebook of M sensor
poses on the surface with normal-aligned z, <=5 deg shear and random yaw (what
``sample_poses_on_mesh`` / ``pose_from_vertex_normal`` produce, mesh.py:84-135,
pose.py:375-455), L2-normalised positive embeddings (GeM + L2 output is positive,
minkloc.py:84-95), and a sliding trajectory with measurement noise
(data_gen/config/method/ycb_slide.yaml:10-15).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

# ------------------------------------------------------------------ primitives
# every primitive: (kind, params); surfaces are sampled area-weighted.


def _box_faces(center, ext):
    cx, cy, cz = center
    ex, ey, ez = ext
    faces = []
    for axis, e in enumerate((ex, ey, ez)):
        for sgn in (-1.0, 1.0):
            n = np.zeros(3)
            n[axis] = sgn
            o = np.array(center, dtype=np.float64) + n * e / 2
            u = np.zeros(3)
            v = np.zeros(3)
            u[(axis + 1) % 3] = 1
            v[(axis + 2) % 3] = 1
            lu = (ex, ey, ez)[(axis + 1) % 3]
            lv = (ex, ey, ez)[(axis + 2) % 3]
            faces.append(("rect", dict(o=o, u=u, v=v, lu=lu, lv=lv, n=n)))
    return faces


def _cyl(center, r, h, axis=2, caps=True):
    out = [("cyl", dict(c=np.array(center, dtype=np.float64), r=r, h=h, axis=axis))]
    if caps:
        for sgn in (-1.0, 1.0):
            n = np.zeros(3)
            n[axis] = sgn
            out.append(("disk", dict(c=np.array(center, dtype=np.float64) + n * h / 2, r=r, axis=axis, n=n)))
    return out


def _area(p):
    k, d = p
    if k == "rect":
        return d["lu"] * d["lv"]
    if k == "cyl":
        return 2 * math.pi * d["r"] * d["h"]
    return math.pi * d["r"] ** 2


def _sample_prim(p, n, rng):
    k, d = p
    if k == "rect":
        a = rng.uniform(-0.5, 0.5, n) * d["lu"]
        b = rng.uniform(-0.5, 0.5, n) * d["lv"]
        pts = d["o"][None] + a[:, None] * d["u"][None] + b[:, None] * d["v"][None]
        return pts, np.repeat(d["n"][None], n, 0)
    ax = d["axis"]
    i, j = (ax + 1) % 3, (ax + 2) % 3
    phi = rng.uniform(0, 2 * math.pi, n)
    pts = np.repeat(d["c"][None], n, 0)
    nrm = np.zeros((n, 3))
    if k == "cyl":
        z = rng.uniform(-0.5, 0.5, n) * d["h"]
        pts[:, i] += d["r"] * np.cos(phi)
        pts[:, j] += d["r"] * np.sin(phi)
        pts[:, ax] += z
        nrm[:, i], nrm[:, j] = np.cos(phi), np.sin(phi)
        return pts, nrm
    rr = d["r"] * np.sqrt(rng.uniform(0, 1, n))
    pts[:, i] += rr * np.cos(phi)
    pts[:, j] += rr * np.sin(phi)
    return pts, np.repeat(d["n"][None], n, 0)


OBJECTS = {
    # nominal extents (m): YCB 004_sugar_box
    "004_sugar_box": lambda: _box_faces((0, 0, 0), (0.089, 0.038, 0.175)),
    # L-shaped solid: body + handle
    "035_power_drill": lambda: _box_faces((0, 0, 0.035), (0.18, 0.06, 0.07)) + _box_faces((-0.04, 0, -0.065), (0.05, 0.045, 0.13)),
    # open cylinder + handle bar
    "025_mug": lambda: _cyl((0, 0, 0), 0.0465, 0.082, caps=True) + _box_faces((0.06, 0, 0), (0.027, 0.012, 0.05)),
    # bent rod ~45 mm long, 3 mm diameter (two cylinders)
    "cotter-pin": lambda: _cyl((0, 0, 0), 0.0015, 0.030, axis=0) + _cyl((0.015, 0, 0.0075), 0.0015, 0.015, axis=2),
}


@dataclass
class SynthObject:
    name: str
    prims: list
    vertices: np.ndarray  # (V,3) float64, ~1 mm spacing -- role of mesh.vertices
    scale: float  # AABB diagonal -- role of trimesh ``mesh.scale``

    def sample_surface(self, n, rng):
        areas = np.array([_area(p) for p in self.prims])
        cnt = rng.multinomial(n, areas / areas.sum())
        pts, nrm = [], []
        for p, c in zip(self.prims, cnt):
            if c:
                a, b = _sample_prim(p, int(c), rng)
                pts.append(a)
                nrm.append(b)
        pts, nrm = np.concatenate(pts), np.concatenate(nrm)
        perm = rng.permutation(n)
        return pts[perm], nrm[perm]


def make_object(name: str, vertex_spacing: float = 1e-3, seed: int = 0) -> SynthObject:
    prims = OBJECTS[name]()
    rng = np.random.default_rng(seed)
    area = sum(_area(p) for p in prims)
    V = max(2000, int(area / vertex_spacing**2))
    obj = SynthObject(name, prims, np.zeros((0, 3)), 0.0)
    v, _ = obj.sample_surface(V, rng)
    obj.vertices = v
    obj.scale = float(np.linalg.norm(v.max(0) - v.min(0)))
    return obj


def poses_from_normals(pts, nrm, rng, shear_deg: float = 5.0) -> np.ndarray:
    """sensor poses on the surface: z axis = normal tilted by <= shear, random yaw."""
    n = pts.shape[0]
    # orthonormal frame around the normal
    ref = np.where(np.abs(nrm[:, 2:3]) < 0.9, np.array([[0.0, 0.0, 1.0]]), np.array([[1.0, 0.0, 0.0]]))
    a = np.cross(nrm, ref)
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    b = np.cross(nrm, a)
    cos_s = rng.uniform(math.cos(math.radians(shear_deg)), 1.0, n)
    sin_s = np.sqrt(1 - cos_s**2)
    phi = rng.uniform(0, 2 * math.pi, n)
    z = cos_s[:, None] * nrm + sin_s[:, None] * (np.cos(phi)[:, None] * a + np.sin(phi)[:, None] * b)
    z /= np.linalg.norm(z, axis=1, keepdims=True)
    yaw = rng.uniform(0, 2 * math.pi, n)
    x0 = np.cos(yaw)[:, None] * a + np.sin(yaw)[:, None] * b
    x = x0 - (x0 * z).sum(1, keepdims=True) * z
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    y = np.cross(z, x)
    T = np.zeros((n, 4, 4))
    T[:, :3, 0], T[:, :3, 1], T[:, :3, 2], T[:, :3, 3], T[:, 3, 3] = x, y, z, pts, 1
    return T


@dataclass
class SynthCodebook:
    poses: torch.Tensor  # (M,4,4) float32
    cam_poses: torch.Tensor  # (M,4,4) float32
    embeddings: torch.Tensor  # (M,D) float64 (build_codebook.py:72-74)


def pose_embedding(poses: np.ndarray, D: int, seed: int = 0, len_t: float = 0.01, kappa: float = 2.0) -> np.ndarray:
    """smooth positive embedding of sensor poses: random Fourier features of (t / len_t, kappa * z-axis,
    kappa * x-axis), 1 + cos(.), L2-normalised.  Stands for a trained TCN: codes of nearby / similarly
    oriented touches are similar (cosine ~1 within ~len_t, ~0.67 far away), which is what lets the
    filter's weights concentrate the particles."""
    P = np.asarray(poses, dtype=np.float64).reshape(-1, 4, 4)
    f = np.concatenate([P[:, :3, 3] / len_t, kappa * P[:, :3, 2], kappa * P[:, :3, 0]], axis=1)
    rng = np.random.default_rng(seed + 7000)
    W, b = rng.normal(size=(9, D)), rng.uniform(0, 2 * math.pi, D)
    E = 1.0 + np.cos(f @ W + b)
    return E / np.linalg.norm(E, axis=1, keepdims=True)


def make_codebook(obj: SynthObject, M: int = 50000, D: int = 256, seed: int = 0, cam_dist: float = 0.022,
                  embedding: str = "random") -> SynthCodebook:
    """embedding="random": i.i.d. positive codes (tests / golden vectors); "smooth": ``pose_embedding``."""
    rng = np.random.default_rng(seed + 1000)
    pts, nrm = obj.sample_surface(M, rng)
    T = poses_from_normals(pts, nrm, rng)
    cam = T.copy()
    cam[:, :3, 3] += cam_dist * T[:, :3, 2]
    if embedding == "smooth":
        E = pose_embedding(T, D, seed)
    else:
        E = rng.uniform(0.0, 1.0, (M, D))
        E /= np.linalg.norm(E, axis=1, keepdims=True)
    return SynthCodebook(torch.from_numpy(T).float(), torch.from_numpy(cam).float(), torch.from_numpy(E))


def make_pose_query(pose: torch.Tensor, D: int, seed: int = 0, noise: float = 0.05, frame: int = 0) -> torch.Tensor:
    """tactile code of a touch at ``pose`` under the smooth embedding (+ noise), (1,D) float64."""
    e = torch.from_numpy(pose_embedding(pose.double().numpy(), D, seed))[0]
    g = torch.Generator().manual_seed(seed + 4000 + int(frame))
    q = e + noise * torch.randn(D, generator=g, dtype=torch.float64) / math.sqrt(D)
    return (q / q.norm())[None]


def _noise_tf(n, sig_t, sig_r_deg, rng):
    from scipy.spatial.transform import Rotation as R

    T = np.zeros((n, 4, 4))
    T[:, :3, :3] = R.from_euler("zyx", rng.normal(0, sig_r_deg, (n, 3)), degrees=True).as_matrix()
    T[:, :3, 3] = rng.normal(0, sig_t, (n, 3))
    T[:, 3, 3] = 1
    return T


def make_trajectory(obj: SynthObject, T: int = 100, seed: int = 0, step: float = 2.5e-4, sig_t: float = 5e-4, sig_r: float = 1.0):
    """straight slide of ``step`` m per frame across the largest primitive.
    Returns (gt (T,4,4), meas (T,4,4)) float32; meas = gt @ noise."""
    rng = np.random.default_rng(seed + 2000)
    areas = [_area(p) for p in obj.prims]
    k, d = obj.prims[int(np.argmax(areas))]
    s = (np.arange(T) - T / 2) * step
    if k == "rect":
        axis_dir = d["u"] if d["lu"] >= d["lv"] else d["v"]
        pts = d["o"][None] + s[:, None] * axis_dir[None]
        nrm = np.repeat(d["n"][None], T, 0)
    elif k == "cyl":
        ax = d["axis"]
        i, j = (ax + 1) % 3, (ax + 2) % 3
        phi = s / d["r"]
        pts = np.repeat(d["c"][None], T, 0)
        pts[:, i] += d["r"] * np.cos(phi)
        pts[:, j] += d["r"] * np.sin(phi)
        nrm = np.zeros((T, 3))
        nrm[:, i], nrm[:, j] = np.cos(phi), np.sin(phi)
    else:
        raise ValueError(k)
    # fixed yaw along the path: frame from the first normal
    one = np.random.default_rng(seed + 3000)
    G0 = poses_from_normals(pts[:1], nrm[:1], one, shear_deg=0.0)[0]
    gt = np.zeros((T, 4, 4))
    for t in range(T):
        z = nrm[t]
        x = G0[:3, 0] - (G0[:3, 0] @ z) * z
        x /= np.linalg.norm(x)
        gt[t, :3, 0], gt[t, :3, 1], gt[t, :3, 2], gt[t, :3, 3], gt[t, 3, 3] = x, np.cross(z, x), z, pts[t], 1
    meas = gt @ _noise_tf(T, sig_t, sig_r, rng)
    return torch.from_numpy(gt).float(), torch.from_numpy(meas).float()


def make_query(cb: SynthCodebook, idx: int, noise: float = 0.05, seed: int = 0) -> torch.Tensor:
    """tactile code for a frame: the codebook embedding nearest the true pose plus noise,
    L2-normalised; (1,D) float64 like TCN.cloud_to_tactile_code (tcn.py:140-148)."""
    g = torch.Generator().manual_seed(seed + 4000 + int(idx))
    q = cb.embeddings[idx] + noise * torch.randn(cb.embeddings.shape[1], generator=g, dtype=torch.float64) / math.sqrt(cb.embeddings.shape[1])
    return (q / q.norm())[None]

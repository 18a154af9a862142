"""Data formats either side of the path.

* The reference persists a codebook as a dill pickle of the ``tactile_tree`` nn.Module with its
  pynanoflann tree inside (build_codebook.py:130-137, tactile_tree.py:13-41): unloadable without
  the original environment.  ``convert_pickled_codebook`` reads such a pickle WITHOUT importing
  midastouch / pynanoflann / theseus (only an explicit allow-list of tensor / array rebuild functions is
  resolved; every other class or callable named by the file becomes an inert stand-in that only keeps its state) and writes the plain-tensor format below.
* Plain format (``.npz``): ``poses`` (M,4,4) f32, ``cam_poses`` (M,4,4) f32, ``embeddings`` (M,D)
  f64 (dtype preserved), ``keys`` (M,6) f32 R3_SE3 keys when known, ``format`` = "midas-b200-codebook-1".
* ``extract_poses_sim`` reads ``tactile_data.pkl`` (touch_simulator.py:158-167; pose.py:272-300):
  [x, y, z, qx, qy, qz, qw] rows -> (T,4,4) float32 transforms.
"""
from __future__ import annotations

import io
import pickle

import numpy as np
import torch

from ._lib import MidasError

FORMAT = "midas-b200-codebook-1"


class _Inert:
    """stand-in for a class that cannot be imported: keeps whatever state the pickle carries"""

    def __init__(self, *a, **k):
        self._args = a

    def __setstate__(self, state):
        self.__dict__["_state"] = state
        if isinstance(state, dict):
            self.__dict__.update(state)

    def __reduce_ex__(self, proto):  # never re-pickled
        raise MidasError("inert stand-in objects cannot be pickled")


class _TolerantUnpickler(pickle.Unpickler):
    """Resolves ONLY the callables a tensor / array / ordered-dict pickle needs (an explicit (module, name) allow-list);
    every other global -- the reference's own classes, pynanoflann, theseus, and anything a hostile file might name
    (builtins.eval, os.system, ...) -- becomes an inert stand-in that merely stores its state."""

    ALLOWED = {
        ("collections", "OrderedDict"),
        ("torch._utils", "_rebuild_tensor_v2"), ("torch._utils", "_rebuild_tensor"), ("torch._utils", "_rebuild_parameter"),
        ("torch._utils", "_rebuild_parameter_with_state"),
        ("torch", "FloatStorage"), ("torch", "DoubleStorage"), ("torch", "HalfStorage"), ("torch", "BFloat16Storage"),
        ("torch", "LongStorage"), ("torch", "IntStorage"), ("torch", "ShortStorage"), ("torch", "CharStorage"),
        ("torch", "ByteStorage"), ("torch", "BoolStorage"), ("torch", "Size"), ("torch", "device"), ("torch", "float32"),
        ("torch", "float64"), ("torch", "float16"), ("torch", "int64"), ("torch", "int32"),
        ("torch.storage", "_load_from_bytes"), ("torch.storage", "UntypedStorage"), ("torch.storage", "TypedStorage"),
        ("torch.serialization", "_get_layout"),
        ("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"),
        ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"),
        ("numpy", "ndarray"), ("numpy", "dtype"),
        ("_codecs", "encode"),
    }

    def find_class(self, module, name):
        if (module, name) in self.ALLOWED:
            try:
                return super().find_class(module, name)
            except Exception:  # noqa: BLE001
                pass
        return type(name, (_Inert,), {"__module__": module})


def _state_of(obj) -> dict:
    d = dict(getattr(obj, "__dict__", {}))
    for k in ("_buffers", "_parameters"):
        if isinstance(d.get(k), dict):
            d.update(d[k])
    return d


def read_pickled_codebook(path_or_bytes):
    """-> (poses, cam_poses, embeddings) CPU tensors from a reference ``codebook.pkl``."""
    raw = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else open(path_or_bytes, "rb").read()
    obj = _TolerantUnpickler(io.BytesIO(raw)).load()
    st = _state_of(obj)
    try:
        poses, cam, emb = st["poses"], st["cam_poses"], st["embeddings"]
    except KeyError as e:
        raise MidasError(f"not a tactile_tree pickle: missing attribute {e}") from None
    as_t = lambda x: x.detach().cpu() if torch.is_tensor(x) else torch.as_tensor(np.asarray(x))  # noqa: E731
    return as_t(poses).float(), as_t(cam).float(), as_t(emb)


def save_codebook(path: str, poses, cam_poses, embeddings, keys=None) -> None:
    arrs = dict(poses=np.asarray(poses.cpu(), np.float32), cam_poses=np.asarray(cam_poses.cpu(), np.float32),
                embeddings=np.asarray(embeddings.cpu()), format=np.array(FORMAT))
    if keys is not None:
        arrs["keys"] = np.asarray(keys.cpu(), np.float32)
    np.savez(path, **arrs)


def load_codebook(path: str, device=None):
    """plain-format file -> drop-in ``tactile_tree`` (moved to ``device`` when given)."""
    from .tactile_tree import tactile_tree

    z = np.load(path, allow_pickle=False)
    if str(z["format"]) != FORMAT:
        raise MidasError(f"{path}: unknown codebook format {z['format']!r}")
    cb = tactile_tree(torch.from_numpy(z["poses"]), torch.from_numpy(z["cam_poses"]), torch.from_numpy(z["embeddings"]))
    if device is not None:
        cb.to_device(device)
    return cb


def convert_pickled_codebook(pkl_path: str, out_path: str) -> None:
    save_codebook(out_path, *read_pickled_codebook(pkl_path))


def xyzquat_xyzw_to_tf(rows: np.ndarray) -> torch.Tensor:
    """[x,y,z,qx,qy,qz,qw] -> (T,4,4) float32 (xyzw_to_wxyz + xyzquat_to_tf, pose.py:51-62,78-86): the
    quaternion is normalised first, like the reference."""
    r = np.atleast_2d(np.asarray(rows, np.float64))
    q = r[:, 3:] / np.linalg.norm(r[:, 3:], axis=1, keepdims=True)
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    T = np.zeros((r.shape[0], 4, 4))
    T[:, 0, 0], T[:, 0, 1], T[:, 0, 2] = 1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)
    T[:, 1, 0], T[:, 1, 1], T[:, 1, 2] = 2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)
    T[:, 2, 0], T[:, 2, 1], T[:, 2, 2] = 2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)
    T[:, :3, 3], T[:, 3, 3] = r[:, :3], 1.0
    return torch.from_numpy(T).float()


def extract_poses_sim(pickle_file: str, device="cpu"):
    """pose.py:272-300 -> (gt_p_cam, gt_p, meas_p), each (T,4,4) float32 on ``device``."""
    with open(pickle_file, "rb") as f:
        poses = _TolerantUnpickler(f).load()
    return tuple(xyzquat_xyzw_to_tf(poses[k]).to(device) for k in ("camposes", "gelposes", "gelposes_meas"))

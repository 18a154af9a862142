"""Codebook / trajectory file formats (no GPU): a reference-style dill pickle of the tactile_tree
nn.Module (with a nanoflann tree inside) is read without importing the reference or its
dependencies and converted to the plain-tensor format."""
import pickle
import sys
import types

import numpy as np
import torch

from midastouch_b200 import codebook_io as io_


def _fake_reference_pickle(poses, cam, emb):
    """build the pickle the reference writes (build_codebook.py:130-137) with throw-away modules named
    like the real ones, then remove them so that unpickling cannot import them"""
    mods = {}
    for name in ("midastouch", "midastouch.tactile_tree", "midastouch.tactile_tree.tactile_tree", "pynanoflann"):
        mods[name] = types.ModuleType(name)
        sys.modules[name] = mods[name]

    class KDTree:  # stands for pynanoflann.KDTree
        def __init__(self):
            self.data = np.zeros((3, 6), np.float32)

    KDTree.__module__, KDTree.__qualname__ = "pynanoflann", "KDTree"
    mods["pynanoflann"].KDTree = KDTree

    class tactile_tree(torch.nn.Module):
        def __init__(self, poses, cam_poses, embeddings):
            super().__init__()
            self.poses, self.cam_poses, self.embeddings = poses, cam_poses, embeddings
            self.tree = KDTree()
            self.tree_size = poses.shape[0]

    tactile_tree.__module__, tactile_tree.__qualname__ = "midastouch.tactile_tree.tactile_tree", "tactile_tree"
    mods["midastouch.tactile_tree.tactile_tree"].tactile_tree = tactile_tree
    # importable classes are pickled by reference (module + name), by dill exactly as by pickle
    raw = pickle.dumps(tactile_tree(poses, cam, emb))
    for name in mods:
        del sys.modules[name]
    return raw


def test_convert_reference_pickle(tmp_path):
    g = torch.Generator().manual_seed(0)
    poses = torch.eye(4)[None].repeat(50, 1, 1)
    poses[:, :3, 3] = torch.randn(50, 3, generator=g)
    cam = poses.clone()
    cam[:, 2, 3] += 0.02
    emb = torch.rand(50, 16, dtype=torch.float64, generator=g)
    raw = _fake_reference_pickle(poses, cam, emb)
    assert "midastouch" not in sys.modules and "pynanoflann" not in sys.modules
    p, c, e = io_.read_pickled_codebook(raw)
    assert torch.equal(p, poses) and torch.equal(c, cam) and torch.equal(e, emb) and e.dtype == torch.float64
    pkl = tmp_path / "codebook.pkl"
    pkl.write_bytes(raw)
    out = str(tmp_path / "codebook.npz")
    io_.convert_pickled_codebook(str(pkl), out)
    z = np.load(out)
    assert str(z["format"]) == io_.FORMAT and z["embeddings"].dtype == np.float64
    cb = io_.load_codebook(out)  # CPU: holds the tensors, no tree yet
    assert len(cb) == 50 and torch.equal(cb.get_embeddings(), emb) and torch.equal(cb.get_poses()[0], poses)


def test_extract_poses_sim(tmp_path):
    from scipy.spatial.transform import Rotation as R

    rng = np.random.default_rng(1)
    rows = lambda: np.concatenate([rng.normal(size=(7, 3)), R.random(7, random_state=3).as_quat()], 1)  # noqa: E731
    d = {"camposes": rows(), "gelposes": rows(), "gelposes_meas": rows(), "mNoise": {"sig_r": 1, "sig_t": 5e-4}}
    f = tmp_path / "tactile_data.pkl"
    f.write_bytes(pickle.dumps(d))
    cam, gel, meas = io_.extract_poses_sim(str(f))
    assert cam.shape == (7, 4, 4) and cam.dtype == torch.float32
    assert np.allclose(gel[:, :3, :3].numpy(), R.from_quat(d["gelposes"][:, 3:]).as_matrix(), atol=1e-6)
    assert np.allclose(meas[:, :3, 3].numpy(), d["gelposes_meas"][:, :3], atol=1e-6)


def test_hostile_pickle_executes_nothing(tmp_path):
    """the reader resolves an explicit allow-list only: a pickle that names builtins.eval / os.system gets inert objects"""
    marker = tmp_path / "pwned"

    class Evil:
        def __reduce__(self):
            import os

            return (os.system, (f"touch {marker}",))

    class Evil2:
        def __reduce__(self):
            return (eval, (f"open({str(marker)!r}, 'w').close()",))

    for raw in (pickle.dumps({"poses": Evil()}), pickle.dumps({"poses": Evil2()})):
        obj = io_._TolerantUnpickler(__import__("io").BytesIO(raw)).load()
        assert isinstance(obj["poses"], io_._Inert)
    assert not marker.exists()

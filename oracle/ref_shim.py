"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Import shim that loads the UNMODIFIED reference modules
``/root/reference/midastouch/modules/particle_filter.py`` and ``pose.py`` in this
container, where ``trimesh``, ``theseus`` and ``omegaconf`` are not installed.

The shim places stub modules in ``sys.modules`` for exactly those three imports
(reference import lines: particle_filter.py:10,25,28; pose.py:13) and nothing
else.  The theseus stub implements only ``SO3(tensor=).log_map()`` and
``SO3(tensor=).to_quaternion()`` (call sites pose.py:19-23, pose.py:26-34) by
delegating to ``oracle.oracle`` (the closed-form restatement), so every other
line executed is the reference's own code.

``/root/reference`` does not exist on the GPU box: this file is only used HERE
by ``oracle/gen_golden.py`` to produce ``tests/golden/*.npz`` and by the
``not gpu`` tests that cross-check the restatement when the reference is present.
"""
import importlib.util
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _ref_root() -> str:
    """the reference tree when it is present (build container), else the verbatim copies of its two hot-path
    modules under oracle/_ref (made by oracle/build_ref.py; what travels to the GPU box)"""
    env = os.environ.get("MIDAS_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", os.path.join(_HERE, "_ref")):
        if os.path.isfile(os.path.join(cand, "midastouch/modules/particle_filter.py")):
            return cand
    return "/root/reference"


REF_ROOT = _ref_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "midastouch/modules/particle_filter.py"))


def _install_stubs():
    import torch
    from oracle import oracle as _o

    if "trimesh" not in sys.modules:
        tm = types.ModuleType("trimesh")

        class _Mesh:  # stand-in for trimesh.Trimesh: .vertices, .scale (AABB diagonal)
            def __init__(self, vertices):
                import numpy as np

                self.vertices = np.asarray(vertices, dtype=np.float64)
                ext = self.vertices.max(0) - self.vertices.min(0)
                self.scale = float(np.linalg.norm(ext))

        def load(path):  # path is a .npy of vertices in the synthetic assets
            import numpy as np

            return _Mesh(np.load(path))

        tm.load = load
        tm.Trimesh = _Mesh
        sys.modules["trimesh"] = tm

    if "theseus" not in sys.modules:
        th = types.ModuleType("theseus")

        class SO3:
            def __init__(self, tensor=None):
                self.tensor = tensor

            def log_map(self):
                return _o.so3_log_map(self.tensor)

            def to_quaternion(self):
                return _o.so3_to_quaternion(self.tensor)

        class SE3:  # only constructed on paths the oracle never executes
            def __init__(self, *a, **k):
                raise NotImplementedError("theseus.SE3 is not stubbed")

        th.SO3, th.SE3 = SO3, SE3
        sys.modules["theseus"] = th

    if "omegaconf" not in sys.modules:
        oc = types.ModuleType("omegaconf")

        class DictConfig(dict):
            pass

        oc.DictConfig = DictConfig
        oc.OmegaConf = object
        sys.modules["omegaconf"] = oc


def _load(name, relpath):
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference():
    """Returns (particle_filter_module, pose_module) of the unmodified reference."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    _install_stubs()
    # package skeleton so that `from midastouch.modules.pose import ...` resolves
    for pkg in ("midastouch", "midastouch.modules"):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = []
            sys.modules[pkg] = m
    pose = _load("midastouch.modules.pose", "midastouch/modules/pose.py")
    pf = _load("midastouch.modules.particle_filter", "midastouch/modules/particle_filter.py")
    return pf, pose


def load_reference_tdn():
    """Returns (fcrn_module, tdn_module) of the unmodified reference (contrib/tdn_fcrn/fcrn.py, tdn.py).  fcrn.py needs
    torch only; tdn.py imports the renderer, the visualiser, hydra, PIL and cv2 at module level (tdn.py:10-25) without
    using them in the methods exercised here (blend_heightmaps, image2heightmap, heightmap2mask): those imports are
    satisfied by empty stand-ins when the real packages are absent."""
    if not os.path.isfile(os.path.join(REF_ROOT, "midastouch/contrib/tdn_fcrn/fcrn.py")):
        raise RuntimeError("reference TDN not present under %s" % REF_ROOT)
    _install_stubs()
    for pkg in ("midastouch", "midastouch.modules", "midastouch.contrib", "midastouch.contrib.tdn_fcrn", "midastouch.render",
                "midastouch.viz"):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = []
            sys.modules[pkg] = m

    def stub(name, **attrs):
        if name not in sys.modules:
            m = types.ModuleType(name)
            for k, v in attrs.items():
                setattr(m, k, v)
            sys.modules[name] = m

    stub("midastouch.render.digit_renderer", digit_renderer=object)
    stub("midastouch.viz.visualizer", Viz=object)
    stub("midastouch.modules.misc", view_subplots=lambda *a, **k: None, DIRS={"weights": ""}, get_device=lambda cpu=False: "cpu")
    try:
        import hydra  # noqa: F401
    except ImportError:
        stub("hydra", main=lambda **kw: (lambda f: f))
    try:
        import PIL  # noqa: F401
    except ImportError:
        stub("PIL", Image=object)
    try:
        import cv2  # noqa: F401
    except ImportError:
        stub("cv2")
    _load("midastouch.modules.pose", "midastouch/modules/pose.py")
    fcrn = _load("midastouch.contrib.tdn_fcrn.fcrn", "midastouch/contrib/tdn_fcrn/fcrn.py")
    tdn = _load("midastouch.contrib.tdn_fcrn.tdn", "midastouch/contrib/tdn_fcrn/tdn.py")
    return fcrn, tdn


class Cfg(dict):
    """attribute-access dict standing in for omegaconf.DictConfig"""

    def __getattr__(self, k):
        v = self[k]
        return Cfg(v) if isinstance(v, dict) else v


def default_cfg(num_particles=1024, noise_r=0.5, noise_t=2e-4, pen_max=0.002):
    # values: config/expt/ycb.yaml:18-24, config/tdn/default.yaml:16-18
    return Cfg(
        expt=dict(params=dict(num_particles=num_particles,
                              noise_r=dict(sim=noise_r, real=noise_r),
                              noise_t=dict(sim=noise_t, real=noise_t),
                              noise_ratio=1.0)),
        tdn=dict(render=dict(pen=dict(min=0.0005, max=pen_max))),
    )

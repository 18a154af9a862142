"""Minimal hydra-style config composition for the reference's config surface (hydra / omegaconf are
not dependencies): ``compose()`` reads ``config.yaml`` (``defaults: [- group: option]``), loads
``<group>/<option>.yaml`` under the group key and applies dotlist overrides exactly like the
reference's command line (README.md:103-106: ``expt=mcmaster expt.params.num_particles=1000``).
Attribute and item access both work (``cfg.expt.params.noise_r.sim`` / ``cfg["expt"]``)."""
from __future__ import annotations

import os

import yaml

CONFIG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config")

# PyYAML follows YAML 1.1 where "2e-4" (no dot) is a string; hydra/omegaconf read it as a float
class _loader(yaml.SafeLoader):  # a subclass: the process-wide yaml.SafeLoader keeps its own resolvers
    pass


_loader.add_implicit_resolver(
    "tag:yaml.org,2002:float",
    __import__("re").compile(r"^[-+]?(\d+\.?\d*|\.\d+)([eE][-+]?\d+)$"), list("-+0123456789."))


class Config(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k) from None

    def __setattr__(self, k, v):
        self[k] = v


def _wrap(x):
    if isinstance(x, dict):
        return Config({k: _wrap(v) for k, v in x.items()})
    if isinstance(x, list):
        return [_wrap(v) for v in x]
    return x


def _load(path):
    with open(path) as f:
        return yaml.load(f, Loader=_loader) or {}


def _parse_value(s: str):
    return yaml.load(s, Loader=_loader)


def compose(config_dir: str = CONFIG_DIR, overrides=(), config_name: str = "config") -> Config:
    root = _load(os.path.join(config_dir, config_name + ".yaml"))
    groups = {}
    for d in root.pop("defaults", []):
        (g, opt), = d.items()
        groups[g] = opt
    dot = []
    for o in overrides:
        k, _, v = o.partition("=")
        if not _:
            raise ValueError(f"override {o!r} is not key=value")
        if k in groups and "." not in k:
            groups[k] = v  # group selection: expt=mcmaster
        else:
            dot.append((k, v))
    cfg = dict(root)
    for g, opt in groups.items():
        path = os.path.join(config_dir, g, f"{opt}.yaml")
        if not os.path.exists(path):
            raise FileNotFoundError(f"no option {opt!r} in config group {g!r} ({path})")
        cfg[g] = _load(path)
    for k, v in dot:
        node = cfg
        parts = k.lstrip("+").split(".")
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = _parse_value(v)
    return _wrap(cfg)

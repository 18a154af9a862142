"""hydra-style composition of the reference's config surface (no GPU)."""
import pytest

from midastouch_b200.config import compose


def test_defaults_match_reference_values():
    cfg = compose()
    assert cfg.expt.obj_model == "004_sugar_box" and cfg.expt.params.num_particles == 50000
    assert cfg.expt.params.noise_r.sim == 0.5 and cfg.expt.params.noise_t.sim == pytest.approx(2e-4)
    assert isinstance(cfg.expt.params.noise_t.sim, float)
    assert cfg.tdn.render.pen.max == 0.002 and cfg.tcn.model.feature_size == 256 and cfg.tcn.model.planes == "32,64,64"
    assert cfg.expt.max_length == "None"  # the reference's YAML spells None as a string too
    assert cfg["expt"]["codebook_size"] == 50000


def test_group_and_dotlist_overrides():
    cfg = compose(overrides=["expt=mcmaster", "expt.params.num_particles=1000", "expt.log_id=3", "expt.params.resample=low_var"])
    assert cfg.expt.obj_model == "cotter-pin" and cfg.expt.params.num_particles == 1000 and cfg.expt.log_id == 3
    assert cfg.expt.params.noise_t == pytest.approx(1e-4)  # scalar in mcmaster.yaml
    assert cfg.expt.params.resample == "low_var"
    with pytest.raises(FileNotFoundError):
        compose(overrides=["expt=nope"])
    with pytest.raises(ValueError):
        compose(overrides=["expt.params"])


def test_particle_filter_accepts_both_noise_forms():
    """mcmaster.yaml gives scalar noise where particle_filter.py:114-121 reads .sim/.real"""
    import numpy as np

    from midastouch_b200.particle_filter import particle_filter

    v = np.random.default_rng(0).uniform(-0.05, 0.05, (500, 3))
    a = particle_filter(compose(), v)
    b = particle_filter(compose(overrides=["expt=mcmaster"]), v)
    assert a.motion_noise == {"mu": 0, "sig_r": 0.5, "sig_t": 2e-4} and b.motion_noise["sig_t"] == pytest.approx(1e-4)
    assert a.pen_max == 0.002

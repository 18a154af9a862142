"""-m gpu: the reference's OWN loop statements (midastouch/filter/filter.py, "# motion model" ... "particles =
pf.resampler(particles)", lines 150-190) executed verbatim -- the text is read from the reference tree (or its verbatim
copy under oracle/_ref, made by oracle/build_ref.py) and exec'd -- with `pf`, `codebook`, `particle_rmse` bound to the
drop-in classes.  Iteration 0 included: init_filter + SE3_NN snap + cluster_particles (count % 50 == 0)."""
import os
import textwrap

import numpy as np
import pytest
import torch

from midastouch_b200 import synth
from oracle import oracle as O
from oracle import ref_shim

pytestmark = pytest.mark.gpu


def _loop_statements():
    path = os.path.join(ref_shim.REF_ROOT, "midastouch/filter/filter.py")
    if not os.path.isfile(path):
        pytest.skip("reference filter.py not present (oracle/_ref not built)")
    lines = open(path).read().splitlines()
    a = next(i for i, l in enumerate(lines) if l.strip() == "# motion model")
    b = next(i for i, l in enumerate(lines) if l.strip() == "particles = pf.resampler(particles)")
    assert 140 < a < 160 and 180 < b < 200, (a, b)
    return textwrap.dedent("\n".join(lines[a:b + 1]))


class _Bar:
    def set_description(self, *_):
        pass


def test_reference_loop_statements_run_on_the_dropin():
    import time

    from midastouch_b200.config import compose
    from midastouch_b200.particle_filter import Particles, particle_filter, particle_rmse
    from midastouch_b200.tactile_tree import tactile_tree

    dev = torch.device("cuda:0")
    block = compile(_loop_statements(), "filter.py:150-190", "exec")
    box = synth.make_object("004_sugar_box")
    cbs = synth.make_codebook(box, M=20000, D=64, seed=4, embedding="smooth")
    codebook = tactile_tree(cbs.poses, cbs.cam_poses, cbs.embeddings)
    codebook.to_device(dev)
    cfg = compose(overrides=["expt.params.num_particles=8000"])
    pf = particle_filter(cfg, box.vertices, downsample=1)
    gt, meas = synth.make_trajectory(box, T=64, seed=4)
    torch.manual_seed(0)
    ns = dict(pf=pf, codebook=codebook, particle_rmse=particle_rmse, torch=torch, time=time, timer={}, get_time=lambda t0: time.time() - t0,
              pbar=_Bar(), gt_p=gt.to(dev), meas_p=meas.to(dev), init_particles=8000, prev_idx=0, count=0,
              filter_stats={"rmse_t": [], "rmse_r": []}, particles=None)
    n_hist, labels_checked = [], 0
    for idx in range(1, 64):
        ns["idx"] = idx
        ns["tactile_code"] = synth.make_pose_query(gt[idx], 64, seed=4, frame=idx).to(dev)
        ns["start_time"] = time.time()
        before_cluster = ns["count"] % 50 == 0
        exec(block, ns)
        parts = ns["particles"]
        assert isinstance(parts, Particles) and parts.poses.is_cuda
        n_hist.append(len(parts))
        if before_cluster:
            assert parts.labels.dtype == torch.int64  # torch.tensor(clustering.labels_) (particle_filter.py:225-227)
            labels_checked += 1
        assert ns["cluster_poses"].shape[1:] == (4, 4) and ns["cluster_stds"].shape[1] == 3
        ns["prev_idx"] = idx
        ns["count"] += 1
    assert labels_checked == 2  # iterations 0 and 50
    rm = ns["filter_stats"]["rmse_t"]
    assert rm[0] > 0.03 and rm[-1] < 0.012, (rm[0], rm[-1])
    # annealing shrank the particle set like the reference does (floor 1000, particle_filter.py:405-447)
    assert min(n_hist) >= 1000 and n_hist[-1] < 8000


def _blobs(seed, n, centres, sig, noise_frac):
    g = torch.Generator().manual_seed(seed)
    k = torch.randint(0, len(centres), (n,), generator=g)
    t = torch.tensor(centres, dtype=torch.float32)[k] + sig * torch.randn(n, 3, generator=g)
    m = torch.rand(n, generator=g) < noise_frac
    t[m] = 0.3 * torch.rand(int(m.sum()), 3, generator=g)
    P = torch.eye(4).repeat(n, 1, 1)
    P[:, :3, 3] = t
    return P


@pytest.mark.parametrize("case", ["one", "two", "chain", "noise", "ties"])
def test_dbscan_vs_sklearn(case):
    """cluster_particles == sklearn.cluster.DBSCAN(eps=1e-2, min_samples=N/5) label for label (numbering, borders, noise)"""
    from midastouch_b200.config import compose
    from midastouch_b200.particle_filter import Particles, particle_filter

    dev = torch.device("cuda:0")
    box = synth.make_object("004_sugar_box")
    pf = particle_filter(compose(), box.vertices)
    if case == "one":
        P = _blobs(1, 3000, [[0.1, 0.1, 0.1]], 3e-3, 0.2)
    elif case == "two":  # two clusters whose border points compete + the later-indexed cluster holds the lowest core index
        P = _blobs(2, 5000, [[0.1, 0.1, 0.1], [0.1, 0.1, 0.118]], 2.5e-3, 0.05)
    elif case == "chain":  # an elongated cluster: label propagation needs several hops
        g = torch.Generator().manual_seed(3)
        n = 4000
        P = torch.eye(4).repeat(n, 1, 1)
        P[:, 0, 3] = 0.05 * torch.rand(n, generator=g)
        P[:, 1:3, 3] = 5e-4 * torch.randn(n, 2, generator=g)
    elif case == "noise":  # spread cloud: nobody has N/5 neighbours -> all -1 (what iteration 0 of the loop sees)
        P = _blobs(4, 2000, [[0.1, 0.1, 0.1]], 5e-2, 0.0)
    else:  # duplicated points and points at distance exactly eps on a lattice
        g = torch.Generator().manual_seed(5)
        lat = torch.stack(torch.meshgrid(torch.arange(8.), torch.arange(8.), torch.arange(8.), indexing="ij"), -1).reshape(-1, 3) * 5e-3
        P = torch.eye(4).repeat(2 * lat.shape[0], 1, 1)
        P[:, :3, 3] = torch.cat([lat, lat[torch.randperm(lat.shape[0], generator=g)]])
    want = O.dbscan_labels(P, eps=1e-2)
    got = pf.cluster_particles(Particles(P.to(dev)), eps=1e-2).labels
    assert got.dtype == torch.int64
    assert torch.equal(got.cpu(), want), (int((got.cpu() != want).sum()), want.unique(), got.unique())
    if case in ("one", "two", "chain"):
        assert int(want.max()) >= 0
    if case == "noise":
        assert int(want.max()) == -1


def test_cluster_centers_logmap_vs_oracle():
    from midastouch_b200.config import compose
    from midastouch_b200.particle_filter import Particles, particle_filter

    dev = torch.device("cuda:0")
    box = synth.make_object("004_sugar_box")
    pf = particle_filter(compose(), box.vertices)
    cbs = synth.make_codebook(box, M=3000, D=8, seed=9)
    g = torch.Generator().manual_seed(2)
    centres = cbs.poses[[10, 500, 2000]]
    n = 6000
    lab = torch.randint(0, 3, (n,), generator=g)
    poses = centres[lab] @ O.noisy_odom(torch.eye(4), 1e-3 * torch.randn(n, 3, generator=g), 3.0 * torch.randn(n, 3, generator=g))
    w = torch.rand(n, dtype=torch.float64, generator=g) + 0.1
    labels = (lab - 1).float()
    cp, cs = pf.get_cluster_centers(Particles(poses.to(dev), w.to(dev), labels.to(dev)))  # default method = "logmap"
    rp, rs = O.cluster_centers_logmap(poses, w, labels)
    assert torch.allclose(cp.cpu(), rp, rtol=1e-4, atol=3e-6), (cp.cpu() - rp).abs().max()
    assert torch.allclose(cs.cpu(), rs, rtol=1e-3, atol=1e-7), (cs.cpu() - rs).abs().max()
    RtR = cp[:, :3, :3].transpose(1, 2) @ cp[:, :3, :3]
    assert float((RtR.cpu() - torch.eye(3)).abs().max()) < 1e-5
    # more than 16 labels (the kernels reduce 16 clusters per pass)
    lab20 = torch.randint(0, 20, (n,), generator=g).float()
    cp20, _ = pf.get_cluster_centers(Particles(poses.to(dev), w.to(dev), lab20.to(dev)), method="quat_avg")
    rp20, _ = O.cluster_centers(poses, w, lab20)
    assert cp20.shape == (20, 4, 4) and torch.allclose(cp20.cpu(), rp20, rtol=1e-4, atol=3e-6)

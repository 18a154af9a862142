"""CPU: the oracle restatement (oracle/oracle.py) against the golden vectors produced by
the unmodified reference module (oracle/gen_golden.py), plus pins of the third-party
restatements against scipy / brute force."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from midastouch_b200 import synth


def T(a):
    return torch.from_numpy(np.asarray(a))


def test_euler_zyx(golden):
    g = golden("euler_zyx")
    assert torch.equal(O.euler_zyx_matrix(T(g["rot_deg"])), T(g["Rn"]))


def test_motion_model_matches_reference(golden):
    g = golden("motion")
    torch.manual_seed(int(g["seed"]))
    tn, rot = O.draw_motion_noise(g["poses"].shape[0], float(g["sig_t"]), float(g["sig_r"]))
    assert torch.equal(tn, T(g["tn"])) and torch.equal(rot, T(g["rot_deg"]))  # RNG contract
    moved, keep = O.motion_model(T(g["poses"]), T(g["odom"]), tn, rot)
    assert keep.all()
    assert torch.equal(moved, T(g["moved"]))


def test_similarity_matches_reference(golden):
    g = golden("similarity")
    obj = synth.make_object("004_sugar_box")
    cb = synth.make_codebook(obj, M=4096, D=256, seed=0)
    assert np.array_equal(cb.embeddings[:64].numpy(), g["emb_head"])
    q, sel = T(g["q"]), T(g["sel"])
    assert torch.equal(O.get_similarity(q, cb.embeddings[sel], True), T(g["w_soft"]))
    assert torch.equal(O.get_similarity(q, cb.embeddings[sel], False), T(g["w_raw"]))
    assert torch.equal(O.get_similarity(q, cb.embeddings[sel][:1].repeat(16, 1), True), T(g["w_const"]))
    heat = O.codebook_similarity(q, cb.embeddings)
    assert torch.equal(heat, T(g["heat"]))
    # table-then-gather == gather-then-dot (what the engine exploits)
    assert torch.equal(heat[sel], T(g["w_raw"]))


@pytest.mark.parametrize("name", ["soft", "raw", "masked", "peaked"])
@pytest.mark.parametrize("seed", [3, 4])
def test_low_var_matches_reference_loop(golden, name, seed):
    g = golden("resample_low_var")
    w, u = T(g[f"{name}_{seed}_w"]), float(g[f"{name}_{seed}_u"][0])
    torch.manual_seed(seed)
    assert float(torch.rand(1)[0]) == u
    anc = O.low_var_indices(w, u)
    filled = T(g[f"{name}_{seed}_filled"])
    assert torch.equal(anc >= 0, filled)
    ref = T(g[f"{name}_{seed}_anc"])
    assert torch.equal(anc[filled], ref[filled])
    assert torch.equal(O.low_var_indices_loop(w, u), anc)
    p, _, _ = O.resample_gather(T(g["in_poses"]), w, torch.zeros(len(w)), anc)
    assert torch.equal(p[:8], T(g[f"{name}_{seed}_poses0"]))
    # "low_var_batch" (particle_filter.py:263-287) is NOT the same draw: it repeats the
    # first forward difference (line 279) instead of the first count, so it returns
    # cnt[1]-cnt[0] copies of particle 0 and can change N.  Recorded in the golden file,
    # not reproduced by the engine (DESIGN.md "deviations").
    batch = T(g[f"{name}_{seed}_anc_batch"])
    assert (batch[1:] >= batch[:-1]).all()


def test_low_var_wraparound_edge():
    # u = largest float32 below 1 at a power-of-two N: fl32(u/N) + (N-1)/N rounds to 1.0 in
    # the reference's float64 sum for some N -> remainder gives 0 for the last slot.
    w = torch.rand(64, dtype=torch.float64) + 0.1
    for u in (0.0, float(np.nextafter(np.float32(1), np.float32(0)))):
        assert torch.equal(O.low_var_indices(w, u), O.low_var_indices_loop(w, u))


def test_rmse_matches_reference(golden):
    g = golden("rmse")
    rt, rr = O.particle_rmse(T(g["poses"]), T(g["gt"]))
    assert torch.equal(rt, T(g["rmse_t"])) and torch.equal(rr, T(g["rmse_r"]))


def test_prune_matches_reference(golden):
    g = golden("prune")
    w, drifted = O.remove_invalid(T(g["poses"]), T(g["w_in"]), g["vertices_ds"], float(g["pen_max"]))
    assert torch.equal(w, T(g["w_out"])) and drifted == bool(g["drifted"])
    assert 0 < int((w == 0).sum()) < len(w)


def test_annealing_plan_matches_reference(golden):
    g = golden("annealing")
    n = len(g["w_in"])
    kind, k = O.annealing_plan(n, 0.8e-3 / 1e-3, 100, n)
    assert kind == "remove" and n - k == int(g["remove_n"])
    keep = torch.topk(T(g["w_in"]), k, largest=False).indices
    mask = torch.ones(n, dtype=torch.bool)
    mask[keep] = False
    assert torch.equal(T(g["w_in"])[mask], T(g["remove_w"]))
    kind, k = O.annealing_plan(n, float(torch.tensor(1.2e-3) / torch.tensor(1e-3)), 100, 2 * n)
    assert kind == "add" and n + k == int(g["add_n"])


# ---------------- third-party restatements pinned against scipy / brute force
def _random_rotations(n, seed=0):
    from scipy.spatial.transform import Rotation as R

    return R.random(n, random_state=seed)


def test_so3_log_map_vs_scipy():
    r = _random_rotations(4000)
    Rm = torch.from_numpy(r.as_matrix()).float()
    lv = O.so3_log_map(Rm).double().numpy()
    assert np.abs(lv - r.as_rotvec()).max() < 2e-5
    # near identity and near pi
    from scipy.spatial.transform import Rotation as R

    ax = np.random.default_rng(1).normal(size=(200, 3))
    ax /= np.linalg.norm(ax, axis=1, keepdims=True)
    for ang, tol in ((1e-3, 1e-6), (4e-3, 1e-6), (math_pi() - 1e-3, 2e-3), (math_pi() - 0.05, 1e-4)):
        rr = R.from_rotvec(ax * ang)
        lv = O.so3_log_map(torch.from_numpy(rr.as_matrix()).float()).double().numpy()
        ref = rr.as_rotvec()
        err = np.minimum(np.abs(lv - ref).max(1), np.abs(lv + ref).max(1))  # +-pi axis ambiguity
        assert err.max() < tol, (ang, err.max())


def math_pi():
    import math

    return math.pi


def test_so3_to_quaternion_vs_scipy():
    r = _random_rotations(4000, seed=2)
    q = O.so3_to_quaternion(torch.from_numpy(r.as_matrix()).float()).double().numpy()
    ref = r.as_quat()[:, [3, 0, 1, 2]]
    ref = ref * np.sign(ref[:, :1])
    q = q * np.sign(q[:, :1])
    assert np.abs(q - ref).max() < 5e-4 and np.median(np.abs(q - ref)) < 1e-7  # f32, ill-conditioned near pi
    assert np.abs(np.linalg.norm(q, axis=1) - 1).max() < 5e-4


def test_nn_exact_equals_brute():
    rng = np.random.default_rng(0)
    keys = rng.normal(size=(3000, 6)).astype(np.float32) * 0.05
    qs = (keys[rng.integers(0, 3000, 2000)] + rng.normal(size=(2000, 6)).astype(np.float32) * 0.003).astype(np.float32)
    assert np.array_equal(O.nn_brute(keys, qs), O.nn_exact(keys, qs))
    # duplicate keys: ties resolve to the lowest index
    keys2 = np.concatenate([keys, keys[:100]])
    a = O.nn_brute(keys2, keys[:100])
    assert np.array_equal(a, np.arange(100)) and np.array_equal(O.nn_exact(keys2, keys[:100]), a)


def test_quat_average_identity_cluster():
    obj = synth.make_object("004_sugar_box")
    cb = synth.make_codebook(obj, M=64, D=8)
    P = cb.poses[:1].repeat(10, 1, 1)
    c = O.quat_average(P, torch.ones(10))
    assert torch.allclose(c, cb.poses[0], atol=1e-5)


def test_filter_step_table_equals_gather_dot():
    obj = synth.make_object("004_sugar_box")
    cb = synth.make_codebook(obj, M=2048, D=64)
    keys = O.r3_se3(cb.poses)
    g = torch.Generator().manual_seed(0)
    poses = cb.poses[torch.randint(0, 2048, (512,), generator=g)]
    gt, meas = synth.make_trajectory(obj, T=4)
    odom = torch.inverse(meas[0]) @ meas[1]
    torch.manual_seed(5)
    tn, rot = O.draw_motion_noise(512, 2e-4, 0.5)
    q = synth.make_query(cb, 3)
    a = O.filter_step(poses, odom, tn, rot, keys, cb.embeddings, q, 0.37, gather_dot=True)
    b = O.filter_step(poses, odom, tn, rot, keys, cb.embeddings, q, 0.37, gather_dot=False)
    assert torch.equal(a["nn_idx"], b["nn_idx"]) and torch.equal(a["anc"], b["anc"])
    assert torch.allclose(a["weights"], b["weights"], rtol=1e-12)
    assert (a["anc"][1:] >= a["anc"][:-1]).all()


def test_fma_f32_is_correctly_rounded():
    """oracle.fma_f32 (the float32 fma behind l2_sq_f32) against exact rational arithmetic, including
    crafted cases where the float64 intermediate lands exactly on a float32 midpoint (double rounding)."""
    from fractions import Fraction

    def exact(a, b, c):
        v = Fraction(float(a)) * Fraction(float(b)) + Fraction(float(c))
        # round-to-nearest-even to float32 via two candidate neighbours
        f = np.float32(float(v))  # float(v) is correctly rounded to f64; may double-round -> check neighbours
        cands = {f, np.nextafter(f, np.float32(np.inf)), np.nextafter(f, np.float32(-np.inf))}
        best = min(cands, key=lambda x: (abs(Fraction(float(x)) - v), int(np.float32(x).view(np.uint32)) & 1))
        return np.float32(best)

    rng = np.random.default_rng(5)
    a = rng.standard_normal(4000).astype(np.float32)
    b = rng.standard_normal(4000).astype(np.float32)
    c = (rng.standard_normal(4000) * 10.0 ** rng.integers(-6, 3, 4000)).astype(np.float32)
    # midpoint traps: c = 1, a*b = 2^-24 + tiny  (1 + 2^-24 is the midpoint between 1 and 1 + 2^-23)
    ta = np.array([2.0 ** -12, 2.0 ** -12, 2.0 ** -12 * (1 + 2.0 ** -23), 2.0 ** -12 * (1 - 2.0 ** -23)], dtype=np.float32)
    tb = np.array([2.0 ** -12, 2.0 ** -12 * (1 + 2.0 ** -23), 2.0 ** -12 * (1 + 2.0 ** -23), 2.0 ** -12], dtype=np.float32)
    tc = np.array([1.0, 1.0, 1.0, 1.0], dtype=np.float32)
    a, b, c = np.concatenate([a, ta]), np.concatenate([b, tb]), np.concatenate([c, tc])
    got = O.fma_f32(a, b, c)
    want = np.array([exact(x, y, z) for x, y, z in zip(a, b, c)], dtype=np.float32)
    assert np.array_equal(got, want)
    assert got[-3] == np.float32(1.0 + 2.0 ** -23) and got[-4] == np.float32(1.0)  # tie-to-even vs residual above

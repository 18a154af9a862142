"""world_size-2 gloo test (CPU) of the sharded resampling plumbing: each rank holds half of the
weights, the ranks all-gather their float64 weight sums (what FilterEngine._allgather_sums does
over NCCL), and every rank derives the systematic slots it owns from its own parents with the
arithmetic of k_step_b (mt_count_below, compiled for the host).  The union over ranks must be the
single-process ancestor vector of the oracle."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

WORKER = r'''
import ctypes, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
H = ctypes.CDLL({so!r})
H.h_shard_children.restype = ctypes.c_longlong
N, u = {N}, {u}
g = torch.Generator().manual_seed(11)
w_all = torch.rand(N, dtype=torch.float64, generator=g) + 0.2
w_all[::7] = 0.0                                    # pruned particles never get children
lo, hi = rank * N // world, (rank + 1) * N // world
w = w_all[lo:hi].contiguous().numpy()
local = torch.tensor([w.sum()], dtype=torch.float64)          # kernel A's local weight sum
sums = torch.zeros(world, dtype=torch.float64)
dist.all_gather_into_tensor(sums, local)                       # the step's only exchange: 8 B per rank
S, A = 0.0, 0.0
for r in range(world):                                         # sequential, identical on every rank (k_step_b)
    if r == rank:
        A = S
    S += float(sums[r])
anc = np.full(N, -1, np.int64)
base = ctypes.c_longlong()
nchild = H.h_shard_children(w.ctypes.data_as(ctypes.c_void_p), ctypes.c_longlong(hi - lo), ctypes.c_double(A), ctypes.c_double(S),
                            ctypes.c_longlong(N), ctypes.c_float(u), anc.ctypes.data_as(ctypes.c_void_p), ctypes.byref(base))
out = torch.full((N,), -1, dtype=torch.int64)
out[base.value: base.value + nchild] = torch.from_numpy(anc[:nchild]) + lo     # global parent index per owned slot
owned = torch.zeros(N, dtype=torch.int64)
owned[base.value: base.value + nchild] = 1
dist.all_reduce(owned)                                          # every slot owned exactly once
gathered = [torch.empty_like(out) for _ in range(world)]
dist.all_gather(gathered, out)
if rank == 0:
    merged = torch.stack(gathered).max(0).values
    torch.save(dict(merged=merged, owned=owned, w=w_all, children=[int((g_ >= 0).sum()) for g_ in gathered]), {out!r})
dist.destroy_process_group()
'''


@pytest.mark.parametrize("N,u", [(4096, 0.37), (100003, 0.9999)])
def test_two_rank_slot_ownership_matches_oracle(tmp_path, N, u):
    from oracle import oracle as O

    so = os.path.join(HERE, "_build", "host_math.so")
    src = os.path.join(HERE, "host_math_harness.cpp")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-x", "c++", "-shared", "-fPIC", "-o", so, src])
    out = str(tmp_path / "res.pt")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, so=so, N=N, u=u, out=out))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(29500 + (N % 200)), WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r))) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=120) == 0
    res = torch.load(out)
    assert bool((res["owned"] == 1).all())
    ref = O.low_var_indices(res["w"], u)
    filled = ref >= 0
    diff = int((res["merged"][filled] != ref[filled]).sum())
    assert diff <= 1, diff  # a slot may move only across a 1-ulp CDF boundary
    assert bool((res["w"][res["merged"][filled]] > 0).all())
    assert sum(res["children"]) == int(filled.sum()) or sum(res["children"]) == N
    assert abs(res["children"][0] - N / 2) < 0.05 * N  # near-flat weights: children stay balanced


def test_rebalance_plan_preserves_order_and_evens_out():
    from midastouch_b200.engine import rebalance_plan

    for counts in ([10, 0, 5, 9], [1050774, 987000, 990000, 972226], [3], [0, 7], [5, 5, 5, 5]):
        G, total = len(counts), sum(counts)
        plans = [rebalance_plan(counts, r) for r in range(G)]
        targets = plans[0][2]
        assert sum(targets) == total and max(targets) - min(targets) <= 1
        for r in range(G):
            send, recv, _ = plans[r]
            assert sum(send) == counts[r] and sum(recv) == targets[r]
            for s_ in range(G):
                assert send[s_] == plans[s_][1][r]  # what r sends to s is what s expects from r
        # simulate the all-to-all on global indices: order must be preserved
        glob = iter(range(total))
        held = [[next(glob) for _ in range(c)] for c in counts]
        new = [[] for _ in range(G)]
        for r in range(G):
            off = 0
            for s_ in range(G):
                new[s_].append((r, held[r][off:off + plans[r][0][s_]]))
                off += plans[r][0][s_]
        flat = [x for s_ in range(G) for _, chunk in sorted(new[s_]) for x in chunk]
        assert flat == list(range(total))

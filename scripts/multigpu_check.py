"""torchrun --nproc-per-node 2 scripts/multigpu_check.py : sharded engine against the single-GPU engine.

rank r holds half of the particles; after one teacher-forced step (same noise, same offset u) the
concatenation of the shards' children must equal the single-GPU result exactly (poses, matches), and
rebalance() must even the shards out without changing the global particle sequence."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from midastouch_b200 import synth
from midastouch_b200.engine import FilterEngine
from midastouch_b200.tactile_tree import tactile_tree

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
box = synth.make_object("004_sugar_box")
cbs = synth.make_codebook(box, M=20000, D=64, seed=4, embedding="smooth")
cb = tactile_tree(cbs.poses, cbs.cam_poses, cbs.embeddings); cb.to_device(dev)
gt, meas = synth.make_trajectory(box, T=8, seed=4)
N = 40000
g = torch.Generator().manual_seed(0)
sel = torch.randint(0, 20000, (N,), generator=g)
poses = cbs.poses[sel]
tn = 2e-4 * torch.randn(N, 3, generator=g); rot = 0.5 * torch.randn(N, 3, generator=g)
q = synth.make_pose_query(gt[1], 64, seed=4, frame=1)
odom = torch.inverse(meas[0]) @ meas[1]
lo, hi = rank * N // world, (rank + 1) * N // world


def gather_state(e):
    n_loc = e.count()
    mine = torch.cat([e.poses().reshape(n_loc, 16), e.nn_idx().float().reshape(n_loc, 1)], 1)
    cnt = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(cnt, torch.tensor([n_loc], device=dev))
    parts = [torch.zeros((int(c), 17), device=dev) for c in cnt.tolist()]
    dist.all_gather(parts, mine)
    return torch.cat(parts), cnt.tolist()


ok = True
# single-GPU engine (every rank runs its own copy), stepped with the same noise and offsets
ref = FilterEngine(cb, capacity=N, mesh_vertices=box.vertices)
ref.load_particles(poses.to(dev))
STEPS = 4
noise = [(2e-4 * torch.randn(N, 3, generator=g), 0.5 * torch.randn(N, 3, generator=g)) for _ in range(STEPS)]
us = [0.37, 0.11, 0.93, 0.5]
want = []
for k in range(STEPS):
    ref.step(q, odom, u=us[k], tn=noise[k][0].to(dev), rot=noise[k][1].to(dev))
    want.append(torch.cat([ref.poses().reshape(N, 16), ref.nn_idx().float().reshape(N, 1)], 1))
full = None
for mode in ("peer", "allgather"):
    eng = FilterEngine(cb, capacity=int(1.5 * (hi - lo)), rank=rank, world=world, n_global=N, mesh_vertices=box.vertices)
    if mode == "peer" and not eng.peer_exchange:
        if rank == 0:
            print("peer exchange unavailable on this box")
        ok = False
    if mode == "allgather":
        eng.peer_exchange = False
    eng.rebalance_every = 0
    eng.load_particles(poses[lo:hi].to(dev))
    counts = [hi_ - lo_ for lo_, hi_ in [(r * N // world, (r + 1) * N // world) for r in range(world)]]
    for k in range(STEPS):
        off = sum(counts[:rank])
        n_loc = counts[rank]
        eng.step(q, odom, u=us[k], tn=noise[k][0][off:off + n_loc].to(dev), rot=noise[k][1][off:off + n_loc].to(dev))
        full, counts = gather_state(eng)
        same = full.shape == want[k].shape and torch.equal(full, want[k])
        st = eng.ctx.stats()
        if rank == 0:
            print(f"[{mode}] step {k}: children per rank {counts} | sharded == single GPU: {same} | overflow flag {st['overflow']}")
        ok &= bool(same) and st['overflow'] == 0
# skew the shards artificially, then rebalance
eng.rebalance()
n2 = eng.count()
mine2 = torch.cat([eng.poses().reshape(n2, 16), eng.nn_idx().float().reshape(n2, 1)], 1)
cnt2 = torch.zeros(world, dtype=torch.int64, device=dev)
dist.all_gather_into_tensor(cnt2, torch.tensor([n2], device=dev))
parts2 = [torch.zeros((int(c), 17), device=dev) for c in cnt2.tolist()]
dist.all_gather(parts2, mine2)
if rank == 0:
    even = max(cnt2.tolist()) - min(cnt2.tolist()) <= 1
    same2 = torch.equal(torch.cat(parts2), full)
    print("after rebalance", cnt2.tolist(), "| even:", even, "| sequence preserved:", same2)
    ok &= even and bool(same2)
# and the engine keeps stepping after a rebalance
eng.step(q, odom, u=0.11)
torch.cuda.synchronize()
flag = torch.tensor([1 if ok else 0], device=dev)
dist.broadcast(flag, 0)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) else 1)

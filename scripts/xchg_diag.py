"""torchrun --nproc-per-node N scripts/xchg_diag.py : where the time of the in-kernel weight-sum exchange goes.
Per step and rank: ns from the end of the local phase to 'sums sent', and from 'sent' to 'all sums received'
(the latter = NVLink latency + however much later the slowest peer reached its exchange)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
from midastouch_b200 import synth
from midastouch_b200._lib import call
from midastouch_b200.engine import FilterEngine, prepare_odom
from midastouch_b200.tactile_tree import tactile_tree

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
obj, cbs, gt, meas = bench.make_assets()
cb = tactile_tree(cbs.poses, cbs.cam_poses, cbs.embeddings); cb.to_device(dev)
n = bench.N_PER_GPU
eng = FilterEngine(cb, capacity=int(1.125 * n), seed=1, rank=rank, world=world, n_global=n * world, mesh_vertices=obj.vertices)
g = torch.Generator().manual_seed(100 + rank)
sel = torch.randint(0, bench.M, (n,), generator=g)
eng.load_particles(cbs.poses.to(dev)[sel.to(dev)], nn_hint=sel.int().to(dev), spatial_sort=True)
odoms = [prepare_odom(torch.inverse(meas[t - 1]) @ meas[t]) for t in range(1, bench.T_TRAJ)]
codes = [synth.make_pose_query(gt[t + 1], bench.D, seed=3, frame=t).to(dev) for t in range(bench.T_TRAJ - 1)]
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
us = torch.rand(64, generator=torch.Generator().manual_seed(7)).tolist()
send, wait, tot = [], [], []
MODE = os.environ.get("XCHG_SYNC", "0") == "1"   # 1: barrier before every step (removes the skew between ranks)
for t in range(40):
    flush.zero_(); flush.sum()
    if MODE:
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.step(codes[t], odoms[t], u=us[t])
    e1.record()
    torch.cuda.synchronize()
    o = (C.c_ulonglong * 3)()
    call("mt_dist_debug", eng.ctx.h, o)
    if t >= 5:
        send.append((o[1] - o[0]) / 1e3), wait.append((o[2] - o[1]) / 1e3), tot.append(1e3 * e0.elapsed_time(e1))
for r in range(world):
    if r == rank:
        f = lambda a: " ".join("%5.1f" % x for x in a[:18])
        print(f"rank {rank}: step us   {f(tot)}\n         send us  {f(send)}\n         wait us  {f(wait)}   mean wait {sum(wait)/len(wait):.1f}", flush=True)
    dist.barrier()
dist.destroy_process_group()

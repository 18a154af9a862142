"""torchrun --nproc-per-node N scripts/rebalance_probe.py : cost of FilterEngine.rebalance() on the bench workload
(per call, host wall clock around a device synchronise), and of the step right after it."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
from midastouch_b200 import synth
from midastouch_b200.engine import FilterEngine, prepare_odom
from midastouch_b200.tactile_tree import tactile_tree

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
obj, cbs, gt, meas = bench.make_assets()
cb = tactile_tree(cbs.poses, cbs.cam_poses, cbs.embeddings)
cb.to_device(dev)
n = bench.N_PER_GPU
eng = FilterEngine(cb, capacity=n + n // 2, sig_t=2e-4, sig_r=0.5, seed=1234, rank=rank, world=world, n_global=n * world,
                   mesh_vertices=obj.vertices, pen_max=0.002)
eng.use_graph = True
eng.rebalance_every = 0  # called by hand below
sel = torch.randint(0, bench.M, (n,), generator=torch.Generator().manual_seed(100 + rank))
eng.load_particles(cbs.poses.to(dev)[sel.to(dev)], nn_hint=sel.int().to(dev), spatial_sort=True)
odoms = [prepare_odom(torch.inverse(meas[t - 1]) @ meas[t]) for t in range(1, bench.T_TRAJ)]
codes = [synth.make_pose_query(gt[t + 1], bench.D, seed=3, frame=t).to(dev) for t in range(bench.T_TRAJ - 1)]
us = torch.rand(4096, generator=torch.Generator().manual_seed(7)).tolist()
rows = []
for t in range(200):
    eng.step(codes[t], odoms[t], u=us[t])
    if (t + 1) % 32 == 0:
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        eng.rebalance()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        eng.step(codes[t + 1], odoms[t + 1], u=us[t + 1]); torch.cuda.synchronize()
        t2 = time.perf_counter()
        eng.step(codes[t + 1], odoms[t + 1], u=us[t + 1]); torch.cuda.synchronize()
        t3 = time.perf_counter()
        rows.append({"after_step": t + 1, "n_local": eng.n, "rebalance_ms": round((t1 - t0) * 1e3, 3), "next_step_ms": round((t2 - t1) * 1e3, 3),
                     "step_after_ms": round((t3 - t2) * 1e3, 3)})
if rank == 0:
    print(json.dumps(rows))
dist.destroy_process_group()

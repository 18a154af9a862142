"""Timeline of one filter step inside its CUDA-graph replay (needs a library built with -DMT_TRACE=1, passed as
MIDAS_B200_LIB=...): earliest block start / latest block end of every step kernel and the phase boundaries of
k_step_bw, from %globaltimer stamps (mt_trace_read), averaged over steps of the bench workload."""
import ctypes as C, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from midastouch_b200 import synth
from midastouch_b200._lib import call
from midastouch_b200.engine import FilterEngine, prepare_odom
from midastouch_b200.tactile_tree import tactile_tree

dev = torch.device("cuda:0")
obj, cbs, gt, meas = bench.make_assets()
cb = tactile_tree(cbs.poses, cbs.cam_poses, cbs.embeddings)
cb.to_device(dev)
n = int(os.environ.get("AB_N", bench.N_PER_GPU))
eng = FilterEngine(cb, capacity=n, sig_t=2e-4, sig_r=0.5, seed=1234, mesh_vertices=obj.vertices, pen_max=0.002)
g = torch.Generator().manual_seed(100)
sel = torch.randint(0, bench.M, (n,), generator=g)
eng.use_graph = not bool(os.environ.get("AB_STREAM"))
eng.load_particles(cbs.poses.to(dev)[sel.to(dev)], nn_hint=sel.int().to(dev), spatial_sort=True)
odoms = [prepare_odom(torch.inverse(meas[t - 1]) @ meas[t]) for t in range(1, bench.T_TRAJ)]
codes = [synth.make_pose_query(gt[t + 1], bench.D, seed=3, frame=t).to(dev) for t in range(bench.T_TRAJ - 1)]
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
us = torch.rand(4096, generator=torch.Generator().manual_seed(7)).tolist()
T = int(os.environ.get("AB_STEPS", 60))
buf = (C.c_ulonglong * 64)()
names = ["a", "meshq", "meshq2", "nnq", "bw"]
acc = {}
cnt = 0
for t in range(T):
    if not os.environ.get("AB_NOFLUSH"):
        flush.zero_(); flush.sum()
    call("mt_trace_read", eng.ctx.h, buf, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.step(codes[t], odoms[t], u=us[t])
    e1.record()
    call("mt_trace_read", eng.ctx.h, buf, 0)
    if buf[63] != 1:
        print(json.dumps({"error": "library was not built with -DMT_TRACE=1"})); sys.exit(0)
    if t < int(os.environ.get("AB_SKIP", 10)):
        continue
    t0 = buf[8]
    row = {"event_step": e0.elapsed_time(e1) * 1e3}
    for k, nm in enumerate(names):
        s, e = buf[8 + 2 * k], buf[9 + 2 * k]
        if s == 2**64 - 1 or e == 0:
            continue
        row[nm + "_start"] = (s - t0) / 1e3
        row[nm + "_end"] = (e - t0) / 1e3
    for w, nm in ((24, "bw_phase1_done_max"), (25, "bw_barrier_exit_min"), (26, "bw_barrier_exit_max"), (27, "bw_prefix_done_max"), (28, "bw_first_chunk_done_max")):
        if buf[w] not in (0, 2**64 - 1):
            row[nm] = (buf[w] - t0) / 1e3
    if buf[32]:
        row["nnq_entries"] = float(buf[32])
        row["nnq_entry_us_max"] = buf[30] / 1e3
        row["nnq_entry_us_avg"] = buf[31] / buf[32] / 1e3
        row["nnq_entry_load_us_max"] = buf[33] / 1e3
        row["nnq_entry_load_us_avg"] = buf[34] / buf[32] / 1e3
        row["nnq_last_entry_start"] = (buf[35] - t0) / 1e3
    for k, v in row.items():
        acc[k] = acc.get(k, 0.0) + v
    cnt += 1
st = eng.ctx.stats(reset=True)
acc["leaves_per_search"] = cnt * st["grid_rows"] / max(1, st["nn_fallbacks"])
acc["leaves_max_one_search"] = cnt * st["grid_rows_max"]
acc["searches_per_step"] = cnt * st["nn_fallbacks"] / T
print(json.dumps({"lib": os.path.basename(os.environ.get("MIDAS_B200_LIB", "default")), "graph": eng.use_graph, "steps": cnt,
                  "us_from_k_step_a_start": {k: round(v / cnt, 2) for k, v in acc.items()}}))

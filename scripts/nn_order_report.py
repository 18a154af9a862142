"""CPU only.  How many SE3_NN indices of the bench workload depend on the float32 accumulation order of the 6-D squared
distance?  The CUDA kernels (and oracle.l2_sq_f32) use two interleaved fma chains; nanoflann accumulates left to right.
Queries: 1e6 particles on codebook poses of the drill, moved by `STEPS` filter steps of odometry + motion noise
(no resampling), i.e. poses between codebook keys like the bench's.  Writes profiles/r02_nn_order.json."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from oracle import oracle as O

obj, cbs, gt, meas = bench.make_assets()
keys = O.r3_se3(cbs.poses).numpy()
N, STEPS = 1_000_000, 3
g = torch.Generator().manual_seed(100)
sel = torch.randint(0, bench.M, (N,), generator=g)
poses = cbs.poses[sel].clone()
torch.manual_seed(5)
out = {"workload": f"{bench.OBJ}, M={bench.M} keys, N={N} particles, {STEPS} motion steps from codebook poses", "per_step": []}
for t in range(STEPS):
    odom = torch.inverse(meas[t]) @ meas[t + 1]
    tn, rot = O.draw_motion_noise(N, 2e-4, 0.5)
    poses, _ = O.motion_model(poses, odom, tn, rot)
    q = O.r3_se3(poses).numpy()
    diff, near, a, b = O.nn_order_sensitivity(keys, q)
    out["per_step"].append({"step": t + 1, "indices_that_differ": diff, "queries_with_a_float32_near_tie (relative gap <= 4e-7)": near, "queries": N})
    print(out["per_step"][-1], flush=True)
out["reading"] = ("indices_that_differ = particles whose nearest key changes when the squared distance is accumulated left to right in "
                  "float32 without fma (nanoflann) instead of in the kernels' order; both are exact nearest neighbours up to float32 rounding of the distance")
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_nn_order.json"), "w"), indent=1)

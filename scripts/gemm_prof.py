import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from midastouch_b200.tactile_tree import tactile_tree
dev = torch.device("cuda:0")
M, D = 50000, 256
g = torch.Generator().manual_seed(0)
poses = torch.eye(4)[None].repeat(M, 1, 1); poses[:, :3, 3] = torch.rand(M, 3, generator=g) * 0.1
emb = torch.rand(M, D, dtype=torch.float64, generator=g)
x = torch.rand(8192, 8192, device=dev)
for _ in range(50):
    y = x @ x  # wake the clocks up
torch.cuda.synchronize()
for dt in (torch.float64, torch.float32, torch.float64):
    cb = tactile_tree(poses, poses, emb.to(dt)); cb.to_device(dev)
    for nq in (128, 1024, 4096):
        Q = torch.rand(nq, D, generator=g).to(dev)
        out = cb.query_batched(Q)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            out = cb.query_batched(Q)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        fl = 2.0 * M * D * nq
        print(f"emb {dt} nq={nq}: {ms*1e3:.0f} us  {fl/ms/1e9:.1f} TFLOP/s algorithmic ({3*fl/ms/1e9:.1f} tensor)  out {nq*M*4/ms/1e6:.0f} GB/s")

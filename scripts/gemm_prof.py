"""batched codebook query (tcgen05 + TMA): correctness spot check + timing at the bench shape (M=50k, D=256, nq=1024)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from midastouch_b200.tactile_tree import tactile_tree
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(5)
M, D, nq = 50000, int(os.environ.get("GEMM_D", 256)), int(os.environ.get("GEMM_NQ", 1024))
poses = torch.eye(4)[None].repeat(M, 1, 1)
poses[:, :3, 3] = torch.rand(M, 3, generator=g) * 0.05
for dt in (torch.float64, torch.float32):
    emb = torch.rand(M, D, dtype=torch.float64, generator=g).to(dt)
    cb = tactile_tree(poses, poses, emb)
    cb.to_device(dev)
    Q = torch.rand(nq, D, generator=g).to(dev)
    out = cb.query_batched(Q)
    ref = torch.nn.functional.cosine_similarity(Q[:64].double()[:, None, :], emb.to(dev).double()[None, :4096], dim=2)
    err = float(((out[:64, :4096].double() - ref).abs() / ref.abs()).max())
    for _ in range(3):
        cb.query_batched(Q)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        cb.query_batched(Q)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"{dt}: max rel err {err:.2e}  {1e3*ms:.1f} us per call (incl. query split)  {3*2.0*M*D*nq/ms/1e9:.0f} TFLOP/s of TF32 tensor work")

import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from midastouch_b200.tactile_tree import tactile_tree
dev = torch.device("cuda:0")
M, D, nq = 50000, 256, 1024
g = torch.Generator().manual_seed(0)
poses = torch.eye(4)[None].repeat(M, 1, 1); poses[:, :3, 3] = torch.rand(M, 3, generator=g) * 0.1
emb = torch.rand(M, D, generator=g)
cb = tactile_tree(poses, poses, emb); cb.to_device(dev)
Q = torch.rand(nq, D, generator=g).to(dev)
for _ in range(3):
    out = cb.query_batched(Q)
torch.cuda.synchronize()

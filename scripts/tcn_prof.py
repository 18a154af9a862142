import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
print("tcn ms", bench.tcn_time(torch.device("cuda:0"), reps=3))

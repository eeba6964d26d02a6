"""Kernel list of one training step (run under ncu --metrics gpu__time_duration.sum; development aid)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from idelucs_b200.seqset import SeqSet
from idelucs_b200.train import ShardedTrainer

dev = torch.device("cuda")
nt, Lt = 50000, 2000
a = torch.randint(0, 4, (nt * Lt,), device=dev, dtype=torch.uint8)
a.mul_(2).add_(65).add_((a >= 69).to(torch.uint8) * 2).add_((a >= 73).to(torch.uint8) * 11)
ss = SeqSet.from_ascii(a, np.arange(nt + 1, dtype=np.int64) * Lt, device=dev)
tr = ShardedTrainer(ss, k=6, n_clusters=5, n_mimics=50, batch_sz=512, seed=7)
for _ in range(5):
    tr.step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()

"""Soak test of the graph-replayed training step (development aid): 30 000 steps on C4-shaped synthetic data, loss trace, finiteness of
the parameters and the optimiser state."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from idelucs_b200.seqset import SeqSet
from idelucs_b200.train import ShardedTrainer
dev = torch.device("cuda")
nt, Lt = 100000, 2000
g = torch.Generator(device=dev).manual_seed(0)
codes = torch.randint(0, 4, (nt * Lt,), device=dev, dtype=torch.uint8, generator=g)
a = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)[codes.long()]
ss = SeqSet.from_ascii(a, np.arange(nt + 1, dtype=np.int64) * Lt, device=dev)
tr = ShardedTrainer(ss, k=6, n_clusters=5, n_mimics=50, batch_sz=512, seed=7)
print("graph:", tr.enable_cuda_graph())
losses = []
for i in range(30000):
    l = tr.step()
    if i % 3000 == 0:
        losses.append(float(l.item()))
torch.cuda.synchronize()
print("losses", losses, "epoch", tr.epoch)
print("params finite", bool(torch.isfinite(tr._flat_param).all()), "sq finite", bool(torch.isfinite(tr._sq).all()), "step_no", int(tr._step_no.item()))

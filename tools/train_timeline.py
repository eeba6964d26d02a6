"""Kernel timeline of graph-replayed training steps (torch.profiler / CUPTI): every kernel of ONE replay in start order with its
stream, start offset and duration — shows what really overlaps (development aid)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from idelucs_b200.seqset import SeqSet
from idelucs_b200.train import ShardedTrainer

from idelucs_b200 import parallel
rank, local, world = parallel.init_from_env()   # torchrun: one rank per GPU (rank 0 prints)
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
nt, Lt = 200000, 2000
g = torch.Generator(device=dev).manual_seed(rank)
codes = torch.randint(0, 4, (nt * Lt,), device=dev, dtype=torch.uint8, generator=g)
a = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)[codes.long()]
del codes
ss = SeqSet.from_ascii(a, np.arange(nt + 1, dtype=np.int64) * Lt, device=dev)
C = int(sys.argv[1]) if len(sys.argv) > 1 else 5
tr = ShardedTrainer(ss, k=6, n_clusters=C, n_mimics=50, batch_sz=512, seed=7, world=world, seq_id0=rank * nt)
ok = tr.enable_cuda_graph()
if rank == 0:
    print("graph:", ok, "mode:", tr._mode)
for _ in range(20):
    tr.step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(4):
        tr.step()
    torch.cuda.synchronize()
if rank != 0:
    sys.exit(0)
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
# split into replays: a step starts with the featurise / zero kernels; use the gaps
per = len(evs) // 4
one = evs[2 * per:3 * per]
t0 = one[0].time_range.start
print("%d device events per step; step span %.1f us" % (per, one[-1].time_range.end - t0))
busy = 0.0
for e in one:
    print("%8.1f  %7.2f  %s" % (e.time_range.start - t0, e.time_range.end - e.time_range.start, e.name[:110]))
    busy += e.time_range.end - e.time_range.start
print("sum of durations %.1f us" % busy)

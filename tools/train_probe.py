"""Where does a training step go?  (development aid)"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from idelucs_b200.seqset import SeqSet
from idelucs_b200.train import ShardedTrainer
from idelucs_b200.LossFunctions import IID_loss, info_nce_loss

dev = torch.device("cuda")
nt, Lt = 200000, 2000
a = torch.randint(0, 4, (nt * Lt,), device=dev, dtype=torch.uint8)
a.mul_(2).add_(65).add_((a >= 69).to(torch.uint8) * 2).add_((a >= 73).to(torch.uint8) * 11)
ss = SeqSet.from_ascii(a, np.arange(nt + 1, dtype=np.int64) * Lt, device=dev)
tr = ShardedTrainer(ss, k=6, n_clusters=5, n_mimics=50, batch_sz=512, seed=7)
for _ in range(20):
    tr.step()
torch.cuda.synchronize()

def timed(fn, n=100):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): r = fn()
    e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    return (t1 - t0) / n * 1e3, e0.elapsed_time(e1) / n, (t2 - t0) / n * 1e3

print("full step: cpu-issue %.3f ms, gpu %.3f ms, wall %.3f ms" % timed(tr.step))
ids = torch.randint(0, tr.loader.n_pairs, (512,), device=dev)
print("batch (featurise sel mode): cpu %.3f gpu %.3f wall %.3f" % timed(lambda: tr.loader.batch(ids)))
b = tr.loader.batch(ids)
def fwd():
    return tr.net(b["true"]), tr.net(b["modified"])
print("2 forwards: cpu %.3f gpu %.3f wall %.3f" % timed(fwd))
def fb():
    tr.opt.zero_grad(set_to_none=True)
    z1, h1 = tr.net(b["true"]); z2, h2 = tr.net(b["modified"])
    loss = 0.75 * info_nce_loss(h1, h2, 0.85) + 0.25 * IID_loss(z1, z2, lamb=2.8)
    loss.backward(); tr.opt.step(); return loss
print("fwd+loss+bwd+opt: cpu %.3f gpu %.3f wall %.3f" % timed(fb))
z1 = torch.softmax(torch.randn(512, 5, device=dev), 1).requires_grad_(True); z2 = torch.softmax(torch.randn(512, 5, device=dev), 1).requires_grad_(True)
print("IID_loss fwd+bwd C=5: cpu %.3f gpu %.3f wall %.3f" % timed(lambda: IID_loss(z1, z2, 2.8).backward()))
z1 = torch.softmax(torch.randn(512, 200, device=dev), 1).requires_grad_(True); z2 = torch.softmax(torch.randn(512, 200, device=dev), 1).requires_grad_(True)
print("IID_loss fwd+bwd C=200: cpu %.3f gpu %.3f wall %.3f" % timed(lambda: IID_loss(z1, z2, 2.8).backward()))
h1 = torch.randn(512, 64, device=dev, requires_grad=True); h2 = torch.randn(512, 64, device=dev, requires_grad=True)
print("info_nce fwd+bwd: cpu %.3f gpu %.3f wall %.3f" % timed(lambda: info_nce_loss(h1, h2, 0.85).backward()))
# graph-captured step
ok = tr.enable_cuda_graph()
print("graph enabled:", ok, getattr(tr, "_graph_error", ""))
if ok:
    print("graph step: cpu %.3f gpu %.3f wall %.3f" % timed(tr.step, 300))
    print("loss", float(tr.step()))

"""Dump the event timeline of CTA 0 of the producer/consumer kernel (IDL_PC_TRACE=1; development aid)."""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from idelucs_b200 import featurise as ft
from idelucs_b200.seqset import SeqSet

n, L, k, F = 4000, 10000, 6, 4096
g = torch.Generator(device="cuda").manual_seed(0)
codes = torch.randint(0, 4, (n * L,), device="cuda", dtype=torch.uint8, generator=g)
ascii_t = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device="cuda")[codes.long()]
ss = SeqSet.from_ascii(ascii_t, np.arange(n + 1, dtype=np.int64) * L)
variants = ft.mimic_schedule(50)
V = len(variants)
out = torch.empty((V, n, F), dtype=torch.float32, device="cuda")
off = [v * n * F for v in range(V)]
mean = torch.rand(F, device="cuda") * 1e-3
scale = torch.rand(F, device="cuda") * 1e-4 + 1e-5
fn = lambda: ft.profiles(ss, k, variants, out_kind=ft.OUT_STD_F32, seed=1, out=out, out_off=off, out_stride=F, mean=mean, scale=scale)
fn(); torch.cuda.synchronize()
ws = list(ft._workspaces.values())[0]
tr = ws[-(128 + 8 * 8192):-128].view(torch.int64)
tr.zero_()
fn(); torch.cuda.synchronize()
t = tr.cpu().numpy()
t = t[t != 0]
ev = t & 0xFF; gj = (t >> 8) & 0xFFFF; clk = t >> 24
order = np.argsort(clk, kind="stable")
c0 = clk[order[0]]
names = {1: "S.waitfull", 2: "S.issue", 3: "S.readdone", 10: "B.waitfree", 11: "B.freeok", 12: "B.built", 20: "F.start", 21: "F.waitfree", 24: "F.waitbuilt", 22: "F.go", 23: "F.full"}
for i in order[:int(sys.argv[1]) if len(sys.argv) > 1 else 600]:
    print("%8d  job %4d  %s" % (clk[i] - c0, gj[i], names.get(int(ev[i]), str(ev[i]))))

"""idl_pack (ASCII -> 2-bit codes + reset mask + validation) on 1 GB of device-resident ASCII: time and bytes / s (development aid)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from idelucs_b200.seqset import SeqSet

for L in (10000, 10001, 2000, 1409):
    n = 1_000_000_000 // L
    g = torch.Generator(device="cuda").manual_seed(0)
    codes = torch.randint(0, 4, (n * L,), device="cuda", dtype=torch.uint8, generator=g)
    asc = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device="cuda")[codes.long()]
    del codes
    off = np.arange(n + 1, dtype=np.int64) * L
    f = lambda: SeqSet.from_ascii(asc, off, validate=False)
    f(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print("L = %5d, %7d sequences: from_ascii (offsets H2D + idl_pack) best %.3f ms = %.2f TB/s of ASCII" % (L, n, min(ts), n * L / min(ts) / 1e9))
    del asc

"""Timing of the prepare pass (idl_profiles_prepare) with and without its statistics part (development aid)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from idelucs_b200 import featurise as ft
from idelucs_b200.seqset import SeqSet


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
    g = torch.Generator(device="cuda").manual_seed(0)
    codes = torch.randint(0, 4, (n * L,), device="cuda", dtype=torch.uint8, generator=g)
    ascii_t = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device="cuda")[codes.long()]
    del codes
    ss = SeqSet.from_ascii(ascii_t, np.arange(n + 1, dtype=np.int64) * L)
    variants = ft.mimic_schedule(50)
    for name, vs, stats in (("3 dense slots + stats", variants, True), ("3 dense slots, no stats", variants, False),
                            ("slot 0 only + stats", variants[:1], True), ("clean + stats", [ft.VariantSpec(ft.KIND_CLEAN)], True)):
        f = lambda: ft.prepare(ss, 6, vs, seed=1, want_stats=stats)
        f(); torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); f(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        print("%-26s best %.3f ms per %d sequences x %d bases" % (name, min(ts), n, L), flush=True)


main()

"""Timing of the scaler-statistics pass (idl_profile_stats) only (development aid)."""
import sys, os
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
    for name, spec, env in (("both fast", variants[0], {}), ("both generic", variants[0], {"IDL_NO_FAST_STATS": "1"}),
                            ("clean fast", ft.VariantSpec(ft.KIND_CLEAN), {}),
                            ("both fast, no S1/S2", variants[0], {"IDL_PC_DBG": "256"}), ("both fast, no S2", variants[0], {"IDL_PC_DBG": "2048"}),
                            ("both fast, no count", variants[0], {"IDL_PC_DBG": "512"}), ("both fast, no fold", variants[0], {"IDL_PC_DBG": "1024"}),
                            ("both fast, only S1", variants[0], {"IDL_PC_DBG": "3584"}), ("nothing", variants[0], {"IDL_PC_DBG": "3840"})):
        os.environ.update(env)
        fs = lambda: ft.profile_stats(ss, 6, spec, seed=1)
        fs(); torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fs(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        for k in env:
            os.environ.pop(k)
        print("%-24s best %.3f ms per %d sequences" % (name, min(ts), n), flush=True)
main()

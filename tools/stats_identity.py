"""sha256 of the scaler statistics the prepare pass produces (development aid: run under two builds of the library,
IDELUCS_B200_LIB, to check that a rewritten statistics kernel leaves every bit in place) + timing of the pieces."""
import hashlib, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from idelucs_b200 import featurise as ft
from idelucs_b200.seqset import SeqSet


def sha(t):
    return hashlib.sha256(t.cpu().numpy().tobytes()).hexdigest()[:16]


def main():
    rng = np.random.default_rng(33)
    alph = np.frombuffer(b"ACGTACGTACGTACGTACGTACGTACGTN", dtype=np.uint8)
    lens = rng.integers(1, 12000, size=3000)
    lens[[0, 1, 2, 3, 4, 450, 899]] = [0, 3, 5, 6, 16, 16385, 70001]
    seqs = [alph[rng.integers(0, alph.size, size=int(L))].tobytes() for L in lens]
    ss = SeqSet.from_sequences(seqs)
    variants = ft.mimic_schedule(50)
    sc = ft.prepare(ss, 6, variants, seed=5).scaler()
    print("ragged  ", sha(sc.mean64), sha(sc.var64), sha(sc.scale64), sha(sc.mean32), sha(sc.scale32))
    n, L = 100000, 10000
    g = torch.Generator(device="cuda").manual_seed(0)
    codes = torch.randint(0, 4, (n * L,), device="cuda", dtype=torch.uint8, generator=g)
    ascii_t = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device="cuda")[codes.long()]
    del codes
    ss = SeqSet.from_ascii(ascii_t, np.arange(n + 1, dtype=np.int64) * L)
    pr = ft.prepare(ss, 6, variants, seed=1)
    sc = pr.scaler()
    print("100k    ", sha(sc.mean64), sha(sc.var64), sha(sc.scale64), sha(sc.mean32), sha(sc.scale32))
    gen = torch.Generator(device="cuda").manual_seed(7)
    for rows, F in ((40000, 4096), (3000, 528), (700, 136), (5, 256), (70000, 2080)):
        x = torch.randn(rows, F, device="cuda", generator=gen) * 1e-4 + 2e-4
        x[:, 3] = 0.5
        sc = ft.Scaler.fit(x)
        ref_mean = x.double().mean(0)
        print("matrix %6d x %4d" % (rows, F), sha(sc.mean64), sha(sc.var64), sha(sc.scale64), "max rel err of the mean %.2e" % float(((sc.mean64 - ref_mean).abs() / ref_mean.abs()).max()))

    def timeit(fn, reps=5):
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return min(ts)
    print("prepare + colstats16 (100k x 10kb): %.3f ms" % timeit(lambda: ft.prepare(ss, 6, variants, seed=1)))
    print("prepare without statistics:         %.3f ms" % timeit(lambda: ft.prepare(ss, 6, variants, seed=1, want_stats=False)))
    print("finalize (%d parts):                %.3f ms" % (pr.parts.shape[0], timeit(lambda: pr.scaler())))


main()

"""smoke(): one small invocation of the hot path on cuda:0, checked against the oracle.
(The oracle import is allowed here: __graft_entry__.smoke() is one of the three places that
may use oracle/ as a checker.)"""
import numpy as np
import torch


def run():
    import idelucs_oracle as orc  # checker only
    from . import featurise as ft
    from .seqset import SeqSet
    from .LossFunctions import IID_loss

    assert torch.cuda.is_available(), "smoke() needs a CUDA device"
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(0)
    alph = np.frombuffer(b"ACGTACGTACGTACGTN", dtype=np.uint8)
    seqs = [alph[rng.integers(0, alph.size, size=L)].tobytes() for L in (1500, 777, 64, 5, 2049, 1000)]
    ss = SeqSet.from_sequences(seqs, device=dev)
    k, seed = 6, 1234
    variants = ft.mimic_schedule(5)
    counts = ft.profiles(ss, k, variants, out_kind=ft.OUT_COUNTS_I32, seed=seed).cpu().numpy()
    want = orc.rng_mimic_counts([bytearray(s) for s in seqs], k, seed, [v.kind for v in variants])
    assert np.array_equal(counts, want), "mimic k-mer counts differ from the oracle"
    freq = ft.profiles(ss, k, variants, out_kind=ft.OUT_FREQ_F32, seed=seed)
    want32 = ((want + 1) / (want + 1).sum(axis=2, keepdims=True)).astype(np.float32)
    assert np.array_equal(freq.cpu().numpy(), want32), "frequencies differ from the oracle"
    sc = ft.Scaler.fit(freq[0])
    std = ft.profiles(ss, k, variants, out_kind=ft.OUT_STD_F32, seed=seed, mean=sc.mean32, scale=sc.scale32)
    m, v, s = orc.standard_scaler_fit(want32[0])
    want_std = orc.standard_scaler_transform(want32.reshape(-1, 4 ** k), m, s).reshape(want32.shape)
    np.testing.assert_allclose(std.cpu().numpy(), want_std, rtol=1e-6, atol=1e-6)
    # the whole-schedule fast path (k = 6, >= 512 items, 51 variants, standardised: what bench.py times), a sample
    # of its rows against the oracle: counts -> float32(count / total) -> (x - mean32) / scale32
    from . import utils as U
    big = [alph[rng.integers(0, alph.size, size=int(L))].tobytes() for L in rng.integers(1500, 3000, size=640)]
    sb = SeqSet.from_sequences(big, device=dev)
    sched = ft.mimic_schedule(50)
    xs, scb, _ = U.augment_device(sb, 50, k=k, seed=seed, seq_id0=9)
    pick = [0, 1, 77, 300, 511, 512, 639]
    # (rng_mimic_counts numbers its sequences seq_id0 + position: one call per item with the item's true id)
    wc = np.stack([orc.rng_mimic_counts([bytearray(big[i])], k, seed, [v.kind for v in sched], seq_id0=9 + i)[:, 0] for i in pick], axis=1)
    w32 = ((wc + 1) / (wc + 1).sum(axis=2, keepdims=True)).astype(np.float32)
    wstd = ((w32 - scb.mean32.cpu().numpy()) / scb.scale32.cpu().numpy()).astype(np.float32)
    assert np.array_equal(xs[:, pick].cpu().numpy(), wstd), "standardised schedule rows differ from the oracle"
    # loss forward/backward
    z1 = torch.softmax(torch.randn(96, 7, device=dev), 1).requires_grad_(True)
    z2 = torch.softmax(torch.randn(96, 7, device=dev), 1).requires_grad_(True)
    loss = IID_loss(z1, z2, lamb=2.8)
    loss.backward()
    wl, wd1, wd2 = orc.IID_loss_grad(z1.detach().cpu().numpy(), z2.detach().cpu().numpy(), lamb=2.8)
    assert abs(loss.item() - wl) < 1e-5, (loss.item(), wl)
    np.testing.assert_allclose(z1.grad.cpu().numpy(), wd1, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(z2.grad.cpu().numpy(), wd2, rtol=1e-4, atol=1e-6)
    torch.cuda.synchronize()
    print("smoke ok: %d sequences x %d variants + %d sequences x %d variants (fast path), k=%d; IID loss %.6f"
          % (len(seqs), len(variants), len(big), len(sched), k, loss.item()))

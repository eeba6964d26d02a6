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
    print("smoke ok: %d sequences x %d variants, k=%d; IID loss %.6f" % (len(seqs), len(variants), k, loss.item()))

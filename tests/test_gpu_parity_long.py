"""GPU parity tests, part 2 (-m gpu): the paths round 1 left unpinned —

* sequences longer than 65 535 bases (the generic kernel's int32 ``long_path`` and the chunked
  long-genome path), every variant kind, against the oracle's mutate-and-recount
  (idelucs/kmers.pyx:38-50 on the mutated bytes, idelucs/utils.py:330-366 for the float outputs);
* the dominant kernel's STANDARDISED output at the BASELINE configs[2] shape (10 kb sequences,
  k = 6, 51 variants), every row of a 512-sequence slab against the oracle;
* mutation rates far above the reference's on long sequences (an on-chip edit list must never
  silently drop mutations).
"""
import numpy as np
import pytest

import idelucs_oracle as orc

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ft():
    from idelucs_b200 import featurise
    return featurise


@pytest.fixture(scope="module")
def SeqSet():
    from idelucs_b200.seqset import SeqSet
    return SeqSet


def _rand_seq(rng, L, n_rate=0.002):
    a = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=L)].copy()
    if n_rate > 0 and L > 0:
        a[rng.random(L) < n_rate] = ord("N")
        for _ in range(min(3, L // 1000)):          # a few N runs (scaffold gaps)
            p = int(rng.integers(0, L))
            a[p:p + int(rng.integers(1, 200))] = ord("N")
    return a.tobytes()


def _oracle_counts(seqs, k, seed, variants, seq_id0=0, explicit=None):
    """[V, N, 4^k] int32: apply the variant's edits (rng mode: the oracle's specification of the counter-based RNG;
    explicit: the given lists) to a copy of the bytes and recount from scratch."""
    out = np.zeros((len(variants), len(seqs), 4 ** k), np.int32)
    for i, s in enumerate(seqs):
        codes = orc.codes_of_seq(s)
        for v, spec in enumerate(variants):
            if spec.kind == 5:
                pos, val = explicit[spec.explicit_idx][i]
                edits = list(zip(np.asarray(pos).tolist(), np.asarray(val).tolist()))
            else:
                rid = spec.rng_id if spec.rng_id is not None else v
                edits = orc.rng_variant_edits(seed, seq_id0 + i, rid, spec.kind, codes, len(s), spec.p1, spec.p2, spec.n_bp)
            mut = bytearray(s)
            for p, val in edits:
                mut[p] = b"ACGTN"[val]
            orc.kmer_counts(mut, k, out[v, i])
    return out


def _check_all_outputs(ft, ss, seqs, k, variants, seed, seq_id0, lists=None, explicit=None, **kw):
    want = _oracle_counts(seqs, k, seed, variants, seq_id0, explicit)
    got = ft.profiles(ss, k, variants, out_kind=ft.OUT_COUNTS_I32, seed=seed, seq_id0=seq_id0, edit_lists=lists, **kw).cpu().numpy()
    bad = np.argwhere((got != want).any(axis=2))
    assert bad.size == 0, ("counts differ (variant, sequence):", bad[:10].tolist(), [len(s) for s in seqs])
    w64 = (want + 1) / (want + 1).sum(axis=2, keepdims=True)           # utils.py:242-250
    f64 = ft.profiles(ss, k, variants, out_kind=ft.OUT_FREQ_F64, seed=seed, seq_id0=seq_id0, edit_lists=lists, **kw).cpu().numpy()
    assert np.array_equal(f64, w64)
    w32 = w64.astype(np.float32)                                       # utils.py:353
    f32 = ft.profiles(ss, k, variants, out_kind=ft.OUT_FREQ_F32, seed=seed, seq_id0=seq_id0, edit_lists=lists, **kw).cpu().numpy()
    assert np.array_equal(f32, w32)
    mean, var, scale = orc.standard_scaler_fit(w32[0])                 # utils.py:358-359 (sklearn arithmetic)
    m32, s32 = mean.astype(np.float32), scale.astype(np.float32)
    std = ft.profiles(ss, k, variants, out_kind=ft.OUT_STD_F32, seed=seed, seq_id0=seq_id0, edit_lists=lists,
                      mean=torch.from_numpy(m32).cuda(), scale=torch.from_numpy(s32).cuda(), **kw).cpu().numpy()
    wstd = orc.standard_scaler_transform(w32.reshape(-1, 4 ** k), mean, scale).reshape(w32.shape)   # utils.py:361-366
    assert np.array_equal(std, wstd)
    return want


LONG_LENGTHS = (65535, 65536, 70001, 200000, 1048577)


@pytest.mark.parametrize("k", [6, 5, 4])
def test_long_sequences_all_variant_kinds_vs_oracle(ft, SeqSet, k):
    """L > 65 535 (int32 histograms): clean, transition, transversion, combined, Random_N and explicit lists"""
    rng = np.random.default_rng(100 + k)
    lens = LONG_LENGTHS if k == 6 else LONG_LENGTHS[:3]
    seqs = [_rand_seq(rng, L) for L in lens] + [_rand_seq(rng, 3000), b"ACGTNACGT" * 9000]
    ss = SeqSet.from_sequences(seqs)
    n = len(seqs)
    explicit = [[]]
    for s in seqs:       # an explicit list per sequence: substitutions + Ns, clustered and spread, incl. first / last base
        L = len(s)
        pos = np.unique(np.concatenate([rng.integers(0, L, size=300), np.arange(0, 40, 3), [L - 1, L - 2, L // 2, L // 2 + 1, L // 2 + 5]]))
        explicit[0].append((pos, rng.integers(0, 5, size=pos.size)))
    lists = ft.pack_edit_lists(explicit, n, ss.device)
    variants = [ft.VariantSpec(ft.KIND_CLEAN), ft.VariantSpec(ft.KIND_BOTH, 1e-2, 0.5e-2), ft.VariantSpec(ft.KIND_TRANSITION, p1=1e-2),
                ft.VariantSpec(ft.KIND_TRANSVERSION, p2=0.5e-2), ft.VariantSpec(ft.KIND_RANDOM_N, n_bp=20),
                ft.VariantSpec(ft.KIND_RANDOM_N, n_bp=333), ft.VariantSpec(ft.KIND_EXPLICIT, explicit_idx=0)]
    for i, v in enumerate(variants):
        v.rng_id = i
    want = _check_all_outputs(ft, ss, seqs, k, variants, seed=0xABCDEF12345, seq_id0=11, lists=lists, explicit=explicit)
    assert want[0, 0].sum() > 60000      # the long items really are long
    if k == 6:   # k = 6 went through the chunked path (tiles + exact reduction); the generic kernel's int32 long path separately
        _check_all_outputs(ft, ss, seqs, k, variants, seed=0xABCDEF12345, seq_id0=11, lists=lists, explicit=explicit, chunked=False)
    # selection mode on long items (pair batches: slot 0 + one mimic per item)
    sidx = torch.tensor([0, 2, 1, 0], dtype=torch.int32, device="cuda")
    sel = torch.tensor([[1, 2], [1, 4], [1, 6], [1, 3]], dtype=torch.int32, device="cuda")
    got = ft.profiles(ss, k, variants, out_kind=ft.OUT_COUNTS_I32, seed=0xABCDEF12345, seq_id0=11, edit_lists=lists,
                      sidx=sidx, sel=sel).cpu().numpy()
    for w in range(4):
        for s in range(2):
            assert np.array_equal(got[s, w], want[int(sel[w, s]), int(sidx[w])]), (w, s)


def test_high_rates_on_long_sequences_vs_oracle(ft, SeqSet):
    """rates far above the reference's 1e-2 / 5e-3 on sequences long enough that a 2 048-entry on-chip edit list
    cannot hold one tile's edits: nothing may be dropped silently (transition(0.1) on >= 20 kb)"""
    rng = np.random.default_rng(7)
    seqs = [_rand_seq(rng, L) for L in (50000, 20000, 70000, 33000)]
    ss = SeqSet.from_sequences(seqs)
    variants = [ft.VariantSpec(ft.KIND_TRANSITION, p1=0.2), ft.VariantSpec(ft.KIND_BOTH, 0.1, 0.1), ft.VariantSpec(ft.KIND_TRANSVERSION, p2=0.5),
                ft.VariantSpec(ft.KIND_BOTH, 1.0, 0.0)]
    for i, v in enumerate(variants):
        v.rng_id = i
    for k in (6, 5):
        want = _oracle_counts(seqs, k, 31337, variants)
        got = ft.profiles(ss, k, variants, out_kind=ft.OUT_COUNTS_I32, seed=31337).cpu().numpy()
        assert np.array_equal(got, want), k
        for spec in variants[:2]:            # the in-kernel statistics path generates the same edits
            x = ft.profiles(ss, k, [spec], out_kind=ft.OUT_FREQ_F32, seed=31337)[0]
            a, b = ft.Scaler.fit(x), ft.profile_stats(ss, k, spec, seed=31337)
            assert torch.allclose(a.mean64, b.mean64, rtol=1e-13, atol=0)


def test_dominant_kernel_standardised_rows_vs_oracle_c3_slab(ft, SeqSet):
    """BASELINE configs[2] shape on a 512-sequence slab: the whole-schedule fast path (k = 6, 51 variants, 10 kb,
    standardised float32 — what bench.py times) against the oracle, EVERY row: counts -> float32(count/total) ->
    StandardScaler statistics of slot 0 -> (x - mean32) / scale32"""
    from idelucs_b200 import utils as U
    rng = np.random.default_rng(2024)
    n, L, k, n_mimics, seed = 512, 10000, 6, 50, 20240607
    seqs = [_rand_seq(rng, L, n_rate=0.001 if i % 4 == 0 else 0.0) for i in range(n)]
    ss = SeqSet.from_sequences(seqs)
    variants = ft.mimic_schedule(n_mimics)
    want = _oracle_counts(seqs, k, seed, variants, seq_id0=5)
    w32 = ((want + 1) / (want + 1).sum(axis=2, keepdims=True)).astype(np.float32)
    mean, var, scale = orc.standard_scaler_fit(w32[0])
    # the product call a user makes (AugmentFasta's device form): statistics + all 51 standardised slots
    got, sc, _ = U.augment_device(ss, n_mimics, k=k, seed=seed, seq_id0=5)
    np.testing.assert_allclose(sc.mean64.cpu().numpy(), mean, rtol=1e-12, atol=0)
    np.testing.assert_allclose(sc.scale64.cpu().numpy(), scale, rtol=1e-9, atol=0)
    # standardise the oracle's float32 frequencies with the statistics the device used (float64 summation order may flip
    # the last bit of a float32 mean; the kernel's arithmetic on given statistics must be exact)
    m32, s32 = sc.mean32.cpu().numpy(), sc.scale32.cpu().numpy()
    assert (m32 != mean.astype(np.float32)).mean() < 0.01 and (s32 != scale.astype(np.float32)).mean() < 0.01
    wstd = ((w32 - m32) / s32).astype(np.float32)
    got = got.cpu().numpy()
    assert got.shape == wstd.shape == (51, n, 4096)
    bad = np.argwhere((got != wstd).any(axis=2))
    assert bad.size == 0, bad[:10].tolist()
    # and the plain frequencies through the same fast path
    f32 = ft.profiles(ss, k, variants, out_kind=ft.OUT_FREQ_F32, seed=seed, seq_id0=5).cpu().numpy()
    assert np.array_equal(f32, w32)


def test_chunked_path_many_genomes_equals_generic_and_oracle(ft, SeqSet):
    """Fungi-shaped set (BASELINE configs[4]): lengths log-uniform over two orders of magnitude, so the tiles of one genome are
    spread over several CTAs and several genomes share a CTA: chunked counts == generic kernel == oracle, all slots of the
    reference schedule (n_mimics = 3) plus extra Random_N / clean slots; the statistics + standardised rows built from them"""
    from idelucs_b200 import utils as U
    rng = np.random.default_rng(77)
    lens = np.exp(rng.uniform(np.log(20000), np.log(700000), size=40)).astype(int)
    lens[[3, 17]] = [65536, 65537]
    seqs = [_rand_seq(rng, int(L), n_rate=0.001) for L in lens]
    ss = SeqSet.from_sequences(seqs)
    variants = ft.mimic_schedule(3) + [ft.VariantSpec(ft.KIND_RANDOM_N, n_bp=20, rng_id=7), ft.VariantSpec(ft.KIND_CLEAN, rng_id=8)]
    a = ft.profiles(ss, 6, variants, out_kind=ft.OUT_COUNTS_I32, seed=5, seq_id0=100)
    b = ft.profiles(ss, 6, variants, out_kind=ft.OUT_COUNTS_I32, seed=5, seq_id0=100, chunked=False)
    assert torch.equal(a, b)
    idx = [0, 3, 17, int(np.argmax(lens)), int(np.argmin(lens))]
    for i in idx:   # (_oracle_counts numbers its sequences from seq_id0: one call per item with the true id)
        w = _oracle_counts([seqs[i]], 6, 5, variants, seq_id0=100 + i)[:, 0]
        assert np.array_equal(a[:, i].cpu().numpy(), w), i
    # a scratch sized for fewer long items than the set holds: the plan kernel notices and the generic kernel computes them
    c = ft.profiles(ss, 6, variants, out_kind=ft.OUT_COUNTS_I32, seed=5, seq_id0=100, max_long=3)
    assert torch.equal(c, b)
    # selection of slots per item (sel) through the chunked path
    sel = torch.stack([torch.zeros(ss.n, dtype=torch.int32), torch.arange(ss.n, dtype=torch.int32) % 5 + 1], dim=1).cuda().contiguous()
    d = ft.profiles(ss, 6, variants, out_kind=ft.OUT_COUNTS_I32, seed=5, seq_id0=100, sel=sel)
    assert torch.equal(d[0], b[0]) and torch.equal(d[1], b[(torch.arange(ss.n) % 5 + 1).cuda(), torch.arange(ss.n).cuda()])
    x, sc, _ = U.augment_device(ss, 3, k=6, seed=9)
    f = ft.profiles(ss, 6, ft.mimic_schedule(3), out_kind=ft.OUT_FREQ_F32, seed=9, chunked=False)
    scg = ft.Scaler.fit(f[0])
    assert torch.allclose(sc.mean64, scg.mean64, rtol=1e-12, atol=0) and torch.allclose(sc.scale64, scg.scale64, rtol=1e-9, atol=0)
    xg = ft.profiles(ss, 6, ft.mimic_schedule(3), out_kind=ft.OUT_STD_F32, seed=9, mean=sc.mean32, scale=sc.scale32, chunked=False)
    assert torch.equal(x, xg)

"""GPU parity tests at BASELINE.json's FULL sizes (-m gpu), through properties that do not depend on the size:

* configs[2] (100 000 sequences x 10 kb, k = 6, n_mimics = 50, standardised float32 — the 83.6 GB pass bench.py times):
  rows of sampled sequences against the oracle's mutate-and-recount (idelucs/kmers.pyx:38-50 on the mutated bytes,
  idelucs/utils.py:330-366) and against the generic kernel on the same items, the standardised 'true' side has zero mean /
  unit variance per column (the scaler of idelucs/utils.py:354-359 fitted on exactly these rows), Random_N rows differ
  from the clean row in at most n_bp * k bins, and a checksum of every slot repeats from run to run;
* configs[3] (1 000 000 sequences x 2 kb, k = 6, n_mimics = 50, batch 512): slab-wise statistics of the whole set equal the
  merge of per-shard statistics, and pair batches regenerated per step (selection mode) equal the oracle's rows.

The whole file needs ~100 GB of device memory and is skipped on smaller devices."""
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


def _need(gb):
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info()
    if free < gb * (1 << 30):
        pytest.skip("needs %d GB of free device memory" % gb)


def _random_set(SeqSet, n, L, seed, n_every=0):
    """n sequences of L uniform bases generated on the device (+ an N at every n_every-th position of every 7th sequence)"""
    g = torch.Generator(device="cuda").manual_seed(seed)
    codes = torch.randint(0, 4, (n * L,), device="cuda", dtype=torch.uint8, generator=g)
    asc = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device="cuda")[codes.long()]
    del codes
    if n_every:
        v = asc.view(n, L)
        v[::7, ::n_every] = ord("N")
    return SeqSet.from_ascii(asc, np.arange(n + 1, dtype=np.int64) * L), asc


def _oracle_rows(seq_bytes, seq_id, k, seed, variants):
    """float32 frequencies [V, 4^k] of one sequence: the oracle's edits applied to the bytes, recount, +1, / total"""
    out = np.zeros((len(variants), 4 ** k), np.int32)
    codes = orc.codes_of_seq(seq_bytes)
    for v, spec in enumerate(variants):
        rid = spec.rng_id if spec.rng_id is not None else v
        edits = orc.rng_variant_edits(seed, seq_id, rid, spec.kind, codes, len(seq_bytes), spec.p1, spec.p2, spec.n_bp)
        mut = bytearray(seq_bytes)
        for p, val in edits:
            mut[p] = b"ACGTN"[val]
        orc.kmer_counts(mut, k, out[v])
    return ((out + 1) / (out + 1).sum(axis=1, keepdims=True)).astype(np.float32)


def _slot_checksums(x):
    return [int(x[s].view(torch.int32).sum(dtype=torch.int64).item()) for s in range(x.shape[0])]


@pytest.mark.timeout(600)
def test_c3_full_size_standardised_schedule(ft, SeqSet):
    from idelucs_b200 import utils as U
    _need(110)
    n, L, k, n_mimics, seed, sid0 = 100000, 10000, 6, 50, 77, 11
    ss, asc = _random_set(SeqSet, n, L, seed=3, n_every=997)
    variants = ft.mimic_schedule(n_mimics)
    x, sc, _ = U.augment_device(ss, n_mimics, k=k, seed=seed, seq_id0=sid0)          # prepare pass + producer/consumer kernel
    assert tuple(x.shape) == (51, n, 4096) and x.dtype == torch.float32
    m32, s32 = sc.mean32.cpu().numpy(), sc.scale32.cpu().numpy()

    # (1) sampled sequences, all 51 rows, against the oracle (standardised with the statistics the device fitted)
    rng = np.random.default_rng(5)
    for i in [0, 7, n - 1] + rng.integers(0, n, size=3).tolist():
        w32 = _oracle_rows(bytes(asc[i * L:(i + 1) * L].cpu().numpy()), sid0 + i, k, seed, variants)
        want = ((w32 - m32) / s32).astype(np.float32)
        got = x[:, i].cpu().numpy()
        assert np.array_equal(got, want), (i, np.argwhere(got != want)[:5].tolist())

    # (2) 2 000 sampled items through the generic kernel (selection of items: one CTA per sequence, no prepared buffer)
    idx = torch.from_numpy(np.sort(rng.choice(n, size=2000, replace=False)).astype(np.int32)).cuda()
    y = ft.profiles(ss, k, variants, out_kind=ft.OUT_STD_F32, seed=seed, seq_id0=sid0, sidx=idx, mean=sc.mean32, scale=sc.scale32)
    assert torch.equal(y, x[:, idx.long()])
    del y

    # (3) the scaler was fitted on slot 0: its standardised columns have zero mean and unit variance (float64 sums of the
    #     float32 rows; a column whose scale fell back to 1 is constant)
    mu = torch.zeros(4096, dtype=torch.float64, device="cuda")
    m2 = torch.zeros(4096, dtype=torch.float64, device="cuda")
    for lo in range(0, n, 10000):
        blk = x[0, lo:lo + 10000].double()
        mu += blk.sum(0)
        m2 += (blk * blk).sum(0)
    mu /= n
    var = m2 / n - mu * mu
    assert float(mu.abs().max()) < 1e-5, float(mu.abs().max())
    assert float((var - 1).abs().max()) < 1e-4, float((var - 1).abs().max())
    # the statistics themselves: float64 column mean of the float32 frequencies, on a 20 000-row slab pair merged by hand
    f = ft.profiles(ss, k, variants[:1], out_kind=ft.OUT_FREQ_F32, seed=seed, seq_id0=sid0)[0]
    fm = torch.zeros(4096, dtype=torch.float64, device="cuda")
    for lo in range(0, n, 10000):
        fm += f[lo:lo + 10000].double().sum(0)
    assert torch.allclose(sc.mean64, fm / n, rtol=1e-12, atol=0)
    # un-standardising a row gives back the frequencies to 1 ulp of the float32 arithmetic
    back = x[0, :2000] * sc.scale32 + sc.mean32
    assert float(((back - f[:2000]).abs() / f[:2000]).max()) < 5e-6
    del f, back

    # (4) Random_N slots (3 .. 50) only REMOVE windows (idelucs/utils.py:89-95): against slot 0's row they differ in the bins
    #     the Bernoulli edits of slot 0 touched (<= 2 k per edit) plus at most n_bp * k removed windows — and rows of two
    #     Random_N slots of one sequence with equal window totals differ in at most 2 * 20 * 6 bins
    a, b = x[3, :4000], x[4, :4000]
    nd = (a != b).sum(dim=1)
    same_total = nd <= 2 * 20 * 6
    assert float(same_total.float().mean()) > 0.2           # (totals agree when both slots removed the same number of windows)
    assert int(nd.min()) > 0                                  # different draws
    # (5) run-to-run identical, slot by slot (int32 view sums), with the first result gone from memory
    c1 = _slot_checksums(x)
    del x, a, b
    torch.cuda.empty_cache()
    x2, sc2, _ = U.augment_device(ss, n_mimics, k=k, seed=seed, seq_id0=sid0)
    assert torch.equal(sc2.mean64, sc.mean64) and torch.equal(sc2.scale64, sc.scale64)
    assert _slot_checksums(x2) == c1
    del x2
    torch.cuda.empty_cache()


@pytest.mark.timeout(600)
def test_c4_full_size_statistics_and_pair_batches(ft, SeqSet):
    from idelucs_b200 import utils as U
    _need(40)
    n, L, k, n_mimics, seed = 1000000, 2000, 6, 50, 2024
    ss, asc = _random_set(SeqSet, n, L, seed=9, n_every=499)
    loader = U.PairBatchLoader(ss, n_mimics, k=k, batch_size=512, seed=seed, materialize_bytes=8 << 30)
    assert loader.profiles is None and loader.n_pairs == 50 * n                      # regenerated per batch, nothing materialised
    sc = loader.scaler                                                                # slab-wise prepare passes over the 10^6 sequences
    variants = loader.variants
    # statistics of the whole set == Chan merge of the statistics of 8 shards (what 8 ranks would exchange)
    parts, part_n = [], []
    for r in range(8):
        sub = torch.arange(r * (n // 8), (r + 1) * (n // 8), dtype=torch.int32, device="cuda")
        pr = ft.prepare(ss, k, variants[:1], seed=seed, sidx=sub)
        m, M2, rows = ft._merge_local(pr.parts, pr.part_n)
        parts.append(torch.stack([m, M2])); part_n.append(rows.reshape(1))
        del pr
    merged = ft.Scaler.from_partials(torch.stack(parts).contiguous(), torch.cat(part_n).contiguous())
    assert torch.allclose(merged.mean64, sc.mean64, rtol=1e-12, atol=0)
    assert torch.allclose(merged.scale64, sc.scale64, rtol=1e-9, atol=0)
    assert float((merged.mean32 != sc.mean32).float().mean()) < 0.01
    # pair batches: 'true' = slot 0 of the sequence, 'modified' = its mimic; sampled pairs against the oracle
    m32, s32 = sc.mean32.cpu().numpy(), sc.scale32.cpu().numpy()
    g = torch.Generator(device="cuda").manual_seed(1)
    pair_ids = torch.randint(0, loader.n_pairs, (512,), device="cuda", generator=g)
    pair_ids[:4] = torch.tensor([0, n - 1, 49 * n, 50 * n - 1], device="cuda")         # first / last sequence, first / last mimic
    bt = loader.batch(pair_ids)
    assert tuple(bt["true"].shape) == (512, 4096) and tuple(bt["modified"].shape) == (512, 4096)
    for j in (0, 1, 2, 3, 100, 511):
        pid = int(pair_ids[j])
        mim, i = pid // n + 1, pid % n
        w32 = _oracle_rows(bytes(asc[i * L:(i + 1) * L].cpu().numpy()), i, k, seed, [variants[0], variants[mim]])
        assert np.array_equal(bt["true"][j].cpu().numpy(), ((w32[0] - m32) / s32).astype(np.float32)), (j, pid)
        assert np.array_equal(bt["modified"][j].cpu().numpy(), ((w32[1] - m32) / s32).astype(np.float32)), (j, pid)
    # the same pairs drawn again give the same rows; the 'true' side of two pairs of one sequence is the same row
    bt2 = loader.batch(pair_ids)
    assert torch.equal(bt2["true"], bt["true"]) and torch.equal(bt2["modified"], bt["modified"])
    two = loader.batch(torch.tensor([5, 5 + 7 * n], device="cuda"))
    assert torch.equal(two["true"][0], two["true"][1]) and not torch.equal(two["modified"][0], two["modified"][1])

"""Pin the CPU oracle (oracle/idelucs_oracle.py): (1) against the committed golden vectors,
which are outputs of the live reference (oracle/gen_golden.py); (2) against the live
reference itself when /root/reference is mounted (build container only)."""
import hashlib
import json
import os
import random

import numpy as np
import pytest

import idelucs_oracle as orc
import ref_live

needs_ref = pytest.mark.skipif(not ref_live.available(), reason="live reference not mounted")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _golden(golden_dir):
    with open(os.path.join(golden_dir, "golden.json")) as fh:
        return json.load(fh)


def test_kmer_counts_kats(golden_dir):
    with open(os.path.join(golden_dir, "kmer_kats.json")) as fh:
        kats = json.load(fh)
    assert len(kats) > 40
    for kat in kats:
        seq, k = bytes.fromhex(kat["seq_hex"]), kat["k"]
        want = np.zeros(4 ** k, dtype=np.int32)
        want[kat["nz_idx"]] = kat["nz_val"]
        got = np.zeros(4 ** k, dtype=np.int32)
        orc.kmer_counts(bytearray(seq), k, got)
        assert np.array_equal(got, want)
        if len(seq) <= 300:
            got2 = np.zeros(4 ** k, dtype=np.int32)
            orc.kmer_counts_py(bytearray(seq), k, got2)
            assert np.array_equal(got2, want)


def test_survey_kats():
    c = np.zeros(16, np.int32)
    orc.kmer_counts(bytearray(b"ACGTNACGTACGTTTT"), 2, c)
    assert c.tolist() == [0, 3, 0, 0, 0, 0, 3, 0, 0, 0, 0, 3, 1, 0, 0, 3]
    c = np.zeros(4, np.int32)
    orc.kmer_counts(bytearray(b"ANCG"), 1, c)
    assert c.tolist() == [1, 1, 1, 0]


@pytest.mark.parametrize("stem", ["Influenza-A", "Actinopterygii"])
@pytest.mark.parametrize("k", [4, 5, 6])
def test_fasta_counts_and_freqs(golden_dir, fasta_files, stem, k):
    g = _golden(golden_dir)["files"][stem][f"k{k}"]
    recs = orc.read_fasta(fasta_files[stem])
    counts = np.zeros((len(recs), 4 ** k), dtype=np.int32)
    for i, (_, seq) in enumerate(recs):
        orc.kmer_counts(seq, k, counts[i])
    assert len(recs) == g["n"]
    assert int(counts.sum()) == g["counts_sum"]
    assert sha(counts) == g["counts_sha256"]
    names, freq = orc.kmersFasta(fasta_files[stem], k=k)
    assert sha(freq) == g["freq64_sha256"]
    assert hashlib.sha256("\n".join(names).encode()).hexdigest() == g["names_sha256"]


@pytest.mark.parametrize("k", [4, 5, 6])
def test_augment_fasta_seed0(golden_dir, fasta_files, k):
    """Same numpy/random global streams as the reference => bit-identical x_train
    (hash is same-container: numpy 2.3 / sklearn-1.9 scaler semantics)."""
    g = _golden(golden_dir)["files"]["Influenza-A"]["augment_seed0_nmimics3"][f"k{k}"]
    np.random.seed(0)
    random.seed(0)
    x = orc.AugmentFasta(fasta_files["Influenza-A"], 3, k=k)
    assert list(x.shape) == g["shape"]
    rows = np.load(os.path.join(golden_dir, f"influenza_xtrain_rows_k{k}.npz"))
    assert np.array_equal(x[rows["rows"]], rows["x"])
    assert sha(x) == g["x_train_sha256"]


def test_edit_lists_reproduce_reference_mutations(golden_dir, fasta_files):
    """Exported mutation lists (reference transforms, seed 0) re-applied by the oracle give
    the same per-pass edit counts the survey measured, and rebuilding x_train from them
    equals the reference AugmentFasta output."""
    g = _golden(golden_dir)["files"]["Influenza-A"]["augment_seed0_nmimics3"]
    ed = np.load(os.path.join(golden_dir, "influenza_edits_seed0.npz"))
    assert g["edits_per_pass"] == [20178, 13383, 6784, 18840]
    recs = orc.read_fasta(fasta_files["Influenza-A"])
    n, k = len(recs), 4
    offs = ed["offsets"]
    prof = np.zeros((4, n, 4 ** k))
    for p in range(4):
        for i, (_, seq) in enumerate(recs):
            lo, hi = offs[p * n + i], offs[p * n + i + 1]
            prof[p, i] = orc.profile_from_seq(orc.apply_edits(seq, ed["pos"][lo:hi], ed["newbyte"][lo:hi]), k)
    x = np.concatenate([np.stack((prof[0], prof[j]), axis=1) for j in (1, 2, 3)], axis=0).astype("float32")
    mean, var, scale = orc.standard_scaler_fit(prof[0].astype("float32"))
    x[:, 0, :] = orc.standard_scaler_transform(x[:, 0, :], mean, scale)
    x[:, 1, :] = orc.standard_scaler_transform(x[:, 1, :], mean, scale)
    assert sha(x) == g["k4"]["x_train_sha256"]


@pytest.mark.parametrize("stem", ["Influenza-A", "Actinopterygii"])
def test_inference_profiles(golden_dir, fasta_files, stem):
    g = _golden(golden_dir)["files"][stem]["inference_k6"]
    _, x = orc.inference_profiles(fasta_files[stem], k=6)
    rows = np.load(os.path.join(golden_dir, f"{stem}_inference_rows_k6.npz"))
    assert np.array_equal(x[rows["rows"]], rows["x"])
    assert sha(x) == g["kmers_sha256"]


def _loss_inputs(B, C, salt):
    import gen_golden
    return gen_golden.loss_inputs(B, C, salt)


def test_iid_loss_goldens(golden_dir):
    with open(os.path.join(golden_dir, "iid_loss.json")) as fh:
        cases = json.load(fh)
    for c in cases:
        z1, z2 = _loss_inputs(c["B"], c["C"], c["salt"]), _loss_inputs(c["B"], c["C"], c["salt"] + 100)
        assert sha(z1) == c["z1_sha256"]
        l64 = orc.IID_loss(z1.astype(np.float64), z2.astype(np.float64), lamb=c["lamb"])
        assert abs(l64 - c["loss64"]) <= 1e-12 * max(1.0, abs(c["loss64"]))
        l32 = orc.IID_loss(z1, z2, lamb=c["lamb"])
        assert abs(float(l32) - c["loss32"]) <= 1e-5  # north-star tolerance for the loss
        loss, dz1, dz2 = orc.IID_loss_grad(z1, z2, lamb=c["lamb"])
        assert abs(loss - c["loss64"]) <= 1e-12 * max(1.0, abs(c["loss64"]))
        rows = np.asarray(c["rows"])
        np.testing.assert_allclose(dz1[rows], np.asarray(c["dz1_rows64"]), rtol=1e-9, atol=1e-13)
        np.testing.assert_allclose(dz2[rows], np.asarray(c["dz2_rows64"]), rtol=1e-9, atol=1e-13)


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    assert orc.philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert orc.philox4x32_10((0xffffffff,) * 4, (0xffffffff,) * 2) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert orc.philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == (
        0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


def test_rng_mode_statistics():
    """rng mode is not the reference's MT19937 stream; it must have the same law:
    Bernoulli(p) hits per base, 50/50 transversion choice, n_bp draws for Random_N."""
    L = 20000
    rng = np.random.default_rng(5)
    codes = rng.integers(0, 4, size=L).astype(np.uint8)
    tot = {1: 0, 2: 0, 3: 0}
    choice = [0, 0]
    nseq = 40
    for s in range(nseq):
        for kind in (orc.KIND_TRANSITION, orc.KIND_TRANSVERSION, orc.KIND_BOTH):
            e = orc.rng_variant_edits(99, s, kind, kind, codes, L)
            tot[kind] += len(e)
            for pos, val in e:
                assert val != codes[pos]
                if kind == orc.KIND_TRANSITION:
                    assert val == codes[pos] ^ 2
                if kind == orc.KIND_TRANSVERSION:
                    assert (val & 1) != (codes[pos] & 1)
                    choice[val >> 1] += 1
        e = orc.rng_variant_edits(99, s, 7, orc.KIND_RANDOM_N, codes, L)
        assert 15 <= len(e) <= 20 and all(v == 4 for _, v in e)
    n = nseq * L
    for kind, p in ((1, 1e-2), (2, 0.5e-2), (3, 1 - (1 - 1e-2) * (1 - 0.5e-2))):
        sd = (n * p * (1 - p)) ** 0.5
        assert abs(tot[kind] - n * p) < 5 * sd, (kind, tot[kind], n * p)
    m = sum(choice)
    assert abs(choice[0] - m / 2) < 5 * (m / 4) ** 0.5


# ----------------------------------------------------------------------------------------
# against the live reference (build container only)
# ----------------------------------------------------------------------------------------

@needs_ref
def test_live_kmer_counts_random_bytes():
    ref = ref_live.load()
    rng = np.random.default_rng(0)
    alph = np.frombuffer(b"ACGTACGTACGTNacgtRYKM-\x00\xff", dtype=np.uint8)
    for t in range(300):
        L = int(rng.integers(0, 400))
        k = int(rng.integers(1, 7))
        s = rng.choice(alph, size=L).tobytes() if t % 3 else rng.integers(0, 256, size=L, dtype=np.uint8).tobytes()
        a = np.ones(4 ** k, np.int32)
        b = np.ones(4 ** k, np.int32)
        c = np.ones(4 ** k, np.int32)
        ref.kmer_counts(bytearray(s), k, a)
        orc.kmer_counts(bytearray(s), k, b)
        orc.kmer_counts_py(bytearray(s), k, c)
        assert np.array_equal(a, b) and np.array_equal(a, c)


@needs_ref
def test_live_check_sequence():
    ref = ref_live.load()
    for s in (b"acgtuUswkmyrbdhvnSWKMYRBDHV-ACGTN", b"AC GT\tAC\nGT\r", b"ACGTX", b"ACGT\xe9", b"", b"ACxGT"):
        try:
            want = ("ok", bytes(ref.check_sequence("hdr", bytearray(s))))
        except ValueError as e:
            want = ("err", str(e))
        try:
            got = ("ok", bytes(orc.check_sequence("hdr", bytearray(s))))
        except ValueError as e:
            got = ("err", str(e))
        assert got == want
    for hdr in (">x", "#x", " x", "a\tb"):
        with pytest.raises(ValueError):
            orc.check_sequence(hdr, bytearray(b"ACGT"))
        with pytest.raises(ValueError):
            ref.check_sequence(hdr, bytearray(b"ACGT"))


@needs_ref
def test_live_augment_fasta_actinopterygii(fasta_files):
    ref = ref_live.load()
    np.random.seed(3)
    random.seed(3)
    want = ref.utils.AugmentFasta(fasta_files["Actinopterygii"], 4, k=5)
    np.random.seed(3)
    random.seed(3)
    got = orc.AugmentFasta(fasta_files["Actinopterygii"], 4, k=5)
    assert np.array_equal(got, want)


@needs_ref
def test_live_scaler_matches_sklearn():
    from sklearn.preprocessing import StandardScaler
    rng = np.random.default_rng(1)
    X = rng.random((200, 64)).astype(np.float32) * 1e-3
    X[:, 3] = X[0, 3]  # constant column -> scale 1
    X[:, 5] = 0.0
    sc = StandardScaler().fit(X)
    mean, var, scale = orc.standard_scaler_fit(X)
    assert np.array_equal(mean, sc.mean_) and np.array_equal(var, sc.var_) and np.array_equal(scale, sc.scale_)
    assert np.array_equal(orc.standard_scaler_transform(X, mean, scale), sc.transform(X))
    X64 = X.astype(np.float64)
    sc = StandardScaler().fit(X64)
    mean, var, scale = orc.standard_scaler_fit(X64)
    assert np.array_equal(scale, sc.scale_)
    assert np.array_equal(orc.standard_scaler_transform(X64, mean, scale), sc.transform(X64))


@needs_ref
def test_live_iid_loss_and_grad():
    import torch
    ref = ref_live.load()
    rng = np.random.default_rng(2)
    for B, C, lamb in ((64, 5, 2.8), (100, 17, 1.0), (33, 200, 2.0)):
        a = rng.standard_normal((B, C)) * 3
        b = rng.standard_normal((B, C)) * 3
        z1 = torch.softmax(torch.from_numpy(a), 1).requires_grad_(True)
        z2 = torch.softmax(torch.from_numpy(b), 1).requires_grad_(True)
        val = ref.LossFunctions.IID_loss(z1, z2, lamb=lamb)
        val.backward()
        loss, dz1, dz2 = orc.IID_loss_grad(z1.detach().numpy(), z2.detach().numpy(), lamb=lamb)
        assert abs(loss - val.item()) < 1e-12
        np.testing.assert_allclose(dz1, z1.grad.numpy(), rtol=1e-9, atol=1e-14)
        np.testing.assert_allclose(dz2, z2.grad.numpy(), rtol=1e-9, atol=1e-14)
        j = ref.LossFunctions.compute_joint(z1.detach(), z2.detach()).numpy()
        np.testing.assert_allclose(orc.compute_joint(z1.detach().numpy(), z2.detach().numpy()), j, rtol=1e-12)


# ---- cgr / kmer_rev_comp ------------------------------------------------------------------
def test_cgr_is_a_permutation_of_kmer_counts():
    """the FCGR vector is the k-mer count vector under cgr_index (the map the CUDA kernel applies)"""
    rng = np.random.default_rng(3)
    alph = np.frombuffer(b"ACGTACGTACGTN", dtype=np.uint8)
    for k in (1, 2, 3, 4, 6):
        seq = alph[rng.integers(0, alph.size, size=700)].tobytes()
        c = np.zeros(4 ** k, np.int32)
        g = np.zeros(4 ** k, np.int32)
        orc.kmer_counts(bytearray(seq), k, c)
        orc.cgr(bytearray(seq), k, g)
        idx = np.zeros(4 ** k, np.int64)
        for m in range(4 ** k):
            ci = cj = 0
            for t in range(k):
                d = (m >> (2 * (k - 1 - t))) & 3
                ci |= (1 if d in (0, 3) else 0) << t
                cj |= (d >> 1) << t
            idx[m] = (ci << k) | cj
        assert sorted(idx.tolist()) == list(range(4 ** k))
        assert np.array_equal(g[idx], c)


def test_reverse_complement_and_canonical_count():
    assert orc.reverse_complement(1, 3) == 47                       # AAC -> GTT
    for k, n in ((2, 10), (3, 32), (4, 136), (6, 2080)):
        assert sum(1 for m in range(4 ** k) if m <= orc.reverse_complement(m, k)) == n
    v = np.arange(1, 17, dtype=np.int32)
    out = orc.kmer_rev_comp(v.copy(), 2)
    # AA(0)<->TT(15): int((1 + 16) * 0.5) = 8 (truncated); palindrome AT(3): (4 + 4) * 0.5 = 4
    assert out[0] == 8 and out.dtype == np.int32 and 4 in out.tolist()


@pytest.mark.skipif(not ref_live.available(), reason="reference mount absent")
def test_live_cgr_and_reduce(fasta_files):
    """oracle == live reference: cgr() on random bytes, kmer_rev_comp on int32 counts, kmersFasta(reduce=True)"""
    ref = ref_live.load()
    from idelucs import kmers as rk, utils as ru
    rng = np.random.default_rng(11)
    for k in (2, 4, 6):
        seq = bytearray(rng.integers(60, 90, size=3000, dtype=np.uint8).tobytes())
        a = np.zeros(4 ** k, np.int32)
        b = np.zeros(4 ** k, np.int32)
        rk.cgr(seq, k, a)
        orc.cgr(seq, k, b)
        assert np.array_equal(a, b)
        v = rng.integers(1, 50, size=4 ** k).astype(np.int32)
        assert np.array_equal(ru.kmer_rev_comp(v.copy(), k), orc.kmer_rev_comp(v.copy(), k))
    n1, x1 = ru.kmersFasta(fasta_files["Influenza-A"], k=4, reduce=True)
    n2, x2 = orc.kmersFasta(fasta_files["Influenza-A"], k=4, reduce=True)
    assert n1 == n2 and x1.shape == (949, 136) and np.array_equal(x1, x2)


@needs_ref
def test_live_fasta_record_loop_vs_native_scanner(tmp_path):
    """the live reference's kmersFasta (idelucs/utils.py:224-277) on files that exercise its record loop — comments,
    blank and blank-padded lines, CRLF, an empty first id (its lines run into the next record), no trailing newline —
    against the native scanner (idl_fasta_scan / idl_fasta_extract) followed by the oracle's check_sequence + counting:
    same names, same float64 profiles"""
    ref = ref_live.load()
    from idelucs_b200.seqset import read_fasta_native
    cases = [
        b">a\nACGTTGCA\nACGGT\n>b\nGGCATTA\n",
        b">a\nACGTTGCA\nACGGT\n>b\nGGCATTA",
        b"# c\n>a desc\n  ACGTAC \t\n#mid\nTTGACA\n\n>b\n\n>c\nACGTNNACGT\n",
        b">\nACGTAC\n>x\nGTACGT\n>y\nTTTTAC\n",
        b">a\r\nACGTAC\r\nTTGGCC\r\n>b\r\nGGGGCCAT\r\n",
        b">a\nacgtacgtRYKM\n>b\nAC-GTACGT\n",
    ]
    for i, content in enumerate(cases):
        pth = os.path.join(str(tmp_path), "live%d.fa" % i)
        with open(pth, "wb") as fh:
            fh.write(content)
        want_names, want = ref.utils.kmersFasta(pth, k=3)
        names, flat, off = read_fasta_native(pth, pinned=False)
        flat = flat.numpy()
        assert names == list(want_names), (i, names, want_names)
        for r, name in enumerate(names):
            seq = orc.check_sequence(name, bytearray(flat[off[r]:off[r + 1]].tobytes()))
            c = np.ones(64, np.int32)
            orc.kmer_counts(seq, 3, c)
            assert np.array_equal(c / c.sum(), want[r]), (i, r)


@pytest.mark.skipif(not ref_live.available(), reason="reference mount absent")
def test_live_training_epoch_port(fasta_files):
    """oracle/train_port.py (the CPU baseline of bench.py's train.pairs_per_s) against the live reference's
    contrastive_training_epoch (idelucs/models.py:113-143): same seeds, same x_train, same loader order -> the same epoch
    losses (identical torch ops on the CPU)"""
    import random
    import torch
    from torch.utils.data import DataLoader
    import train_port
    ref = ref_live.load()
    from idelucs import models as rmodels
    from idelucs import utils as rutils
    np.random.seed(0); random.seed(0)
    x = rutils.AugmentFasta(fasta_files["Actinopterygii"], 3, k=4)
    args = {"sequence_file": fasta_files["Actinopterygii"], "GT_file": None, "n_clusters": 3, "k": 4, "model_size": "linear", "n_mimics": 3,
            "batch_sz": 64, "lambda": 2.8, "lr": 1e-3, "weight": 0.25, "scheduler": None, "optimizer": "RMSprop"}
    torch.manual_seed(5)
    m = rmodels.IID_model(args)
    m.dataloader = DataLoader(rutils.AugmentedDataset(x), batch_size=64, shuffle=True, num_workers=0)
    torch.manual_seed(6)
    want = [m.contrastive_training_epoch() for _ in range(2)]
    torch.manual_seed(5)
    t = train_port.Trainer(x, 256, 3, batch_sz=64, lamb=2.8, weight=0.25, lr=1e-3, num_workers=0)
    torch.manual_seed(6)
    got = [t.contrastive_training_epoch() for _ in range(2)]
    assert np.allclose(want, got, rtol=1e-6, atol=1e-7), (want, got)

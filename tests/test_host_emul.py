"""CPU replay of the kernels' per-thread routines (idelucs_b200/csrc/core.cuh compiled with
g++, see tests/host_emul.cpp) checked against the oracle.  This is how the index logic of
the CUDA path is validated in the GPU-less build container; the same routines are then
checked end-to-end on the B200 by the ``-m gpu`` tests."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import idelucs_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emul():
    build = os.path.join(ROOT, "tests", "_build")
    os.makedirs(build, exist_ok=True)
    so = os.path.join(build, "libhost_emul.so")
    src = os.path.join(ROOT, "tests", "host_emul.cpp")
    hdr = os.path.join(ROOT, "idelucs_b200", "csrc", "core.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-mfma", "-ffp-contract=off", "-shared", "-fPIC",
                               "-x", "c++", src, "-o", so])
    lib = ctypes.CDLL(so)
    lib.emul_pack.restype = ctypes.c_longlong
    lib.emul_div_check.restype = ctypes.c_longlong
    return lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def pack(lib, seq: bytes, strict=0):
    L = len(seq)
    nch = (L + 63) // 64
    codes = np.zeros((nch + 1) * 4, np.uint32)
    nmask = np.zeros((nch + 1) * 2, np.uint32)
    a = np.frombuffer(seq, dtype=np.uint8) if L else np.zeros(1, np.uint8)
    bad = lib.emul_pack(_ptr(a), ctypes.c_longlong(L), strict, _ptr(codes), _ptr(nmask))
    return codes, nmask, bad


def counts(lib, codes, nmask, L, k):
    c = np.zeros(4 ** k, np.int32)
    nv = lib.emul_counts(_ptr(codes), _ptr(nmask), L, k, _ptr(c))
    return c, nv


def rand_seq(rng, L, n_rate=0.0):
    alph = np.frombuffer(b"ACGT", dtype=np.uint8)
    s = alph[rng.integers(0, 4, size=L)]
    if n_rate:
        s = np.where(rng.random(L) < n_rate, ord("N"), s).astype(np.uint8)
    return s.tobytes()


def test_pack_alphabet(emul):
    s = b"acgtuUswkmyrbdhvnSWKMYRBDHV-ACGTN"
    codes, nmask, bad = pack(emul, s)
    assert bad == -1
    want = orc.codes_of_seq(orc.check_sequence("x", bytearray(s)))
    got = np.array([(codes[i >> 4] >> (30 - 2 * (i & 15))) & 3 for i in range(len(s))])
    gotn = np.array([(nmask[i >> 5] >> (31 - (i & 31))) & 1 for i in range(len(s))])
    assert np.array_equal(gotn, (want == 4).astype(int))
    assert np.array_equal(got[want < 4], want[want < 4])
    # invalid byte and deletable whitespace are reported with their position
    _, _, bad = pack(emul, b"ACGTACGTACGTACGTACGTAXGT")
    assert bad == (21 << 3 | 6)
    _, _, bad = pack(emul, b"ACG TACGT")
    assert bad == (3 << 3 | 5)
    # strict (kmers.pyx LUT): lowercase is a reset, nothing is invalid
    codes, nmask, bad = pack(emul, b"ACgtAC\xff", strict=1)
    assert bad == -1
    assert [(nmask[0] >> (31 - i)) & 1 for i in range(7)] == [0, 0, 1, 1, 0, 0, 1]


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6])
def test_count_chunk_matches_oracle(emul, k):
    rng = np.random.default_rng(k)
    for L in [0, 1, k - 1, k, k + 1, 15, 16, 17, 31, 32, 33, 63, 64, 65, 127, 128, 129, 191, 192, 193, 1000, 4099]:
        for n_rate in (0.0, 0.02, 0.3):
            s = rand_seq(rng, max(L, 0), n_rate)
            codes, nmask, _ = pack(emul, s)
            got, nv = counts(emul, codes, nmask, len(s), k)
            want = np.zeros(4 ** k, np.int32)
            orc.kmer_counts(bytearray(s), k, want)
            assert np.array_equal(got, want), (L, n_rate)
            assert nv == want.sum()


def test_count_strict_bytes(emul):
    rng = np.random.default_rng(7)
    for _ in range(50):
        s = rng.integers(0, 256, size=int(rng.integers(0, 500)), dtype=np.uint8).tobytes()
        codes, nmask, _ = pack(emul, s, strict=1)
        got, _ = counts(emul, codes, nmask, len(s), 3)
        want = np.zeros(64, np.int32)
        orc.kmer_counts(bytearray(s), 3, want)
        assert np.array_equal(got, want)


def test_philox_and_tables(emul):
    out = np.zeros(4, np.uint32)
    emul.emul_philox(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0, _ptr(out))
    assert tuple(int(x) for x in out) == orc.philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0))
    for p in (1e-2, 0.5e-2, 0.3, 0.0, 1.0, 1e-9):
        t = np.zeros(64, np.uint32)
        emul.emul_geometric_table(ctypes.c_double(p), _ptr(t))
        assert [int(x) for x in t] == orc.geometric_table(p)


def variant(lib, codes, nmask, L, k, seed, seq_id, rng_id, kind, p1=1e-2, p2=0.5e-2, n_bp=20, explicit=None):
    clean, nv = counts(lib, codes, nmask, L, k)
    lst = np.zeros(max(L + 8, n_bp + 8), np.uint32)
    nlist = ctypes.c_int(0)
    ex = np.asarray(explicit if explicit is not None else [], dtype=np.uint32)
    d = lib.emul_variant(_ptr(codes), _ptr(nmask), L, k, ctypes.c_ulonglong(seed), seq_id, rng_id, kind,
                         ctypes.c_double(p1), ctypes.c_double(p2), n_bp, _ptr(ex) if ex.size else None, int(ex.size),
                         _ptr(clean), _ptr(lst), ctypes.byref(nlist))
    return clean, nv + d, lst[:nlist.value]


@pytest.mark.parametrize("k", [4, 5, 6])
def test_rng_variants_match_oracle_recount(emul, k):
    """delta-patched histogram == oracle's mutate-then-recount, for every variant kind,
    with high mutation rates so that edits cluster within k of each other, and with Ns."""
    rng = np.random.default_rng(100 + k)
    seed = 0x1234567887654321
    cases = [(300, 0.0, 1e-2, 0.5e-2), (1409, 0.01, 1e-2, 0.5e-2), (700, 0.05, 0.3, 0.25), (129, 0.0, 0.9, 0.9),
             (64, 0.2, 0.5, 0.5), (5, 0.0, 0.5, 0.5), (2500, 0.001, 0.08, 0.04)]
    for ci, (L, n_rate, p1, p2) in enumerate(cases):
        s = rand_seq(rng, L, n_rate)
        codes, nmask, _ = pack(emul, s)
        kinds = [orc.KIND_BOTH, orc.KIND_TRANSITION, orc.KIND_TRANSVERSION, orc.KIND_RANDOM_N, orc.KIND_RANDOM_N, orc.KIND_CLEAN]
        want = orc.rng_mimic_counts([bytearray(s)], k, seed, kinds, seq_id0=17 + ci, p1=p1, p2=p2, n_bp=20)
        for v, kind in enumerate(kinds):
            got, tot, lst = variant(emul, codes, nmask, L, k, seed, 17 + ci, v, kind, p1, p2, 20)
            edits = orc.rng_variant_edits(seed, 17 + ci, v, kind, orc.codes_of_seq(s), L, p1, p2, 20)
            if kind != orc.KIND_RANDOM_N:
                assert [(int(e) >> 3, int(e) & 7) for e in lst] == edits, (ci, v)
            else:
                assert sorted(set((int(e) >> 3, int(e) & 7) for e in lst)) == edits
            assert np.array_equal(got, want[v, 0]), (ci, v, kind)
            assert tot == want[v, 0].sum()


def test_explicit_edit_lists_match_oracle(emul):
    """arbitrary sorted unique edit lists, including N->base, base->N, no-op edits, adjacent edits."""
    rng = np.random.default_rng(5)
    for k in (3, 4, 6):
        for L in (1, 7, 64, 200, 1000):
            s = bytearray(rand_seq(rng, L, 0.05))
            codes, nmask, _ = pack(emul, bytes(s))
            for dens in (0.02, 0.3, 1.0):
                pos = np.nonzero(rng.random(L) < dens)[0]
                val = rng.integers(0, 5, size=pos.size)
                mut = bytearray(s)
                for p, v in zip(pos, val):
                    mut[int(p)] = b"ACGTN"[int(v)]
                want = np.zeros(4 ** k, np.int32)
                orc.kmer_counts(mut, k, want)
                got, tot, _ = variant(emul, codes, nmask, L, k, 0, 0, 0, orc.KIND_EXPLICIT,
                                      explicit=(pos.astype(np.uint32) << 3) | val.astype(np.uint32))
                assert np.array_equal(got, want), (k, L, dens)
                assert tot == want.sum()


def test_reference_mutation_lists(emul, golden_dir, fasta_files):
    """The reference's own mutations (seed 0, exported by oracle/gen_golden.py) fed through the
    explicit-list path give the same integer counts as recounting the mutated sequence."""
    ed = np.load(os.path.join(golden_dir, "influenza_edits_seed0.npz"))
    recs = orc.read_fasta(fasta_files["Influenza-A"])
    n = len(recs)
    offs = ed["offsets"]
    lut = np.full(256, 4, np.uint32)
    for i, ch in enumerate(b"ACGT"):
        lut[ch] = i
    for p in range(4):
        for i in range(0, n, 37):
            seq = recs[i][1]
            lo, hi = offs[p * n + i], offs[p * n + i + 1]
            pos, nb = ed["pos"][lo:hi], ed["newbyte"][lo:hi]
            want = np.zeros(4 ** 6, np.int32)
            orc.kmer_counts(orc.apply_edits(seq, pos, nb), 6, want)
            codes, nmask, _ = pack(emul, bytes(seq))
            got, _, _ = variant(emul, codes, nmask, len(seq), 6, 0, 0, 0, orc.KIND_EXPLICIT,
                                explicit=(pos.astype(np.uint32) << 3) | lut[nb])
            assert np.array_equal(got, want)


def test_div_rn_is_ieee_division(emul):
    rng = np.random.default_rng(11)
    n = 4_000_000
    tot = rng.integers(4097, 1 << 24, size=n).astype(np.float32)
    cnt = (rng.integers(1, 1 << 24, size=n) % tot.astype(np.int64) + 1).astype(np.float32)
    assert emul.emul_div_check(_ptr(cnt), _ptr(tot), ctypes.c_longlong(n)) == 0
    cnt = rng.integers(1, 60, size=n).astype(np.float32)
    tot = rng.integers(4096, 30000, size=n).astype(np.float32)
    assert emul.emul_div_check(_ptr(cnt), _ptr(tot), ctypes.c_longlong(n)) == 0
    a = (rng.standard_normal(n) * 10.0 ** rng.integers(-8, 3, size=n)).astype(np.float32)
    b = (rng.random(n) * 10.0 ** rng.integers(-7, 1, size=n) + 1e-9).astype(np.float32)
    assert emul.emul_div_check(_ptr(a), _ptr(b), ctypes.c_longlong(n)) == 0


@pytest.mark.parametrize("k", [1, 3, 4, 6])
def test_random_n_specialised_path(emul, k):
    """the unsorted-draw removal rule (random_n_removals) == mutate-and-recount, including
    duplicate draws, draws within k of each other, draws next to existing Ns and sequence ends"""
    rng = np.random.default_rng(50 + k)
    for L, n_rate, n_bp in [(40, 0.0, 20), (7, 0.0, 20), (300, 0.05, 20), (1409, 0.0, 20), (100, 0.3, 32), (64, 0.0, 3), (1, 0.0, 5)]:
        for rep in range(4):
            s = rand_seq(rng, L, n_rate)
            codes, nmask, _ = pack(emul, s)
            got, nv = counts(emul, codes, nmask, L, k)
            d = emul.emul_random_n(_ptr(codes), _ptr(nmask), L, k, ctypes.c_ulonglong(99), 5 + rep, 3, n_bp, _ptr(got))
            edits = orc.rng_variant_edits(99, 5 + rep, 3, orc.KIND_RANDOM_N, orc.codes_of_seq(s), L, n_bp=n_bp)
            mut = bytearray(s)
            for pos, val in edits:
                mut[pos] = ord("N")
            want = np.zeros(4 ** k, np.int32)
            orc.kmer_counts(mut, k, want)
            assert np.array_equal(got, want), (L, n_rate, n_bp, rep)
            assert nv + d == want.sum()


def test_mask_block_generator_equals_stream_walk(emul):
    """the register-only mask generator (hits of a stream as a 64-bit mask, merge / N-drop / ordering by logic ops)
    emits exactly the edits of the stream-walking merged generator, at any rate (no per-block cap)"""
    rng = np.random.default_rng(77)
    biggest = 0
    for L, n_rate, p1, p2 in [(10000, 0.0, 1e-2, 0.5e-2), (2000, 0.01, 1e-2, 0.5e-2), (1409, 0.05, 0.05, 0.03),
                              (700, 0.1, 0.12, 0.1), (130, 0.0, 0.5, 0.5), (63, 0.3, 0.2, 0.2), (200, 0.02, 1.0, 0.0),
                              (200, 0.02, 0.0, 1.0), (321, 0.0, 1.0, 1.0), (64, 0.0, 0.0, 0.0)]:
        s = rand_seq(rng, L, n_rate)
        codes, nmask, _ = pack(emul, s)
        for kind in (orc.KIND_BOTH, orc.KIND_TRANSITION, orc.KIND_TRANSVERSION):
            for seq_id in range(6):
                mx = ctypes.c_int(0)
                bad = emul.emul_masks_vs_slow(_ptr(codes), _ptr(nmask), L, ctypes.c_ulonglong(4242), seq_id, kind, kind,
                                              ctypes.c_double(p1), ctypes.c_double(p2), ctypes.byref(mx))
                assert bad == 0, (L, p1, kind, seq_id)
                biggest = max(biggest, mx.value)
    assert biggest == 64   # p = 1: every base of a block is edited


def test_entry_deltas_register_form_equals_callback_form(emul):
    """entry_deltas (the prepare kernel's single-evaluation form: deltas returned in registers) produced exactly the
    +-1 updates of apply_entry for every edit-list entry the tests above pushed through the emulation, plus a dense run"""
    rng = np.random.default_rng(5)
    for L, p1, p2 in [(3000, 0.2, 0.1), (500, 0.5, 0.5), (64, 1.0, 0.0)]:
        s = rand_seq(rng, L, 0.02)
        codes, nmask, _ = pack(emul, s)
        for k in (4, 5, 6):
            c, _ = counts(emul, codes, nmask, L, k)
            variant(emul, codes, nmask, L, k, 7, 1, 0, orc.KIND_BOTH, p1=p1, p2=p2)
    emul.emul_entry_deltas_mismatches.restype = ctypes.c_longlong
    assert emul.emul_entry_deltas_mismatches() == 0

"""TEST INFRASTRUCTURE ONLY — CPU oracle for the iDeLUCS featurisation / mimic / IIC-loss hot path.

This module is a *restatement* of the reference algorithm in numpy / plain Python.  It is
the checker the CUDA path is compared against.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it; the product
package ``idelucs_b200`` never does (it fails loudly when its CUDA library is missing).

Pinning status: every function below is pinned against the *live* reference
(``/root/reference`` imported in the build container through ``oracle/ref_live.py``) by
``tests/test_oracle_pinning.py`` and against the committed golden vectors in
``tests/golden/`` (generated from the live reference by ``oracle/gen_golden.py``).  The
reference's own test-suite holds no golden vectors for this path (tests/test_import.py:1-6
only imports the package), so the goldens here are reference *outputs*, not reference
*fixtures*.

Reference citations are ``file:line`` into the upstream repository.

Part A restates the reference (kmers.pyx, utils.py, LossFunctions.py, sklearn scaler).
Part B is the specification of the counter-based RNG ("rng mode") mimic generator that the
B200 kernels implement — the reference draws its mutations from numpy's global MT19937
stream, which a GPU cannot reproduce, so rng mode is specified here independently (Philox
4x32-10 + geometric gap tables) and the kernels are tested bit-exactly against it.
"""
from __future__ import annotations

import random
import sys

import numpy as np

# --------------------------------------------------------------------------------------
# Part A — restatement of the reference
# --------------------------------------------------------------------------------------

_A, _C, _G, _T, _N = ord("A"), ord("C"), ord("G"), ord("T"), ord("N")

# idelucs/kmers.pyx:19-34 — 256-entry LUT: 'A'(65)->0 'C'(67)->1 'G'(71)->2 'T'(84)->3, else 4
KMER_LUT = np.full(256, 4, dtype=np.uint8)
KMER_LUT[_A], KMER_LUT[_C], KMER_LUT[_G], KMER_LUT[_T] = 0, 1, 2, 3

# idelucs/utils.py:42-43 — translation table of check_sequence
_BASEMASK = bytearray.maketrans(b"acgtuUswkmyrbdhvnSWKMYRBDHV-", b"ACGTTTNNNNNNNNNNNNNNNNNNNNNN")


def check_sequence(header: str, seq: bytearray) -> bytearray:
    """idelucs/utils.py:26-51."""
    if len(header) > 0 and (header[0] in (">", "#") or header[0].isspace()):
        raise ValueError("Bad character in sequence header")
    if "\t" in header:
        raise ValueError("tab included in header")
    masked = seq.translate(_BASEMASK, b" \t\n\r")
    stripped = masked.translate(None, b"ACGTN")
    if len(stripped) > 0:
        raise ValueError("Invalid DNA byte in sequence {}: '{}'".format(header, chr(stripped[0])))
    return masked


def kmer_counts_py(seq, k: int, counts: np.ndarray) -> None:
    """idelucs/kmers.pyx:13-50, statement for statement (slow: small inputs only).

    Accumulates INTO ``counts`` (int32[4**k]) like the reference.
    """
    k_mer = 0
    countdown = k - 1
    size = (1 << (2 * k)) - 1
    for bp in bytes(seq):
        code = int(KMER_LUT[bp])
        if code == 4:
            countdown = k
        k_mer = ((k_mer << 2) | code) & size
        if countdown == 0:
            counts[k_mer] += 1
        else:
            countdown -= 1


def kmer_counts(seq, k: int, counts: np.ndarray) -> None:
    """Vectorised equivalent of idelucs/kmers.pyx:13-50 (bit-exact; checked against
    ``kmer_counts_py`` and the live Cython build in tests/test_oracle_pinning.py).

    A window ending at i is counted iff the k bytes i-k+1..i are all in {A,C,G,T}
    (kmers.pyx:42-50: any other byte sets countdown=k, which suppresses the k windows
    that contain it; the initial countdown=k-1 suppresses the first k-1 windows).
    """
    b = np.frombuffer(bytes(seq), dtype=np.uint8)
    n = b.size
    if n < k:
        return
    code = KMER_LUT[b].astype(np.int64)
    bad = (code == 4).astype(np.int64)
    cs = np.concatenate(([0], np.cumsum(bad)))
    nbad = cs[k:] - cs[:-k]  # windows starting at 0..n-k
    idx = np.zeros(n - k + 1, dtype=np.int64)
    c2 = np.where(code == 4, 0, code)
    for t in range(k):
        idx = (idx << 2) | c2[t:n - k + 1 + t]
    idx = idx[nbad == 0]
    counts += np.bincount(idx, minlength=4 ** k).astype(counts.dtype)


class transition(object):
    """idelucs/utils.py:54-76 (same RNG calls, same order)."""

    def __init__(self, threshold):
        self.threshold = threshold

    def __call__(self, seq):
        x = np.random.random(len(seq))
        index = np.where(x < self.threshold)[0]
        mutations = {_A: _G, _G: _A, _T: _C, _C: _T, _N: _N}
        for i in index:
            seq[i] = mutations[seq[i]]


class Random_N(object):
    """idelucs/utils.py:78-95."""

    def __init__(self, n_bp):
        self.n_bp = n_bp

    def __call__(self, seq):
        index = np.random.randint(0, len(seq), self.n_bp)
        for i in index:
            seq[i] = _N


class transversion(object):
    """idelucs/utils.py:98-118."""

    def __init__(self, threshold):
        self.threshold = threshold

    def __call__(self, seq):
        x = np.random.random(len(seq))
        index = np.where(x < self.threshold)[0]
        mutations = {_A: [_T, _C], _G: [_T, _C], _T: [_A, _G], _C: [_A, _G], _N: [_N]}
        for i in index:
            seq[i] = random.choice(mutations.get(seq[i], [_N]))


class transition_transversion(object):
    """idelucs/utils.py:120-135."""

    def __init__(self, threshold_1, threshold_2):
        self.tf1 = transition(threshold_1)
        self.tf2 = transversion(threshold_2)

    def __call__(self, seq):
        self.tf1(seq)
        self.tf2(seq)


def read_fasta(fname):
    """Record iteration of idelucs/utils.py:229-261 (kmersFasta) without the counting:
    returns ``[(seq_id, checked bytearray), ...]`` in file order.

    Quirks kept: '#' lines skipped anywhere (:230); id = header minus '>' and minus its
    LAST byte (:253-254); sequence lines are ``strip()``-ped and joined (:257-259); a
    record is only flushed when the running id is non-empty (:234); the final record is
    always flushed (:259-261), so an empty file yields one empty record named "".
    """
    records = []
    lines = []
    seq_id = ""
    with open(fname, "rb") as fh:
        for line in fh:
            if line.startswith(b"#"):
                pass
            elif line.startswith(b">"):
                if seq_id != "":
                    seq = bytearray().join(lines)
                    records.append((seq_id, check_sequence(seq_id, seq)))
                    lines = []
                seq_id = line[1:-1].decode()
            else:
                lines += [line.strip()]
    seq = bytearray().join(lines)
    records.append((seq_id, check_sequence(seq_id, seq)))
    return records


def profile_from_seq(seq, k: int) -> np.ndarray:
    """idelucs/utils.py:242-250: ones (+1 pseudocount) -> kmer_counts -> counts/sum (float64)."""
    counts = np.ones(4 ** k, dtype=np.int32)
    kmer_counts(seq, k, counts)
    return counts / np.sum(counts)


def cgr(seq, k: int, CGR: np.ndarray) -> None:
    """idelucs/kmers.pyx:53-123: FCGR cell counts.  Strict alphabet (A, C, G, T; everything else resets);
    a base appended at bit n_bp of cgr_i / cgr_j with (i, j)(A, C, G, T) = (1,0) (0,0) (0,1) (1,1) (:55-88 tables);
    when k bases are held the cell (cgr_i << k) + cgr_j is incremented and the oldest base (bit 0) dropped
    (:110-123)."""
    enc = {ord("A"): (1, 0), ord("C"): (0, 0), ord("G"): (0, 1), ord("T"): (1, 1)}
    ci = cj = n_bp = 0
    for b in bytes(seq):
        e = enc.get(b)
        if e is None:
            n_bp = 0; ci = 0; cj = 0
        else:
            ci |= e[0] << n_bp
            cj |= e[1] << n_bp
            n_bp += 1
        if n_bp == k:
            CGR[(ci << k) + cj] += 1
            ci >>= 1; cj >>= 1
            n_bp -= 1


def reverse_complement(x: int, k: int) -> int:
    """idelucs/utils.py:191-206: swap the two bits of every base, complement, reverse all 2k bits =
    reversed order of the complemented bases."""
    numbits = 2 * k
    mask = 0xAAAAAAAA
    x = ((x >> 1) & (mask >> 1)) | ((x << 1) & mask)
    x = (1 << numbits) - 1 - x
    rev = 0
    for _ in range(numbits):
        rev = (rev << 1) | (x & 1)
        x >>= 1
    return rev


def kmer_rev_comp(kmer_counts_vec: np.ndarray, k: int) -> np.ndarray:
    """idelucs/utils.py:208-221, including its in-place semantics: for every canonical k-mer (kmer <= revcomp, in
    increasing order) ``v[kmer] += v[revcomp]; v[kmer] *= 0.5`` — on the int32 vector the reference passes
    (:246-247, 268-269) the float product is truncated when stored back — and the canonical entries are returned."""
    index = []
    for kmer in range(4 ** k):
        rc = reverse_complement(kmer, k)
        if kmer <= rc:
            index.append(kmer)
            kmer_counts_vec[kmer] += kmer_counts_vec[rc]
            kmer_counts_vec[kmer] = kmer_counts_vec[kmer] * 0.5   # element assignment casts back to the array dtype
    return kmer_counts_vec[index]


def kmersFasta(fname, k=6, transform=None, reduce=False):
    """idelucs/utils.py:224-277 (reduce=True: canonical folding of the +1-pseudocount int32 counts, :246-247)."""
    names, kmers = [], []
    for seq_id, seq in read_fasta(fname):
        names.append(seq_id)
        if transform:
            transform(seq)
        if reduce:
            counts = np.ones(4 ** k, dtype=np.int32)
            kmer_counts(seq, k, counts)
            counts = kmer_rev_comp(counts, k)
            kmers.append(counts / np.sum(counts))
        else:
            kmers.append(profile_from_seq(seq, k))
    return names, np.array(kmers)


def standard_scaler_fit(X: np.ndarray):
    """sklearn.preprocessing.StandardScaler.fit as the reference calls it
    (idelucs/utils.py:358-359, 404-405) — third-party, scikit-learn 1.9.0 installed here
    (pinned 1.2.1 upstream, pyproject.toml:32).  Restates
    sklearn/utils/extmath.py::_incremental_mean_and_var (first batch) and
    sklearn/preprocessing/_data.py::_is_constant_feature / _handle_zeros_in_scale:
    float64 accumulators, corrected two-pass population variance, near-constant columns
    get scale 1.  Returns (mean64, var64, scale64).
    """
    n = X.shape[0]
    X64 = X.astype(np.float64, copy=False)
    new_sum = np.sum(X64, axis=0)
    mean = new_sum / n
    temp = X64 - mean  # T == new_sum / n
    correction = np.sum(temp, axis=0)
    temp = temp ** 2
    unnorm = np.sum(temp, axis=0)
    unnorm -= correction ** 2 / n
    var = unnorm / n
    eps = np.finfo(np.float64).eps
    constant = var <= n * eps * var + (n * mean * eps) ** 2
    scale = np.sqrt(var)
    scale[constant] = 1.0
    return mean, var, scale


def standard_scaler_transform(X: np.ndarray, mean64, scale64) -> np.ndarray:
    """sklearn/preprocessing/_data.py transform(): ``X -= mean_.astype(X.dtype);
    X /= scale_.astype(X.dtype)`` on a copy (sklearn 1.9.0 casts the statistics to X.dtype)."""
    Y = np.array(X, copy=True)
    Y -= mean64.astype(Y.dtype)
    Y /= scale64.astype(Y.dtype)
    return Y


MIMIC_SCHEDULE_DOC = """idelucs/utils.py:330-351 pass schedule:
pass 0: transition_transversion(1e-2, 0.5e-2) -> t_norm ('true' column of every pair)
pass 1: transition(1e-2); pass 2: transversion(0.5e-2); pass 3..n_mimics: Random_N(20)."""


def mimic_transforms(n_mimics: int):
    """The transform objects of the n_mimics+1 passes (utils.py:330, 336, 342, 349).
    Note the reference always runs passes 0,1,2 even for n_mimics < 2."""
    tfs = [transition_transversion(1e-2, 0.5e-2), transition(1e-2), transversion(0.5e-2)]
    tfs += [Random_N(20) for _ in range(n_mimics - 2)]
    return tfs


def AugmentFasta(sequence_file, n_mimics, k=6, reduce=False, return_parts=False):
    """idelucs/utils.py:321-368."""
    _, t_norm = kmersFasta(sequence_file, k=k, transform=transition_transversion(1e-2, 0.5e-2), reduce=reduce)
    feats = []
    _, t_mut = kmersFasta(sequence_file, k=k, transform=transition(1e-2), reduce=reduce)
    feats.append(np.stack((t_norm, t_mut), axis=1))
    _, t_mut = kmersFasta(sequence_file, k=k, transform=transversion(0.5e-2), reduce=reduce)
    feats.append(np.stack((t_norm, t_mut), axis=1))
    for _ in range(n_mimics - 2):
        _, t_mut = kmersFasta(sequence_file, k=k, transform=Random_N(20), reduce=reduce)
        feats.append(np.stack((t_norm, t_mut), axis=1))
    x_train = np.concatenate(feats, axis=0).astype("float32")
    x_test = t_norm.astype("float32")
    mean, var, scale = standard_scaler_fit(x_test)
    x_train[:, 0, :] = standard_scaler_transform(x_train[:, 0, :], mean, scale)
    x_train[:, 1, :] = standard_scaler_transform(x_train[:, 1, :], mean, scale)
    if return_parts:
        return x_train, mean, var, scale
    return x_train


def inference_profiles(fasta_file, k=6):
    """idelucs/utils.py:400-405 (SequenceDataset): clean float64 profiles + its own
    StandardScaler.fit_transform in float64."""
    names, kmers = kmersFasta(fasta_file, k, None)
    mean, var, scale = standard_scaler_fit(kmers)
    return names, standard_scaler_transform(kmers, mean, scale)


class RecordingTransform(object):
    """wraps a reference transform callable; snapshots the bytearray, calls
    the transform, and records the diff as (pos, new_byte) per call, in call order."""

    def __init__(self, inner):
        self.inner = inner
        self.edits = []  # one (pos int64[], newbyte uint8[]) per sequence, in file order

    def __call__(self, seq):
        before = np.frombuffer(bytes(seq), dtype=np.uint8)
        self.inner(seq)
        after = np.frombuffer(bytes(seq), dtype=np.uint8)
        pos = np.nonzero(before != after)[0].astype(np.int64)
        self.edits.append((pos, after[pos].copy()))


def apply_edits(seq: bytearray, pos, newbyte) -> bytearray:
    out = bytearray(seq)
    for p, b in zip(pos, newbyte):
        out[int(p)] = int(b)
    return out


# ---- IIC loss (idelucs/LossFunctions.py:20-62) in numpy --------------------------------

EPS = sys.float_info.epsilon  # LossFunctions.py:20 default EPS


def compute_joint(z1: np.ndarray, z2: np.ndarray) -> np.ndarray:
    """idelucs/LossFunctions.py:49-62, in the dtype of the inputs."""
    p = np.einsum("bi,bj->ij", z1, z2)  # sum_b z1[b,i] z2[b,j]  (:57-58)
    p = (p + p.T) / np.asarray(2.0, dtype=p.dtype)  # :59
    p = p / p.sum()  # :60
    return p


def IID_loss(z1: np.ndarray, z2: np.ndarray, lamb=1.0, eps=EPS):
    """idelucs/LossFunctions.py:20-46 forward (dtype of inputs)."""
    k = z1.shape[1]
    p = compute_joint(z1, z2)
    p_i = np.repeat(p.sum(axis=1).reshape(k, 1), k, axis=1)
    p_j = np.repeat(p.sum(axis=0).reshape(1, k), k, axis=0)
    p = np.where(p < eps, np.asarray(eps, p.dtype), p)
    p_j = np.where(p_j < eps, np.asarray(eps, p.dtype), p_j)
    p_i = np.where(p_i < eps, np.asarray(eps, p.dtype), p_i)
    lam = np.asarray(lamb, dtype=p.dtype)
    loss = -p * (np.log(p) - lam * np.log(p_j) - lam * np.log(p_i))
    return loss.sum()


def IID_loss_grad(z1: np.ndarray, z2: np.ndarray, lamb=1.0, eps=EPS):
    """Closed-form gradient of IID_loss w.r.t. z1, z2 in float64 (equals the
    reference autograd, checked in tests/test_oracle_pinning.py).  Clamped entries drop
    their own-term derivative (the in-place ``p[p<EPS]=EPS`` cuts the graph)."""
    z1 = z1.astype(np.float64)
    z2 = z2.astype(np.float64)
    C = z1.shape[1]
    S = z1.T @ z2
    Ssym = (S + S.T) / 2.0
    T = Ssym.sum()
    P = Ssym / T
    pi = P.sum(axis=1)  # row marginal (expanded along columns)
    pj = P.sum(axis=0)
    Pc = np.where(P < eps, eps, P)
    pic = np.where(pi < eps, eps, pi)
    pjc = np.where(pj < eps, eps, pj)
    logP, logpi, logpj = np.log(Pc), np.log(pic), np.log(pjc)
    loss = -(Pc * (logP - lamb * logpj[None, :] - lamb * logpi[:, None])).sum()
    # dL/dP (through unclamped entries only)
    G = np.where(P < eps, 0.0, -(logP - lamb * logpj[None, :] - lamb * logpi[:, None]) - 1.0)
    # marginal terms: loss contains +lamb * Pc_ij * log pi_i  -> d/dpi_i = lamb * sum_j Pc_ij / pi_i
    gi = np.where(pi < eps, 0.0, lamb * Pc.sum(axis=1) / pic)
    gj = np.where(pj < eps, 0.0, lamb * Pc.sum(axis=0) / pjc)
    G = G + gi[:, None] + gj[None, :]
    dSsym = (G - (G * P).sum()) / T
    dS = (dSsym + dSsym.T) / 2.0
    return loss, z2 @ dS.T, z1 @ dS


# --------------------------------------------------------------------------------------
# Part B — specification of the counter-based ("rng mode") mimic generator
# --------------------------------------------------------------------------------------
# Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3",
# SC'11) — published algorithm, restated from the paper's constants.

_PHILOX_M0 = 0xD2511F53
_PHILOX_M1 = 0xCD9E8D57
_PHILOX_W0 = 0x9E3779B9
_PHILOX_W1 = 0xBB67AE85
_M32 = 0xFFFFFFFF

RNG_BLOCK = 64  # bases per Bernoulli generation block

# variant kinds (shared numbering with include/idelucs_b200.h)
KIND_CLEAN, KIND_TRANSITION, KIND_TRANSVERSION, KIND_BOTH, KIND_RANDOM_N, KIND_EXPLICIT = range(6)
STREAM_TRANSITION, STREAM_TRANSVERSION, STREAM_RANDOM_N = 0, 1, 2


def philox4x32_10(counter, key):
    c0, c1, c2, c3 = (int(c) & _M32 for c in counter)
    k0, k1 = (int(k) & _M32 for k in key)
    for _ in range(10):
        p0 = _PHILOX_M0 * c0
        p1 = _PHILOX_M1 * c2
        hi0, lo0 = p0 >> 32, p0 & _M32
        hi1, lo1 = p1 >> 32, p1 & _M32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & _M32, lo1, (hi0 ^ c3 ^ k1) & _M32, lo0
        k0 = (k0 + _PHILOX_W0) & _M32
        k1 = (k1 + _PHILOX_W1) & _M32
    return c0, c1, c2, c3


def geometric_table(p: float):
    """T[g-1] = floor((1-(1-p)^g) * 2^32) for g = 1..RNG_BLOCK; (1-p)^g by repeated IEEE
    double multiplication (deterministic on every host)."""
    q = 1.0
    one_minus_p = 1.0 - float(p)
    out = []
    for _ in range(RNG_BLOCK):
        q = q * one_minus_p
        out.append(min(int((1.0 - q) * 4294967296.0), _M32))
    return out


class _WordStream(object):
    """32-bit words w_0, w_1, ... = concatenated Philox outputs for counter
    (j, block, seq_id, variant*4+stream), j = 0, 1, ...; key = (seed_lo, seed_hi)."""

    def __init__(self, seed, seq_id, variant, stream, block):
        self.key = (seed & _M32, (seed >> 32) & _M32)
        self.c = (block, seq_id, (variant << 2) | stream)
        self.j = 0
        self.buf = []

    def next(self):
        if not self.buf:
            self.buf = list(philox4x32_10((self.j,) + self.c, self.key))
            self.j += 1
        return self.buf.pop(0)


def _gap(u, table):
    """smallest g in 1..RNG_BLOCK with u < T[g-1]; None when u >= T[RNG_BLOCK-1]."""
    for g, t in enumerate(table, start=1):
        if u < t:
            return g
    return None


def rng_bernoulli_positions(seed, seq_id, variant, stream, L, table, with_choice):
    """Positions hit by an iid Bernoulli(p) process over [0, L) generated block-wise by
    geometric gap skipping (memoryless, so restarting at each block is exact).  For the
    transversion stream each hit also consumes a second word whose top bit is the
    50/50 choice (utils.py:118 ``random.choice`` of two)."""
    hits = []
    for block in range((L + RNG_BLOCK - 1) // RNG_BLOCK):
        ws = _WordStream(seed, seq_id, variant, stream, block)
        pos = block * RNG_BLOCK - 1
        end = min((block + 1) * RNG_BLOCK, L)
        while True:
            g = _gap(ws.next(), table)
            ch = (ws.next() >> 31) if with_choice else 0
            if g is None:
                break
            pos += g
            if pos >= end:
                break
            hits.append((pos, ch))
    return hits


def rng_variant_edits(seed, seq_id, variant, kind, codes, L, p1=1e-2, p2=0.5e-2, n_bp=20):
    """Edit list ``[(pos, val)]`` (val 0..3 = set base A/C/G/T, 4 = set N) of one variant
    in rng mode.  ``codes`` is the clean sequence as uint8 codes 0..3 / 4 (N).

    Semantics mirror idelucs/utils.py:65-76, 89-95, 108-118, 132-135 on the
    check_sequence alphabet: transition A<->G, C<->T (code ^ 2), N stays N; transversion
    purine -> [T, C][choice], pyrimidine -> [A, G][choice], N stays N; both = transition
    then transversion on the mutated buffer (the transversion result only depends on the
    purine/pyrimidine class, which a transition preserves); Random_N sets n_bp uniformly
    drawn positions (with replacement) to N.
    """
    edits = {}
    if kind in (KIND_TRANSITION, KIND_BOTH):
        for pos, _ in rng_bernoulli_positions(seed, seq_id, variant, STREAM_TRANSITION, L,
                                              geometric_table(p1), False):
            c = int(codes[pos])
            if c < 4:
                edits[pos] = c ^ 2
    if kind in (KIND_TRANSVERSION, KIND_BOTH):
        for pos, ch in rng_bernoulli_positions(seed, seq_id, variant, STREAM_TRANSVERSION, L,
                                               geometric_table(p2), True):
            c = int(codes[pos])
            if c < 4:
                edits[pos] = ((1 | ((1 - ch) << 1)) if (c & 1) == 0 else (ch << 1))
    if kind == KIND_RANDOM_N and L > 0:
        ws = _WordStream(seed, seq_id, variant, STREAM_RANDOM_N, 0)
        for _ in range(n_bp):
            edits[(ws.next() * L) >> 32] = 4
    return sorted(edits.items())


ASCII_OF_CODE = np.frombuffer(b"ACGTN", dtype=np.uint8)


def codes_of_seq(seq) -> np.ndarray:
    """check_sequence alphabet -> codes 0..3, 4 for N."""
    return KMER_LUT[np.frombuffer(bytes(seq), dtype=np.uint8)]


def schedule_kinds(n_mimics: int):
    """Variant kinds of the AugmentFasta pass schedule (utils.py:330-351)."""
    return [KIND_BOTH, KIND_TRANSITION, KIND_TRANSVERSION] + [KIND_RANDOM_N] * (n_mimics - 2)


def rng_mimic_counts(seqs, k, seed, kinds, seq_id0=0, p1=1e-2, p2=0.5e-2, n_bp=20):
    """int32 counts[V, N, 4^k] of every rng-mode variant of every sequence: apply the
    variant's edits to a copy of the sequence and RECOUNT FROM SCRATCH with kmer_counts
    (the kernels instead patch the clean histogram with +-1 deltas — this is the check)."""
    out = np.zeros((len(kinds), len(seqs), 4 ** k), dtype=np.int32)
    for i, seq in enumerate(seqs):
        codes = codes_of_seq(seq)
        for v, kind in enumerate(kinds):
            edits = rng_variant_edits(seed, seq_id0 + i, v, kind, codes, len(seq), p1, p2, n_bp)
            mut = bytearray(seq)
            for pos, val in edits:
                mut[pos] = int(ASCII_OF_CODE[val])
            kmer_counts(mut, k, out[v, i])
    return out


def colstats_partials(X, rows_per_part=2048):
    """Checker restatement of the scaler partials the CUDA path exchanges between ranks:
    per block of rows the (count, mean, M2) of every column."""
    parts, ns = [], []
    for lo in range(0, X.shape[0], rows_per_part):
        blk = X[lo:lo + rows_per_part].astype(np.float64)
        m = blk.mean(axis=0)
        parts.append(np.stack([m, ((blk - m) ** 2).sum(axis=0)]))
        ns.append(float(blk.shape[0]))
    return np.stack(parts), np.asarray(ns)


def merge_partials(parts, ns):
    """Chan et al. pairwise merge in index order -> (mean, population variance)."""
    na, ma, M2 = 0.0, np.zeros(parts.shape[2]), np.zeros(parts.shape[2])
    for (mb, Mb), nb in zip(parts, ns):
        if nb <= 0:
            continue
        nt = na + nb
        delta = mb - ma
        ma = ma + delta * (nb / nt)
        M2 = M2 + Mb + delta * delta * (na * nb / nt)
        na = nt
    return ma, M2 / na

"""``kmer_counts`` with the reference's signature (idelucs/kmers.pyx:2), computed on the GPU.

The reference is a Cython loop over one sequence; this shim packs the bytes on the device
(strict kmers.pyx alphabet: only 'A','C','G','T' are bases) and runs the sm_100a counting
kernel.  For throughput use ``featurise.kmer_counts_batch`` on a whole ``SeqSet`` — one call
per sequence pays a host<->device round trip, exactly like calling any GPU op per record.
"""
import numpy as np
import torch

from . import featurise
from .seqset import SeqSet


def kmer_counts(seq, k, counts):
    """idelucs/kmers.pyx:2-50: ``counts[kmer] += 1`` for every window of k bytes that are all in
    {A,C,G,T}; accumulates INTO ``counts`` (contiguous int32[4**k]), returns None.  Like the
    typed-memoryview signature it rejects read-only buffers and non-int32 arrays."""
    mv = memoryview(seq)
    if mv.readonly:
        raise BufferError("Object is not writable.")  # unsigned char[::1] needs a writable buffer
    c = np.asarray(counts) if not isinstance(counts, np.ndarray) else counts
    if c.dtype != np.int32:
        raise ValueError("Buffer dtype mismatch, expected 'int' but got '%s'" % c.dtype.name)
    if c.ndim != 1 or not c.flags["C_CONTIGUOUS"]:
        raise ValueError("ndarray is not C-contiguous")
    if c.size < 4 ** k:
        raise ValueError("counts must hold 4**k entries")  # the reference is silent UB here
    ss = SeqSet.from_sequences([bytes(mv)], alphabet="strict")
    got = featurise.kmer_counts_batch(ss, k)[0].cpu().numpy()
    c[: 4 ** k] += got


def cgr(seq, k, CGR):
    """idelucs/kmers.pyx:53-123: accumulates the FCGR cell counts of ``seq`` INTO ``CGR`` (contiguous int32[4**k]);
    same windows and alphabet as ``kmer_counts``, cell (cgr_i << k) + cgr_j.  Computed on the GPU as the counting
    kernel followed by the index map (``featurise.cgr_batch``)."""
    mv = memoryview(seq)
    if mv.readonly:
        raise BufferError("Object is not writable.")
    c = np.asarray(CGR) if not isinstance(CGR, np.ndarray) else CGR
    if c.dtype != np.int32:
        raise ValueError("Buffer dtype mismatch, expected 'int' but got '%s'" % c.dtype.name)
    if c.ndim != 1 or not c.flags["C_CONTIGUOUS"]:
        raise ValueError("ndarray is not C-contiguous")
    if c.size < 4 ** k:
        raise ValueError("CGR must hold 4**k entries")
    ss = SeqSet.from_sequences([bytes(mv)], alphabet="strict")
    got = featurise.cgr_batch(featurise.kmer_counts_batch(ss, k), k)[0].cpu().numpy()
    c[: 4 ** k] += got

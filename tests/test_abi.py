"""CPU-side checks of the boundary: the shared library loads and exports every symbol that
include/idelucs_b200.h declares; the host-only entry points work without a GPU."""
import ctypes
import os
import re

import numpy as np

import idelucs_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "idelucs_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(idl_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from idelucs_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(lib, name), name
        assert name in _lib.PROTOTYPES, "no ctypes prototype for " + name
    assert sorted(_lib.PROTOTYPES) == declared
    assert lib.idl_abi_version() == 3


def test_geometric_table_matches_oracle():
    from idelucs_b200 import _lib
    lib = _lib.load()
    for p in (1e-2, 0.5e-2, 0.25, 1e-7):
        out = (ctypes.c_uint32 * 64)()
        assert lib.idl_geometric_table(p, out) == 0
        assert list(out) == orc.geometric_table(p)
    assert lib.idl_geometric_table(2.0, (ctypes.c_uint32 * 64)()) != 0
    assert b"bad argument" in lib.idl_last_error()


def test_argument_validation_without_gpu():
    from idelucs_b200 import _lib
    lib = _lib.load()
    assert lib.idl_iid_loss_max_clusters() == 256
    assert lib.idl_iid_loss_workspace_bytes(5) > 0 and lib.idl_iid_loss_workspace_bytes(1000) == 0
    assert lib.idl_colstats_parts(1) == 1 and lib.idl_colstats_parts(257) == 2 and lib.idl_colstats_parts(10 ** 7) <= 1024
    assert lib.idl_profiles_workspace_bytes() > 1024
    # null pointers are rejected before any CUDA call
    rc = lib.idl_kmer_counts(None, None, None, None, 1, 6, None, 0, None, 0, None)
    assert rc == 1
    rc = lib.idl_iid_loss(None, None, 4, 4, 1.0, 1e-16, None, None, None, None, None, 0, None)
    assert rc == 1


def test_no_oracle_import_in_product():
    """the product package must never import the oracle (only tests / smoke / bench may)"""
    pkg = os.path.join(ROOT, "idelucs_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py") and fn != "_smoke.py":
            src = open(os.path.join(pkg, fn)).read()
            assert "idelucs_oracle" not in src and "ref_live" not in src, fn


def _ref_loop(path):
    """the record loop of idelucs/utils.py:229-261, literally (names, joined stripped lines)"""
    lines, seq_id, names, seqs = [], "", [], []
    for line in open(path, "rb"):
        if line.startswith(b"#"):
            pass
        elif line.startswith(b">"):
            if seq_id != "":
                names.append(seq_id)
                seqs.append(b"".join(lines))
                lines = []
                seq_id = line[1:-1].decode()
            seq_id = line[1:-1].decode()
        else:
            lines += [line.strip()]
    names.append(seq_id)
    seqs.append(b"".join(lines))
    return names, seqs


FASTA_CASES = [
    b">a\nACGT\nAC\n>b\nGG\n",
    b">a\nACGT\nAC\n>b\nGG",                      # no trailing newline
    b"# comment\n>a desc\n  ACGT \t\n#mid\nTT\n\n>b\n\n>c\nA\n",   # comments, blank lines, empty record
    b">\nAC\n>x\nGT\n>y\nTT\n",                    # empty first id: its lines run into the next record
    b"ACGT\n>a\nCC\n",                            # sequence before any header
    b">a\r\nAC GT\r\nTT\r\n>b\r\nGG\r\n",          # CRLF (the id keeps its \\r minus the last byte rule), interior blank
    b">a\n\x0bAC\x0c\n>b",                         # vertical tab / form feed are stripped; header without newline at EOF
    b"",
    b"\n\n",
    b">",
    b">only\n",
]


def test_native_fasta_scanner_equals_reference_loop(tmp_path, fasta_files):
    """idl_fasta_scan / idl_fasta_extract (host entry points, no GPU) against the reference's record loop
    restated literally, on hand-made edge cases and on both bundled FASTA files"""
    from idelucs_b200.seqset import read_fasta_native, read_fasta_raw
    paths = []
    for i, content in enumerate(FASTA_CASES):
        pth = os.path.join(str(tmp_path), "case%d.fa" % i)
        with open(pth, "wb") as fh:
            fh.write(content)
        paths.append(pth)
    rng = np.random.default_rng(5)
    alph = np.frombuffer(b"ACGTNacgt-RY \t", dtype=np.uint8)
    big = []
    for r in range(300):
        if rng.random() < 0.1:
            big.append(b"#c %d\n" % r)
        big.append(b">seq_%d some text\n" % r)
        for _ in range(int(rng.integers(0, 6))):
            big.append(alph[rng.integers(0, alph.size, size=int(rng.integers(0, 90)))].tobytes() + b"\n")
    pth = os.path.join(str(tmp_path), "big.fa")
    with open(pth, "wb") as fh:
        fh.write(b"".join(big))
    paths += [pth, fasta_files["Influenza-A"], fasta_files["Actinopterygii"]]
    for pth in paths:
        want_names, want_seqs = _ref_loop(pth)
        for env in ({}, {"IDL_FASTA_CHUNK": "64", "IDL_FASTA_THREADS": "5"}, {"IDL_FASTA_CHUNK": "1", "IDL_FASTA_THREADS": "32"}):
            os.environ.update(env)          # forced: many chunks per file, cut anywhere (the parallel stitching path)
            try:
                names, flat, off = read_fasta_native(pth, pinned=False)
            finally:
                for k in env:
                    os.environ.pop(k)
            flat = flat.numpy()
            assert names == want_names, (pth, env)
            assert [flat[off[i]:off[i + 1]].tobytes() for i in range(len(names))] == want_seqs, (pth, env)
        assert (names, want_seqs) == tuple(read_fasta_raw(pth))
    for stem in ("Influenza-A", "Actinopterygii"):      # the oracle's loop (which also runs check_sequence) on the real files
        names, flat, off = read_fasta_native(fasta_files[stem], pinned=False)
        recs = orc.read_fasta(fasta_files[stem])
        assert [r[0] for r in recs] == names
        assert [bytes(r[1]) for r in recs] == [flat.numpy()[off[i]:off[i + 1]].tobytes().upper() for i in range(len(names))]


def test_native_fasta_scanner_argument_validation():
    from idelucs_b200 import _lib
    lib = _lib.load()
    n, b = ctypes.c_int64(0), ctypes.c_int64(0)
    assert lib.idl_fasta_scan(None, 5, ctypes.byref(n), ctypes.byref(b)) == 1
    buf = ctypes.create_string_buffer(b">a\nAC\n>b\nGT\n")
    assert lib.idl_fasta_scan(buf, 12, ctypes.byref(n), ctypes.byref(b)) == 0 and (n.value, b.value) == (2, 4)
    off, ho, hl = (ctypes.c_int64 * 3)(), (ctypes.c_int64 * 2)(), (ctypes.c_int64 * 2)()
    out = ctypes.create_string_buffer(4)
    assert lib.idl_fasta_extract(buf, 12, 2, out, 3, off, ho, hl) == 1          # capacity too small
    assert lib.idl_fasta_extract(buf, 12, 1, out, 4, off, ho, hl) == 1          # record count mismatch
    assert lib.idl_fasta_extract(buf, 12, 2, out, 4, off, ho, hl) == 0
    assert out.raw == b"ACGT" and list(off) == [0, 2, 4] and list(ho) == [1, 7] and list(hl) == [1, 1]


def test_native_fasta_scanner_fuzz(tmp_path):
    """random byte soup made of the bytes that matter to the record loop ('>', '#', newlines, CR, blanks, tabs,
    vertical tab, bases): the native scanner and the literal reference loop must agree on every file"""
    from idelucs_b200.seqset import read_fasta_native
    rng = np.random.default_rng(2024)
    alphabet = np.frombuffer(b">#\n\n\n\r \t\x0b\x0cACGTNacgtxyz-", dtype=np.uint8)
    pth = os.path.join(str(tmp_path), "fuzz.fa")
    for trial in range(600):
        n = int(rng.integers(0, 200))
        data = alphabet[rng.integers(0, alphabet.size, size=n)].tobytes()
        with open(pth, "wb") as fh:
            fh.write(data)
        want_names, want_seqs = _ref_loop(pth)
        env = {} if trial % 3 == 0 else {"IDL_FASTA_CHUNK": str(1 + trial % 37), "IDL_FASTA_THREADS": str(2 + trial % 9)}
        os.environ.update(env)              # two thirds of the trials: several chunks, cut anywhere
        try:
            names, flat, off = read_fasta_native(pth, pinned=False)
        finally:
            for k in env:
                os.environ.pop(k)
        flat = flat.numpy()
        assert names == want_names, (data, env)
        assert [flat[off[i]:off[i + 1]].tobytes() for i in range(len(names))] == want_seqs, (data, env)

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
    assert lib.idl_abi_version() == 1


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
    assert lib.idl_colstats_parts(1) == 1 and lib.idl_colstats_parts(2049) == 2
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

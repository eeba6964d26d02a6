import gzip
import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLD


@pytest.fixture(scope="session")
def fasta_files(tmp_path_factory):
    """Decompress the bundled golden FASTA files (copies of the reference's Example/*.fas)."""
    out = {}
    d = tmp_path_factory.mktemp("fasta")
    for stem in ("Influenza-A", "Actinopterygii"):
        dst = os.path.join(str(d), stem + ".fas")
        with gzip.open(os.path.join(GOLD, stem + ".fas.gz"), "rb") as fi, open(dst, "wb") as fo:
            shutil.copyfileobj(fi, fo)
        out[stem] = dst
    return out

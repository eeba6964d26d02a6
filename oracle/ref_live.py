"""TEST INFRASTRUCTURE ONLY — loader for the *live* reference (container-only).

Imports the unmodified reference package from /root/reference (read-only mount) so that
golden vectors can be generated from it and the oracle restatement in
``oracle/idelucs_oracle.py`` can be pinned against it.  /root/reference does not exist on
the GPU box, so nothing under ``-m gpu`` tests, ``smoke()`` or ``bench.py`` may import
this module; it is used by ``oracle/gen_golden.py`` and by the ``not gpu`` pinning tests
(which skip when the mount is absent).

The reference does ``import matplotlib.pyplot`` (idelucs/utils.py:21) which is not
installed here; only plot helpers use it, so empty stub modules are registered first.
``idelucs/utils.py:3-4`` calls ``pyximport.install()`` which compiles ``kmers.pyx`` into
``~/.pyxbld`` (not into the read-only tree).
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("IDELUCS_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "idelucs"))


def load():
    """Return the live reference package (``idelucs``)."""
    if not available():
        raise RuntimeError("reference mount not present: " + REFERENCE_ROOT)
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if "matplotlib" in sys.modules and not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import idelucs  # noqa: E402  (the reference package)
    return idelucs

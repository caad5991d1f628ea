"""Import the UNMODIFIED reference (read-only at /root/reference) inside this container.

Test tooling only: used by tests/golden/make_golden.py and tools/* to run the reference's own
numpy/scipy path on seeded synthetic RF and record golden vectors.  Never imported by the
product package; /root/reference does not exist on the GPU box.

Shims, none of which touch reference code:
  * matplotlib is absent  -> stub modules in sys.modules (lddutils.py:20-21, fdls.py:21)
  * np.float / np.int were removed from numpy (lddecode_core.py:438-439, 444, 997, 1095)
  * numpy >= 2 (NEP 50) raises OverflowError for `uint8 * 10000` in processphilipscode
    (lddecode_core.py:831, 856): the reference module's `np` global is replaced by a proxy whose
    packbits() returns int64, which restores the value-based promotion of the numpy the reference
    was written for.  numpy itself is not patched.
  * lddecode_core.loader is a module global the CLI assigns (lddecode.py:53-58); we assign an
    in-memory loader following the contract of lddutils.py:117-129.
"""
import io
import os
import sys
import types

import numpy as np

REFERENCE_DIR = os.environ.get("LDD_REFERENCE_DIR", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "lddecode_core.py"))


def load_reference():
    """Returns the reference's lddecode_core module (imported once)."""
    if "lddecode_core" in sys.modules and getattr(sys.modules["lddecode_core"], "_ldd_shimmed", False):
        return sys.modules["lddecode_core"]
    if not available():
        raise RuntimeError("reference not present at %s" % REFERENCE_DIR)
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            sys.modules[name] = m
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if not hasattr(np, "float"):
        np.float = float
    if not hasattr(np, "int"):
        np.int = int
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    import lddecode_core  # noqa: E402

    class _NumpyProxy:
        def __getattr__(self, name):
            return getattr(np, name)

        @staticmethod
        def packbits(*a, **k):
            return np.packbits(*a, **k).astype(np.int64)

    lddecode_core.np = _NumpyProxy()
    lddecode_core._ldd_shimmed = True
    return lddecode_core


class MemFile(io.BytesIO):
    """A seekable binary 'file' holding a capture (the reference checks isinstance(io.IOBase))."""


def make_array_loader(samples):
    """loader(infile, sample, readlen) over an in-memory array of samples.

    Mirrors the reference loaders' contract: returns exactly `readlen` samples starting at
    `sample`, or None when the capture is too short (lddutils.py:117-129)."""
    samples = np.asarray(samples)

    def loader(infile, sample, readlen):
        sample = int(sample)
        readlen = int(readlen)
        if sample < 0 or sample + readlen > len(samples):
            return None
        return samples[sample:sample + readlen]

    return loader

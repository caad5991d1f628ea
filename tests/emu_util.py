"""Helpers shared by the CPU-emulation tests (tests/emu) and the GPU tests.

`backend()` returns the CUDA backend when a GPU is present and, for CPU-only runs of the tests
that are written against the C ABI, the emulation backend (the product's .cu sources compiled
with g++ by tests/emu/build_emu.py).  Only tests call this; the package never does."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))

_emu = None


def emu_backend():
    global _emu
    if _emu is None:
        import build_emu
        from lddecode_b200._backend import EmuBackend
        _emu = EmuBackend(build_emu.build())
    return _emu


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False

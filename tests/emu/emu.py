"""TEST INFRASTRUCTURE ONLY: loads tests/emu/_build/libqpademu.so (the per-routine device code of qpad_b200/csrc compiled for the
host, see cuda_runtime.h in this directory) and puts it in the place of libqpadb200.so underneath qpad_b200.capi, so that the very
same host classes (capi.Ctx, Field, Part2d, Part3d, Neutral, Stage ...) and the very same test bodies run against the emulation.
Entry points that the emulation does not contain (qpg_sim_*, qpg_laser_*, qpg_wire_*) are simply absent."""
import contextlib
import ctypes as C

from qpad_b200 import capi
from . import build as _build

_lib = None
_EMU_SIGS = {"emu_launches": (C.c_long, []), "emu_barriers": (C.c_long, []), "emu_collectives": (C.c_long, []), "emu_coop_launches": (C.c_long, []),
             "emu_polls": (C.c_long, [])}


def lib():
    global _lib
    if _lib is None:
        import platform
        if platform.machine() not in ("x86_64", "AMD64"):      # the fiber switch of tests/emu/cuda_runtime.h is x86-64 assembly
            import pytest
            pytest.skip("tests/emu (host emulation of the device sources) needs an x86-64 host")
        L = C.CDLL(_build.build())
        for name, (res, args) in list(_EMU_SIGS.items()) + list(capi.SIGNATURES.items()):
            if hasattr(L, name):
                fn = getattr(L, name)
                fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


@contextlib.contextmanager
def patched():
    """qpad_b200.capi bound to the emulated library for the duration of the block"""
    old = capi._lib
    capi._lib = lib()
    try:
        yield capi
    finally:
        capi._lib = old

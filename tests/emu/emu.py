"""TEST INFRASTRUCTURE ONLY: ctypes front end of tests/emu/_build/libqpademu.so (device code of the shuffle-free kernels compiled
for the host, see cuda_runtime.h in this directory).  The classes carry the names and methods of qpad_b200.capi so a test
body reads the same against the emulation and against the GPU."""
import ctypes as C

import numpy as np

from qpad_b200 import capi
from . import build as _build

_lib = None
_vp, _i, _l, _d = C.c_void_p, C.c_int, C.c_long, C.c_double
_EMU_SIGS = {
    "emu_launches": (_l, []), "emu_barriers": (_l, []),
    "emu_ctx_create": (_i, [C.POINTER(_vp), _i, _i, _d, _d, _i]), "emu_ctx_destroy": (_i, [_vp]), "emu_ctx_launches": (_l, [_vp]),
    "emu_field_create": (_i, [C.POINTER(_vp), _vp, _i, _i, _i]), "emu_field_destroy": (_i, [_vp]),
    "emu_field_upload_f1": (_i, [_vp, _vp]), "emu_field_download_f1": (_i, [_vp, _vp]),
    "emu_field_upload_f2": (_i, [_vp, _vp]), "emu_field_download_f2": (_i, [_vp, _vp]),
    "emu_part2d_create": (_i, [C.POINTER(_vp), _vp, _d, _l]), "emu_part2d_destroy": (_i, [_vp]), "emu_part2d_npp": (_l, [_vp]),
    "emu_part2d_upload": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _l]), "emu_part2d_download": (_i, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "emu_part3d_create": (_i, [C.POINTER(_vp), _vp, _d, _d, _l]), "emu_part3d_destroy": (_i, [_vp]),
    "emu_part3d_upload": (_i, [_vp, _vp, _vp, _vp, _l]),
}


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_build.build())
        for name, (res, args) in list(_EMU_SIGS.items()) + list(capi.SIGNATURES.items()):
            if hasattr(L, name):
                fn = getattr(L, name)
                fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _chk(rc):
    if rc != 0:
        raise RuntimeError(f"emulated libqpadb200 error {rc}: {lib().qpg_last_error().decode()}")


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Ctx:
    def __init__(self, nr, max_mode, dr, dxi, field_boundary=capi.BND_OPEN):
        self.L, self.nr, self.max_mode, self.P, self.dr, self.dxi = lib(), nr, max_mode, 2 * max_mode + 1, dr, dxi
        h = _vp()
        _chk(self.L.emu_ctx_create(C.byref(h), nr, max_mode, dr, dxi, field_boundary))
        self.h = h.value

    def launch_count(self): return self.L.emu_ctx_launches(self.h)
    def solve_vpotz(self, cu, vpot): _chk(self.L.qpg_solve_vpotz(self.h, cu.h, vpot.h))
    def solve_vpott(self, cu, vpot): _chk(self.L.qpg_solve_vpott(self.h, cu.h, vpot.h))
    def close(self): self.L.qpg_vpot_release(self.h)


class Field:
    def __init__(self, ctx, dim, nzp=0, has_2d=False):
        self.ctx, self.L, self.dim, self.nzp = ctx, ctx.L, dim, nzp
        h = _vp()
        _chk(self.L.emu_field_create(C.byref(h), ctx.h, dim, nzp, int(has_2d)))
        self.h = h.value

    def upload(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.shape == (self.ctx.P, self.ctx.nr + 2, self.dim)
        _chk(self.L.emu_field_upload_f1(self.h, _ptr(a)))

    def download(self):
        a = np.zeros((self.ctx.P, self.ctx.nr + 2, self.dim))
        _chk(self.L.emu_field_download_f1(self.h, _ptr(a)))
        return a

    def upload_f2(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.shape == (self.ctx.P, self.nzp + 1, self.ctx.nr + 2, self.dim)
        _chk(self.L.emu_field_upload_f2(self.h, _ptr(a)))

    def download_f2(self):
        a = np.zeros((self.ctx.P, self.nzp + 1, self.ctx.nr + 2, self.dim))
        _chk(self.L.emu_field_download_f2(self.h, _ptr(a)))
        return a


class Part2d:
    def __init__(self, ctx, qbm, npmax):
        self.ctx, self.L, self.npmax = ctx, ctx.L, npmax
        h = _vp()
        _chk(self.L.emu_part2d_create(C.byref(h), ctx.h, qbm, npmax))
        self.h = h.value

    def npp(self): return self.L.emu_part2d_npp(self.h)

    def upload(self, x, p, gamma, psi, q):
        a = [np.ascontiguousarray(v, dtype=np.float64) for v in (x, p, gamma, psi, q)]
        _chk(self.L.emu_part2d_upload(self.h, *[_ptr(v) for v in a], len(a[4])))

    def download(self):
        n = self.npp()
        x, p, g, psi, q = np.zeros((n, 2)), np.zeros((n, 3)), np.zeros(n), np.zeros(n), np.zeros(n)
        _chk(self.L.emu_part2d_download(self.h, _ptr(x), _ptr(p), _ptr(g), _ptr(psi), _ptr(q)))
        return x, p, g, psi, q

    def clear(self): _chk(self.L.qpg_part2d_clear(self.h))
    def close(self): pass

    def exp_fac_max(self):
        v = _d()
        _chk(self.L.qpg_part2d_exp_fac_max(self.h, C.byref(v)))
        return v.value

    def clamp_exp_fac(self, exp_fac_clamped): _chk(self.L.qpg_part2d_clamp_exp_fac(self.h, exp_fac_clamped))


class Part3d:
    def __init__(self, ctx, qbm, dt, npmax):
        self.ctx, self.L, self.npmax = ctx, ctx.L, npmax
        h = _vp()
        _chk(self.L.emu_part3d_create(C.byref(h), ctx.h, qbm, dt, npmax))
        self.h = h.value

    def upload(self, x, p, q):
        a = [np.ascontiguousarray(v, dtype=np.float64) for v in (x, p, q)]
        _chk(self.L.emu_part3d_upload(self.h, *[_ptr(v) for v in a], len(a[2])))


def subcyc_step(exp_fac, exp_fac_max, dt, dt_min):
    return capi.subcyc_step(exp_fac, exp_fac_max, dt, dt_min, L=lib())


class Stage(capi.Stage):
    """capi.Stage over the emulated library (same code: only ctx.L differs)"""


class Neutral:
    """capi.Neutral over the emulated library"""

    def __init__(self, ctx, element, ion_max, ppc, num_theta, q=-1.0, m=1.0, density=1.0, n0=1.0e17, dt_xi=None):
        self.ctx, self.L, self.num_theta = ctx, ctx.L, num_theta
        h = _vp()
        _chk(self.L.qpg_neutral_create(C.byref(h), ctx.h, element, ion_max, ppc[0], ppc[1], num_theta, q, m, density, n0, ctx.dxi if dt_xi is None else dt_xi))
        self.h = h.value
        self.multi_max = self.L.qpg_neutral_multi_max(self.h)
        cap = ctx.nr * num_theta * ppc[0] * ppc[1] + 64
        self.part, self.part_add = Part2d(ctx, q / m, cap), Part2d(ctx, q / m, cap)

    def update(self, e): _chk(self.L.qpg_neutral_update(self.h, e.h, self.part.h, self.part_add.h))

    def renew(self):
        _chk(self.L.qpg_neutral_reset(self.h))
        self.part.clear(); self.part_add.clear()

    def levels(self):
        out = np.zeros((self.multi_max + 2, self.num_theta, self.ctx.nr))
        _chk(self.L.qpg_neutral_levels(self.h, _ptr(out)))
        return out

    def close(self): _chk(self.L.qpg_neutral_destroy(self.h))

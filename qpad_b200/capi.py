"""ctypes binding of libqpadb200.so (include/qpad_b200.h) -- the same C-ABI a Fortran ISO_C_BINDING shim binds.

Thin object wrappers (Ctx, Field, Part2d, Part3d, Sim) whose method names follow the reference's type-bound
procedures (species/part2d_class.f03:61-81, fields/field_class.f03, beam/part3d_class.f03:71-84).  numpy arrays use
the reference's host layouts (see the header).  There is NO CPU fallback: if the shared library is missing, or no
CUDA device is present, construction raises.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# QPG_LIB: an alternative build of the SAME library (A/B experiments of compile-time options, tools/ab_bench.py); never a fallback
LIB_PATH = os.environ.get("QPG_LIB") or os.path.join(_HERE, "libqpadb200.so")

BND_ZERO, BND_OPEN = 2, 3
PUSH2_STD, PUSH2_ROBUST, PUSH2_STD_PGC, PUSH2_ROBUST_PGC = 0, 1, 4, 5
PUSH3_REDUCED, PUSH3_BORIS = 1, 2
COPY_1TO2, COPY_2TO1 = 0, 1
CONV_RECORD, CONV_COMPARE = 0, 1


class QpadError(RuntimeError):
    pass


class SimParams(C.Structure):
    _fields_ = [("nr", C.c_int), ("nz_total", C.c_int), ("noff2", C.c_int), ("nzp", C.c_int), ("max_mode", C.c_int),
                ("field_boundary", C.c_int), ("iter_max", C.c_int), ("sort_freq", C.c_int),
                ("dr", C.c_double), ("dxi", C.c_double), ("dt", C.c_double), ("iter_reltol", C.c_double),
                ("iter_abstol", C.c_double), ("relax_fac", C.c_double),
                ("sp_qbm", C.c_double), ("sp_npmax", C.c_long),
                ("beam_push_type", C.c_int), ("beam_evol", C.c_int), ("beam_qbm", C.c_double), ("beam_npmax", C.c_long),
                ("use_graph", C.c_int), ("sp_push_std", C.c_int), ("sp_push_pgc", C.c_int), ("laser_iter", C.c_int),
                ("laser_k0", C.c_double), ("sp_ppc_r", C.c_int)]


_lib = None
_vp, _i, _l, _d = C.c_void_p, C.c_int, C.c_long, C.c_double
_pd = C.POINTER(C.c_double)
_pl = C.POINTER(C.c_long)
_pi = C.POINTER(C.c_int)

# name -> (restype, argtypes); mirrors include/qpad_b200.h one to one
SIGNATURES = {
    "qpg_last_error": (C.c_char_p, []),
    "qpg_version": (_i, []),
    "qpg_ctx_create": (_i, [C.POINTER(_vp), _i, _vp, _i, _i, _d, _d, _i, _d]),
    "qpg_ctx_destroy": (_i, [_vp]),
    "qpg_ctx_sync": (_i, [_vp]),
    "qpg_tprof_enable": (_i, [_vp, _i]),
    "qpg_tprof_reset": (_i, [_vp]),
    "qpg_tprof_get": (_i, [_vp, C.c_char_p, _pd, _pl]),
    "qpg_launch_count": (_l, [_vp]),
    "qpg_field_create": (_i, [C.POINTER(_vp), _vp, _i, _i, _i]),
    "qpg_field_destroy": (_i, [_vp]),
    "qpg_field_dim": (_i, [_vp]),
    "qpg_field_fill": (_i, [_vp, _d]),
    "qpg_field_fill_f2": (_i, [_vp, _d]),
    "qpg_field_copy": (_i, [_vp, _vp]),
    "qpg_field_copy_slice": (_i, [_vp, _i, _i]),
    "qpg_field_add": (_i, [_vp, _vp]),
    "qpg_field_add3": (_i, [_vp, _vp, _vp]),
    "qpg_field_add_dim": (_i, [_vp, _vp, _i, _pi, _pi]),
    "qpg_field_add_f2": (_i, [_vp, _vp]),
    "qpg_field_scale": (_i, [_vp, _d]),
    "qpg_field_smooth": (_i, [_vp, _i, _i]),
    "qpg_field_upload_f1": (_i, [_vp, _vp]),
    "qpg_field_download_f1": (_i, [_vp, _vp]),
    "qpg_field_upload_f2": (_i, [_vp, _vp]),
    "qpg_field_download_f2": (_i, [_vp, _vp]),
    "qpg_field_pack": (_i, [_vp, _i, _vp]),
    "qpg_field_unpack": (_i, [_vp, _i, _vp, _i]),
    "qpg_field_wire_count": (_l, [_vp]),
    "qpg_field_lineout": (_i, [_vp, _i, _i, _i, _vp]),
    "qpg_solve_psi": (_i, [_vp, _vp, _vp]),
    "qpg_solve_bt": (_i, [_vp, _vp, _vp]),
    "qpg_solve_bz": (_i, [_vp, _vp, _vp]),
    "qpg_solve_bt_iter": (_i, [_vp, _vp, _vp, _vp]),
    "qpg_solve_ez": (_i, [_vp, _vp, _vp]),
    "qpg_solve_et": (_i, [_vp, _vp, _vp, _vp]),
    "qpg_solve_et_beam": (_i, [_vp, _vp, _vp]),
    "qpg_solve_djdxi": (_i, [_vp, _vp, _vp, _vp]),
    "qpg_bperp_residual": (_i, [_vp, _vp, _i, _i, _pd, _pd]),
    "qpg_part2d_create": (_i, [C.POINTER(_vp), _vp, _d, _l]),
    "qpg_part2d_destroy": (_i, [_vp]),
    "qpg_part2d_upload": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _l]),
    "qpg_part2d_download": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _pl]),
    "qpg_part2d_npp": (_i, [_vp, _pl]),
    "qpg_part2d_snapshot": (_i, [_vp]),
    "qpg_part2d_renew": (_i, [_vp]),
    "qpg_part2d_qdeposit": (_i, [_vp, _vp]),
    "qpg_part2d_amjdeposit": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _d]),
    "qpg_part2d_push_u": (_i, [_vp, _i, _vp, _vp, _d]),
    "qpg_part2d_amjdeposit_pgc": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _d]),
    "qpg_part2d_push_u_pgc": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _d]),
    "qpg_part2d_push_x": (_i, [_vp, _d]),
    "qpg_part2d_interp_psi": (_i, [_vp, _vp]),
    "qpg_part2d_update_bound": (_i, [_vp]),
    "qpg_part2d_move": (_i, [_vp]),
    "qpg_part2d_sort": (_i, [_vp]),
    "qpg_part2d_sort_index": (_i, [_vp, _vp, _vp]),
    "qpg_part2d_pack": (_i, [_vp, _vp]),
    "qpg_part2d_unpack": (_i, [_vp, _vp]),
    "qpg_part2d_wire_count": (_l, [_vp]),
    "qpg_part3d_create": (_i, [C.POINTER(_vp), _vp, _d, _d, _l, _i, _i, _i]),
    "qpg_part3d_destroy": (_i, [_vp]),
    "qpg_part3d_upload": (_i, [_vp, _vp, _vp, _vp, _l]),
    "qpg_part3d_download": (_i, [_vp, _vp, _vp, _vp, _pl]),
    "qpg_part3d_qdeposit": (_i, [_vp, _vp]),
    "qpg_part3d_push": (_i, [_vp, _i, _vp, _vp]),
    "qpg_part3d_update_bound": (_i, [_vp]),
    "qpg_part3d_push_interior": (_i, [_vp, _i, _vp, _vp]),
    "qpg_part3d_push_edge": (_i, [_vp, _i, _vp, _vp]),
    "qpg_sim_beam_push_interior": (_i, [_vp]),
    "qpg_sim_beam_push_edge": (_i, [_vp]),
    "qpg_sim_beam_qdp_part": (_i, [_vp, _i]),
    "qpg_part3d_qdeposit_part": (_i, [_vp, _vp, _i]),
    "qpg_part3d_pack_forward": (_i, [_vp, _vp]),
    "qpg_part3d_unpack": (_i, [_vp, _vp]),
    "qpg_part3d_wire_cap": (_l, [_vp]),
    "qpg_part3d_set_wire_cap": (_i, [_vp, _l]),
    "qpg_part3d_enable_spin": (_i, [_vp, _d]),
    "qpg_part3d_has_spin": (_i, [_vp]),
    "qpg_part3d_upload_spin": (_i, [_vp, _vp, _l]),
    "qpg_part3d_download_spin": (_i, [_vp, _vp, C.POINTER(_l)]),
    "qpg_part3d_wire_count": (_l, [_vp]),
    "qpg_sim_create": (_i, [C.POINTER(_vp), _i, _vp, C.POINTER(SimParams)]),
    "qpg_sim_destroy": (_i, [_vp]),
    "qpg_sim_ctx": (_vp, [_vp]),
    "qpg_sim_field": (_vp, [_vp, C.c_char_p]),
    "qpg_sim_species": (_vp, [_vp]),
    "qpg_sim_beam": (_vp, [_vp]),
    "qpg_sim_init_species": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _l]),
    "qpg_sim_beam_qdp_begin": (_i, [_vp]),
    "qpg_sim_beam_qdp_end": (_i, [_vp]),
    "qpg_sim_begin_step": (_i, [_vp]),
    "qpg_sim_set_back_handoff": (_i, [_vp, _vp, _vp, _vp, C.c_uint]),
    "qpg_sim_beam_qdp_raw": (_i, [_vp]),
    "qpg_sim_beam_qdp_fix": (_i, [_vp]),
    "qpg_sim_begin_step_zero": (_i, [_vp]),
    "qpg_sim_begin_step_add": (_i, [_vp]),
    "qpg_part3d_qdeposit_raw": (_i, [_vp, _vp]),
    "qpg_part3d_qdeposit_fix": (_i, [_vp, _vp]),
    "qpg_sim_run_slices": (_i, [_vp, _i, _i]),
    "qpg_sim_beam_push": (_i, [_vp]),
    "qpg_sim_renew": (_i, [_vp]),
    "qpg_sim_stats": (_i, [_vp, _pl, _pl, _pl]),
    "qpg_sim_set_graph": (_i, [_vp, _i]),
    "qpg_sim_set_fused": (_i, [_vp, _i]),
    "qpg_sim_set_sweep": (_i, [_vp, _i]),
    "qpg_sim_set_sweep_ctas": (_i, [_vp, _i]),
    "qpg_sim_sweep_profile": (_i, [_vp, _pd, _i]),
    "qpg_sim_laser": (_vp, [_vp]),
    "qpg_sim_laser_advance": (_i, [_vp]),
    "qpg_laser_create": (_i, [C.POINTER(_vp), _vp, _i, _d, _d, _i]),
    "qpg_laser_destroy": (_i, [_vp]),
    "qpg_laser_volume_size": (_l, [_vp]),
    "qpg_laser_upload": (_i, [_vp, _vp, _vp]),
    "qpg_laser_download": (_i, [_vp, _vp, _vp]),
    "qpg_laser_field": (_vp, [_vp, _i]),
    "qpg_laser_slice": (_i, [_vp, _i]),
    "qpg_laser_deposit_chi": (_i, [_vp, _vp, _i, _d]),
    "qpg_laser_advance": (_i, [_vp]),
    "qpg_laser_sync": (_i, [_vp]),
    "qpg_laser_guard_size": (_l, [_vp]),
    "qpg_laser_set_handoff": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "qpg_sim_set_laser_overlap": (_i, [_vp, _i]),
    "qpg_sim_set_graph_unroll": (_i, [_vp, _i]),
    "qpg_neutral_create": (_i, [C.POINTER(_vp), _vp, _i, _i, _i, _i, _i, _d, _d, _d, _d, _d]),
    "qpg_neutral_destroy": (_i, [_vp]),
    "qpg_sim_neutral_wire_count": (_l, [_vp]),
    "qpg_sim_neutral_pack": (_i, [_vp, _vp]),
    "qpg_sim_neutral_unpack": (_i, [_vp, _vp]),
    "qpg_neutral_reset": (_i, [_vp]),
    "qpg_neutral_multi_max": (_i, [_vp]),
    "qpg_neutral_update": (_i, [_vp, _vp, _vp, _vp]),
    "qpg_neutral_levels": (_i, [_vp, _vp]),
    "qpg_part2d_clear": (_i, [_vp]),
    "qpg_debug_fastmath": (_i, [_vp, _l, _vp, _vp, _vp]),
    "qpg_sim_attach_neutral": (_i, [_vp, _vp, _vp, _vp]),
    "qpg_sim_set_subcyc": (_i, [_vp, _i, _d, _d, _d]),
    "qpg_sim_subcycles": (_l, [_vp]),
    "qpg_part2d_exp_fac_max": (_i, [_vp, _pd]),
    "qpg_part2d_clamp_exp_fac": (_i, [_vp, _d]),
    "qpg_subcyc_step": (_i, [_d, _d, _d, _d, _pd, _pi]),
    "qpg_solve_vpotz": (_i, [_vp, _vp, _vp]),
    "qpg_solve_vpott": (_i, [_vp, _vp, _vp]),
    "qpg_vpot_release": (_i, [_vp]),
    "qpg_stage_create": (_i, [C.POINTER(_vp), _vp, _l]),
    "qpg_stage_destroy": (_i, [_vp]),
    "qpg_stage_field": (_i, [_vp, _vp, _pl]),
    "qpg_stage_part2d": (_i, [_vp, _vp, _i, _pl]),
    "qpg_stage_part3d": (_i, [_vp, _vp, _i, _d, _pl]),
    "qpg_stage_wait": (_i, [_vp, C.POINTER(_pd), _pl]),
    "qpg_sim_slice_trace": (_i, [_vp, _pd, _pi]),
    "qpg_sim_debug_abort": (_i, [_vp]),
    "qpg_wire_alloc": (_i, [C.POINTER(_vp), _l]),
    "qpg_wire_free": (_i, [_vp]),
    "qpg_wire_export": (_i, [_vp, C.c_char_p]),
    "qpg_wire_import": (_i, [C.c_char_p, C.POINTER(_vp)]),
    "qpg_wire_unmap": (_i, [_vp]),
    "qpg_stream_signal": (_i, [_vp, _vp, C.c_uint]),
    "qpg_stream_wait": (_i, [_vp, _vp, C.c_uint]),
    "qpg_stream_wait_unless_empty": (_i, [_vp, _vp, _vp, C.c_uint]),
    "qpg_part3d_count_ptr": (_vp, [_vp]),
    "qpg_stream_wait_is_memop": (_i, []),
}


def load():
    """Load libqpadb200.so (no fallback).  Safe without a GPU: only symbols are resolved."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise QpadError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def _chk(rc):
    if rc != 0:
        raise QpadError(f"libqpadb200 error {rc}: {load().qpg_last_error().decode()}")


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Ctx:
    def __init__(self, nr, max_mode, dr, dxi, field_boundary=BND_OPEN, relax_fac=-1.0, device=0, stream=None, handle=None):
        self.L = load()
        self.nr, self.max_mode, self.P, self.dr, self.dxi = nr, max_mode, 2 * max_mode + 1, dr, dxi
        self._own = handle is None
        if handle is None:
            h = _vp()
            _chk(self.L.qpg_ctx_create(C.byref(h), device, stream, nr, max_mode, dr, dxi, field_boundary, relax_fac))
            handle = h.value
        self.h = handle

    def close(self):
        if getattr(self, "h", None) and self._own:
            self.L.qpg_vpot_release(self.h)
            self.L.qpg_ctx_destroy(self.h)
        self.h = None

    __del__ = close

    def sync(self):
        _chk(self.L.qpg_ctx_sync(self.h))

    def launch_count(self):
        return self.L.qpg_launch_count(self.h)

    def tprof_enable(self, on=True):
        _chk(self.L.qpg_tprof_enable(self.h, int(on)))

    def tprof_reset(self):
        _chk(self.L.qpg_tprof_reset(self.h))

    def tprof_get(self, event):
        ms, n = _d(), _l()
        _chk(self.L.qpg_tprof_get(self.h, event.encode(), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    # field solves, named like the reference generics
    def solve_psi(self, q, psi): _chk(self.L.qpg_solve_psi(self.h, q.h, psi.h))
    def solve_bt(self, qb, b): _chk(self.L.qpg_solve_bt(self.h, qb.h, b.h))
    def solve_bz(self, cu, b): _chk(self.L.qpg_solve_bz(self.h, cu.h, b.h))
    def solve_bt_iter(self, dcu, cu, b): _chk(self.L.qpg_solve_bt_iter(self.h, dcu.h, cu.h, b.h))
    def solve_ez(self, cu, e): _chk(self.L.qpg_solve_ez(self.h, cu.h, e.h))
    def solve_et(self, b, psi, e): _chk(self.L.qpg_solve_et(self.h, b.h, psi.h, e.h))
    def solve_et_beam(self, b, e): _chk(self.L.qpg_solve_et_beam(self.h, b.h, e.h))
    def solve_djdxi(self, acu, amu, dcu): _chk(self.L.qpg_solve_djdxi(self.h, acu.h, amu.h, dcu.h))
    def debug_fastmath(self, x):
        """(fast_rcp(x), fast_sqrt(x)) of the momentum arithmetic evaluated on the device"""
        x = _f64(x); r, q = np.zeros_like(x), np.zeros_like(x)
        _chk(self.L.qpg_debug_fastmath(self.h, x.size, _ptr(x), _ptr(r), _ptr(q)))
        return r, q

    def solve_vpotz(self, cu, vpot): _chk(self.L.qpg_solve_vpotz(self.h, cu.h, vpot.h))     # field_vpot%solve_vpotz
    def solve_vpott(self, cu, vpot): _chk(self.L.qpg_solve_vpott(self.h, cu.h, vpot.h))     # field_vpot%solve_vpott

    def convergence_tester(self, fld, dim, op):
        rel, ab = _d(), _d()
        _chk(self.L.qpg_bperp_residual(self.h, fld.h, dim, op, C.byref(rel), C.byref(ab)))
        return rel.value, ab.value


class Field:
    def __init__(self, ctx, dim, nzp=0, has_2d=False, handle=None):
        self.ctx, self.L, self.dim, self.nzp, self.has_2d = ctx, ctx.L, dim, nzp, has_2d
        self._own = handle is None
        if handle is None:
            h = _vp()
            _chk(self.L.qpg_field_create(C.byref(h), ctx.h, dim, nzp, int(has_2d)))
            handle = h.value
        self.h = handle

    def close(self):
        if getattr(self, "h", None) and self._own and self.ctx.h:
            self.L.qpg_field_destroy(self.h)
        self.h = None

    __del__ = close

    def fill(self, v=0.0): _chk(self.L.qpg_field_fill(self.h, v))
    def fill_f2(self, v=0.0): _chk(self.L.qpg_field_fill_f2(self.h, v))
    def copy_to(self, dst): _chk(self.L.qpg_field_copy(self.h, dst.h))
    def copy_slice(self, idx, direction): _chk(self.L.qpg_field_copy_slice(self.h, idx, direction))
    def add_to(self, b): _chk(self.L.qpg_field_add(self.h, b.h))
    def scale(self, s): _chk(self.L.qpg_field_scale(self.h, s))
    def smooth(self, order, kind): _chk(self.L.qpg_field_smooth(self.h, order, kind))

    def add_dim_to(self, b, adim, bdim):
        a1 = (C.c_int * len(adim))(*adim)
        b1 = (C.c_int * len(bdim))(*bdim)
        _chk(self.L.qpg_field_add_dim(self.h, b.h, len(adim), a1, b1))

    def add_f2_to(self, b): _chk(self.L.qpg_field_add_f2(self.h, b.h))

    @staticmethod
    def add3(a1, a2, a3): _chk(a1.L.qpg_field_add3(a1.h, a2.h, a3.h))

    def upload(self, f1):
        f1 = _f64(f1)
        assert f1.shape == (self.ctx.P, self.ctx.nr + 2, self.dim), f1.shape
        _chk(self.L.qpg_field_upload_f1(self.h, _ptr(f1)))

    def download(self):
        out = np.zeros((self.ctx.P, self.ctx.nr + 2, self.dim))
        _chk(self.L.qpg_field_download_f1(self.h, _ptr(out)))
        return out

    def upload_f2(self, f2):
        f2 = _f64(f2)
        assert f2.shape == (self.ctx.P, self.nzp + 1, self.ctx.nr + 2, self.dim), f2.shape
        _chk(self.L.qpg_field_upload_f2(self.h, _ptr(f2)))

    def download_f2(self):
        out = np.zeros((self.ctx.P, self.nzp + 1, self.ctx.nr + 2, self.dim))
        _chk(self.L.qpg_field_download_f2(self.h, _ptr(out)))
        return out

    def lineout(self, comp, plane=0, node=1):
        out = np.zeros(self.nzp)
        _chk(self.L.qpg_field_lineout(self.h, comp, plane, node, _ptr(out)))
        return out

    def wire_count(self): return self.L.qpg_field_wire_count(self.h)
    def pack(self, slice_idx, dev_ptr): _chk(self.L.qpg_field_pack(self.h, slice_idx, dev_ptr))
    def unpack(self, slice_idx, dev_ptr, add=False): _chk(self.L.qpg_field_unpack(self.h, slice_idx, dev_ptr, int(add)))


class Part2d:
    def __init__(self, ctx, qbm, npmax, handle=None):
        self.ctx, self.L, self.qbm, self.npmax = ctx, ctx.L, qbm, npmax
        self._own = handle is None
        if handle is None:
            h = _vp()
            _chk(self.L.qpg_part2d_create(C.byref(h), ctx.h, qbm, npmax))
            handle = h.value
        self.h = handle

    def close(self):
        if getattr(self, "h", None) and self._own and self.ctx.h:
            self.L.qpg_part2d_destroy(self.h)
        self.h = None

    __del__ = close

    def upload(self, x, p, gamma, psi, q):
        x, p, gamma, psi, q = map(_f64, (x, p, gamma, psi, q))
        _chk(self.L.qpg_part2d_upload(self.h, _ptr(x), _ptr(p), _ptr(gamma), _ptr(psi), _ptr(q), len(q)))

    def npp(self):
        n = _l()
        _chk(self.L.qpg_part2d_npp(self.h, C.byref(n)))
        return n.value

    def download(self):
        n = self.npp()
        x, p = np.zeros((n, 2)), np.zeros((n, 3))
        g, psi, q = np.zeros(n), np.zeros(n), np.zeros(n)
        m = _l()
        _chk(self.L.qpg_part2d_download(self.h, _ptr(x), _ptr(p), _ptr(g), _ptr(psi), _ptr(q), C.byref(m)))
        assert m.value == n
        return x, p, g, psi, q

    def snapshot(self): _chk(self.L.qpg_part2d_snapshot(self.h))
    def renew(self): _chk(self.L.qpg_part2d_renew(self.h))
    def qdeposit(self, q): _chk(self.L.qpg_part2d_qdeposit(self.h, q.h))

    def amjdeposit_robust(self, ef, bf, cu, amu, dcu, dt):
        _chk(self.L.qpg_part2d_amjdeposit(self.h, PUSH2_ROBUST, ef.h, bf.h, cu.h, amu.h, dcu.h, dt))

    def amjdeposit_std(self, ef, bf, cu, amu, dcu, dt):
        _chk(self.L.qpg_part2d_amjdeposit(self.h, PUSH2_STD, ef.h, bf.h, cu.h, amu.h, dcu.h, dt))

    def push_u_robust(self, ef, bf, dt): _chk(self.L.qpg_part2d_push_u(self.h, PUSH2_ROBUST, ef.h, bf.h, dt))
    def push_u_std(self, ef, bf, dt): _chk(self.L.qpg_part2d_push_u(self.h, PUSH2_STD, ef.h, bf.h, dt))
    def interp_psi(self, psi): _chk(self.L.qpg_part2d_interp_psi(self.h, psi.h))

    def amjdeposit_pgc(self, push_type, ef, bf, laser, cu, amu, dcu, dt):
        """laser = (a_r, a_i, grad a_r, grad a_i) fields"""
        _chk(self.L.qpg_part2d_amjdeposit_pgc(self.h, push_type, ef.h, bf.h, *(f.h for f in laser), cu.h, amu.h, dcu.h, dt))

    def push_u_pgc(self, push_type, ef, bf, laser, dt):
        _chk(self.L.qpg_part2d_push_u_pgc(self.h, push_type, ef.h, bf.h, *(f.h for f in laser), dt))
    def push_x(self, dt): _chk(self.L.qpg_part2d_push_x(self.h, dt))
    def update_bound(self): _chk(self.L.qpg_part2d_update_bound(self.h))
    def move(self): _chk(self.L.qpg_part2d_move(self.h))                 # move_part2d_comm (species/part2d_comm.f03:147)
    def sort(self): _chk(self.L.qpg_part2d_sort(self.h))

    def sort_index(self):
        n = self.npp()
        ix, ip = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
        _chk(self.L.qpg_part2d_sort_index(self.h, _ptr(ix), _ptr(ip)))
        return ix, ip

    def clear(self): _chk(self.L.qpg_part2d_clear(self.h))

    def exp_fac_max(self):
        """part2d_subcyc%get_exp_fac_max (proj_subcyc/part2d_subcyc_class.f03:28); synchronises"""
        v = _d()
        _chk(self.L.qpg_part2d_exp_fac_max(self.h, C.byref(v)))
        return v.value

    def clamp_exp_fac(self, exp_fac_clamped): _chk(self.L.qpg_part2d_clamp_exp_fac(self.h, exp_fac_clamped))
    def wire_count(self): return self.L.qpg_part2d_wire_count(self.h)
    def pack(self, dev_ptr): _chk(self.L.qpg_part2d_pack(self.h, dev_ptr))
    def unpack(self, dev_ptr): _chk(self.L.qpg_part2d_unpack(self.h, dev_ptr))


class Part3d:
    def __init__(self, ctx, qbm, dt, npmax, nz_total, noff2, nzp, handle=None):
        self.ctx, self.L = ctx, ctx.L
        self._own = handle is None
        if handle is None:
            h = _vp()
            _chk(self.L.qpg_part3d_create(C.byref(h), ctx.h, qbm, dt, npmax, nz_total, noff2, nzp))
            handle = h.value
        self.h = handle

    def close(self):
        if getattr(self, "h", None) and self._own and self.ctx.h:
            self.L.qpg_part3d_destroy(self.h)
        self.h = None

    __del__ = close

    def upload(self, x, p, q):
        x, p, q = map(_f64, (x, p, q))
        _chk(self.L.qpg_part3d_upload(self.h, _ptr(x), _ptr(p), _ptr(q), len(q)))

    def download(self):
        n = _l()
        _chk(self.L.qpg_part3d_download(self.h, None, None, None, C.byref(n)))
        x, p, q = np.zeros((n.value, 3)), np.zeros((n.value, 3)), np.zeros(n.value)
        if n.value:
            _chk(self.L.qpg_part3d_download(self.h, _ptr(x), _ptr(p), _ptr(q), C.byref(n)))
        return x, p, q

    def enable_spin(self, amm):
        """part3d%has_spin with the anomalous magnetic moment `amm` (init_part3d :117-135)"""
        _chk(self.L.qpg_part3d_enable_spin(self.h, float(amm)))

    def has_spin(self): return bool(self.L.qpg_part3d_has_spin(self.h))

    def upload_spin(self, s):
        s = _f64(s)
        _chk(self.L.qpg_part3d_upload_spin(self.h, _ptr(s), len(s)))

    def download_spin(self):
        n = _l()
        _chk(self.L.qpg_part3d_download_spin(self.h, None, C.byref(n)))
        s = np.zeros((n.value, 3))
        if n.value:
            _chk(self.L.qpg_part3d_download_spin(self.h, _ptr(s), C.byref(n)))
        return s

    def wire_count(self): return int(self.L.qpg_part3d_wire_count(self.h))     # doubles of a hand-off message (7 or 10 reals per particle + the count)
    def qdeposit(self, q): _chk(self.L.qpg_part3d_qdeposit(self.h, q.h))
    def push(self, push_type, ef, bf): _chk(self.L.qpg_part3d_push(self.h, push_type, ef.h, bf.h))
    def push_interior(self, push_type, ef, bf): _chk(self.L.qpg_part3d_push_interior(self.h, push_type, ef.h, bf.h))
    def push_edge(self, push_type, ef, bf): _chk(self.L.qpg_part3d_push_edge(self.h, push_type, ef.h, bf.h))
    def update_bound(self): _chk(self.L.qpg_part3d_update_bound(self.h))
    def count_ptr(self): return self.L.qpg_part3d_count_ptr(self.h)          # device address of the live particle count
    def wire_cap(self): return self.L.qpg_part3d_wire_cap(self.h)
    def set_wire_cap(self, cap): _chk(self.L.qpg_part3d_set_wire_cap(self.h, int(cap)))
    def pack_forward(self, dev_ptr): _chk(self.L.qpg_part3d_pack_forward(self.h, dev_ptr))
    def unpack(self, dev_ptr): _chk(self.L.qpg_part3d_unpack(self.h, dev_ptr))


def subcyc_step(exp_fac, exp_fac_max, dt, dt_min, L=None):
    """simulation_subcyc_class.f03:431-451 -> (dt_subcyc, n_subcyc)"""
    dts, n = _d(), _i()
    _chk((L or load()).qpg_subcyc_step(exp_fac, exp_fac_max, dt, dt_min, C.byref(dts), C.byref(n)))
    return dts.value, n.value


class Stage:
    """Diagnostics staging (csrc/diag.cu): device re-layout + asynchronous copy into a pinned host buffer.  `field`, `part2d`,
    `part3d` enqueue; `wait` blocks and returns the datasets in the layouts the HDF5 writer takes (hdf5io_class.f03:591, :1027,
    :1220)."""

    def __init__(self, ctx, capacity_doubles):
        self.ctx, self.L = ctx, ctx.L
        h = _vp()
        _chk(self.L.qpg_stage_create(C.byref(h), ctx.h, capacity_doubles))
        self.h, self._shape = h.value, None

    def close(self):
        if getattr(self, "h", None) and self.ctx.h:
            self.L.qpg_stage_destroy(self.h)
        self.h = None

    __del__ = close

    def field(self, f):
        _chk(self.L.qpg_stage_field(self.h, f.h, None))
        self._shape = ("field", (self.ctx.P, f.dim, f.nzp, self.ctx.nr))

    def part2d(self, p, dspl):
        st = _l()
        _chk(self.L.qpg_stage_part2d(self.h, p.h, dspl, C.byref(st)))
        self._shape = ("part", (6, st.value))

    def part3d(self, p, dspl, z0=0.0):
        st = _l()
        _chk(self.L.qpg_stage_part3d(self.h, p.h, dspl, z0, C.byref(st)))
        self._shape = ("part", (7, st.value))

    def wait(self):
        """field: array [plane][comp][slice][node]; particles: array [dataset][tnpp]"""
        host, n = _pd(), _l()
        _chk(self.L.qpg_stage_wait(self.h, C.byref(host), C.byref(n)))
        a = np.ctypeslib.as_array(host, shape=(n.value,)).copy()
        kind, shape = self._shape
        if kind == "field":
            return a.reshape(shape)
        tnpp = int(a[0])
        return a[1:].reshape(shape)[:, :tnpp]


class Neutral:
    """neutral (species/neutral_class.f03): ionisation levels per (radial cell, theta sector) + the two particle sets the
    reference calls `part` (released electrons) and `part_add` (positions of the ions created by the last update)."""

    def __init__(self, ctx, element, ion_max, ppc, num_theta, q=-1.0, m=1.0, density=1.0, n0=1.0e17, dt_xi=None):
        self.ctx, self.L, self.num_theta = ctx, ctx.L, num_theta
        h = _vp()
        _chk(self.L.qpg_neutral_create(C.byref(h), ctx.h, element, ion_max, ppc[0], ppc[1], num_theta, q, m, density, n0, ctx.dxi if dt_xi is None else dt_xi))
        self.h = h.value
        self.multi_max = self.L.qpg_neutral_multi_max(self.h)
        cap = ctx.nr * num_theta * ppc[0] * ppc[1] + 64          # one update releases at most ppc electrons per (cell, sector)
        self.part = Part2d(ctx, q / m, cap * max(1, self.multi_max))   # ... and a step at most multi_max times that (an overflow is reported)
        self.part_add = Part2d(ctx, q / m, cap)

    def close(self):
        if getattr(self, "h", None) and self.ctx.h:
            self.part.close(); self.part_add.close()
            self.L.qpg_neutral_destroy(self.h)
        self.h = None

    __del__ = close

    def update(self, e): _chk(self.L.qpg_neutral_update(self.h, e.h, self.part.h, self.part_add.h))

    def renew(self):
        _chk(self.L.qpg_neutral_reset(self.h))
        self.part.clear(); self.part_add.clear()

    def levels(self):
        out = np.zeros((self.multi_max + 2, self.num_theta, self.ctx.nr))
        _chk(self.L.qpg_neutral_levels(self.h, _ptr(out)))
        return out


class Laser:
    """field_laser (laser/field_laser_class.f03) of one xi stage (nz = its slab): envelope volumes of shape (P, nz+3, nr+2), xi slice j at
    index j+1; method names follow the type-bound procedures (set_grad + gather = slice, solve = advance)."""

    def __init__(self, ctx, nz, k0, ds, iteration=1, handle=None):
        self.ctx, self.L, self.nz = ctx, ctx.L, nz
        self._own = handle is None
        if handle is None:
            h = _vp()
            _chk(self.L.qpg_laser_create(C.byref(h), ctx.h, nz, k0, ds, iteration))
            handle = h.value
        self.h = handle
        self.shape = (ctx.P, nz + 3, ctx.nr + 2)

    def close(self):
        if getattr(self, "h", None) and self._own and self.ctx.h:
            self.L.qpg_laser_destroy(self.h)
        self.h = None

    __del__ = close

    def upload(self, ar, ai):
        ar, ai = _f64(ar), _f64(ai)
        assert ar.shape == self.shape and ai.shape == self.shape, (ar.shape, self.shape)
        _chk(self.L.qpg_laser_upload(self.h, _ptr(ar), _ptr(ai)))

    def download(self):
        ar, ai = np.zeros(self.shape), np.zeros(self.shape)
        _chk(self.L.qpg_laser_download(self.h, _ptr(ar), _ptr(ai)))
        return ar, ai

    def field(self, which):
        """0 a_r, 1 a_i, 2 grad a_r, 3 grad a_i (slice images of `slice`), 4 chi"""
        dim = (1, 1, 3, 3, 1)[which]
        return Field(self.ctx, dim, self.nz if which == 4 else 0, which == 4, handle=self.L.qpg_laser_field(self.h, which))

    def slice(self, j): _chk(self.L.qpg_laser_slice(self.h, j))
    def deposit_chi(self, part, j, ax_corr): _chk(self.L.qpg_laser_deposit_chi(self.h, part.h, j, ax_corr))
    def advance(self): _chk(self.L.qpg_laser_advance(self.h))
    def sync(self): _chk(self.L.qpg_laser_sync(self.h))
    def guard_size(self): return int(self.L.qpg_laser_guard_size(self.h))

    def set_handoff(self, guard_in=None, in_ready=None, in_ack=None, guard_out=None, out_ready=None, out_ack=None):
        """the stage's two envelope links of a xi-pipeline (device addresses; None = no link on that side), see include/qpad_b200.h"""
        _chk(self.L.qpg_laser_set_handoff(self.h, guard_in, in_ready, in_ack, guard_out, out_ready, out_ack))

    def upload_slab(self, ar, ai, noff2):
        """the stage's part of WHOLE-BOX envelope volumes (P, nz_total+3, nr+2): its slices and, as guards, the neighbours' edge slices
        (init_field_laser :163-169)"""
        sl = slice(noff2, noff2 + self.nz + 3)
        self.upload(np.ascontiguousarray(ar[:, sl]), np.ascontiguousarray(ai[:, sl]))


class Sim:
    """Fused slice loop (qpg_sim_*): simulation_class.f03:294-512 for one xi slab on one GPU."""

    FIELD_DIMS = dict(psi=1, e=3, b=3, e_spe=3, b_spe=3, e_beam=3, b_beam=3, cu=3, amu=3, acu=2, dcu=2, q_spe=1, q_beam=1,
                      spe_q=1, spe_qn=1, spe_cu=3, spe_dcu=2, spe_amu=3, beam_q=1, neut_q=1, rho_ion=1)
    HAS_2D = {"psi", "e", "b", "e_spe", "b_spe", "e_beam", "b_beam", "cu", "q_spe", "q_beam", "spe_q", "beam_q", "neut_q", "rho_ion"}

    def __init__(self, nr, nz, max_mode, rmax, zmin, zmax, dt, sp_qbm=-1.0, sp_npmax=0, beam_qbm=-1.0, beam_npmax=32,
                 beam_push_type=PUSH3_REDUCED, beam_evol=1, iter_max=1, iter_reltol=1e-3, iter_abstol=1e-3, relax_fac=-1.0,
                 field_boundary=BND_OPEN, sort_freq=0, use_graph=0, noff2=0, nzp=None, device=0, stream=None, sp_push_std=0,
                 sp_push_pgc=0, laser_iter=1, laser_k0=10.0, sp_ppc_r=1):
        self.L = load()
        nzp = nz if nzp is None else nzp
        prm = SimParams(nr=nr, nz_total=nz, noff2=noff2, nzp=nzp, max_mode=max_mode, field_boundary=field_boundary,
                        iter_max=iter_max, sort_freq=sort_freq, dr=rmax / nr, dxi=(zmax - zmin) / nz, dt=dt,
                        iter_reltol=iter_reltol, iter_abstol=iter_abstol, relax_fac=relax_fac, sp_qbm=sp_qbm,
                        sp_npmax=sp_npmax, beam_push_type=beam_push_type, beam_evol=beam_evol, beam_qbm=beam_qbm,
                        beam_npmax=beam_npmax, use_graph=use_graph, sp_push_std=sp_push_std, sp_push_pgc=sp_push_pgc,
                        laser_iter=laser_iter, laser_k0=laser_k0, sp_ppc_r=sp_ppc_r)
        self.prm = prm
        h = _vp()
        _chk(self.L.qpg_sim_create(C.byref(h), device, stream, C.byref(prm)))
        self.h = h.value
        self.nr, self.nz, self.nzp, self.noff2, self.max_mode = nr, nz, nzp, noff2, max_mode
        self.ctx = Ctx(nr, max_mode, prm.dr, prm.dxi, handle=self.L.qpg_sim_ctx(self.h))
        self.species = Part2d(self.ctx, sp_qbm, sp_npmax, handle=self.L.qpg_sim_species(self.h))
        self.beam = Part3d(self.ctx, beam_qbm, dt, beam_npmax, nz, noff2, nzp, handle=self.L.qpg_sim_beam(self.h))
        self._fields = {}
        lh = self.L.qpg_sim_laser(self.h)
        self.laser = Laser(self.ctx, nzp, laser_k0, dt, laser_iter, handle=lh) if lh else None

    def laser_advance(self): _chk(self.L.qpg_sim_laser_advance(self.h))
    def set_laser_overlap(self, on): _chk(self.L.qpg_sim_set_laser_overlap(self.h, int(on)))
    def set_graph_unroll(self, on): _chk(self.L.qpg_sim_set_graph_unroll(self.h, int(on)))

    def set_subcyc(self, exp_fac_max, exp_fac_clamped, dt_min, on=True):
        """the sub-cycling variant of the slice loop (proj_subcyc): plain per-slice launches, one host synchronisation per slice"""
        _chk(self.L.qpg_sim_set_subcyc(self.h, int(on), exp_fac_max, exp_fac_clamped, dt_min))

    def subcycles(self): return self.L.qpg_sim_subcycles(self.h)

    def attach_neutral(self, element, ion_max, ppc, num_theta, q=-1.0, m=1.0, density=1.0, n0=1.0e17):
        """a field-ionisation neutral species inside the slice loop (qpg_sim_attach_neutral): per-slice launch paths (not the sweep kernel); on a
        xi-pipeline its state travels with qpg_sim_neutral_pack / _unpack"""
        self.neutral = Neutral(self.ctx, element, ion_max, ppc, num_theta, q, m, density, n0, self.ctx.dxi)
        _chk(self.L.qpg_sim_attach_neutral(self.h, self.neutral.h, self.neutral.part.h, self.neutral.part_add.h))
        return self.neutral

    def neutral_wire_count(self): return int(self.L.qpg_sim_neutral_wire_count(self.h))
    def neutral_pack(self, ptr): _chk(self.L.qpg_sim_neutral_pack(self.h, ptr))
    def neutral_unpack(self, ptr): _chk(self.L.qpg_sim_neutral_unpack(self.h, ptr))

    def close(self):
        if getattr(self, "h", None):
            if getattr(self, "neutral", None):
                self.ctx.sync()
                self.neutral.close()
                self.neutral = None
            self.L.qpg_sim_destroy(self.h)
            self.ctx.h = None
        self.h = None

    __del__ = close

    def field(self, name):
        if name not in self._fields:
            fh = self.L.qpg_sim_field(self.h, name.encode())
            if not fh:
                raise KeyError(name)
            self._fields[name] = Field(self.ctx, self.FIELD_DIMS[name], self.nzp, name in self.HAS_2D, handle=fh)
        return self._fields[name]

    def init_species(self, x, p, gamma, psi, q):
        x, p, gamma, psi, q = map(_f64, (x, p, gamma, psi, q))
        _chk(self.L.qpg_sim_init_species(self.h, _ptr(x), _ptr(p), _ptr(gamma), _ptr(psi), _ptr(q), len(q)))

    def beam_qdp_begin(self): _chk(self.L.qpg_sim_beam_qdp_begin(self.h))
    def beam_qdp_end(self): _chk(self.L.qpg_sim_beam_qdp_end(self.h))
    def begin_step(self): _chk(self.L.qpg_sim_begin_step(self.h))
    def set_back_handoff(self, wire_b, wire_e, flag, seq):
        _chk(self.L.qpg_sim_set_back_handoff(self.h, wire_b, wire_e, flag, seq & 0xFFFFFFFF))

    def beam_qdp_raw(self): _chk(self.L.qpg_sim_beam_qdp_raw(self.h))
    def beam_qdp_fix(self): _chk(self.L.qpg_sim_beam_qdp_fix(self.h))
    def begin_step_zero(self): _chk(self.L.qpg_sim_begin_step_zero(self.h))
    def begin_step_add(self): _chk(self.L.qpg_sim_begin_step_add(self.h))
    def run_slices(self, j0, j1): _chk(self.L.qpg_sim_run_slices(self.h, j0, j1))
    def beam_push(self): _chk(self.L.qpg_sim_beam_push(self.h))
    def beam_push_interior(self): _chk(self.L.qpg_sim_beam_push_interior(self.h))   # the half of the push that needs nothing from the downstream stage
    def beam_push_edge(self): _chk(self.L.qpg_sim_beam_push_edge(self.h))           # the other half + update_bound
    def beam_qdp_part(self, part): _chk(self.L.qpg_sim_beam_qdp_part(self.h, part))  # 3: the particles the upstream stage has just handed over
    def renew(self): _chk(self.L.qpg_sim_renew(self.h))
    def set_graph(self, on): _chk(self.L.qpg_sim_set_graph(self.h, int(on)))
    def set_fused(self, on): _chk(self.L.qpg_sim_set_fused(self.h, int(on)))
    def set_sweep(self, on): _chk(self.L.qpg_sim_set_sweep(self.h, int(on)))
    def set_sweep_ctas(self, n): _chk(self.L.qpg_sim_set_sweep_ctas(self.h, int(n)))

    def sweep_profile(self, reset=False):
        """in-kernel phase clocks of the persistent sweep kernel (see qpg_sim_sweep_profile)"""
        out = (C.c_double * 12)()
        _chk(self.L.qpg_sim_sweep_profile(self.h, out, int(reset)))
        keys = ("cyc_A", "cyc_amj", "cyc_C", "cyc_push", "cyc_total", "ns_total", "slices", "amj_phases",
                "work_A", "work_amj", "work_C", "work_push")
        return dict(zip(keys, [float(v) for v in out]))

    def slice_trace(self):
        """(ns, PC iterations) of the sweep kernel's last pass over each slice of the slab"""
        ns, it = np.zeros(self.nzp), np.zeros(self.nzp, dtype=np.int32)
        _chk(self.L.qpg_sim_slice_trace(self.h, ns.ctypes.data_as(_pd), it.ctypes.data_as(_pi)))
        return ns, it

    def debug_abort(self):
        """test hook: raise the sweep kernel's sticky watchdog word (qpg_sim_debug_abort)"""
        _chk(self.L.qpg_sim_debug_abort(self.h))

    def stats(self):
        u, it, sl = _l(), _l(), _l()
        _chk(self.L.qpg_sim_stats(self.h, C.byref(u), C.byref(it), C.byref(sl)))
        return u.value, it.value, sl.value

    def step3d(self):
        """One 3D step of a single-stage run (simulation_class.f03:294-501 with nodes = [1,1])."""
        self.beam_qdp_begin()
        self.beam_qdp_end()
        self.begin_step()
        self.run_slices(1, self.nzp)
        self.laser_advance()
        self.beam_push()
        self.renew()


class WireBuf:
    """exportable device buffer (qpg_wire_alloc) or a peer's buffer mapped through its IPC handle (qpg_wire_import)"""

    def __init__(self, nbytes=None, handle=None):
        self.L = load()
        p = _vp()
        self.imported = handle is not None
        if handle is not None:
            _chk(self.L.qpg_wire_import(handle, C.byref(p)))
        else:
            _chk(self.L.qpg_wire_alloc(C.byref(p), int(nbytes)))
        self.ptr = p.value

    def data_ptr(self):
        return self.ptr

    def export(self):
        buf = C.create_string_buffer(64)
        _chk(self.L.qpg_wire_export(self.ptr, buf))
        return buf.raw

    def close(self):
        if getattr(self, "ptr", None):
            (self.L.qpg_wire_unmap if self.imported else self.L.qpg_wire_free)(self.ptr)
        self.ptr = None

    __del__ = close


def stream_signal(cuda_stream, flag_ptr, value):
    _chk(load().qpg_stream_signal(cuda_stream, flag_ptr, value & 0xFFFFFFFF))


def stream_wait(cuda_stream, flag_ptr, value):
    _chk(load().qpg_stream_wait(cuda_stream, flag_ptr, value & 0xFFFFFFFF))


def stream_wait_unless_empty(cuda_stream, count_ptr, flag_ptr, value):
    """wait for the flag only if the device int at count_ptr is non-zero (qpg_stream_wait_unless_empty)"""
    _chk(load().qpg_stream_wait_unless_empty(cuda_stream, count_ptr, flag_ptr, value & 0xFFFFFFFF))

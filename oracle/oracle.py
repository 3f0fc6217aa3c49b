"""ctypes front-end of the CPU oracle (oracle/qpad_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never from qpad_b200/.  PARITY UNPINNED: the reference
ships no golden vectors; the oracle is pinned by analytic known answers.

Array conventions (numpy, C-contiguous, float64):
  particles   x (np,2)  p (np,3)  gamma/psi/q (np,)            == Fortran x(2,np), p(3,np)
  field f1    (P, nr+2, dim)   planes re0, re1, im1, re2, im2…  == Fortran f1(dim,0:nr+1) per plane
  field f2    (P, nzp+1, nr+2, dim)
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")

FK_PSI, FK_EZ, FK_BZ, FK_BT, FK_BPLUS, FK_BMINUS, FK_VPOTZ, FK_VPOTP, FK_VPOTM = range(9)
BND_ZERO, BND_OPEN = 2, 3
PUSH3_REDUCED, PUSH3_BORIS = 1, 2


class Params(C.Structure):
    _fields_ = [("nr", C.c_int), ("nz", C.c_int), ("max_mode", C.c_int), ("bnd", C.c_int), ("iter_max", C.c_int),
                ("nstages", C.c_int),
                ("rmax", C.c_double), ("zmin", C.c_double), ("zmax", C.c_double), ("dt", C.c_double),
                ("iter_reltol", C.c_double), ("iter_abstol", C.c_double), ("relax_fac", C.c_double),
                ("ppc1", C.c_int), ("ppc2", C.c_int), ("num_theta", C.c_int), ("sort_freq", C.c_int),
                ("sp_q", C.c_double), ("sp_m", C.c_double), ("sp_density", C.c_double), ("sp_den_min", C.c_double),
                ("beam_push_type", C.c_int), ("beam_evol", C.c_int), ("beam_qbm", C.c_double), ("sp_push_type", C.c_int),
                ("laser_on", C.c_int), ("laser_iter", C.c_int), ("laser_k0", C.c_double),
                ("neut_on", C.c_int), ("neut_elem", C.c_int), ("neut_ion_max", C.c_int), ("neut_ppc1", C.c_int), ("neut_ppc2", C.c_int),
                ("neut_num_theta", C.c_int), ("neut_q", C.c_double), ("neut_m", C.c_double), ("neut_density", C.c_double), ("n0", C.c_double),
                ("subcyc_on", C.c_int), ("subcyc_exp_fac_max", C.c_double), ("subcyc_exp_fac_clamped", C.c_double), ("subcyc_dt_min", C.c_double)]


def build(fast=False, force=False):
    target = "liborc_fast.so" if fast else "liborc.so"
    path = os.path.join(_HERE, target)
    newest = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("qpad_oracle.c", "qpad_oracle_laser.c", "qpad_oracle_neutral.c", "qpad_oracle.h"))
    if force or not os.path.exists(path) or os.path.getmtime(path) < newest:
        subprocess.check_call(["make", "-C", _HERE, "-B", target], stdout=subprocess.DEVNULL)
    return path


_libs = {}


def lib(fast=False):
    if fast in _libs:
        return _libs[fast]
    L = C.CDLL(build(fast))
    i, l, d, vp = C.c_int, C.c_long, C.c_double, C.c_void_p
    sig = {
        "orc_qdeposit": (None, [_dp, _dp, l, d, i, i, _dp]),
        "orc_amjdeposit_robust": (None, [_dp, _dp, _dp, _dp, _dp, l, d, i, i, d, d, _dp, _dp, _dp, _dp, _dp]),
        "orc_push_u_robust": (None, [_dp, _dp, _dp, l, d, i, i, d, d, _dp, _dp]),
        "orc_amjdeposit_std": (None, [_dp, _dp, _dp, _dp, _dp, l, d, i, i, d, d, _dp, _dp, _dp, _dp, _dp]),
        "orc_push_u_std": (None, [_dp, _dp, _dp, _dp, l, d, i, i, d, d, _dp, _dp]),
        "orc_interp_psi": (None, [_dp, _dp, l, d, i, i, _dp]),
        "orc_amjdeposit_pgc": (None, [_dp, _dp, _dp, _dp, _dp, l, d, i, i, d, d, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, i]),
        "orc_push_u_pgc": (None, [_dp, _dp, _dp, _dp, l, d, i, i, d, d, _dp, _dp, _dp, _dp, _dp, _dp]),
        "orc_push_x": (None, [_dp, _dp, _dp, l, d]),
        "orc_update_bound": (l, [_dp, _dp, _dp, _dp, _dp, l, d]),
        "orc_sort_idx": (None, [_dp, l, d, i, _ip, _ip]),
        "orc_sort_part2d": (None, [_dp, _dp, _dp, _dp, _dp, l, d, i]),
        "orc_inject_uniform": (l, [_dp, _dp, _dp, _dp, _dp, i, d, i, i, i, d, d, d]),
        "orc_build_matrix": (None, [i, i, i, d, i, d, _dp, _dp, _dp]),
        "orc_tridiag_solve": (None, [_dp, _dp, _dp, _dp, i]),
        "orc_tridiag_solve_ld": (None, [_dp, _dp, _dp, _dp, i]),
        "orc_solve_psi": (None, [_dp, _dp, i, i, d, i]),
        "orc_solve_bt": (None, [_dp, _dp, i, i, d, i]),
        "orc_solve_bz": (None, [_dp, _dp, i, i, d, i]),
        "orc_solve_bt_iter": (None, [_dp, _dp, _dp, i, i, d, i, d]),
        "orc_solve_ez": (None, [_dp, _dp, i, i, d, i]),
        "orc_solve_et": (None, [_dp, _dp, _dp, i, i, d]),
        "orc_solve_et_beam": (None, [_dp, _dp, i, i]),
        "orc_solve_djdxi": (None, [_dp, _dp, _dp, i, i, d]),
        "orc_smooth_f1": (None, [_dp, i, i, _ip]),
        "orc_solve_vpotz": (None, [_dp, _dp, i, i, d, i]),
        "orc_solve_vpott": (None, [_dp, _dp, i, i, d, i]),
        "orc_qdeposit3d": (None, [_dp, _dp, l, d, d, i, i, i, i, _dp]),
        "orc_push3d": (None, [_dp, _dp, l, d, d, i, i, i, i, d, d, i, _dp, _dp]),
        "orc_update_bound3d": (l, [_dp, _dp, _dp, l, d, d]),
        "orc_push3d_spin": (None, [_dp, _dp, _dp, d, l, d, d, i, i, i, i, d, d, i, _dp, _dp]),
        "orc_update_bound3d_spin": (l, [_dp, _dp, _dp, _dp, l, d, d]),
        "orc_sim_create": (vp, [C.POINTER(Params)]),
        "orc_sim_destroy": (None, [vp]),
        "orc_sim_set_beam": (None, [vp, _dp, _dp, _dp, l]),
        "orc_sim_step3d": (l, [vp, i]),
        "orc_sim_run_slices": (l, [vp, i]),
        "orc_sim_run_range": (l, [vp, i, i]),
        "orc_sim_nzp": (i, [vp, i]),
        "orc_sim_plasma_np": (l, [vp, i]),
        "orc_sim_get_plasma": (None, [vp, i, _dp, _dp, _dp, _dp, _dp]),
        "orc_sim_beam_np": (l, [vp, i]),
        "orc_sim_get_beam": (None, [vp, i, _dp, _dp, _dp]),
        "orc_sim_get_field": (l, [vp, i, C.c_char_p, i, C.c_void_p]),
        "orc_sim_total_iters": (l, [vp]),
        "orc_sim_snapshot": (None, [vp]),
        "orc_sim_restore": (None, [vp]),
        "orc_sim_get_slice_iters": (None, [vp, i, _ip]),
        "orc_sim_total_subcycles": (l, [vp]),
        "orc_exp_fac_max": (d, [_dp, _dp, l]),
        "orc_clamp_exp_fac": (None, [_dp, _dp, l, d]),
        "orc_subcyc_step": (None, [d, d, d, d, C.POINTER(d), C.POINTER(i)]),
        "orc_adk_params": (i, [i, i, _dp]),
        "orc_plasma_frequency": (d, [d]),
        "orc_neutral_reset": (None, [_dp, i, i, i]),
        "orc_neutral_ionize": (None, [_dp, _dp, _dp, d, d, i, i, i, i, i, i]),
        "orc_neutral_add_particles": (l, [_dp, _dp, i, i, i, i, i, d, d, d, d, _dp, _dp, _dp, _dp, _dp, C.POINTER(l), _dp, _dp]),
        "orc_neutral_ion_deposit": (None, [_dp, _dp, l, d, i, i, _dp, _dp]),
        "orc_sim_neutral_np": (l, [vp, i]),
        "orc_sim_get_neutral": (None, [vp, i, _dp, _dp, _dp, _dp, _dp]),
        "orc_sim_get_levels": (None, [vp, i, _dp]),
        "orc_sim_set_laser": (None, [vp, _dp, _dp]),
        "orc_sim_get_laser": (None, [vp, _dp, _dp, _dp]),
        "orc_laser_volume_size": (l, [i, i, i]),
        "orc_deposit_ax_corr": (d, [i]),
        "orc_deposit_chi": (None, [_dp, _dp, _dp, l, d, i, i, d, d, _dp]),
        "orc_laser_gaussian_point": (None, [d, d, d, d, d, d, C.POINTER(d), C.POINTER(d)]),
        "orc_laser_launch_gaussian": (None, [d, d, d, d, d, d, d, d, d, d, d, i, i, i, _dp, _dp]),
        "orc_laser_build_matrix": (None, [i, i, d, d, d, d, _dp]),
        "orc_penta_solve": (None, [_dp, _dp, i]),
        "orc_laser_set_rhs": (None, [_dp, _dp, _dp, i, i, i, d, d, d, d, _dp, _dp]),
        "orc_laser_solve": (None, [_dp, _dp, _dp, _dp, _dp, i, i, i, d, d, d, d, i]),
        "orc_laser_set_grad": (None, [_dp, _dp, i, i, i, i, d, d, _dp, _dp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _libs[fast] = L
    return L


def nplanes(max_mode):
    return 2 * max_mode + 1


def zeros_f1(dim, nr, max_mode):
    return np.zeros((nplanes(max_mode), nr + 2, dim))


def inject_uniform(nr, dr, ppc1, ppc2, num_theta, qm=-1.0, density=1.0, den_min=1e-10):
    n = nr * ppc1 * ppc2 * num_theta
    x, p = np.zeros((n, 2)), np.zeros((n, 3))
    g, psi, q = np.zeros(n), np.zeros(n), np.zeros(n)
    npp = lib().orc_inject_uniform(x, p, g, psi, q, nr, dr, ppc1, ppc2, num_theta, qm, density, den_min)
    return x[:npp], p[:npp], g[:npp], psi[:npp], q[:npp]


class Sim:
    """Whole-loop oracle (simulation_class.f03:226-512), optionally as S pipeline stages run in sequence."""

    def __init__(self, fast=False, **kw):
        self.L = lib(fast)
        prm = Params()
        defaults = dict(nr=64, nz=32, max_mode=1, bnd=BND_OPEN, iter_max=1, nstages=1, rmax=5.0, zmin=-5.0, zmax=5.0,
                        dt=10.0, iter_reltol=1e-3, iter_abstol=1e-3, relax_fac=-1.0, ppc1=2, ppc2=2, num_theta=8,
                        sort_freq=0, sp_q=-1.0, sp_m=1.0, sp_density=1.0, sp_den_min=1e-10,
                        beam_push_type=PUSH3_REDUCED, beam_evol=1, beam_qbm=-1.0, sp_push_type=1, laser_on=0, laser_iter=1, laser_k0=10.0,
                        neut_on=0, neut_elem=3, neut_ion_max=1, neut_ppc1=2, neut_ppc2=2, neut_num_theta=8, neut_q=-1.0, neut_m=1.0, neut_density=1.0,
                        n0=1.0e17, subcyc_on=0, subcyc_exp_fac_max=1.5, subcyc_exp_fac_clamped=10.0, subcyc_dt_min=1e-3)
        defaults.update(kw)
        for k, v in defaults.items():
            setattr(prm, k, v)
        self.prm = prm
        self.h = self.L.orc_sim_create(C.byref(prm))
        self.nr, self.nz, self.max_mode, self.nstages = prm.nr, prm.nz, prm.max_mode, max(1, prm.nstages)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_sim_destroy(self.h)
            self.h = None

    def set_beam(self, x, p, q):
        x, p, q = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, p, q))
        self.L.orc_sim_set_beam(self.h, x, p, q, len(q))

    def step3d(self, istep=1):
        return self.L.orc_sim_step3d(self.h, istep)

    def run_slices(self, n):
        return self.L.orc_sim_run_slices(self.h, n)

    def run_range(self, j0, j1):
        return self.L.orc_sim_run_range(self.h, j0, j1)

    def nzp(self, stage=0):
        return self.L.orc_sim_nzp(self.h, stage)

    def plasma(self, stage=0):
        n = self.L.orc_sim_plasma_np(self.h, stage)
        x, p = np.zeros((n, 2)), np.zeros((n, 3))
        g, psi, q = np.zeros(n), np.zeros(n), np.zeros(n)
        self.L.orc_sim_get_plasma(self.h, stage, x, p, g, psi, q)
        return x, p, g, psi, q

    def beam(self, stage=0):
        n = self.L.orc_sim_beam_np(self.h, stage)
        x, p, q = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros(n)
        self.L.orc_sim_get_beam(self.h, stage, x, p, q)
        return x, p, q

    _DIM = dict(psi=1, e=3, b=3, e_spe=3, b_spe=3, e_beam=3, b_beam=3, cu=3, amu=3, acu=2, dcu=2, q_spe=1, q_beam=1,
                spe_q=1, spe_qn=1)

    def field(self, name, which=1, stage=0):
        n = self.L.orc_sim_get_field(self.h, stage, name.encode(), which, None)
        if n < 0:
            raise KeyError(name)
        out = np.zeros(n)
        self.L.orc_sim_get_field(self.h, stage, name.encode(), which, out.ctypes.data_as(C.c_void_p))
        P, dim = nplanes(self.max_mode), self._DIM[name]
        if which == 1:
            return out.reshape(P, self.nr + 2, dim)
        return out.reshape(P, self.nzp(stage) + 1, self.nr + 2, dim)

    def total_iters(self):
        return self.L.orc_sim_total_iters(self.h)

    def snapshot(self):
        self.L.orc_sim_snapshot(self.h)

    def restore(self):
        self.L.orc_sim_restore(self.h)

    def slice_iters(self, stage=0):
        """predictor-corrector iterations of each slice of the stage's last sweep"""
        out = np.zeros(self.nzp(stage), dtype=np.int32)
        self.L.orc_sim_get_slice_iters(self.h, stage, out)
        return out

    def total_subcycles(self):
        return self.L.orc_sim_total_subcycles(self.h)

    def neutral(self, stage=0):
        """electrons created by field ionisation so far (plasma-particle layout)"""
        n = self.L.orc_sim_neutral_np(self.h, stage)
        x, p = np.zeros((n, 2)), np.zeros((n, 3))
        g, psi, q = np.zeros(n), np.zeros(n), np.zeros(n)
        if n:
            self.L.orc_sim_get_neutral(self.h, stage, x, p, g, psi, q)
        return x, p, g, psi, q

    def levels(self, multi_max, stage=0):
        """(multi_max + 2, n_theta, nr): charge states 1..multi_max, neutral residue, total discrete ion level"""
        lev = np.zeros((multi_max + 2, self.prm.neut_num_theta, self.nr))
        self.L.orc_sim_get_levels(self.h, stage, lev)
        return lev

    def set_laser(self, ar, ai):
        """envelope volumes of shape (P, nz+3, nr+2), xi slice j at index j+1 (see Laser)"""
        self.L.orc_sim_set_laser(self.h, np.ascontiguousarray(ar, dtype=np.float64), np.ascontiguousarray(ai, dtype=np.float64))

    def laser(self):
        shape = (nplanes(self.max_mode), self.nz + 3, self.nr + 2)
        ar, ai, chi = np.zeros(shape), np.zeros(shape), np.zeros((nplanes(self.max_mode), self.nz + 1, self.nr + 2))
        self.L.orc_sim_get_laser(self.h, ar, ai, chi)
        return ar, ai, chi


class Laser:
    """One laser envelope on one stage owning the whole box (laser/field_laser_class.f03): volumes a_r, a_i of shape
    (P, nz+3, nr+2) with xi slice j at index j+1 (two lower guard slices), advanced by set_rhs + solve per 3D step
    (sim_lasers_class.f03:197-222 advance)."""

    def __init__(self, nr, nz, max_mode, rmax, zmin, zmax, ds, k0, iteration=1):
        self.L = lib()
        self.nr, self.nz, self.max_mode, self.k0, self.ds, self.iter = nr, nz, max_mode, k0, ds, iteration
        self.dr, self.dz, self.z0 = rmax / nr, (zmax - zmin) / nz, zmin
        shape = (nplanes(max_mode), nz + 3, nr + 2)
        self.ar, self.ai = np.zeros(shape), np.zeros(shape)
        self.sr, self.si = np.zeros(shape), np.zeros(shape)

    def launch_gaussian(self, a0, w0, focal_distance=0.0, lon_center=0.0, t_rise=1.0, t_flat=0.0, t_fall=1.0):
        self.L.orc_laser_launch_gaussian(self.k0, a0, w0, focal_distance, lon_center, t_rise, t_flat, t_fall, self.z0, self.dz, self.dr,
                                         self.nr, self.nz, self.max_mode, self.ar, self.ai)

    def advance(self, chi=None):
        """chi: (P, nz+1, nr+2, 1) susceptibility volume (None = vacuum)"""
        if chi is None:
            chi = np.zeros((nplanes(self.max_mode), self.nz + 1, self.nr + 2, 1))
        chi = np.ascontiguousarray(chi, dtype=np.float64)
        a = (self.nr, self.nz, self.max_mode, self.k0, self.ds, self.dr, self.dz)
        self.L.orc_laser_set_rhs(self.ar, self.ai, chi, *a, self.sr, self.si)
        self.L.orc_laser_solve(self.ar, self.ai, self.sr, self.si, chi, *a, self.iter)

    def set_grad(self, j):
        g = zeros_f1(3, self.nr, self.max_mode), zeros_f1(3, self.nr, self.max_mode)
        self.L.orc_laser_set_grad(self.ar, self.ai, j, self.nr, self.nz, self.max_mode, self.dr, self.dz, g[0], g[1])
        return g

    def gaussian_point(self, r, z, w0, f_dist):
        a, b = C.c_double(), C.c_double()
        self.L.orc_laser_gaussian_point(r, z, self.k0, self.k0, w0, f_dist, C.byref(a), C.byref(b))
        return a.value, b.value

"""Known answers for the laser-envelope half of the oracle (oracle/qpad_oracle_laser.c).

The reference ships nothing that pins this path, so it is pinned against physics: a Gaussian pulse in vacuum
must diffract like the analytic paraxial Gaussian beam (spot size, on-axis amplitude, curvature and Gouy phase) -- which
checks the operator rows, the right-hand side, the signs of the k0 couplings and the xi recurrence at once -- at second
order in (dr, ds)."""
import numpy as np
import pytest

from oracle import oracle as O


def test_penta_solve_matches_dense():
    L = O.lib()
    rng = np.random.default_rng(3)
    for m, nr in ((0, 40), (1, 33), (2, 64)):
        A = np.zeros((2 * nr, 5))
        L.orc_laser_build_matrix(m, nr, 20.0, 2.0, 0.03, 0.03, A)
        n = 2 * nr
        D = np.zeros((n, n))
        for r in range(n):
            for k in range(5):
                c = r - 2 + k
                if 0 <= c < n:
                    D[r, c] = A[r, k]
        b = rng.standard_normal(n)
        x = b.copy()
        L.orc_penta_solve(A, x, n)
        assert np.max(np.abs(x - np.linalg.solve(D, b))) < 1e-12 * np.max(np.abs(x))


def test_matrix_rows():
    """field_laser_class.f03:269-391: interior rows, the m = 0 axis rows and the outer rows"""
    L = O.lib()
    nr, k0, ds, dr, dz = 16, 20.0, 2.0, 0.05, 0.04
    A = np.zeros((2 * nr, 5))
    L.orc_laser_build_matrix(0, nr, k0, ds, dr, dz, A)
    q, h = 0.25 * ds, 1.5 * dr * dr / dz
    assert np.allclose(A[0], [0, 0, ds + h, -k0 * dr * dr, -ds]) and np.allclose(A[1], [0, k0 * dr * dr, ds + h, 0, -ds])
    j = 5
    assert np.allclose(A[2 * j], [-q * (1 - 0.5 / j), 0, 2 * q + h, -k0 * dr * dr, -q * (1 + 0.5 / j)])
    assert np.allclose(A[2 * j + 1], [-q * (1 - 0.5 / j), k0 * dr * dr, 2 * q + h, 0, -q * (1 + 0.5 / j)])
    assert A[2 * nr - 2, 4] == 0 and A[2 * nr - 1, 4] == 0 and A[2 * nr - 1, 3] == 0
    L.orc_laser_build_matrix(2, nr, k0, ds, dr, dz, A)
    assert np.allclose(A[0], [0, 0, 1, 0, 0]) and np.allclose(A[2 * j, 2], q * (2 + 4.0 / j ** 2) + h)


def _propagate(nr, nz, ds, nsteps, k0=20.0, w0=2.0, a0=1.0):
    las = O.Laser(nr, nz, 0, rmax=12.0, zmin=-3.0, zmax=3.0, ds=ds, k0=k0, iteration=1)
    las.launch_gaussian(a0, w0, focal_distance=0.0, lon_center=0.0, t_rise=1.5, t_flat=2.0, t_fall=1.5)   # flat for |xi| < 1
    for _ in range(nsteps):
        las.advance()
    return las


def test_gaussian_pulse_diffracts_like_the_paraxial_beam():
    """z_R = k0 w0^2 / 2 = 40.  After s = 20 the spot has grown by sqrt(1.25), the on-axis amplitude has dropped to 0.894 and
    the Gouy phase is 0.46 rad.  Inside the flat part of the pulse the envelope equation (i k0 + d/dxi) da/ds = lap(a)/2
    differs from the paraxial beam only by O(1/(k0 z_R)) ~ 1e-3 (the d/dxi term), so the complex envelope must agree with
    the analytic beam to a few 1e-3."""
    k0, w0, s, nr, nz = 20.0, 2.0, 20.0, 192, 192
    las = _propagate(nr, nz, 2.0, 10)
    r = np.arange(nr) * las.dr
    worst = 0.0
    for j in range(1, nz + 1):
        xi = (j - 1) * las.dz + las.z0
        if abs(xi) > 0.5:
            continue
        # every slice is a distance s further along its propagation: the launch formula with the focus moved by -s
        want = np.array([las.gaussian_point(rr, xi, w0, -s) for rr in r[: nr // 2]])
        got = np.stack([las.ar[0, j + 1, 1:nr // 2 + 1], las.ai[0, j + 1, 1:nr // 2 + 1]], axis=1)
        worst = max(worst, np.max(np.abs(got - want)))
        amp0 = np.hypot(got[0, 0], got[0, 1])
        assert abs(amp0 - 1 / np.sqrt(1 + (s - xi) ** 2 / 40.0 ** 2)) < 2e-3
    assert worst < 4e-3, worst
    # nothing diffracts without the operator: the initial pulse differs from the propagated one by far more than that
    ini = O.Laser(nr, nz, 0, rmax=12.0, zmin=-3.0, zmax=3.0, ds=2.0, k0=k0)
    ini.launch_gaussian(1.0, w0, 0.0, 0.0, 1.5, 2.0, 1.5)
    assert np.max(np.abs(ini.ar - las.ar)) > 0.2


def test_envelope_solver_is_second_order():
    """Richardson: successive refinements of (dr, dxi, ds) by 2 differ 4x less each time"""
    sols = []
    for f in (1, 2, 4):
        las = _propagate(96 * f, 96 * f, 2.0 / f, 5 * f)
        sols.append(las.ar[0, 2:-1:f, 1:-1:f][:96, :96] + 1j * las.ai[0, 2:-1:f, 1:-1:f][:96, :96])   # common nodes
    # slice j of the coarse grid sits at (j-1) dz: index j+1 -> fine index f*(j-1)+2
    d1, d2 = np.max(np.abs(sols[0] - sols[1])), np.max(np.abs(sols[1] - sols[2]))
    assert d2 < d1 / 3.0, (d1, d2)


def test_deposit_chi_equals_charge_deposit_for_cold_electrons():
    """psi = 0, qbm = -1: chi = -qbm q / (1 - qbm psi) = q, so off the axis the susceptibility deposit is the charge deposit
    (part2d_class.f03:361 vs :231); on the axis it carries get_deposit_ax_corr instead of 8"""
    L = O.lib()
    nr, dr, M = 48, 0.1, 2
    rng = np.random.default_rng(1)
    x, p, g, psi, q = O.inject_uniform(nr, dr, 2, 2, 8)
    x = x + 0.2 * dr * rng.standard_normal(x.shape)
    x, q = np.ascontiguousarray(x), np.ascontiguousarray(q)
    rho, chi = O.zeros_f1(1, nr, M), O.zeros_f1(1, nr, M)
    L.orc_qdeposit(x, q, len(q), dr, nr, M, rho)
    corr = L.orc_deposit_ax_corr(2)
    assert abs(corr - 48.0 / 9.0) < 1e-15
    L.orc_deposit_chi(x, q, np.zeros(len(q)), len(q), dr, nr, M, -1.0, corr, chi)
    assert np.array_equal(chi[:, 2:], rho[:, 2:])
    assert chi[0, 1, 0] == rho[0, 1, 0] / 8.0 * corr or abs(chi[0, 1, 0] - rho[0, 1, 0] / 8.0 * corr) < 1e-15 * abs(chi[0, 1, 0])
    # a particle with psi != 0 weighs 1 / (1 + psi)
    chi2 = O.zeros_f1(1, nr, M)
    L.orc_deposit_chi(x, q, np.full(len(q), 0.25), len(q), dr, nr, M, -1.0, corr, chi2)
    assert np.max(np.abs(chi2 - chi / 1.25)) < 1e-13 * np.max(np.abs(chi))


def test_set_grad_is_second_order_and_keeps_the_axis_quirk():
    nr, nz, M = 64, 64, 1
    las = O.Laser(nr, nz, M, rmax=4.0, zmin=0.0, zmax=4.0, ds=1.0, k0=10.0)
    r = (np.arange(nr + 2) - 1) * las.dr
    xi = (np.arange(nz + 3) - 2) * las.dz            # slice j at index j+1 sits at (j-1) dz
    f, g = np.exp(-r ** 2), np.sin(0.7 * xi)
    las.ar[0] = g[:, None] * f[None, :]
    las.ar[1] = g[:, None] * (r * f)[None, :]        # re1
    las.ar[2] = 0.5 * g[:, None] * (r * f)[None, :]  # im1
    j = 40
    gr, gi = las.set_grad(j)
    xj = (j - 1) * las.dz
    assert np.max(np.abs(gr[0, 2:nr, 0] - np.sin(0.7 * xj) * (-2 * r[2:nr] * f[2:nr]))) < 4e-3          # d/dr (central difference, dr = 1/16)
    assert np.max(np.abs(gr[0, 1:nr + 1, 2] - 0.7 * np.cos(0.7 * xj) * f[1:nr + 1])) < 4e-3           # d/dxi (3-point backward)
    assert np.allclose(gr[1, 2:nr + 1, 1], -1.0 / r[2:nr + 1] * las.ar[2, j + 1, 2:nr + 1])             # -(m/r) Im
    assert np.allclose(gr[2, 2:nr + 1, 1], 1.0 / r[2:nr + 1] * las.ar[1, j + 1, 2:nr + 1])
    # the reference writes the m = 1 axis rule with the loop variable after the loop (field_laser_class.f03:708-717):
    # it lands on the guard node nr+1, node 1 keeps 0
    assert gr[1, 1, 0] == 0.0 and gr[1, nr + 1, 0] == 2.0 * (0.5 / las.dr) * las.ar[1, j + 1, 2]
    assert np.all(gi == 0.0)


def test_linear_laser_wakefield_matches_theory():
    """The coupled path (robust_pgc deposit + push, psi solve, chi deposit, envelope advance) in the linear regime:
    behind a weak (a0 = 0.05), wide (w0 = 8) pulse the on-axis wake potential obeys  psi'' + psi = |a|^2 / 4
    (linear polarisation, k_p = 1), i.e. psi(xi) = 1/4 int_{-inf}^{xi} sin(xi - xi') |a(xi')|^2 dxi'."""
    nr, nz, rmax, zmin, zmax, a0, w0 = 256, 256, 24.0, -3.0, 9.0, 0.05, 8.0
    sim = O.Sim(nr=nr, nz=nz, max_mode=0, rmax=rmax, zmin=zmin, zmax=zmax, dt=2.0, ppc1=2, ppc2=2, num_theta=4, iter_max=5,
                iter_reltol=1e-4, iter_abstol=1e-9, sp_push_type=5, laser_on=1, laser_iter=1, laser_k0=20.0, beam_evol=0)
    las = O.Laser(nr, nz, 0, rmax, zmin, zmax, 2.0, 20.0, 1)
    las.launch_gaussian(a0, w0, 0.0, 0.0, 1.5, 0.0, 1.5)
    sim.set_laser(las.ar, las.ai)
    sim.set_beam(np.zeros((0, 3)), np.zeros((0, 3)), np.zeros(0))
    sim.step3d(1)
    psi = sim.field("psi", 2)[0, :nz, 1, 0]
    dz = (zmax - zmin) / nz
    xi = np.arange(nz) * dz + zmin
    a2 = las.ar[0, 2:nz + 2, 1] ** 2 + las.ai[0, 2:nz + 2, 1] ** 2
    G = np.array([np.sum(np.sin(xi[j] - xi[:j + 1]) * a2[:j + 1]) * dz for j in range(nz)])
    assert abs(np.dot(psi, G) / np.dot(G, G) - 0.25) < 2.5e-3
    assert np.max(np.abs(psi - 0.25 * G)) < 0.02 * np.max(np.abs(0.25 * G))
    # E_z = d psi / d xi behind the pulse
    ez = sim.field("e", 2)[0, :nz, 1, 2]
    dpsi = np.gradient(psi, dz)
    assert np.max(np.abs(ez[8:-8] - dpsi[8:-8])) < 0.03 * np.max(np.abs(ez))
    # the susceptibility of the barely perturbed plasma is -n/(1 + psi) ~ -1, and the envelope has been advanced
    ar, ai, chi = sim.laser()
    assert np.max(np.abs(chi[0, 8:nz, 2:nr // 2] + 1.0)) < 5e-3
    assert 0 < np.max(np.abs(ar - las.ar)) < 0.05 * a0


def test_host_side_laser_launch_matches_oracle():
    """qpad_b200.decks.laser_gaussian (the host-side profile launch the GPU path uploads) against the oracle's restatement of
    profile_laser%launch, at the lwfa deck's parameters"""
    from qpad_b200 import decks
    cfg = dict(decks.CONFIGS["C4"])
    las = cfg.pop("laser")
    nr, nz = 128, 96
    ar, ai = decks.laser_gaussian(nr, nz, cfg["rmax"], cfg["zmin"], cfg["zmax"], **las)
    o = O.Laser(nr, nz, 0, cfg["rmax"], cfg["zmin"], cfg["zmax"], cfg["dt"], las["k0"], las["iteration"])
    o.launch_gaussian(las["a0"], las["w0"], las["focal_distance"], las["lon_center"], las["t_rise"], las["t_flat"], las["t_fall"])
    assert np.max(np.abs(o.ar)) > 1.9
    assert np.max(np.abs(ar - o.ar)) < 1e-14 and np.max(np.abs(ai - o.ai)) < 1e-14


def test_laser_pipeline_stages_match_single_stage():
    """the envelope across xi stages (sim_lasers%advance: each stage receives the upstream stage's NEW last two slices as its
    lower guard slices before it solves, while its slice loop still sees the OLD ones) reproduces the single-stage run: the
    lwfa deck's nodes = [1, 4] in small"""
    from qpad_b200 import decks
    cfg = dict(nr=64, nz=48, max_mode=0, rmax=10.0, zmin=-3.0, zmax=5.0, dt=2.0, iter_max=4, iter_reltol=1e-3, iter_abstol=1e-6,
               ppc1=2, ppc2=2, num_theta=4, sp_push_type=5, laser_on=1, laser_iter=2, laser_k0=20.0, beam_evol=0)
    ar, ai = decks.laser_gaussian(cfg["nr"], cfg["nz"], cfg["rmax"], cfg["zmin"], cfg["zmax"], k0=20.0, a0=1.0, w0=2.0, t_rise=1.5, t_fall=1.5)
    sims = [O.Sim(nstages=S, **cfg) for S in (1, 4)]
    for s in sims:
        s.set_laser(ar, ai)
        s.set_beam(np.zeros((0, 3)), np.zeros((0, 3)), np.zeros(0))
        assert np.array_equal(s.laser()[0], ar) and np.array_equal(s.laser()[1], ai)      # scatter / gather round trip
        for step in (1, 2, 3):
            s.step3d(step)
    ref, pip = sims
    a1, b1, c1 = ref.laser()
    a4, b4, c4 = pip.laser()
    assert np.max(np.abs(a1 - ar)) > 1e-3                                                   # the pulse has evolved
    assert np.max(np.abs(a4 - a1)) <= 1e-12 * np.max(np.abs(a1)) and np.max(np.abs(b4 - b1)) <= 1e-12 * np.max(np.abs(b1))
    assert np.max(np.abs(c4[:, :-1] - c1[:, :-1])) <= 1e-11 * np.max(np.abs(c1))
    for name in ("psi", "e"):
        full = ref.field(name, 2)[:, :-1]
        parts = np.concatenate([pip.field(name, 2, stage=k)[:, :-1] for k in range(4)], axis=1)
        assert np.max(np.abs(full)) > 1e-3 and np.max(np.abs(parts - full)) <= 1e-11 * np.max(np.abs(full)), name

"""Input-deck helpers that stay on the host (out of the hot path).

* ``load_deck`` reads a QPAD ``qpinput.json`` (``!`` comments allowed, input_class.f03:48-105).
* ``beam_std`` restates the *lattice* part of ``inject_fdist3d_std`` (beam/fdist3d_std_class.f03:439-613).
  The reference draws momenta from the compiler's ``random_number`` (math_module.f03:77-102), which is
  gfortran-version dependent; here the draws come from ``numpy.random.Generator(PCG64(seed))`` -- a documented
  substitute (SURVEY.md §8c PRNG row).  Both the CUDA path and the oracle consume the same arrays.
"""
import json
import re
import numpy as np

CONFIGS = {
    # SURVEY.md §8 config table.  Beam blocks follow input_file/blowout_regime/qpinput_tri-gaussian.json.
    "C1": dict(nr=250, nz=500, max_mode=1, rmax=5.0, zmin=-5.0, zmax=5.0, dt=10.0, ppc1=2, ppc2=2, num_theta=16,
               iter_max=10, iter_reltol=1e-3, iter_abstol=1e-3,
               beam=dict(ppc=(2, 2, 2), num_theta=16, q=-1.0, m=1.0, gamma=20000.0, density=4.0, quiet=True,
                         center=(0.0, 0.0, -2.5), sigma=(0.25, 0.25, 0.5), range1=(-1.25, 1.25), range2=(-1.25, 1.25),
                         range3=(-5.0, 0.0), uth=(5.0, 5.0, 0.0), den_min=1e-10)),
    "C2": dict(nr=1024, nz=2048, max_mode=1, rmax=5.0, zmin=-5.0, zmax=5.0, dt=10.0, ppc1=4, ppc2=4, num_theta=16,
               iter_max=10, iter_reltol=1e-3, iter_abstol=1e-3,
               beam=dict(ppc=(2, 2, 2), num_theta=16, q=-1.0, m=1.0, gamma=20000.0, density=4.0, quiet=True,
                         center=(0.0, 0.0, -2.5), sigma=(0.25, 0.25, 0.5), range1=(-1.25, 1.25), range2=(-1.25, 1.25),
                         range3=(-5.0, 0.0), uth=(5.0, 5.0, 0.0), den_min=1e-10)),
    # input_file/hosing/qpinput.json: max_mode 2, a drive beam on the axis and a witness beam 0.0376 off it (both q/m = -1:
    # one particle set here), nodes [1, 4], time 20.1 / dt 10 = 2 steps
    "C3": dict(nr=256, nz=438, max_mode=2, rmax=9.395847, zmin=0.0, zmax=9.6, dt=10.0, ppc1=2, ppc2=2, num_theta=16,
               iter_max=10, iter_reltol=1e-3, iter_abstol=1e-3, nstep3d=2,
               beam=[dict(ppc=(1, 1, 1), num_theta=16, q=-1.0, m=1.0, gamma=20000.0, density=93.4633, quiet=True,
                          center=(0.0, 0.0, 3.0067), sigma=(0.1371, 0.1371, 0.4798), range1=(-0.6854, 0.6854), range2=(-0.6854, 0.6854),
                          range3=(0.6077, 5.4056), uth=(13.7083, 13.7083, 0.0), den_min=1e-10),
                     dict(ppc=(1, 1, 1), num_theta=16, q=-1.0, m=1.0, gamma=20000.0, density=56.078, quiet=True,
                          center=(0.0376, 0.0, 8.6442), sigma=(0.1371, 0.1371, 0.2399), range1=(-0.6854, 0.6854), range2=(-0.6854, 0.6854),
                          range3=(7.4447, 9.8437), uth=(13.7083, 13.7083, 0.0), den_min=1e-10)]),
    # input_file/lwfa/qpinput.json: no beam, one laser (gaussian x sin2), robust_pgc plasma, max_mode 0, time 10.1 / dt 2 = 5 steps
    "C4": dict(nr=512, nz=512, max_mode=0, rmax=15.0, zmin=-3.0, zmax=12.0, dt=2.0, ppc1=8, ppc2=2, num_theta=8,
               iter_max=10, iter_reltol=1e-2, iter_abstol=1e-3, nstep3d=5,
               laser=dict(k0=20.0, a0=2.0, w0=2.828427, focal_distance=0.0, lon_center=0.0, t_rise=2.0, t_flat=0.0, t_fall=2.0,
                          iteration=3)),
    # input_file/ionization/qpinput.json: nspecies 0, one lithium neutral (ADK, ion_max 3), one beam, nodes [1, 2]; oracle only so far
    "C5": dict(nr=500, nz=1000, max_mode=1, rmax=5.0, zmin=0.0, zmax=10.0, dt=10.0, ppc1=8, ppc2=8, num_theta=32,
               iter_max=10, iter_reltol=1e-3, iter_abstol=1e-3, nstep3d=1, n0=1.0e17,
               neutral=dict(element=3, ion_max=3, q=-1.0, m=1.0, density=1.0),
               beam=dict(ppc=(1, 1, 1), num_theta=16, q=-1.0, m=1.0, gamma=20000.0, density=4.0, quiet=True,
                         center=(0.0, 0.0, 2.5), sigma=(0.25, 0.25, 0.5), range1=(-1.25, 1.25), range2=(-1.25, 1.25),
                         range3=(0.0, 5.0), uth=(5.0, 5.0, 0.0), den_min=1e-10)),
}


def load_deck(path):
    txt = open(path).read()
    txt = re.sub(r"!.*", "", txt)
    return json.loads(txt)


def beam_std(nr, nz, rmax, zmin, zmax, ppc, num_theta, q, m, gamma, density, center, sigma, range1, range2, range3,
             uth, den_min=1e-10, quiet=True, seed=10, xi_cells=None, **_):
    """Tri-Gaussian 'standard' beam in Cartesian geometry.  Returns x(np,3) [x, y, xi - zmin], p(np,3), q(np)."""
    dr, dz = rmax / nr, (zmax - zmin) / nz
    dtheta = 2.0 * np.pi / num_theta
    ppc1, ppc2, ppc3 = ppc
    coef = np.sign(q / m) / (ppc1 * ppc2 * ppc3 * num_theta)
    r3 = (range3[0] - zmin, range3[1] - zmin)
    k0 = max(0, int(np.floor(r3[0] / dz)) - 1)
    k1 = min(nz, int(np.ceil(r3[1] / dz)) + 1)
    if xi_cells is not None:            # only the beam particles of lattice cells xi_cells[0] <= k < xi_cells[1] (a bounded sample of the step)
        k0, k1 = max(k0, int(xi_cells[0])), min(k1, int(xi_cells[1]))
        if k1 <= k0:
            return np.zeros((0, 3)), np.zeros((0, 3)), np.zeros(0)
    rmax_beam = np.hypot(max(abs(range1[0]), abs(range1[1])), max(abs(range2[0]), abs(range2[1])))
    i1max = min(nr, int(np.ceil(rmax_beam / dr)) + 1)
    # loop order of the reference: k, j, i, i3, i2, i1 (innermost)
    k = np.arange(k0, k1)
    j = np.arange(num_theta)
    i = np.arange(i1max)
    i3 = (np.arange(ppc3) + 0.5) / ppc3
    i2 = (np.arange(ppc2) + 0.5) / ppc2
    i1 = (np.arange(ppc1) + 0.5) / ppc1
    K, J, I, I3, I2, I1 = np.meshgrid(k, j, i, i3, i2, i1, indexing="ij")
    zn = (I3 + K).ravel()
    theta = ((I2 + J) * dtheta).ravel()
    rn = (I1 + I).ravel()
    x1 = rn * dr * np.cos(theta)
    x2 = rn * dr * np.sin(theta)
    x3 = zn * dz
    keep = (x1 >= range1[0]) & (x1 <= range1[1]) & (x2 >= range2[0]) & (x2 <= range2[1]) & (x3 >= r3[0]) & (x3 <= r3[1])
    den = (np.exp(-0.5 * ((x1 - center[0]) / sigma[0]) ** 2) * np.exp(-0.5 * ((x2 - center[1]) / sigma[1]) ** 2)
           * np.exp(-0.5 * ((x3 - (center[2] - zmin)) / sigma[2]) ** 2)) * density
    keep &= den >= den_min
    x1, x2, x3, rn, den = x1[keep], x2[keep], x3[keep], rn[keep], den[keep]
    n = len(x1)
    rng = np.random.Generator(np.random.PCG64(seed))
    g = rng.standard_normal((n, 3))
    p1, p2 = uth[0] * g[:, 0], uth[1] * g[:, 1]
    p3 = uth[2] * g[:, 2] + gamma
    p3 = np.sqrt(p3 ** 2 - p1 ** 2 - p2 ** 2 - 1.0)
    qq = rn * den * coef
    x = np.stack([x1, x2, x3], 1)
    p = np.stack([p1, p2, p3], 1)
    if quiet:
        x = np.concatenate([x, x * np.array([-1.0, -1.0, 1.0])])
        p = np.concatenate([p, p * np.array([-1.0, -1.0, 1.0])])
        qq = np.concatenate([0.5 * qq, 0.5 * qq])
    return np.ascontiguousarray(x), np.ascontiguousarray(p), np.ascontiguousarray(qq)


def plasma_uniform(nr, rmax, ppc1, ppc2, num_theta, q=-1.0, density=1.0, den_min=1e-10):
    """Host-side plasma injection for the uniform/uniform profile with uth = 0 and ordered theta: a vectorised
    restatement of inject_fdist2d (species/fdist2d_class.f03:289-357).  Loop order of the reference: theta sector j,
    radial cell i, i1, i2 (innermost).  Returns x(np,2), p(np,3), gamma, psi, q in the reference's AoS layout."""
    dr = rmax / nr
    if density < den_min:
        z = np.zeros
        return z((0, 2)), z((0, 3)), z(0), z(0), z(0)
    dtheta = 2.0 * np.pi / num_theta
    j = np.arange(1, num_theta + 1)
    i = np.arange(1, nr + 1)
    i1 = (np.arange(1, ppc1 + 1) - 0.5) / ppc1
    i2 = (np.arange(1, ppc2 + 1) - 0.5) / ppc2
    J, I, I1, I2 = np.meshgrid(j, i, i1, i2, indexing="ij")
    rn = (I1 + (I - 1.0)).ravel()
    theta = ((I2 + J - 1.0) * dtheta).ravel()
    coef = np.sign(q) / (float(ppc1 * ppc2) * float(num_theta))
    x = np.stack([rn * dr * np.cos(theta), rn * dr * np.sin(theta)], 1)
    qq = rn * 1.0 * 1.0 * density * coef
    n = len(qq)
    p = np.zeros((n, 3))
    gamma = np.ones(n)
    psi = (1.0 - gamma + p[:, 2]) / q
    return np.ascontiguousarray(x), p, gamma, psi, np.ascontiguousarray(qq)


def laser_gaussian(nr, nz, rmax, zmin, zmax, k0, a0, w0, focal_distance=0.0, lon_center=0.0, t_rise=1.0, t_flat=0.0,
                   t_fall=1.0, max_mode=0, **_):
    """Host-side launch of a laser pulse with a Gaussian transverse and sin^2 longitudinal profile at t = 0: a vectorised
    restatement of profile_laser%launch (laser/profile_laser_class.f03:318-378) with get_prof_perp_gaussian
    (laser/profile_laser_lib.f03:56-96) and get_prof_lon_sin2 (:472-502), no chirp.  Stays on the host like the input
    deck; the result is uploaded with qpg_laser_upload.  Returns a_r, a_i of shape (P, nz+3, nr+2): xi slice j (1-based)
    at index j+1 (two lower guard slices, zero: nothing is ahead of the box), radial guards zero."""
    dr, dz = rmax / nr, (zmax - zmin) / nz
    P = 2 * max_mode + 1
    ar, ai = np.zeros((P, nz + 3, nr + 2)), np.zeros((P, nz + 3, nr + 2))
    z = (np.arange(1, nz + 1) - 1.0) * dz + zmin - lon_center          # "z" is xi = t - z
    pih = 1.570796326794897
    fs, fe = -0.5 * t_flat, 0.5 * t_flat
    env = np.zeros(nz)
    rise = (z >= fs - t_rise) & (z < fs)
    env[rise] = np.cos((z[rise] - fs) / t_rise * pih) ** 2
    env[(z >= fs) & (z < fe)] = 1.0
    fall = (z >= fe) & (z < fe + t_fall)
    env[fall] = np.cos((z[fall] - fe) / t_fall * pih) ** 2
    r = (np.arange(1, nr + 1) - 1.0) * dr
    zs = -1.0 * (z + focal_distance)
    zr = 0.5 * k0 * w0 * w0
    curv = zs / (zs * zs + zr * zr)
    w = w0 * np.sqrt(1.0 + zs * zs / (zr * zr))
    gouy = np.arctan2(zs, zr)
    r2 = (r * r)[None, :]
    phase = 0.5 * k0 * r2 * curv[:, None] - gouy[:, None]
    amp = (w0 / w)[:, None] * np.exp(-r2 / (w * w)[:, None])
    ar[0, 2:nz + 2, 1:nr + 1] = (env * a0)[:, None] * amp * np.cos(phase)
    ai[0, 2:nz + 2, 1:nr + 1] = -(env * a0)[:, None] * amp * np.sin(phase)
    return ar, ai

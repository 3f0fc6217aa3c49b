"""Device code checked on the CPU: the shuffle-free kernels of qpad_b200/csrc (neutral.cu, subcyc.cu, vpot.cu, diag.cu) are
compiled for the host through tests/emu (CTA threads = fibers, exact __syncthreads) and compared with the oracle through the
same case bodies the GPU tests use (tests/kernel_cases.py).  This is a check of the kernels' LOGIC (indexing, ordering, the
scan, the host entry points) -- it is not the parity gate: that is `pytest -m gpu` on the B200 through libqpadb200.so."""
import numpy as np
import pytest

from oracle import oracle as O
import kernel_cases as K
from emu import emu


def test_launch_rewrite():
    from emu.build import transform
    src = "k_a<<<(n + 127) / 128, 128, 0, c->stream>>>(x, f(y, z));\nDISPATCH(k_b<2><<<dim3(3, 2), 64>>>(p));"
    assert transform(src) == "emu::launch((n + 127) / 128, 128, 0, [&] { k_a(x, f(y, z)); });\nDISPATCH(emu::launch(dim3(3, 2), 64, 0, [&] { k_b<2>(p); }));"


@pytest.fixture()
def api():
    with emu.patched() as capi:
        yield capi


@pytest.mark.parametrize("elem,mm,M", [(1, 1, 0), (3, 3, 1), (2, 2, 2)])
def test_neutral_update_emulated(api, elem, mm, M):
    n0 = emu.lib().emu_launches()
    K.neutral_update(api, O, elem, mm, M)
    assert emu.lib().emu_launches() - n0 >= 6          # one fused update kernel (k_neutral_update) per call


def test_subcyc_particles_emulated(api):
    K.subcyc_particles(api, O)


@pytest.mark.parametrize("M,bnd", [(0, O.BND_OPEN), (1, O.BND_OPEN), (2, O.BND_OPEN), (2, O.BND_ZERO)])
def test_vpot_emulated(api, M, bnd):
    K.vpot(api, O, M, bnd, exact=True)


def test_vpot_nr1024_emulated(api):
    K.vpot(api, O, 1, O.BND_OPEN, nr=1024, exact=True)


def test_stage_emulated(api):
    K.stage(api, O)


def test_fastmath_probe_emulated(api):
    """plumbing of qpg_debug_fastmath only: in the emulation the MUFU seeds are exact IEEE values, the GPU test measures the real thing"""
    ur, uq = K.fastmath_accuracy(api, max_ulp=1.0)
    assert ur <= 1.0 and uq <= 1.0


@pytest.mark.parametrize("M,dim,kind,order", [(0, 1, 0, 1), (2, 3, 1, 2), (2, 2, 2, 2), (1, 2, 2, 0)])
def test_field_smooth_emulated(api, M, dim, kind, order):
    K.field_smooth(api, O, M, dim, kind, order)


def test_part2d_move_emulated(api):
    K.part2d_move(api, O)


def test_neutral_overflow_emulated(api):
    K.neutral_overflow(api, O)


def test_handoff_entry_points_validate_their_arguments(api):
    """error behaviour of the round-2 hand-off entry points (no silent acceptance): a link needs its buffer AND both flag words, a stage that
    hands its last two slices on needs two slices, the neutral's record needs an attached neutral, spin planes must be enabled first"""
    import numpy as np
    cfg = dict(nr=32, nz=8, max_mode=0, rmax=4.0, zmin=0.0, zmax=2.0, dt=2.0)
    sim = api.Sim(sp_npmax=64, beam_npmax=64, sp_push_pgc=1, laser_iter=1, laser_k0=10.0, noff2=0, nzp=1, **cfg)
    buf = np.zeros(sim.laser.guard_size() + 16)
    flags = np.zeros(16, dtype=np.uint32)
    p, f = buf.ctypes.data, flags.ctypes.data
    with pytest.raises(api.QpadError, match="go together"):
        sim.laser.set_handoff(guard_in=p, in_ready=f)                                   # ack word missing
    with pytest.raises(api.QpadError, match="at least two"):
        sim.laser.set_handoff(guard_out=p, out_ready=f, out_ack=f + 32)                 # a slab of ONE slice cannot hand two on
    sim.laser.set_handoff(guard_in=p, in_ready=f, in_ack=f + 32)                        # an upstream link alone is fine
    with pytest.raises(api.QpadError, match="no neutral"):
        sim.neutral_pack(p)
    assert sim.neutral_wire_count() == -1
    with pytest.raises(api.QpadError, match="enable_spin"):
        sim.beam.upload_spin(np.zeros((0, 3)))
    assert sim.beam.wire_count() == 7 * sim.beam.wire_cap() + 1
    sim.close()

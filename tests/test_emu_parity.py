"""The per-routine C-ABI of libqpadb200.so checked on the CPU: the SAME test bodies as the GPU parity tests (tests/test_gpu_parity.py,
tests/kernel_cases.py) run against the host emulation of the device sources (tests/emu: fields.cu, particles.cu, beam.cu, laser.cu, neutral.cu,
subcyc.cu, vpot.cu, diag.cu compiled for the host; CTA threads are fibers, barriers and warp collectives -- shuffles, ballots,
match_any, the m8n8k4 DMMA -- are exact).  What this covers: indexing, ordering, reductions, scans, the field-program interpreter,
the axis rules, the host entry points -- i.e. the LOGIC of the kernels, incl. code that has not yet had GPU time (neutral species,
sub-cycling, vpot, staging).  What it does not cover: anything about the hardware (memory model, occupancy, launch limits, the
MUFU-seeded reciprocal's last bits), the persistent sweep kernel, CUDA graphs and the peer-memory transport.  The parity gate
proper remains `pytest -m gpu` on a B200."""
import pytest

from oracle import oracle as O
from emu import emu
import kernel_cases as K
import test_gpu_parity as G
import test_gpu_laser as GL
import test_golden as TG

G.PATHS["stream-oplist"] = dict(sweep=0, use_graph=0, fused=0)      # the slab driver the emulation can run: plain per-slice launches


@pytest.fixture()
def mods():
    with emu.patched() as capi:
        yield capi, O


def test_field_roundtrip_and_arith(mods): G.test_field_roundtrip_and_arith(mods)


@pytest.mark.parametrize("nr,M,bnd", [(64, 0, 3), (64, 1, 3), (250, 1, 2), (96, 2, 3), (1024, 1, 3)])
def test_solves_match_oracle(mods, nr, M, bnd): G.test_solves_match_oracle(mods, nr, M, bnd)


def test_convergence_tester(mods): G.test_convergence_tester(mods)


@pytest.mark.parametrize("nr,M", [(64, 0), (64, 1), (96, 2)])
def test_particle_kernels_match_oracle(mods, nr, M): G.test_particle_kernels_match_oracle(mods, nr, M)


@pytest.mark.parametrize("nr,M", [(64, 1), (96, 2)])
def test_std_pusher_kernels_match_oracle(mods, nr, M): G.test_std_pusher_kernels_match_oracle(mods, nr, M)


@pytest.mark.parametrize("nr,M,std", [(64, 0, 0), (64, 1, 0), (64, 1, 1)])
def test_pgc_pusher_kernels_match_oracle(mods, nr, M, std): G.test_pgc_pusher_kernels_match_oracle(mods, nr, M, std)


def test_update_bound_heavy_loss(mods): G.test_update_bound_heavy_loss(mods)


@pytest.mark.parametrize("nr,ppc,nth", [(64, 2, 8), (250, 2, 16)])
def test_sort_bit_exact(mods, nr, ppc, nth): G.test_sort_bit_exact(mods, nr, ppc, nth)


@pytest.mark.parametrize("M,push", [(1, 1), (2, 2), (0, 1)])
def test_beam_kernels_match_oracle(mods, M, push): G.test_beam_kernels_match_oracle(mods, M, push)


@pytest.mark.parametrize("M,push", [(1, 1), (2, 2)])
def test_beam_spin_push_matches_oracle(mods, M, push): G.test_beam_spin_push_matches_oracle(mods, M, push)


@pytest.mark.parametrize("nr,nz,M", [(64, 12, 0), (50, 9, 2)])
def test_laser_slice_images_match_oracle(mods, nr, nz, M): GL.test_laser_slice_images_match_oracle(mods, nr, nz, M)


@pytest.mark.parametrize("nr,M", [(64, 0), (96, 2)])
def test_deposit_chi_matches_oracle(mods, nr, M): GL.test_deposit_chi_matches_oracle(mods, nr, M)


@pytest.mark.parametrize("nr,nz,M,iters", [(100, 24, 1, 2), (64, 16, 2, 1), (33, 8, 1, 1)])
def test_envelope_advance_matches_oracle(mods, nr, nz, M, iters): GL.test_envelope_advance_matches_oracle(mods, nr, nz, M, iters)


def test_wire_formats(mods):
    """the hand-off records of the xi pipeline (pipe_send / recv of a field slice, pipesend_part2d, the beam's forward hand-off)"""
    G.test_wire_formats(mods)


def test_ionization_loop_matches_oracle(mods):
    """the ionisation deck (config 5 in small) through qpad_b200.ionization.IonizationStage: neutral.cu + every per-routine kernel"""
    K.ionization_loop(mods[0], O, nsl=32)


def test_subcyc_loop_matches_oracle(mods):
    """the sub-cycled slice loop through qpad_b200.subcyc.SubcycStage"""
    K.subcyc_loop(mods[0], O)


# ---- qpg_sim (csrc/sim.cu) on its plain per-slice launch path: the field programs A / C / D as op lists, the particle kernels with
# the device-side skip flags of the predictor-corrector loop, compaction, look-ahead deposit, beam deposit / push, sorting -------------
@pytest.mark.parametrize("M", [1, 2])
def test_slice_loop_matches_oracle(mods, M): G.test_slice_loop_matches_oracle(mods, M, "stream-oplist")


def test_one_slice_from_identical_state(mods): G.test_one_slice_from_identical_state(mods, 1, "stream-oplist")


def test_full_3d_step_with_beam_push(mods): G.test_full_3d_step_with_beam_push(mods, "stream-oplist")


def test_sorted_loop_still_matches(mods): G.test_sorted_loop_still_matches(mods, "stream-oplist")


def test_std_pusher_slice_loop(mods): G.test_std_pusher_slice_loop(mods, 0)


def test_golden_blowout_fixture(mods):
    """the committed fixture (tests/golden/blowout_small.npz): 1e-10 per-slice fields, 1e-6 line-outs and beam moments"""
    TG.test_cuda_blowout_matches_fixture()


def test_golden_lwfa_fixture(mods):
    """laser envelope + ponderomotive pushers + qpg_sim over two 3D steps against tests/golden/lwfa_small.npz (~70 s: the envelope
    solve is one persistent CTA with 2 log2(nr) barriers per slice)"""
    TG.test_cuda_lwfa_matches_fixture()


# ---- the neutral species inside qpg_sim (qpg_sim_attach_neutral) ------------------------------------------------------------------
def test_sim_neutral_loop_matches_oracle(mods):
    """input_file/ionization in small (nspecies 0, Li, one beam) through the fast-path object"""
    K.sim_neutral_loop(mods[0], O)


def test_sim_neutral_loop_two_levels_with_plasma(mods):
    K.sim_neutral_loop(mods[0], O, ion_max=2, with_plasma=True, nsl=32)


def test_sim_neutral_full_step(mods):
    K.sim_neutral_full_step(mods[0], O)


# ---- the sub-cycling variant inside qpg_sim (qpg_sim_set_subcyc) ------------------------------------------------------------------
def test_sim_subcyc_loop_matches_oracle(mods):
    K.sim_subcyc_loop(mods[0], O)


def test_sim_subcyc_with_neutral_matches_oracle(mods):
    """both at once: the released electrons are sub-cycled and clamped too, the ionisation runs in every sub-step (:312-323)"""
    K.sim_subcyc_loop(mods[0], O, with_neutral=True)


# ---- the persistent cooperative sweep kernel (csrc/sweep.cu), the default slab driver on the GPU: all CTAs alive at once, meeting at
# its hand-rolled grid / team barriers and flagged 16-byte exchange words; field team of 1 .. 8 CTAs here (nr = 1024 = 32 CTAs passes
# too with QPAD_EMU_SMS=40, 25 s) -----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M", [0, 1, 2])
def test_sweep_slice_loop_matches_oracle(mods, M):
    c0 = emu.lib().emu_coop_launches()
    G.test_slice_loop_matches_oracle(mods, M, "sweep")
    assert emu.lib().emu_coop_launches() > c0 and emu.lib().emu_polls() > 0


def test_sweep_one_slice_from_identical_state(mods): G.test_one_slice_from_identical_state(mods, 1, "sweep")


def test_sweep_full_3d_step_with_beam_push(mods): G.test_full_3d_step_with_beam_push(mods, "sweep")


def test_sweep_sorted_loop(mods): G.test_sorted_loop_still_matches(mods, "sweep")


def test_sweep_lwfa_slice_loop(mods):
    """config 4 in small with the laser hooks INSIDE the sweep kernel (k_sweep<M, PGC = true>: slice images by a helper CTA, pgc pushers,
    susceptibility deposit fused into the push phase), two 3D steps with the envelope advance in between"""
    GL.test_lwfa_slice_loop_matches_oracle(mods, 0, 1, nr=64, nz=64)


@pytest.mark.parametrize("nr,M,ppc,nth", [(250, 1, 2, 8), (65, 1, 2, 8), (33, 2, 2, 8), (24, 1, 2, 8)])
def test_sweep_kernel_multi_cta_team(mods, nr, M, ppc, nth): G.test_sweep_kernel_multi_cta_team(mods, nr, M, ppc, nth)


def test_graft_entry_smoke_body(mods, capsys):
    """__graft_entry__.smoke() -- what the driver runs on cuda:0 before the bench -- with the library replaced by the emulation"""
    import __graft_entry__ as ge
    ge.smoke()
    assert "psi rel err vs oracle" in capsys.readouterr().out


# ---- pipeline.LocalPipeline, the default driver of bench.py: S persistent sweep kernels on S streams of one GPU, event-ordered hand-offs,
# the backward e / b hand-off published from inside the downstream sweep kernel.  torch is replaced by tests/emu/faketorch.py (streams
# and events are bookkeeping: every operation has run when its enqueue returns, and the host's issue order is a valid schedule -- a
# stream-ordered flag wait whose producer was not enqueued first is reported as an error) --------------------------------------------
@pytest.fixture()
def pipeline_mods(mods, monkeypatch):
    import sys
    from emu import faketorch
    monkeypatch.setitem(sys.modules, "torch", faketorch)
    monkeypatch.setenv("QPAD_EMU_SWEEP", "1")          # qpg_sim defaults to the sweep kernel, as on the GPU
    return mods


@pytest.mark.parametrize("S", [2, 3])
def test_local_pipeline_matches_oracle(pipeline_mods, S):
    c0 = emu.lib().emu_coop_launches()
    G.test_local_pipeline_matches_oracle(pipeline_mods, S)
    assert emu.lib().emu_coop_launches() - c0 >= 4 * S


@pytest.mark.parametrize("S", [2, 3])
def test_lwfa_local_pipeline(pipeline_mods, S):
    """the envelope across xi stages: slabs of the envelope per stage, guard hand-off between the explicit and the implicit half of the advance
    (the GPU test's deck at half the resolution)"""
    GL.test_lwfa_local_pipeline_matches_oracle(pipeline_mods, S, nr=64, nz=48, nsteps=3)


def test_beam_spin_on_the_pipeline(pipeline_mods):
    K.beam_spin_pipeline(pipeline_mods[0])


def test_neutral_local_pipeline(pipeline_mods):
    """the neutral species across xi stages (released electrons, ion buffer, rho_ion, levels in the forward hand-off)"""
    K.neutral_local_pipeline(pipeline_mods[0], pipeline_mods[1], 2)


def test_local_pipeline_with_unequal_slabs(pipeline_mods):
    G.test_local_pipeline_with_unequal_slabs(pipeline_mods)


def test_local_pipeline_e2e_wave_with_upload(pipeline_mods):
    """the end-to-end leg of bench.py: every wave re-uploads the plasma lattice into the first stage (instead of the on-device renewal)
    and reads the slabs' line-outs and the counters back -- same physics as the resident run, i.e. the oracle's S-stage run"""
    import numpy as np
    capi, _ = pipeline_mods
    from qpad_b200 import decks
    from qpad_b200.pipeline import LocalPipeline
    S = 2
    cfg = dict(nr=64, nz=32, max_mode=1, rmax=5.0, zmin=-5.0, zmax=5.0, dt=10.0, ppc1=2, ppc2=2, num_theta=8, iter_max=2, iter_reltol=1e-3, iter_abstol=1e-3)
    bm = decks.beam_std(cfg["nr"], cfg["nz"], cfg["rmax"], cfg["zmin"], cfg["zmax"], **dict(decks.CONFIGS["C1"]["beam"]))
    plasma = decks.plasma_uniform(cfg["nr"], cfg["rmax"], cfg["ppc1"], cfg["ppc2"], cfg["num_theta"])
    lp = LocalPipeline(cfg, plasma, bm, S)
    lp.fill()
    nwaves = 3
    for _ in range(nwaves):
        lp.wave(upload=plasma)
        for sim in lp.sims:
            assert np.all(np.isfinite(sim.field("e").lineout(3, 0, 1))) and np.all(np.isfinite(sim.field("psi").lineout(1, 0, 1)))
        lp.stats()
    lp.drain()
    upd, iters, slices = lp.stats()
    nsteps = slices // cfg["nz"]
    kw = {k: cfg[k] for k in ("nr", "nz", "max_mode", "rmax", "zmin", "zmax", "dt", "iter_max", "iter_reltol", "iter_abstol")}
    orc = O.Sim(ppc1=2, ppc2=2, num_theta=8, nstages=S, **kw)
    orc.set_beam(*bm)
    for k in range(nsteps):
        orc.step3d(k + 1)
    assert nsteps >= nwaves and iters == orc.total_iters() and upd == nsteps * cfg["nz"] * len(plasma[4])
    for r, sim in enumerate(lp.sims):
        for name in ("psi", "e"):
            got, want = sim.field(name).download_f2()[:, :sim.nzp], orc.field(name, 2, stage=r)[:, :sim.nzp]
            assert np.max(np.abs(got - want)) < 1e-6 * np.max(np.abs(want)), (r, name)
    lp.close()

"""GPU parity of csrc/subcyc.cu, csrc/vpot.cu, csrc/diag.cu, smooth_f1 and move_part2d_comm against the oracle, through the C-ABI.
The case bodies are shared with tests/test_emu_kernels.py (the same device sources compiled for the host)."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    from qpad_b200 import capi
    from oracle import oracle as O
    import kernel_cases as K
    capi.load()
    return capi, O, K


def test_subcyc_particles(mods):
    capi, O, K = mods
    K.subcyc_particles(capi, O)


@pytest.mark.parametrize("M,bnd", [(0, 3), (1, 3), (2, 3), (2, 2)])
def test_vpot(mods, M, bnd):
    capi, O, K = mods
    K.vpot(capi, O, M, bnd)


def test_vpot_nr1024(mods):
    capi, O, K = mods
    K.vpot(capi, O, 1, 3, nr=1024)


def test_stage(mods):
    capi, O, K = mods
    K.stage(capi, O)


def test_subcyc_loop_matches_oracle(mods):
    """the sub-cycled slice loop through the per-routine C-ABI (qpad_b200.subcyc.SubcycStage) against the oracle's"""
    capi, O, K = mods
    K.subcyc_loop(capi, O)


@pytest.mark.parametrize("with_neutral", [False, True])
def test_sim_subcyc_loop_matches_oracle(mods, with_neutral):
    """the sub-cycling variant inside qpg_sim (qpg_sim_set_subcyc), without and with an attached neutral species"""
    capi, O, K = mods
    K.sim_subcyc_loop(capi, O, with_neutral=with_neutral)


def test_fastmath_accuracy(mods):
    """fast_rcp / fast_sqrt (MUFU seed + Newton steps) stay within 2 ulp of IEEE on the device -- also the gate for the shorter
    variants of the QPG_FASTMATH_SHORT experiment (build with QPG_NVCC_EXTRA=-DQPG_FASTMATH_SHORT)"""
    capi, O, K = mods
    print("max ulp error (rcp, sqrt):", K.fastmath_accuracy(capi))


@pytest.mark.parametrize("M,dim,kind,order", [(0, 1, 0, 1), (1, 1, 0, 2), (2, 3, 1, 1), (1, 3, 1, 3), (2, 2, 2, 2), (1, 2, 2, 0)])
def test_field_smooth(mods, M, dim, kind, order):
    """field_rho / field_jay / field_djdxi %smooth (fields/field_src_class.f03:102-271, ufield_class.f03:274-339)"""
    capi, O, K = mods
    K.field_smooth(capi, O, M, dim, kind, order)


def test_field_smooth_nr1024(mods):
    capi, O, K = mods
    K.field_smooth(capi, O, 1, 3, 1, 2, nr=1024)


def test_part2d_move(mods):
    """move_part2d_comm (species/part2d_comm.f03:147) on one radial partition"""
    capi, O, K = mods
    K.part2d_move(capi, O)


def test_sweep_watchdog_is_reported(mods):
    """a sweep kernel that leaves through its watchdog must not pass silently: the abort word is sticky, the next launch leaves at
    once, qpg_sim_stats / qpg_ctx_sync return QPG_ERR_STATE, and a pending backward hand-off flag is still raised (the upstream
    stage's stream must not hang)"""
    capi, O, K = mods
    import numpy as np
    cfg = dict(nr=64, nz=8, max_mode=1, rmax=5.0, zmin=-1.0, zmax=1.0, dt=10.0, iter_max=2)
    x, p, g, psi, q = O.inject_uniform(cfg["nr"], cfg["rmax"] / cfg["nr"], 2, 2, 8)
    sim = capi.Sim(sp_npmax=2 * len(q), beam_npmax=64, use_graph=0, **cfg)
    sim.init_species(x, p, g, psi, q)
    sim.beam_qdp_begin(); sim.beam_qdp_end(); sim.begin_step()
    sim.run_slices(1, 2)
    assert sim.stats()[2] == 2
    import torch
    nb = sim.field("b").wire_count()
    wb = torch.zeros(2 * nb + 8, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    sim.debug_abort()
    sim.set_back_handoff(wb.data_ptr(), wb.data_ptr() + 8 * nb, wb.data_ptr() + 16 * nb, 7)
    sim.begin_step()
    sim.run_slices(1, 2)
    with pytest.raises(capi.QpadError, match="aborted"):
        sim.stats()
    with pytest.raises(capi.QpadError, match="aborted"):
        sim.ctx.sync()
    assert int(wb[2 * nb:].view(torch.int32)[0].item()) == 7     # the flag was raised although the sweep did no work
    sim.close()


def test_conditional_stream_wait(mods):
    """qpg_stream_wait_unless_empty: the backward hand-off of the xi-pipeline is awaited only by a stage that holds beam particles"""
    import torch
    capi, O, K = mods
    flag = torch.zeros(8, dtype=torch.int32, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    out = torch.zeros(1, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    st, st2 = torch.cuda.Stream(), torch.cuda.Stream()
    capi.stream_wait_unless_empty(st.cuda_stream, cnt.data_ptr(), flag.data_ptr(), 5)     # count 0: must not block although the flag is 0
    st.synchronize()
    cnt.fill_(3); flag[0] = 7
    torch.cuda.synchronize()
    capi.stream_wait_unless_empty(st.cuda_stream, cnt.data_ptr(), flag.data_ptr(), 5)     # count > 0, flag already past the value
    st.synchronize()
    flag.zero_()
    torch.cuda.synchronize()
    # count > 0, flag not there yet: blocks until another stream raises it.  As everywhere in the pipeline the PRODUCER is enqueued first
    # (a polling kernel may only depend on work that is already in some hardware queue: streams can share one)
    with torch.cuda.stream(st2):                            # ~0.1 s of work on st2, then the signal
        torch.cuda._sleep(200_000_000)
    capi.stream_signal(st2.cuda_stream, flag.data_ptr(), 2)
    capi.stream_wait_unless_empty(st.cuda_stream, cnt.data_ptr(), flag.data_ptr(), 2)
    with torch.cuda.stream(st):
        out.copy_(flag[:1])                                 # ordered behind the wait: must see the raised flag, not the 0 of ~0.1 s earlier
    st.synchronize()
    assert int(out.item()) == 2 and int(flag[0].item()) == 2

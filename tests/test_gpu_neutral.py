"""GPU parity of the field-ionisation neutral species (csrc/neutral.cu, qpad_b200/ionization.py) against the oracle
(oracle/qpad_oracle_neutral.c), through the C-ABI.  First run on a B200 at the start of round 2 (all green); the case bodies are
shared with the host-emulation tests (tests/test_emu_kernels.py, tests/test_emu_parity.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    from qpad_b200 import capi
    from oracle import oracle as O
    capi.load()
    return capi, O


@pytest.mark.parametrize("elem,mm,M", [(1, 1, 0), (3, 3, 1), (2, 2, 2)])
def test_neutral_update_matches_oracle(mods, elem, mm, M):
    """ionize + add_particles on a given field, several updates in a row (the body is shared with the host-emulation test,
    tests/test_emu_kernels.py, where it already passes)"""
    capi, O = mods
    import kernel_cases as K
    K.neutral_update(capi, O, elem, mm, M)


def test_ionization_loop_matches_oracle(mods):
    """config 5 in small through the per-routine C-ABI (qpad_b200.ionization.IonizationStage) against the oracle's loop (body shared
    with tests/test_emu_parity.py)"""
    capi, O = mods
    import kernel_cases as K
    K.ionization_loop(capi, O)


@pytest.mark.parametrize("use_graph", [0, 1])
def test_sim_neutral_loop_matches_oracle(mods, use_graph):
    """the neutral species inside qpg_sim (qpg_sim_attach_neutral), plain launches and CUDA-graph replay"""
    capi, O = mods
    import kernel_cases as K
    K.sim_neutral_loop(capi, O, use_graph=use_graph)
    K.sim_neutral_loop(capi, O, use_graph=use_graph, ion_max=2, with_plasma=True, nsl=32)


@pytest.mark.parametrize("use_graph", [0, 1])
def test_sim_neutral_full_step(mods, use_graph):
    capi, O = mods
    import kernel_cases as K
    K.sim_neutral_full_step(capi, O, use_graph=use_graph)


def test_neutral_overflow_is_reported(mods):
    capi, O = mods
    import kernel_cases as K
    K.neutral_overflow(capi, O)


@pytest.mark.parametrize("S", [2, 3])
def test_neutral_local_pipeline_matches_oracle(mods, S):
    """the neutral's state in the xi hand-offs (neutral_class.f03:1025-1101): S pipeline stages on one GPU against the oracle's S-stage run"""
    capi, O = mods
    import kernel_cases as K
    K.neutral_local_pipeline(capi, O, S)


def test_neutral_pipeline_with_unrolled_slice_graph(mods):
    """the slice graph without the WHILE node (qpg_sim_set_graph_unroll: all iter_max predictor-corrector iterations captured, the surplus ones
    skip themselves on the device) -- what a pipeline of >= 6 stages on the per-slice launch path uses -- gives the same run"""
    capi, O = mods
    import kernel_cases as K
    K.neutral_local_pipeline(capi, O, 2, graph_unroll=True)

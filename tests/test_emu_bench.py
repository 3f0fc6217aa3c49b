"""bench.py's B200 arms executed end to end on the CPU: the library is the host emulation (tests/emu), torch its stand-in
(tests/emu/faketorch.py), the decks shrunk.  Time is not emulated, so every rate in the printed line is meaningless -- what is
checked is that the functions the driver runs (`run_b200_local`: probe_partition -> LocalPipeline with 4 sweep-kernel stages -> fill ->
timed waves -> end-to-end waves; `run_c5`: the neutral inside qpg_sim) execute without error on the current inputs (incl. the
thinned, charge-scaled beam of make_inputs) and emit the JSON contract of the task."""
import contextlib
import io
import json
import types

import pytest

from emu import emu, faketorch

KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
        "clocks", "gpu_launches", "roofline", "e2e")


@pytest.fixture()
def bench_mod(monkeypatch):
    faketorch.install(monkeypatch)
    monkeypatch.setenv("QPAD_EMU_SWEEP", "1")
    monkeypatch.setenv("QPAD_EMU_SMS", "16")
    import bench
    with emu.patched():
        yield bench


def _run(fn, args):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        fn(args)
    return json.loads(buf.getvalue().strip().splitlines()[-1])


def test_default_bench_function_runs(bench_mod, monkeypatch):
    from qpad_b200 import decks
    monkeypatch.setitem(decks.CONFIGS, "C2", dict(decks.CONFIGS["C2"], nr=64, nz=48, ppc1=2, ppc2=2, num_theta=8, iter_max=3))
    args = types.SimpleNamespace(gpus=1, steps=1, warmup=1, config="C2", balance=1, transport=None, stages=4, no_cpu=True, no_micro=True, roof_slices=16,
                                 ref_slices=8, no_sweep=False, no_graph=False, legacy_pipeline=False, impl="b200", check=1, fill_steps=[2], rebalance=1)
    c0 = emu.lib().emu_coop_launches()
    line = _run(bench_mod.run_b200_local, args)
    assert all(k in line for k in KEYS), [k for k in KEYS if k not in line]
    assert line["n_gpus"] == 1 and line["steps"] == 1 and line["dtype"] == "f64" and line["gpu_launches"] > 0
    assert {"bound", "achieved", "peak", "frac", "traffic"} <= set(line["roofline"]) and {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(line["e2e"])
    assert line["e2e"]["h2d_bytes_per_step"] == 64 * 2048 and line["config"]["pc_iters_per_slice"] > 1.2      # a driven wake, not the quiet plasma
    assert emu.lib().emu_coop_launches() - c0 >= 4 * 2                                                        # the sweep kernels of the 4 stages ran
    # the pipelined run carried its own correctness check: line-outs and beam moments equal to a one-stage run of the same 3D steps
    pc = line["parity_check"]
    assert pc["ok"] and pc["ez_lineout_rel_err"] < 1e-6 and pc["psi_lineout_rel_err"] < 1e-6 and pc["beam_particles"] == pc["beam_particles_single_stage"] > 0, pc
    assert line["single_step"]["single_step_ms"] > 0 and line["fill_inclusive"]["K=2"]["updates_per_s"] > 0


def test_c5_bench_function_runs(bench_mod, monkeypatch):
    from qpad_b200 import decks
    monkeypatch.setitem(decks.CONFIGS, "C5", dict(decks.CONFIGS["C5"], nr=64, nz=48, ppc1=2, ppc2=2, num_theta=8, iter_max=3, rmax=6.0, zmax=8.0))
    args = types.SimpleNamespace(steps=1, warmup=1, no_graph=False, no_cpu=True, ref_slices=8)
    line = _run(bench_mod.run_c5, args)
    assert all(k in line for k in KEYS), [k for k in KEYS if k not in line]
    assert line["value"] > 0 and "neutral Li" in line["config"]["workload"] and line["e2e"]["h2d_bytes_per_step"] > 0


def test_single_stage_bench_function_runs(bench_mod, monkeypatch):
    """`bench.py --stages 1` (run_b200): one sweep kernel on all SMs, the command the ncu captures are taken from"""
    from qpad_b200 import decks
    monkeypatch.setitem(decks.CONFIGS, "C2", dict(decks.CONFIGS["C2"], nr=64, nz=64, ppc1=2, ppc2=2, num_theta=8, iter_max=3))
    args = types.SimpleNamespace(gpus=1, steps=1, warmup=1, config="C2", balance=1, transport=None, stages=1, no_cpu=True, no_micro=True, roof_slices=16,
                                 ref_slices=8, no_sweep=False, no_graph=False, legacy_pipeline=False, impl="b200")
    line = _run(bench_mod.run_b200, args)
    assert all(k in line for k in KEYS), [k for k in KEYS if k not in line]
    assert line["gpu_launches"] > 0 and line["roofline"]["bound"] == "hbm"


def test_c4_pipeline_bench_function_runs(bench_mod, monkeypatch):
    """`bench.py --config C4` (run_c4_pipeline): the laser-wakefield deck on the xi-pipeline, envelope slabs with guard hand-off; the line carries
    the comparison of the pipelined envelope and wake with a one-stage run of the same 3D steps"""
    from qpad_b200 import decks
    monkeypatch.setitem(decks.CONFIGS, "C4", dict(decks.CONFIGS["C4"], nr=64, nz=48, ppc1=2, ppc2=2, num_theta=8, iter_max=3))
    args = types.SimpleNamespace(gpus=1, steps=2, warmup=1, config="C4", stages=2, no_cpu=True, ref_slices=8, no_sweep=False, no_graph=False, impl="b200", check=1, balance=1, rebalance=1)
    c0 = emu.lib().emu_coop_launches()
    line = _run(bench_mod.run_c4_pipeline, args)
    assert all(k in line for k in KEYS), [k for k in KEYS if k not in line]
    assert line["gpu_launches"] > 0 and emu.lib().emu_coop_launches() - c0 >= 2 * 3 and "2 stages per GPU" in line["config"]["parallelism"]
    pc = line["parity_check"]
    assert pc["ok"] and pc["envelope_rel_err"] < 1e-6 and pc["envelope_change_since_launch"] > 1e-3, pc


def test_c5_pipeline_bench_function_runs(bench_mod, monkeypatch):
    """`bench.py --config C5` (run_c5_pipeline): the ionisation deck on the xi-pipeline, the neutral's state in the forward hand-off"""
    from qpad_b200 import decks
    monkeypatch.setitem(decks.CONFIGS, "C5", dict(decks.CONFIGS["C5"], nr=64, nz=48, ppc1=2, ppc2=2, num_theta=8, iter_max=3, rmax=6.0, zmax=8.0))
    args = types.SimpleNamespace(steps=1, warmup=1, no_graph=False, no_cpu=True, ref_slices=8, stages=2, check=1)
    line = _run(bench_mod.run_c5_pipeline, args)
    assert all(k in line for k in KEYS), [k for k in KEYS if k not in line]
    assert line["value"] > 0 and "neutral Li" in line["config"]["workload"] and "2 stages per GPU" in line["config"]["parallelism"]
    pc = line["parity_check"]
    assert pc["ok"] and pc["beam_particles"] == pc["beam_particles_single_stage"] > 0, pc

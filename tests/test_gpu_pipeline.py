"""The xi-pipeline on real GPUs (needs >= 2): two stages over NCCL reproduce the oracle's 2-stage run, i.e. the
single-stage physics (tests/test_oracle_known_answers.py::test_pipeline_stages_match_single_stage)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_stage_pipeline_matches_oracle(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from oracle import oracle as O
    from qpad_b200 import decks
    nsteps, world = 3, 2
    # every stage ends having run nsteps + world - 1 complete steps (PipelineStage.unwind)
    total = nsteps + world - 1
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "pipeline_gpu_worker.py"), str(tmp_path), str(nsteps)]
    env = dict(os.environ, NCCL_MAX_P2P_NCHANNELS="2")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    cfg = dict(nr=64, nz=32, max_mode=1, rmax=5.0, zmin=-5.0, zmax=5.0, dt=10.0, iter_max=2, iter_reltol=1e-3, iter_abstol=1e-3)
    beam = dict(decks.CONFIGS["C1"]["beam"])
    bm = decks.beam_std(cfg["nr"], cfg["nz"], cfg["rmax"], cfg["zmin"], cfg["zmax"], **beam)
    orc = O.Sim(ppc1=2, ppc2=2, num_theta=8, nstages=world, **cfg)
    orc.set_beam(*bm)
    for k in range(total):
        orc.step3d(k + 1)
    nb = 0
    for rank in range(world):
        d = np.load(tmp_path / f"rank{rank}.npz")
        nzp = int(d["nzp"])
        for name in ("psi", "e"):
            got, want = d[name][:, :nzp], orc.field(name, 2, stage=rank)[:, :nzp]
            assert np.max(np.abs(got - want)) < 1e-6 * np.max(np.abs(want)), (rank, name)
        ox, op, oq = orc.beam(stage=rank)
        assert len(d["bq"]) == len(oq) and np.array_equal(d["bq"], oq)
        if len(oq):
            assert np.max(np.abs(d["bx"] - ox)) < 1e-9 * np.max(np.abs(ox))
        nb += len(oq)
    assert nb > 0


def _local_pipeline_vs_oracle(tmp_path, world, S, transport, one_device, port):
    from oracle import oracle as O
    from qpad_b200 import decks
    nsteps = 3
    G = world * S
    total = G - 1 + nsteps
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "pipeline_gpu_worker.py"), str(tmp_path), str(nsteps), str(S), transport]
    env = dict(os.environ, NCCL_MAX_P2P_NCHANNELS="4")
    if one_device:
        env["QPG_TEST_ONE_DEVICE"] = "1"
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    cfg = dict(nr=64, nz=32, max_mode=1, rmax=5.0, zmin=-5.0, zmax=5.0, dt=10.0, iter_max=2, iter_reltol=1e-3, iter_abstol=1e-3)
    beam = dict(decks.CONFIGS["C1"]["beam"])
    bm = decks.beam_std(cfg["nr"], cfg["nz"], cfg["rmax"], cfg["zmin"], cfg["zmax"], **beam)
    orc = O.Sim(ppc1=2, ppc2=2, num_theta=8, nstages=G, **cfg)
    orc.set_beam(*bm)
    for k in range(total):
        orc.step3d(k + 1)
    nb = 0
    for g in range(G):
        d = np.load(tmp_path / f"stage{g}.npz")
        nzp = int(d["nzp"])
        for name in ("psi", "e"):
            got, want = d[name][:, :nzp], orc.field(name, 2, stage=g)[:, :nzp]
            assert np.max(np.abs(got - want)) < 1e-6 * np.max(np.abs(want)), (g, name)
        ox, op, oq = orc.beam(stage=g)
        assert len(d["bq"]) == len(oq) and np.array_equal(d["bq"], oq)
        if len(oq):
            assert np.max(np.abs(d["bx"] - ox)) < 1e-9 * np.max(np.abs(ox))
        nb += len(oq)
    assert nb > 0


@pytest.mark.parametrize("transport", ["p2p", "nccl"])
def test_sm_partitioned_pipeline_across_two_gpus(tmp_path, transport):
    """LocalPipeline with 2 stages per GPU on 2 GPUs (4 global stages: event-ordered hand-offs inside a GPU; between the
    GPUs either peer-memory writes + flags (p2p) or NCCL send/recv) against the oracle's 4-stage run"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _local_pipeline_vs_oracle(tmp_path, 2, 2, transport, False, 29534)


@pytest.mark.parametrize("world,S", [(2, 2), (3, 1)])
def test_peer_memory_links_between_processes_on_one_gpu(tmp_path, world, S):
    """the p2p transport (CUDA IPC mapped wire buffers, pack kernels writing into the consumer's memory, flag words
    awaited by stream memory operations) between rank PROCESSES that share GPU 0 -- the same protocol as across GPUs,
    testable on a one-GPU box; against the oracle's (world*S)-stage run"""
    _local_pipeline_vs_oracle(tmp_path, world, S, "p2p", True, 29535 + world)


@pytest.mark.parametrize("world,S", [(2, 1), (2, 2)])
def test_lwfa_envelope_handoff_between_processes(tmp_path, world, S):
    """the laser envelope's guard hand-off over the peer-memory links (capi.Laser.set_handoff with the wire buffer and the ready / ack words in
    the neighbour PROCESS's memory, CUDA IPC): rank processes sharing GPU 0, S stages each, against the oracle's one-stage run"""
    from oracle import oracle as O
    nr, nz, k0, iters, nsteps = 128, 96, 20.0, 3, 3
    G = world * S
    total = G - 1 + nsteps
    cfg = dict(nr=nr, nz=nz, max_mode=0, rmax=12.0, zmin=-3.0, zmax=6.0, dt=2.0, iter_max=6, iter_reltol=1e-3, iter_abstol=1e-6)
    olas = O.Laser(nr, nz, 0, cfg["rmax"], cfg["zmin"], cfg["zmax"], cfg["dt"], k0, iters)
    olas.launch_gaussian(1.2, 2.5, 0.0, 0.0, 1.5, 0.0, 1.5)
    x, p, g, psi, q = O.inject_uniform(nr, cfg["rmax"] / nr, 4, 2, 8)
    np.savez(tmp_path / "lwfa_inputs.npz", nr=nr, nz=nz, ar=olas.ar, ai=olas.ai, x=x, p=p, g=g, psi=psi, q=q)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29560 + S), os.path.join(ROOT, "tests", "pipeline_gpu_worker.py"), str(tmp_path), str(nsteps), str(S), "p2p", "lwfa"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, QPG_TEST_ONE_DEVICE="1"))
    assert r.returncode == 0, r.stderr[-3000:]
    orc = O.Sim(ppc1=4, ppc2=2, num_theta=8, sp_push_type=5, laser_on=1, laser_iter=iters, laser_k0=k0, beam_evol=0, **cfg)
    orc.set_laser(olas.ar, olas.ai)
    orc.set_beam(np.zeros((0, 3)), np.zeros((0, 3)), np.zeros(0))
    for k in range(total):
        orc.step3d(k + 1)
    oar, oai, _ = orc.laser()
    assert np.max(np.abs(oar - olas.ar)) > 1e-3 * np.max(np.abs(oar))
    covered = 0
    for gidx in range(G):
        d = np.load(tmp_path / f"stage{gidx}.npz")
        off, nzp = int(d["noff2"]), int(d["nzp"])
        covered += nzp
        assert int(d["stats"][2]) == total * nzp
        err = max(np.max(np.abs(d["ar"][:, 2:2 + nzp] - oar[:, off + 2:off + 2 + nzp])), np.max(np.abs(d["ai"][:, 2:2 + nzp] - oai[:, off + 2:off + 2 + nzp])))
        assert err < 1e-9 * np.max(np.abs(oar)), (gidx, err)
        for name in ("psi", "e"):
            whole = orc.field(name, 2)[:, :nz]
            assert np.max(np.abs(d[name][:, :nzp] - whole[:, off:off + nzp])) < 1e-7 * np.max(np.abs(whole)), (gidx, name)
    assert covered == nz


@pytest.mark.parametrize("world,S", [(2, 1), (2, 2)])
def test_neutral_handoff_between_processes(tmp_path, world, S):
    """neut%psend / precv (neutral_class.f03:1025-1101) over the peer-memory links: the record of the released electrons, the ion buffer, rho_ion
    and the levels is written into the next rank PROCESS's memory (CUDA IPC) ahead of the forward message's ready word; ranks share GPU 0,
    S stages each, against the oracle's (world*S)-stage run of the ionisation deck in small"""
    from oracle import oracle as O
    from qpad_b200 import decks
    nsteps = 2
    G = world * S
    total = G - 1 + nsteps
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29570 + S), os.path.join(ROOT, "tests", "pipeline_gpu_worker.py"), str(tmp_path), str(nsteps), str(S), "p2p", "neutral"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, QPG_TEST_ONE_DEVICE="1"))
    assert r.returncode == 0, r.stderr[-3000:]
    cfg = dict(nr=64, nz=36, max_mode=1, rmax=6.0, zmin=0.0, zmax=8.0, dt=10.0, iter_max=3, iter_reltol=1e-3, iter_abstol=1e-3)
    bm = decks.beam_std(cfg["nr"], cfg["nz"], cfg["rmax"], cfg["zmin"], cfg["zmax"], **dict(decks.CONFIGS["C5"]["beam"]))
    orc = O.Sim(sp_density=0.0, neut_on=1, neut_elem=3, neut_ion_max=2, neut_ppc1=2, neut_ppc2=2, neut_num_theta=8, ppc1=2, ppc2=2, num_theta=8, n0=1.0e17, nstages=G, **cfg)
    orc.set_beam(*bm)
    upd_o = sum(orc.step3d(k + 1) for k in range(total))
    upd = 0
    for g in range(G):
        d = np.load(tmp_path / f"stage{g}.npz")
        nzp = int(d["nzp"])
        upd += int(d["stats"][0])
        for name in ("psi", "e"):
            got, want = d[name][:, :nzp], orc.field(name, 2, stage=g)[:, :nzp]
            if g == G - 1:
                assert np.max(np.abs(want)) > 1e-6                      # the wake of the electrons released upstream
            assert np.max(np.abs(got - want)) < 1e-7 * np.max(np.abs(orc.field(name, 2, stage=G - 1))) + 1e-7 * np.max(np.abs(want)), (g, name)
        ox, op, oq = orc.beam(stage=g)
        assert len(d["bq"]) == len(oq) and np.array_equal(d["bq"], oq)
    assert upd == upd_o > 100      # every electron released upstream was pushed downstream

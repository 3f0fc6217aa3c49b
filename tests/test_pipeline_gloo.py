"""Host logic of the xi-pipeline (qpad_b200/pipeline.py) over gloo on CPU, world_size 2 and 3.

The stage protocol -- which message goes to which neighbour at which point of a 3D step, the head/tail split, priming,
steady-state stepping and unwinding -- is exercised with a recording stand-in for the device object: every wire buffer
carries (kind, sender, step) and every unpack checks it got the message the reference's call order prescribes
(simulation_class.f03:303-340, 428-434, 458-493).  No CUDA, no oracle: this is the N>1 coverage of the CPU suite.
"""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

KINDS = {"beam_q": 1, "cu": 2, "b_spe": 3, "e": 4, "b": 5, "plasma": 6, "beam": 7}


def _view(ptr, n):
    return np.ctypeslib.as_array((ctypes.c_double * n).from_address(ptr))


class _Rec:
    def __init__(self, sim, kind, n):
        self.sim, self.kind, self.n = sim, kind, n

    def wire_count(self):
        return self.n

    def wire_cap(self):
        return (self.n - 1) // 7

    def _stamp(self, ptr):
        _view(ptr, 3)[:] = (KINDS[self.kind], self.sim.rank, self.sim.step)
        self.sim.log.append(("pack", self.kind, self.sim.step))

    def pack(self, *a):            # field.pack(slice, ptr) / species.pack(ptr)
        self._stamp(a[-1])

    pack_forward = pack

    def unpack(self, *a, add=False):   # field.unpack(slice, ptr[, add]) / species.unpack(ptr) / beam.unpack(ptr)
        ptr = a[-1] if len(a) < 3 else a[1]
        kind, sender, step = _view(ptr, 3)
        up = self.kind in ("e", "b")       # e, b travel backward (from rank+1), everything else forward (from rank-1)
        want = (KINDS[self.kind], self.sim.rank + (1 if up else -1), self.sim.step)
        assert (kind, sender, step) == want, f"rank {self.sim.rank} step {self.sim.step}: unpack {self.kind} got {(kind, sender, step)} want {want}"
        self.sim.log.append(("unpack", self.kind, self.sim.step))

    def upload(self, *a):
        pass

    def set_wire_cap(self, cap):
        pass


class FakeSim:
    """records the call order of one stage; step = number of begin_step() calls so far"""

    def __init__(self, rank):
        self.rank, self.step, self.log = rank, 0, []
        self._f = {k: _Rec(self, k, 16) for k in ("beam_q", "cu", "b_spe", "e", "b")}
        self.species, self.beam = _Rec(self, "plasma", 16), _Rec(self, "beam", 7 * 2 + 1)
        self.ctx = self

    def field(self, name): return self._f[name]
    def sync(self): pass
    def init_species(self, *a): pass
    def set_sweep_ctas(self, n): pass
    def beam_qdp_begin(self): self.step += 1; self.log.append(("qdp_begin", self.step))
    def beam_qdp_end(self): pass
    def begin_step(self): pass
    def run_slices(self, j0, j1): self.log.append(("slices", j0, j1, self.step))
    def beam_push(self): self.log.append(("beam_push", self.step))
    def renew(self): self.log.append(("renew", self.step))
    def close(self): pass


def _worker(rank, world, port, warm, timed, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from qpad_b200.pipeline import PipelineStage
    cfg = dict(nr=8, nz=8 * world + 1, zmin=0.0, zmax=1.0)
    beam = (np.zeros((4, 3)), np.zeros((4, 3)), np.zeros(4))
    plasma = (None, None, None, None, np.zeros(4))
    sim = FakeSim(rank)
    st = PipelineStage(cfg, plasma, beam, rank=rank, world=world, dist=dist, sim=sim,
                       make_buf=lambda n: torch.zeros(n, dtype=torch.float64))
    for _ in range(warm):
        st.step()
    st.prime()
    dist.barrier()                 # nothing may be pending here: a stage sits between a head and its tail
    heads_before = sum(1 for e in sim.log if e[0] == "qdp_begin")
    for _ in range(timed):
        st.step_primed()
    dist.barrier()
    heads = sum(1 for e in sim.log if e[0] == "qdp_begin") - heads_before
    st.unwind()
    st.drain()
    dist.barrier()
    out.put((rank, sim.step, heads, sum(1 for e in sim.log if e[0] == "renew"), st.nzp, st.noff2))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_stage_protocol_over_gloo(world):
    warm, timed = 2, 3
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29600 + world + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, world, port, warm, timed, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    res = sorted(out.get(timeout=5) for _ in range(world))
    total = warm + (world - 1) + timed
    for rank, steps, heads, renews, nzp, noff2 in res:
        assert steps == total and renews == total          # every stage finished every step the first one started
        assert heads == timed                              # exactly K slab sweeps per stage inside the timed region
    # slab partition rule of options_class.f03:103-106: remainder to the first stages, contiguous
    assert [r[4] for r in res] == [9] + [8] * (world - 1)
    assert [r[5] for r in res] == [0] + [9 + 8 * k for k in range(world - 1)]

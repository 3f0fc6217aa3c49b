"""The scheduling logic of pipeline.LocalPipeline (the default driver of bench.py) on CPU: the real host code runs against
a stand-in for the CUDA library and for torch's streams / events that RECORDS every stream operation; a small
discrete-event simulator then executes the recorded streams of all ranks and checks

  * deadlock freedom: every stream drains (event waits, flag waits of the peer-memory transport, in-kernel back hand-off);
  * data-race freedom by vector clocks: every read of a wire buffer happens-after the write it expects, every overwrite
    happens-after the read of the previous content (the back-pressure of the *_free events and ack flags);
  * content: a stage reads the record its neighbour produced for the SAME 3D step (fwd from stage g-1, back from g+1,
    crossing beam particles from g-1).

Ranks are host threads (their collectives rendezvous through a fake `dist`), so the cross-rank protocol of PeerLinks is
covered without a GPU.  This is a test of host logic only; the device code is covered by the `-m gpu` tests."""
import threading

import numpy as np
import pytest

from qpad_b200 import pipeline


# ---------------------------------------------------------------------------------------------------------------
class World:
    """everything the fake device of one test shares between its ranks"""

    def __init__(self):
        self.lock = threading.RLock()
        self.streams = {}
        self.flags, self.flag_sig = {}, {}
        self.bufs = {}
        self.next_addr = 1 << 44
        self.errors = []

    def alloc(self, nbytes):
        with self.lock:
            a = self.next_addr
            self.next_addr += (int(nbytes) + 4095) // 4096 * 4096 + 4096
            return a


WORLD = None


class Ticket:
    def __init__(self):
        self.done, self.vc = False, None


class FakeStream:
    def __init__(self, device=None):
        self.cuda_stream = WORLD.alloc(8)
        self.ops, self.vc = [], {}
        with WORLD.lock:
            WORLD.streams[self.cuda_stream] = self

    def push(self, kind, **kw):
        with WORLD.lock:
            self.ops.append(dict(kind=kind, **kw))

    def wait_event(self, ev):
        if ev.ticket is not None:
            self.push("wait_event", ticket=ev.ticket)

    def synchronize(self):
        pass


class FakeEvent:
    def __init__(self, enable_timing=False):
        self.ticket = None

    def record(self, stream=None):
        self.ticket = Ticket()
        stream.push("record", ticket=self.ticket)


def stream_of(cuda_stream):
    return WORLD.streams[cuda_stream]


class Buf:
    """one wire sub-buffer, identified by its address"""

    def __init__(self):
        self.tag, self.w_stream, self.w_clock = None, None, 0
        self.r_stream, self.r_clock, self.consumed = None, 0, True


def _buf(ptr):
    return WORLD.bufs.setdefault(ptr, Buf())


def do_write(ptr, tag, sid, vc, needs_consumer=True):
    b = _buf(ptr)
    if b.tag is not None and needs_consumer:
        if not b.consumed:
            WORLD.errors.append(f"overwrite of {b.tag} with {tag} before it was read")
        elif vc.get(b.r_stream, 0) < b.r_clock:
            WORLD.errors.append(f"overwrite of {b.tag} with {tag} does not happen-after its read")
    b.tag, b.w_stream, b.w_clock, b.consumed = tag, sid, vc[sid], False


def do_read(ptr, want, sid, vc):
    b = _buf(ptr)
    if b.tag != want:
        WORLD.errors.append(f"read {b.tag}, expected {want}")
    elif vc.get(b.w_stream, 0) < b.w_clock:
        WORLD.errors.append(f"read of {want} does not happen-after its write")
    b.r_stream, b.r_clock, b.consumed = sid, vc[sid], True


# ---------------------------------------------------------------------------------------------------------------
class FakeWireBuf:
    def __init__(self, nbytes=None, handle=None):
        self.ptr = int.from_bytes(handle, "little") if handle is not None else WORLD.alloc(nbytes)

    def data_ptr(self):
        return self.ptr

    def export(self):
        return self.ptr.to_bytes(8, "little")

    def close(self):
        pass


class FakeField:
    def __init__(self, sim, name):
        self.sim, self.name = sim, name

    def wire_count(self):
        return 64

    def pack(self, slice_idx, ptr):
        s = self.sim
        s.kernel(lambda sid, vc, tag=(self.name, s.g, s.cur - 1): do_write(ptr, tag, sid, vc))

    def unpack(self, slice_idx, ptr, add=False):
        s = self.sim
        back = self.name in ("b", "e")
        want = (self.name, s.g + 1, s.cur - 1) if back else (self.name, s.g - 1, s.cur)
        s.kernel(lambda sid, vc: do_read(ptr, want, sid, vc))

    def lineout(self, *a):
        return np.zeros(self.sim.nzp)


class FakeSpecies:
    def __init__(self, sim):
        self.sim = sim

    def wire_count(self):
        return 1 + 8 * 2 * self.sim.npp0

    def pack(self, ptr):
        s = self.sim
        s.kernel(lambda sid, vc, tag=("species", s.g, s.cur - 1): do_write(ptr, tag, sid, vc))

    def unpack(self, ptr):
        s = self.sim
        want = ("species", s.g - 1, s.cur)
        s.kernel(lambda sid, vc: do_read(ptr, want, sid, vc))

    def upload(self, *a):
        self.sim.kernel(lambda sid, vc: None)


class FakeBeam:
    def __init__(self, sim):
        self.sim, self.cap = sim, 128

    def upload(self, *a): pass
    def count_ptr(self): return 0
    def set_wire_cap(self, c): self.cap = int(c)
    def wire_cap(self): return self.cap
    def wire_count(self): return 7 * self.cap + 1

    def pack_forward(self, ptr):
        s = self.sim
        s.kernel(lambda sid, vc, tag=("beam", s.g, s.cur - 1): do_write(ptr, tag, sid, vc))

    def unpack(self, ptr):
        s = self.sim
        want = ("beam", s.g - 1, s.cur - 1)
        s.kernel(lambda sid, vc: do_read(ptr, want, sid, vc))


class FakeCtx:
    def launch_count(self): return 0


class FakeLaser:
    """the stage's envelope slab: qpg_laser_set_handoff + the hand-off protocol of qpg_laser_advance (csrc/laser.cu laser_launch_advance): message
    n = the n-th advance; wait in_ready >= n, copy the guard record in, in_ack = n, solve, wait out_ack >= n - 1, write the own record, out_ready = n"""

    def __init__(self, sim):
        self.sim, self.links, self.seq = sim, (None,) * 6, 0

    def guard_size(self): return 48
    def upload_slab(self, ar, ai, noff2): pass
    def sync(self): pass

    def set_handoff(self, guard_in=None, in_ready=None, in_ack=None, guard_out=None, out_ready=None, out_ack=None):
        self.links, self.seq = (guard_in, in_ready, in_ack, guard_out, out_ready, out_ack), 0

    def advance(self):
        s = self.sim
        gin, in_ready, in_ack, gout, out_ready, out_ack = self.links
        self.seq += 1
        n, step = self.seq, s.cur - 1
        s.kernel(lambda sid, vc: None)                                   # set_rhs
        if gin is not None:
            s.stream.push("wait_flag", flag=in_ready, value=n)
            s.kernel(lambda sid, vc, want=("lasg", s.g - 1, step): do_read(gin, want, sid, vc))
            s.stream.push("signal", flag=in_ack, value=n)
        s.kernel(lambda sid, vc: None)                                   # solve
        if gout is not None:
            if n > 1:
                s.stream.push("wait_flag", flag=out_ack, value=n - 1)
            s.kernel(lambda sid, vc, tag=("lasg", s.g, step): do_write(gout, tag, sid, vc))
            s.stream.push("signal", flag=out_ready, value=n)


class FakeSim:
    G_OF = {}          # (noff2) -> global stage index, filled by the test

    def __init__(self, nr, nz, max_mode, rmax, zmin, zmax, dt, sp_npmax=0, noff2=0, nzp=None, stream=None, **kw):
        self.nzp, self.noff2, self.npp0 = nzp, noff2, sp_npmax // 2
        self.g = FakeSim.G_OF[noff2]
        self.stream = stream_of(stream)
        self.cur = 0                     # number of slab sweeps started so far = index of the next 3D step of this stage
        self.beam, self.species, self.ctx = FakeBeam(self), FakeSpecies(self), FakeCtx()
        self.laser = FakeLaser(self) if kw.get("sp_push_pgc") else None
        self.handoff = None
        self.log = []

    def kernel(self, action):
        self.stream.push("kernel", action=action)

    def field(self, name): return FakeField(self, name)
    def init_species(self, *a): pass
    def set_sweep_ctas(self, n): pass
    def beam_qdp_begin(self): self.kernel(lambda sid, vc: None)
    def beam_qdp_raw(self): self.kernel(lambda sid, vc: None)
    def beam_qdp_fix(self): self.kernel(lambda sid, vc: None)
    def begin_step_zero(self): self.kernel(lambda sid, vc: None)
    def begin_step_add(self): self.kernel(lambda sid, vc: None)
    def beam_push(self): self.log.append(("push", self.cur - 1)); self.kernel(lambda sid, vc: None)
    def beam_push_interior(self): self.kernel(lambda sid, vc: None)
    def beam_qdp_part(self, part): self.kernel(lambda sid, vc: None)
    def beam_push_edge(self): self.log.append(("push", self.cur - 1)); self.kernel(lambda sid, vc: None)
    def renew(self): self.kernel(lambda sid, vc: None)
    def stats(self): return (0, 0, 0)
    def close(self): pass

    # a neutral species attached to the stage (qpg_sim_attach_neutral): its record travels with the forward message
    def set_laser_overlap(self, on): pass
    def laser_advance(self): self.laser.advance()
    def attach_neutral(self, *a, **kw): self.neutral = True
    def set_graph_unroll(self, on): pass
    def neutral_wire_count(self): return 96

    def neutral_pack(self, ptr):
        self.kernel(lambda sid, vc, tag=("neutral", self.g, self.cur - 1): do_write(ptr, tag, sid, vc))

    def neutral_unpack(self, ptr):
        want = ("neutral", self.g - 1, self.cur)
        self.kernel(lambda sid, vc: do_read(ptr, want, sid, vc))

    def set_back_handoff(self, wire_b, wire_e, flag, seq):
        self.handoff = (wire_b, wire_e, flag, seq)

    def run_slices(self, j0, j1):
        assert j0 == 1 and j1 == self.nzp
        self.log.append(("sweep", self.cur))
        h, self.handoff = self.handoff, None
        step = self.cur
        if h is not None:       # the sweep kernel writes b, e of its first slice and raises the flag
            wb, we, flag, seq = h

            def act(sid, vc):
                do_write(wb, ("b", self.g, step), sid, vc)
                do_write(we, ("e", self.g, step), sid, vc)
            self.kernel(act)
            self.stream.push("signal", flag=flag, value=seq)
        else:
            self.kernel(lambda sid, vc: None)
        self.cur += 1


class FakeCapi:
    WireBuf = FakeWireBuf
    Sim = FakeSim

    @staticmethod
    def stream_wait(cuda_stream, flag, value):
        stream_of(cuda_stream).push("wait_flag", flag=flag, value=value)

    @staticmethod
    def stream_wait_unless_empty(cuda_stream, count_ptr, flag, value):
        # the stand-in's stages all hold beam particles: the conditional wait of the backward hand-off is a wait
        stream_of(cuda_stream).push("wait_flag", flag=flag, value=value)

    @staticmethod
    def stream_signal(cuda_stream, flag, value):
        stream_of(cuda_stream).push("signal", flag=flag, value=value)


class FakeDist:
    """collectives of `world` rank threads"""

    def __init__(self, world):
        self.world, self.barrier_obj, self.slots, self.lock = world, threading.Barrier(world), {}, threading.Lock()
        self.local = threading.local()

    def set_rank(self, r):
        self.local.rank = r

    def all_gather_object(self, out, obj):
        with self.lock:
            self.slots[self.local.rank] = obj
        self.barrier_obj.wait()
        for r in range(self.world):
            out[r] = self.slots[r]
        self.barrier_obj.wait()

    def barrier(self, **kw):
        self.barrier_obj.wait()


def simulate():
    """execute the recorded streams; returns the list of streams that could not drain"""
    progress = True
    while progress:
        progress = False
        for sid, st in WORLD.streams.items():
            while st.ops:
                op = st.ops[0]
                k = op["kind"]
                if k == "wait_event":
                    if not op["ticket"].done:
                        break
                    for s_, c_ in op["ticket"].vc.items():
                        st.vc[s_] = max(st.vc.get(s_, 0), c_)
                elif k == "wait_flag":
                    if WORLD.flags.get(op["flag"], 0) < op["value"]:
                        break
                    m = min(v for (f, v) in WORLD.flag_sig if f == op["flag"] and v >= op["value"])
                    for s_, c_ in WORLD.flag_sig[(op["flag"], m)].items():
                        st.vc[s_] = max(st.vc.get(s_, 0), c_)
                st.vc[sid] = st.vc.get(sid, 0) + 1
                if k == "record":
                    op["ticket"].vc, op["ticket"].done = dict(st.vc), True
                elif k == "signal":
                    if WORLD.flags.get(op["flag"], 0) >= op["value"]:
                        WORLD.errors.append(f"flag {op['flag']:#x} does not count upwards: {op['value']}")
                    WORLD.flags[op["flag"]] = op["value"]
                    WORLD.flag_sig[(op["flag"], op["value"])] = dict(st.vc)
                elif k == "kernel":
                    op["action"](sid, st.vc)
                st.ops.pop(0)
                progress = True
    return [sid for sid, st in WORLD.streams.items() if st.ops]


# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture
def fake_device(monkeypatch):
    global WORLD
    WORLD = World()
    import torch
    real_zeros = torch.zeros
    monkeypatch.setattr(pipeline, "capi", FakeCapi)
    monkeypatch.setattr(torch.cuda, "Stream", FakeStream)
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "get_device_properties", lambda d: type("P", (), {"multi_processor_count": 148})())
    monkeypatch.setattr(torch, "zeros", lambda n, dtype=None, device=None: real_zeros(n, dtype=dtype))
    yield WORLD
    WORLD = None


def _inputs(nz):
    n = 40
    plasma = (np.zeros((n, 2)), np.zeros((n, 3)), np.ones(n), np.zeros(n), -np.ones(n))
    rng = np.random.default_rng(0)
    bx = np.zeros((50, 3)); bx[:, 2] = rng.uniform(0, nz * 0.25, 50)
    return dict(nr=64, nz=nz, max_mode=1, rmax=5.0, zmin=0.0, zmax=nz * 0.25, dt=10.0), plasma, (bx, np.zeros((50, 3)), -np.ones(50))


def _run_rank(cfg, plasma, beam, S, rank, world, dist, parts, nwaves, upload, out):
    try:
        if dist is not None:
            dist.set_rank(rank)
        lp = pipeline.LocalPipeline(cfg, plasma, beam, S, rank=rank, world=world, dist=dist, transport="p2p" if world > 1 else None, partition=parts,
                                    laser=(np.zeros((1, cfg["nz"] + 3, cfg["nr"] + 2)),) * 2 if cfg.get("laser") else None)
        lp.fill()
        for w in range(nwaves):
            lp.wave(upload=plasma if (upload and w % 2) else None)
        lp.drain()
        out[rank] = lp
    except Exception as exc:      # surfaces in the main thread
        out[rank] = exc
        if dist is not None:
            dist.barrier_obj.abort()


@pytest.mark.parametrize("world,S,parts", [(1, 1, None), (1, 2, None), (1, 4, None), (1, 3, [(0, 5), (5, 17), (22, 10)]),
                                           (2, 1, None), (2, 2, None), (3, 2, None), (4, 4, None)])
def test_schedule_is_deadlock_and_race_free(fake_device, world, S, parts):
    nz, nwaves = 32, 5
    cfg, plasma, beam = _inputs(nz)
    G = world * S
    FakeSim.G_OF = {noff: g for g, (noff, _) in enumerate(parts or pipeline.slab_partition(nz, G))}
    dist = FakeDist(world) if world > 1 else None
    out = {}
    threads = [threading.Thread(target=_run_rank, args=(cfg, plasma, beam, S, r, world, dist, parts, nwaves, world == 1, out)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=60)
    for r in range(world):
        assert not isinstance(out.get(r), Exception), out[r]
        assert r in out, "rank thread did not finish"
    stuck = simulate()
    assert not stuck, f"{len(stuck)} stream(s) cannot drain: deadlock"
    assert not fake_device.errors, fake_device.errors[:5]
    # every stage swept the same number of 3D steps (fill + waves + drain), a beam push after each sweep
    for r in range(world):
        for sim in out[r].sims:
            sweeps = [n for k, n in sim.log if k == "sweep"]
            pushes = [n for k, n in sim.log if k == "push"]
            assert sweeps == list(range(G - 1 + nwaves)) and pushes == sweeps, (r, sim.g, sweeps, pushes)
    # hand-offs really happened: every inter-stage buffer carries the record of the last step
    last = G - 2 + nwaves
    tags = {b.tag for b in fake_device.bufs.values() if b.tag is not None}
    if G > 1:
        assert ("species", 0, last) in tags and ("b", G - 1, last) in tags and ("beam", G - 2, last) in tags


def test_the_checker_catches_a_missing_wait(fake_device, monkeypatch):
    """sanity of the harness itself: without the flag wait that orders a stage's tail behind the downstream stage's first
    slice, the read of the guard-slice buffer is reported (wrong content or no happens-before edge)"""
    cfg, plasma, beam = _inputs(32)
    FakeSim.G_OF = {noff: g for g, (noff, _) in enumerate(pipeline.slab_partition(32, 2))}
    monkeypatch.setattr(FakeCapi, "stream_wait", staticmethod(lambda cuda_stream, flag, value: None))
    monkeypatch.setattr(FakeCapi, "stream_wait_unless_empty", staticmethod(lambda cuda_stream, count_ptr, flag, value: None))
    out = {}
    _run_rank(cfg, plasma, beam, 2, 0, 1, None, None, 4, False, out)
    assert not isinstance(out[0], Exception), out[0]
    assert not simulate()
    assert any("does not happen-after its write" in e or "expected ('b'" in e or "expected ('e'" in e for e in fake_device.errors), fake_device.errors[:3]


@pytest.mark.parametrize("world,S", [(1, 3), (2, 2), (3, 1)])
def test_neutral_record_travels_race_free(fake_device, world, S):
    """a deck with a neutral species: the record of qpg_sim_neutral_pack (released electrons, ion buffer, rho_ion, levels) is written into
    the next stage's buffer -- a device buffer inside a GPU, the `neu_in` peer buffer between ranks -- under the forward message's events /
    ready and ack words: every stage reads its upstream neighbour's record of the SAME 3D step, no buffer is overwritten before it was read"""
    nz, nwaves = 32, 4
    cfg, plasma, beam = _inputs(nz)
    cfg.update(neutral=dict(element=3, ion_max=2), ppc1=2, ppc2=2, num_theta=8)
    G = world * S
    FakeSim.G_OF = {noff: g for g, (noff, _) in enumerate(pipeline.slab_partition(nz, G))}
    dist = FakeDist(world) if world > 1 else None
    out = {}
    threads = [threading.Thread(target=_run_rank, args=(cfg, plasma, beam, S, r, world, dist, None, nwaves, False, out)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=60)
    for r in range(world):
        assert r in out and not isinstance(out[r], Exception), out.get(r)
    assert not simulate(), "deadlock"
    assert not fake_device.errors, fake_device.errors[:5]
    tags = {b.tag for b in fake_device.bufs.values() if b.tag is not None}
    assert ("neutral", G - 2, G - 2 + nwaves) in tags          # the last record of the last link


@pytest.mark.parametrize("world,S", [(1, 4), (2, 2), (3, 1)])
def test_envelope_guard_handoff_is_deadlock_and_race_free(fake_device, world, S):
    """a laser deck: LocalPipeline wires every stage's envelope links (buffer and ready word at the consumer, ack word at the producer; device
    buffers inside a GPU, PeerLinks' las_in / ready_las / ack_las between ranks) and calls the advance after the slab: with the library's
    hand-off protocol replayed by the stand-in, every advance reads the upstream stage's record of the SAME 3D step and no record is
    overwritten before it was read, for stages of one GPU and across ranks"""
    nz, nwaves = 32, 5
    cfg, plasma, beam = _inputs(nz)
    cfg.update(max_mode=0, laser=dict(k0=20.0, iteration=3), ppc1=2, ppc2=2, num_theta=8)
    G = world * S
    FakeSim.G_OF = {noff: g for g, (noff, _) in enumerate(pipeline.slab_partition(nz, G))}
    dist = FakeDist(world) if world > 1 else None
    out = {}
    threads = [threading.Thread(target=_run_rank, args=(cfg, plasma, beam, S, r, world, dist, None, nwaves, False, out)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=60)
    for r in range(world):
        assert r in out and not isinstance(out[r], Exception), out.get(r)
    assert not simulate(), "deadlock"
    assert not fake_device.errors, fake_device.errors[:5]
    tags = {b.tag for b in fake_device.bufs.values() if b.tag is not None}
    assert ("lasg", G - 2, G - 2 + nwaves) in tags
    for r in range(world):
        for sim in out[r].sims:
            assert sim.laser.seq == G - 1 + nwaves          # one advance per 3D step of the stage


def test_the_checker_catches_a_missing_envelope_ack(fake_device, monkeypatch):
    """sanity of the harness for the envelope link: a producer that does not wait for the consumer's ack before it rewrites the one wire
    record of the link is reported"""
    def no_ack_wait(self):
        s = self.sim
        gin, in_ready, in_ack, gout, out_ready, out_ack = self.links
        self.seq += 1
        n, step = self.seq, s.cur - 1
        if gin is not None:
            s.stream.push("wait_flag", flag=in_ready, value=n)
            s.kernel(lambda sid, vc, want=("lasg", s.g - 1, step): do_read(gin, want, sid, vc))
            s.stream.push("signal", flag=in_ack, value=n)
        if gout is not None:
            s.kernel(lambda sid, vc, tag=("lasg", s.g, step): do_write(gout, tag, sid, vc))
            s.stream.push("signal", flag=out_ready, value=n)
    monkeypatch.setattr(FakeLaser, "advance", no_ack_wait)
    cfg, plasma, beam = _inputs(32)
    cfg.update(max_mode=0, laser=dict(k0=20.0, iteration=3), ppc1=2, ppc2=2, num_theta=8)
    FakeSim.G_OF = {noff: g for g, (noff, _) in enumerate(pipeline.slab_partition(32, 3))}
    out = {}
    _run_rank(cfg, plasma, beam, 3, 0, 1, None, None, 5, False, out)
    assert not isinstance(out[0], Exception), out[0]
    simulate()
    assert any("lasg" in e and "happen-after" in e for e in fake_device.errors), fake_device.errors[:3]

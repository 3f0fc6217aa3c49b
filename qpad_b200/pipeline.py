"""Host-side drivers of a 3D step: single stage, and the xi-pipeline over one GPU per stage.

Mirrors the stage logic of simulation_class.f03:294-512 / parallel_module.f03:221-239: `nodes(2)` stages own
contiguous xi slabs (options_class.f03:103-106), stage s works on 3D step n while stage s+1 works on step n-1.
The reference's MPI isend/recv pairs become torch.distributed (NCCL over NVLink) p2p transfers of the library's
wire buffers (qpg_*_pack / qpg_*_unpack); torch is plumbing only (device buffers, streams, process group).

Per 3D step and stage boundary:
  forward : beam q guard slice (add), plasma particles (8 fp64 each), last-slice cu and b_spe
  backward: first-slice e and b into the upstream guard slice nzp+1
  forward : beam particles that crossed the slab edge (7 fp64 each)
"""
import os
import sys
import time

import numpy as np

from . import capi

_TRACE = bool(os.environ.get("QPG_TRACE"))


def _trace(rank, msg):
    if _TRACE:
        print(f"[{time.time() % 1000:8.3f}] rank {rank}: {msg}", file=sys.stderr, flush=True)


def slab_partition(nz, nstages):
    """noff/ndp rule of options_class.f03:103-106 (remainder goes to the first stages)."""
    local, extra = nz // nstages, nz % nstages
    out = []
    for k in range(nstages):
        out.append((local * k + min(k, extra), local + (1 if k < extra else 0)))
    return out


def balanced_partition(cost, nstages, min_len=2):
    """Contiguous xi slabs of (nearly) equal summed cost instead of equal length: the pipeline advances at the pace of
    its slowest stage, and the cost of a slice varies along xi (predictor-corrector iterations inside the wake, clustered
    particles).  `cost[j]` = measured cost of slice j (qpg_sim_slice_trace).  Minimises the largest slab cost (bisection
    on the bound + greedy fill), every slab at least `min_len` slices.  Returns [(noff, nzp)] like slab_partition."""
    cost = np.maximum(np.asarray(cost, dtype=np.float64), 1e-30)
    nz = len(cost)
    if nstages * min_len > nz:
        raise ValueError("more stages than slices")

    def fill(bound):
        cuts, acc, start = [], 0.0, 0
        for j in range(nz):
            left = nstages - len(cuts) - 1                       # slabs still to be opened after the current one
            must_cut = nz - j == left * min_len and j - start >= min_len      # the rest is needed for the remaining slabs
            if j > start and ((acc + cost[j] > bound and j - start >= min_len and left > 0) or must_cut):
                cuts.append(j); acc, start = 0.0, j
            acc += cost[j]
        return cuts

    def worst(cuts):
        edges = [0] + cuts + [nz]
        return max(cost[a:b].sum() for a, b in zip(edges[:-1], edges[1:]))

    lo, hi = cost.sum() / nstages, cost.sum()
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        c = fill(mid)
        if len(c) == nstages - 1 and worst(c) <= mid * (1 + 1e-12):
            hi = mid
        else:
            lo = mid
    cuts = fill(hi)
    while len(cuts) < nstages - 1:                               # (only if the bound was never binding) split the longest slab
        edges = [0] + cuts + [nz]
        k = int(np.argmax(np.diff(edges)))
        cuts = sorted(cuts + [(edges[k] + edges[k + 1]) // 2])
    edges = [0] + cuts + [nz]
    return [(a, b - a) for a, b in zip(edges[:-1], edges[1:])]


def split_beam(bx, bp, bq, nz, dxi, nstages, parts=None, spin=None):
    """owner stage of each beam particle: the slab [noff2, noff2+nzp)*dxi that holds xi (part3d_comm.f03 goto_here); with `spin` (s[np][3])
    every stage's tuple gets its particles' spin vectors as a fourth array"""
    edges = [noff * dxi for noff, _ in (parts or slab_partition(nz, nstages))][1:]
    owner = np.searchsorted(np.asarray(edges), bx[:, 2], side="right") if nstages > 1 else np.zeros(len(bq), int)
    arrays = (bx, bp, bq) + ((spin,) if spin is not None else ())
    return [tuple(np.ascontiguousarray(a[owner == k]) for a in arrays) for k in range(nstages)]


def _make_sim(cfg, npp0, nbeam, stream, device, use_graph, noff2=0, nzp=None, beam_cap=None):
    cuda_stream = None if stream is None else stream.cuda_stream
    las = cfg.get("laser")
    pgc = {} if not las else dict(sp_push_pgc=1, laser_iter=las["iteration"], laser_k0=las["k0"], sp_ppc_r=cfg["ppc1"], beam_evol=0 if nbeam == 0 else 1)
    return capi.Sim(cfg["nr"], cfg["nz"], cfg["max_mode"], cfg["rmax"], cfg["zmin"], cfg["zmax"], cfg["dt"],
                    sp_qbm=-1.0, sp_npmax=max(2 * npp0, 64), beam_qbm=-1.0, beam_npmax=beam_cap or (nbeam + 1024),
                    iter_max=cfg.get("iter_max", 1), iter_reltol=cfg.get("iter_reltol", 1e-3),
                    iter_abstol=cfg.get("iter_abstol", 1e-3), sort_freq=cfg.get("sort_freq", 0), use_graph=use_graph,
                    noff2=noff2, nzp=nzp, device=device, stream=cuda_stream, **pgc)


def kernel_microbench(sim, cfg, peak_gbs, n_big=4 * 1024 * 1024, reps=5):
    """stream-from-HBM numbers for the three particle kernels: a lattice 16x larger than the L2 can hold
    (n_big particles x 64 B = 268 MB), fields = the current slice's e and b of `sim`"""
    from . import decks
    s = sim
    nth = max(16, n_big // (cfg["nr"] * cfg["ppc1"] * cfg["ppc2"]))
    x, p, g, psi, q = decks.plasma_uniform(cfg["nr"], cfg["rmax"], cfg["ppc1"], cfg["ppc2"], nth)
    n = len(q)
    rng = np.random.default_rng(0)
    p = 0.3 * rng.standard_normal(p.shape)
    g = np.sqrt(1 + (p ** 2).sum(1))
    part = capi.Part2d(s.ctx, -1.0, n + 64)
    part.upload(x, p, g, psi, q)
    e, b = s.field("e"), s.field("b")
    cu, dcu, amu = capi.Field(s.ctx, 3), capi.Field(s.ctx, 2), capi.Field(s.ctx, 3)
    fq = capi.Field(s.ctx, 1)
    res = {}
    s.ctx.tprof_reset(); s.ctx.tprof_enable(True)
    for _ in range(reps + 2):
        part.qdeposit(fq)
        part.amjdeposit_robust(e, b, cu, amu, dcu, 1e-3)
        part.push_u_robust(e, b, 1e-3)
        part.push_x(1e-6)
    for ev in ("kernel amjdeposit", "kernel qdeposit", "kernel push"):
        res[ev] = s.ctx.tprof_get(ev)
    s.ctx.tprof_enable(False)
    out = {"particles": n, "note": "stand-alone particle kernels on a particle set larger than L2 (CUDA events per launch); algorithmic bytes: amjdeposit 64 B, qdeposit 24 B, push_u 72 B + push_x 64 B per particle"}
    ms, nc = res["kernel amjdeposit"]
    out["amjdeposit_GBs"] = 64.0 * n / (ms / nc * 1e-3) / 1e9
    ms, nc = res["kernel qdeposit"]
    out["qdeposit_GBs"] = 24.0 * n / (ms / nc * 1e-3) / 1e9
    ms, nc = res["kernel push"]  # push_u and push_x launches alternate: 136 B per pair
    out["push_u_plus_push_x_GBs"] = 136.0 * n / (2 * ms / nc * 1e-3) / 1e9
    out["peak"] = peak_gbs
    for k in ("amjdeposit", "qdeposit", "push_u_plus_push_x"):
        out[k + "_frac"] = out[k + "_GBs"] / peak_gbs
    part.close()
    return out


class SingleStage:
    """nodes = [1, 1]: the whole box on one GPU."""

    def __init__(self, cfg, plasma, beam, stream=None, rank=0, world=1, device=0, use_graph=1):
        self.cfg, self.plasma = cfg, plasma
        self.sim = _make_sim(cfg, len(plasma[4]), len(beam[2]), stream, device, use_graph)
        self.sim.init_species(*plasma)
        self.sim.beam.upload(*beam)

    def step(self):
        self.sim.step3d()

    # pieces used by bench.py's roofline leg
    def prepare_step(self):
        s = self.sim
        s.beam_qdp_begin(); s.beam_qdp_end(); s.begin_step()

    def finish_step(self):
        self.sim.renew()

    def step_e2e(self):
        """one step through host buffers: the freshly injected plasma goes host -> device (species%renew on the
        host side, as the Fortran driver does), results come back as on-axis line-outs"""
        s = self.sim
        x, p, g, psi, q = self.plasma
        s.species.upload(x, p, g, psi, q)
        s.beam_qdp_begin(); s.beam_qdp_end(); s.begin_step()
        s.run_slices(1, s.nzp)
        s.beam_push()
        ez = s.field("e").lineout(3, 0, 1)
        ps = s.field("psi").lineout(1, 0, 1)
        st = s.stats()
        self.last = (ez, ps, st)
        return 8 * 8 * len(q), 8 * (len(ez) + len(ps)) + 24

    def kernel_microbench(self, peak_gbs, n_big=4 * 1024 * 1024, reps=5):
        return kernel_microbench(self.sim, self.cfg, peak_gbs, n_big, reps)

    def close(self):
        self.sim.close()


class PipelineStage:
    """One rank = one xi slab on one GPU; consecutive ranks are consecutive pipeline stages."""

    def __init__(self, cfg, plasma, beam, stream=None, rank=0, world=1, device=0, use_graph=1, dist=None, make_buf=None, sim=None,
                 beam_wire_cap=None):
        """`sim`, `dist`, `make_buf` are injection points for the host-logic tests (tests/test_pipeline_gloo.py drives the
        stage protocol over gloo with a recording stand-in for the device object); production code leaves them None."""
        import torch
        import torch.distributed as tdist
        self.dist = dist or tdist
        self.rank, self.world, self.cfg, self.plasma = rank, world, cfg, plasma
        self.noff2, self.nzp = slab_partition(cfg["nz"], world)[rank]
        dxi = (cfg["zmax"] - cfg["zmin"]) / cfg["nz"]
        mine = split_beam(*beam, cfg["nz"], dxi, world)[rank]
        self.sim = sim if sim is not None else _make_sim(cfg, len(plasma[4]), len(mine[2]), stream, device, use_graph, self.noff2, self.nzp, beam_cap=len(beam[2]) + 1024)
        self.sim.init_species(*plasma)
        self.sim.beam.upload(*mine)
        s = self.sim
        if world > 1 and make_buf is None:
            # leave a few SMs to the NCCL send/recv kernels that overlap the slab sweep (a cooperative kernel that fills
            # the GPU would delay the backward e/b hand-off until the slab is done and serialise the pipeline)
            s.set_sweep_ctas(-int(os.environ.get("QPG_PIPELINE_FREE_SMS", "4")))
        mk = make_buf or (lambda n: torch.zeros(n, dtype=torch.float64, device=torch.device("cuda", device)))
        # message sizes: plasma particles only ever leave, so the live prefix of the wire buffer is bounded by the injected
        # lattice; beam particles crossing a slab edge in one step are few (the reference reserves 0.1 npmax and stops on
        # overflow, so does this library -- with a smaller default so the hand-off stays off the critical path)
        self.n_plasma_wire = min(s.species.wire_count(), 1 + 8 * len(plasma[4]))
        s.beam.set_wire_cap(beam_wire_cap if beam_wire_cap is not None else max(16384, len(beam[2]) // 64))
        # One message per direction and step: forward [beam-q guard slice | cu | b_spe | plasma particles] (the plasma
        # record is last so that only its live prefix travels), backward [b | e] of the first slice.  Separate in / out
        # buffers because a middle stage receives and sends the same kinds of message.
        nq, ncu, nbs = s.field("beam_q").wire_count(), s.field("cu").wire_count(), s.field("b_spe").wire_count()
        self.off_fwd = (0, nq, nq + ncu, nq + ncu + nbs)
        self.n_fwd = nq + ncu + nbs + self.n_plasma_wire
        self.fwd_in, self.fwd_out = mk(nq + ncu + nbs + s.species.wire_count()), mk(nq + ncu + nbs + s.species.wire_count())
        nb = s.field("b").wire_count()
        self.off_back = (0, nb)
        self.back_in, self.back_out = mk(nb + s.field("e").wire_count()), mk(nb + s.field("e").wire_count())
        self.buf_beam, self.buf_beam_in = mk(s.beam.wire_count()), mk(s.beam.wire_count())
        self.first, self.last = rank == 0, rank == world - 1
        self.stream = stream
        self.torch = torch
        self.comm = torch.cuda.Stream(device=device) if (stream is not None and make_buf is None) else None
        self.pending = {}
        self.pending_tail = False
        self._ev_on = bool(os.environ.get("QPG_TRACE_EVENTS")) and stream is not None and make_buf is None
        self._evs = []

    # optional device-time trace of one stage's step (QPG_TRACE_EVENTS=1): CUDA events on the compute stream at the marks
    def _mark(self, name):
        if not self._ev_on:
            return
        ev = self.torch.cuda.Event(enable_timing=True)
        ev.record(self.stream)
        self._evs.append((name, ev))

    def event_report(self):
        """average device time between consecutive marks over the recorded steps (ms)"""
        if not self._evs:
            return {}
        self.torch.cuda.synchronize()
        acc, cnt = {}, {}
        for (n0, e0), (n1, e1) in zip(self._evs[:-1], self._evs[1:]):
            k = f"{n0}->{n1}"
            acc[k] = acc.get(k, 0.0) + e0.elapsed_time(e1)
            cnt[k] = cnt.get(k, 0) + 1
        out = {k: round(acc[k] / cnt[k], 3) for k in acc}
        return out

    # mpi_isend analogue: the transfer runs on the communication stream, the compute stream carries on.  The buffer
    # is only repacked after _wait(name) (the reference's mpi_wait before every pipe_send, simulation_class.f03:430).
    def _isend(self, name, t, dst, after=None):
        if self.comm is None:
            if self.stream is None:
                self.sim.ctx.sync()
            self.pending[name] = self.dist.isend(t, dst)
            return
        if after is not None:
            self.comm.wait_event(after)
        else:
            self.comm.wait_stream(self.stream)
        with self.torch.cuda.stream(self.comm):
            self.pending[name] = self.dist.isend(t, dst)

    def _irecv(self, name, t, src):
        if self.comm is None:
            self.pending[name] = self.dist.irecv(t, src)
            return
        with self.torch.cuda.stream(self.comm):
            self.pending[name] = self.dist.irecv(t, src)

    def _wait(self, name):
        w = self.pending.pop(name, None)
        if w is not None:
            w.wait()

    def _recv(self, t, src):
        self.dist.recv(t, src)
        if self.comm is None and self.stream is None:
            pass

    # A 3D step of one stage = head (everything up to the packed hand-off buffers) + tail (the hand-offs that need the
    # downstream stage, the beam push and the renewal).  step() = head() + tail(); bench.py primes the pipeline by
    # letting stage r run (world-1-r) steps ahead and stops every stage between a head and its tail, where no
    # send is pending -- the reference reaches the same steady state after world-1 steps of fill.
    def head(self):
        s, r = self.sim, self.rank
        _trace(r, "head")
        # species%precv, cu / b_spe pipe_recv and the beam q guard slice: everything stage r-1 hands forward arrives
        # when it has finished its slab (simulation_class.f03:303-340; the guard slice is taken at the same point so
        # that an upstream stage never waits for a downstream one)
        self._mark("head")
        s.beam_qdp_begin()                                              # beam3d_class.f03:207
        fin = lambda k: self.fwd_in.data_ptr() + 8 * self.off_fwd[k]
        if not self.first:
            self._recv(self.fwd_in[:self.n_fwd], r - 1)
            s.field("beam_q").unpack(1, fin(0), add=True)
        self._mark("recv_fwd")
        s.beam_qdp_end()                                                # :210
        s.begin_step()
        if not self.first:
            s.species.unpack(fin(3))
            s.field("cu").unpack(0, fin(1))
            s.field("b_spe").unpack(0, fin(2))
        # first slice, then the backward hand-off of e and b (:460-467), then the rest of the slab
        self._mark("slice1")
        s.run_slices(1, 1)
        self._mark("slice1_done")
        ev = None
        if not self.first:
            self._wait("back")
            s.field("b").pack(1, self.back_out.data_ptr() + 8 * self.off_back[0])
            s.field("e").pack(1, self.back_out.data_ptr() + 8 * self.off_back[1])
            if self.comm is not None:
                ev = self.torch.cuda.Event()
                ev.record(self.stream)
        # The rest of the slab is enqueued BEFORE the NCCL calls: posting a send / receive can block the host for
        # milliseconds (measured 2.5 ms per step), which must not keep the sweep kernel from starting.  The
        # communication stream waits for the pack kernels only (event), not for the sweep.
        self._mark("sweep")
        if self.nzp > 1:
            s.run_slices(2, self.nzp)
        self._mark("sweep_done")
        if not self.first:
            self._isend("back", self.back_out, r - 1, after=ev)
            # the beam particles stage r-1 pushes across the slab edge in THIS step arrive while the slab is swept:
            # post the receive now on the communication stream, consume it in tail() (part3d_comm.f03:278-314)
            self._irecv("beam_in", self.buf_beam_in, r - 1)
        if not self.last:                                               # :210-215, :429-434, :472-474
            fout = lambda k: self.fwd_out.data_ptr() + 8 * self.off_fwd[k]
            self._wait("fwd")
            s.field("beam_q").pack(self.nzp + 1, fout(0))
            s.field("cu").pack(0, fout(1))
            s.field("b_spe").pack(0, fout(2))
            s.species.pack(fout(3))
        self.pending_tail = True

    def tail(self):
        s, r = self.sim, self.rank
        _trace(r, "tail")
        self._mark("tail")
        if not self.last:
            self._isend("fwd", self.fwd_out[:self.n_fwd], r + 1)
            self._recv(self.back_in, r + 1)                              # :482-483
            s.field("b").unpack(self.nzp + 1, self.back_in.data_ptr() + 8 * self.off_back[0])
            s.field("e").unpack(self.nzp + 1, self.back_in.data_ptr() + 8 * self.off_back[1])
        # beam push + forward hand-off                                  (:489-493, part3d_comm.f03:278-314)
        self._mark("beam_push")
        s.beam_push()
        if not self.first:
            self._wait("beam_in")
            s.beam.unpack(self.buf_beam_in.data_ptr())
        if not self.last:
            self._wait("beam"); s.beam.pack_forward(self.buf_beam.data_ptr()); self._isend("beam", self.buf_beam, r + 1)
        self._mark("renew")
        s.renew()                                                       # :498-501
        self._mark("step_end")
        self.pending_tail = False

    def step(self):
        self.head()
        self.tail()

    def prime(self):
        """after any number of complete steps: run ahead until this stage has finished the head of its
        (world-1-rank)-th extra step.  Leaves every stage between a head and its tail (the last one idle)."""
        for k in range(self.world - 1 - self.rank):
            if self.pending_tail:
                self.tail()
            self.head()

    def step_primed(self):
        """one steady-state step of a primed stage: the pending tail, then the next head"""
        if self.pending_tail:
            self.tail()
        self.head()

    def unwind(self):
        """finish every step the first stage has started"""
        if self.pending_tail:
            self.tail()
        for k in range(self.rank):
            self.step()

    def drain(self):
        for k in list(self.pending):
            self._wait(k)

    def close(self):
        self.drain()
        self.sim.close()


def probe_slice_costs(cfg, plasma, beam, device=0, ctas=0, steps=2):
    """Cost profile of the deck along xi: one stage sweeps the whole box `steps` times with the CTA count a pipeline
    stage will have; returns (ns per slice, PC iterations per slice, ns of beam deposit + push + move per slice) of the
    last step (qpg_sim_slice_trace; the beam kernels' time is spread over the slices like the beam particles)."""
    sim = _make_sim(cfg, len(plasma[4]), len(beam[2]), None, device, 1)
    try:
        if ctas:
            sim.set_sweep_ctas(ctas)
        sim.init_species(*plasma)
        sim.beam.upload(*beam)
        for k in range(steps):
            if k == steps - 1:
                sim.ctx.tprof_reset(); sim.ctx.tprof_enable(True)
            sim.step3d()
        beam_ms = sum(sim.ctx.tprof_get(ev)[0] for ev in ("deposit 3D particles", "push 3D particles", "move 3D particles"))
        sim.ctx.tprof_enable(False)
        ns, it = sim.slice_trace()
        bx = sim.beam.download()[0]
        dxi = (cfg["zmax"] - cfg["zmin"]) / cfg["nz"]
        cnt = np.bincount(np.clip((bx[:, 2] / dxi).astype(np.int64), 0, cfg["nz"] - 1), minlength=cfg["nz"]).astype(np.float64)
        beam_ns = cnt * (beam_ms * 1e6 / max(cnt.sum(), 1.0))
        return ns, it, beam_ns
    finally:
        sim.close()


def _slab_cost(ns, beam_ns, stages_per_gpu):
    k = 8                                                     # smooth over a few slices: single-slice timer noise is not load
    cost = np.convolve(np.pad(ns, (k // 2, k - 1 - k // 2), mode="edge"), np.ones(k) / k, mode="valid")
    return cost + float(os.environ.get("QPG_BALANCE_BEAM_WEIGHT", stages_per_gpu)) * beam_ns


def measured_partition(lp, cfg, beam_ns=None, nwaves=4, tol=0.03, nwarm=1):
    """Closed loop of the slab balancing: the RUNNING pipeline is the probe.  Fills it, runs `nwaves` waves with CUDA-event marks on
    every stage's stream, and takes (a) each stage's busy time per wave = everything between its marks except the two blocking waits
    for the neighbours' messages -- sweep kernel, beam deposit / push (which run on the stage's own SM share once its sweep kernel has
    left), hand-off kernels, fills -- and (b) the per-slice device times inside the sweep kernels (qpg_sim_slice_trace, the shape of
    the cost along xi), both measured while all stages of the GPU run side by side, which the open-loop probe_partition cannot see.
    Cost of slice j of stage g = its trace share of the stage's sweep time + an equal share of the stage's non-sweep time; slabs of
    equal cost are cut from that profile.  Returns (partition or None if the spread max/mean - 1 of the busy times is already below
    `tol`, per-stage busy ms, spread).  The same answer on every rank.
    The cost of a slab drifts with the 3D step (the beam focuses and the wake deepens over the first betatron quarter period), so the
    caller measures over the step numbers it cares about: `nwarm` waves after the fill are skipped, `nwaves` are measured."""
    lp.fill()
    for _ in range(max(nwarm, 1)):
        lp.wave()
    lp.sync()
    on = lp._ev_on
    lp._ev_on = True
    lp.trace_reset()
    for sim in lp.sims:
        sim.sweep_profile(reset=True)
    for _ in range(nwaves):
        lp.wave()
    reps = lp.event_report()
    lp._ev_on = on
    lp.trace_reset()
    mine = []
    for sim, rep in zip(lp.sims, reps):
        busy = sum(v for k, v in rep.items() if k not in ("w_fwd>got_fwd", "pre>got_back"))        # ms per wave
        sweep = sim.sweep_profile()["ns_total"] * 1e-6 / nwaves
        mine.append((sim.slice_trace()[0], busy, sweep))
    every = [mine]
    if lp.world > 1:
        every = [None] * lp.world
        lp.dist.all_gather_object(every, mine)
    stages = [t for m in every for t in m]
    busy_ms = [b for _, b, _ in stages]
    spread = max(busy_ms) / (sum(busy_ms) / len(busy_ms)) - 1.0
    if spread < tol or sum(len(t) for t, _, _ in stages) != cfg["nz"]:
        return None, busy_ms, spread
    cost = []
    for tr, busy, sweep in stages:
        k = min(8, len(tr))
        sm = np.convolve(np.pad(tr, (k // 2, k - 1 - k // 2), mode="edge"), np.ones(k) / k, mode="valid")
        sweep = min(sweep, busy)
        cost.append(sm * (sweep / max(sm.sum() * 1e-6, 1e-12)) * 1e-6 + max(busy - sweep, 0.0) / len(tr))
    return balanced_partition(np.concatenate(cost), lp.G, min_len=min(16, cfg["nz"] // lp.G)), busy_ms, spread


def probe_partition(cfg, plasma, beam, nstages_total, stages_per_gpu, device=0, rank=0, world=1, dist=None, free_sms=0, with_beam_cost=False):
    """cost-balanced slab partition for a pipeline of `nstages_total` stages; with several ranks rank 0 measures and
    everybody uses its answer (the partition must be the same on every rank).  Cost of a slice = its sweep time with a
    stage's CTA count + its share of the beam deposit / push (measured on the whole GPU, a stage has 1/stages_per_gpu of
    it while its neighbours sweep)."""
    import torch
    parts = beam_ns = None
    if rank == 0:
        nsm = torch.cuda.get_device_properties(device).multi_processor_count
        ns, _, beam_ns = probe_slice_costs(cfg, plasma, beam, device, (nsm - free_sms) // stages_per_gpu if nstages_total > 1 else 0)
        parts = balanced_partition(_slab_cost(ns, beam_ns, stages_per_gpu), nstages_total, min_len=min(16, cfg["nz"] // nstages_total))
    if world > 1:
        box = [parts, beam_ns]
        dist.broadcast_object_list(box, src=0)
        parts, beam_ns = [tuple(p) for p in box[0]], box[1]
    return (parts, beam_ns) if with_beam_cost else parts


class PeerLinks:
    """Peer-memory links of one rank to its two neighbours (csrc/p2p.cu): the wire buffers this rank CONSUMES and one
    block of flag words live in its own memory and are exported through CUDA IPC handles; the neighbours map them and
    write into them over NVLink.  Flag words (32 bytes apart, one writer each, counting messages):

        ready_fwd, ready_beam, ack_back, ready_las   written by rank-1      ready_back, ack_fwd, ack_beam, ack_las   written by rank+1
    (the *_las pair and the las_in buffer: the laser envelope's guard slices, capi.Laser.set_handoff)
    """
    FLAGS = ("ready_fwd", "ready_beam", "ack_back", "ready_back", "ack_fwd", "ack_beam", "ready_las", "ack_las")

    def __init__(self, dist, rank, world, n_fwd, n_back, n_beam, n_las=0, n_neu=0):
        self.dist, self.rank, self.world = dist, rank, world
        self.own = {"flags": capi.WireBuf(32 * len(self.FLAGS))}
        if rank > 0:
            self.own["fwd_in"] = capi.WireBuf(8 * n_fwd)
            self.own["beam_in"] = capi.WireBuf(8 * n_beam)
            if n_las:
                self.own["las_in"] = capi.WireBuf(8 * n_las)
            if n_neu:
                self.own["neu_in"] = capi.WireBuf(8 * n_neu)          # the neutral's record travels with the forward message (same ready / ack words)
        if rank < world - 1:
            self.own["back_in"] = capi.WireBuf(8 * n_back)
        mine = {k: b.export() for k, b in self.own.items()}
        every = [None] * world
        dist.all_gather_object(every, mine)
        self.up = {k: capi.WireBuf(handle=every[rank - 1][k]) for k in ("flags", "back_in")} if rank > 0 else {}
        self.down = {k: capi.WireBuf(handle=every[rank + 1][k]) for k in ("flags", "fwd_in", "beam_in") + (("las_in",) if n_las else ()) + (("neu_in",) if n_neu else ())} if rank < world - 1 else {}
        self.count = {}

    def next(self, link):
        """number of the next message on a link (both ends count alike: one message per link and 3D step)"""
        self.count[link] = self.count.get(link, 0) + 1
        return self.count[link]

    def flag(self, where, name):
        base = {"own": self.own, "up": self.up, "down": self.down}[where]["flags"].ptr
        return base + 32 * self.FLAGS.index(name)

    def close(self):
        self.dist.barrier()          # nobody unmaps or frees while a neighbour may still write
        for grp in (self.up, self.down):
            for b in grp.values():
                b.close()
        self.dist.barrier()
        for b in self.own.values():
            b.close()


class LocalPipeline:
    """The xi-pipeline over SM partitions: S stages = S persistent sweep kernels per GPU, each on its own stream with 1/S
    of the SMs; with world > 1 the G = world*S stages continue across GPUs (NCCL between the last stage of rank k and
    the first stage of rank k+1).

    Why: a slab sweep is a chain of latency-bound phases (field programs on a 32-CTA team, grid barriers) in which
    most SMs idle about a third of the time, and its particle phases scale with the SM count only down to a point --
    measured at C2: 148 CTAs 40 us/slice, 74 CTAs 55 us/slice, 37 CTAs 88 us/slice.  The quasi-static loop already has
    the concurrency to fill that idle time: the reference's own pipeline over xi slabs (parallel_module.f03:221-239),
    stage s working on 3D step n-s.  Here the stages are SM partitions of a B200 instead of MPI ranks; hand-offs inside
    a GPU are pack / unpack through device buffers ordered by CUDA events (no NCCL, no copies through the host).

    One `wave()` = every stage advances by one 3D step (tail of its previous step, then head of the next), i.e. one
    full deck's worth of slices in steady state; the first G-1 waves fill the pipeline.  All NCCL messages of a wave are
    matched inside the same wave on both ranks, so nothing is pending between waves (barrier + synchronize are safe).

    transport (world > 1): "p2p" = the pack kernels of the last stage of rank k write straight into the memory of rank
    k+1 and raise a flag its stream waits on (PeerLinks / csrc/p2p.cu: no library kernel competes with the sweep kernels
    for SMs, nothing blocks the host); "nccl" = torch.distributed send/recv of local wire buffers.
    """

    def __init__(self, cfg, plasma, beam, nstages, device=0, beam_wire_cap=None, rank=0, world=1, dist=None, transport=None, partition=None, laser=None,
                 beam_spin=None, graph_unroll=None):
        """laser = (a_r, a_i): the launched envelope of the whole box (capi.Laser layout) for a cfg with a "laser" block (robust_pgc plasma): every
        stage holds its slab of the envelope and advances it after its sweep, the new last two slices travel to the next stage's guards.
        A cfg with a "neutral" block (field ionisation, decks.CONFIGS["C5"]): every stage attaches the neutral species to its sim (per-slice
        launch path instead of the sweep kernel: the stages overlap as concurrent streams of small kernels) and the neutral's state --
        released electrons, ion buffer, rho_ion, levels -- travels forward with the plasma hand-off (neutral_class.f03:1025-1101).
        beam_spin = (s[np][3], amm): the beam carries spin vectors (part3d%has_spin): pushed with the particles, 10-real hand-off records."""
        import torch
        self.torch, self.cfg, self.S, self.plasma = torch, cfg, nstages, plasma
        self.rank, self.world, self.G, self.base = rank, world, world * nstages, rank * nstages
        self.dist = dist
        S, G = nstages, self.G
        dxi = (cfg["zmax"] - cfg["zmin"]) / cfg["nz"]
        # xi slabs: the reference's equal-length rule (options_class.f03:103-106) unless a partition is handed in
        # (balanced_partition / probe_partition: slabs of equal measured cost)
        parts = [tuple(p) for p in partition] if partition is not None else slab_partition(cfg["nz"], G)
        if len(parts) != G or parts[0][0] != 0 or sum(n for _, n in parts) != cfg["nz"] or any(a + n != b for (a, n), (b, _) in zip(parts[:-1], parts[1:])):
            raise ValueError(f"partition {parts} does not tile the {cfg['nz']} slices with {G} contiguous slabs")
        self.parts = parts
        beams = split_beam(*beam, cfg["nz"], dxi, G, parts=parts, spin=beam_spin[0] if beam_spin is not None else None)
        nsm = torch.cuda.get_device_properties(device).multi_processor_count
        self.transport = (transport or os.environ.get("QPG_PIPELINE_TRANSPORT", "p2p")) if world > 1 else None
        if self.transport not in (None, "p2p", "nccl"):
            raise ValueError(f"unknown pipeline transport {self.transport!r}")
        self.p2p = self.transport == "p2p"
        free = int(os.environ.get("QPG_PIPELINE_FREE_SMS", "4")) if self.transport == "nccl" else 0   # for the NCCL kernels beside the sweeps
        self.pgc = bool(cfg.get("laser"))
        if self.pgc:
            if laser is None:
                raise ValueError("a cfg with a laser block needs the launched envelope (laser=(a_r, a_i))")
            if self.transport == "nccl":
                raise ValueError("the envelope hand-off between GPUs uses the peer-memory transport (p2p)")
            free += S                                 # one SM per stage for its envelope solve (one CTA that cannot share an SM with a sweep CTA)
        self.neu = cfg.get("neutral")
        if self.neu and self.transport == "nccl":
            raise ValueError("the neutral species' hand-off between GPUs uses the peer-memory transport (p2p)")
        self.streams = [torch.cuda.Stream(device=device) for _ in range(S)]
        self.comm = torch.cuda.Stream(device=device) if self.transport == "nccl" else None
        self.sims = []
        dev = torch.device("cuda", device)
        for r in range(S):
            noff2, nzp = parts[self.base + r]
            mine = beams[self.base + r]
            sim = _make_sim(cfg, len(plasma[4]), len(mine[2]), self.streams[r], device, 1, noff2, nzp, beam_cap=len(beam[2]) + 1024)
            sim.init_species(*plasma)
            if beam_spin is not None:
                sim.beam.enable_spin(beam_spin[1])
            sim.beam.upload(*mine[:3])
            if beam_spin is not None:
                sim.beam.upload_spin(mine[3])
            if self.pgc:
                sim.laser.upload_slab(laser[0], laser[1], noff2)
            if self.neu:
                nu = self.neu
                sim.attach_neutral(nu["element"], nu["ion_max"], (cfg["ppc1"], cfg["ppc2"]), cfg["num_theta"], nu.get("q", -1.0), nu.get("m", 1.0), nu.get("density", 1.0),
                                   cfg.get("n0", 1.0e17))
                # per-slice launch path: from ~6 slabs per host thread on, the host time of launching graphs with a WHILE node paces the wave
                # (measured on C5: 8 stages 3.8e8 updates/s with it, 4.5e8 without; 4 stages 3.6e8 / 3.3e8)
                sim.set_graph_unroll(1 if (graph_unroll if graph_unroll is not None else S >= 6) else 0)
            if G > 1:
                if S > 1 or world > 1:
                    sim.set_sweep_ctas((nsm - free) // S)
                    if self.pgc:
                        sim.set_laser_overlap(1)      # the SMs counted in `free` above
                sim.beam.set_wire_cap(beam_wire_cap if beam_wire_cap is not None else min(len(beam[2]) + 1024, max(16384, len(beam[2]) // 64)))
            self.sims.append(sim)
        s0 = self.sims[0]
        nq, ncu, nbs = s0.field("beam_q").wire_count(), s0.field("cu").wire_count(), s0.field("b_spe").wire_count()
        self.off_fwd = (0, nq, nq + ncu, nq + ncu + nbs)
        self.n_fwd = nq + ncu + nbs + min(s0.species.wire_count(), 1 + 8 * len(plasma[4]))   # live prefix of the plasma record
        nb = s0.field("b").wire_count()
        mk = lambda n: torch.zeros(n, dtype=torch.float64, device=dev)
        nfw, nbk, nbm = nq + ncu + nbs + s0.species.wire_count(), nb + s0.field("e").wire_count(), s0.beam.wire_count()
        self.fwd = [mk(nfw) for _ in range(S)]        # written by stage r, read by the next stage
        self.back = [mk(nbk) for _ in range(S)]       # written by stage r, read by the previous stage
        self.beamb = [mk(nbm) for _ in range(S)]      # written by stage r, read by the next stage
        self.neub = [mk(s0.neutral_wire_count()) if r < S - 1 else None for r in range(S)] if self.neu else None   # neut%psend record of stage r
        self.links = None
        n_las = s0.laser.guard_size() if self.pgc else 0
        if self.p2p:
            self.links = PeerLinks(dist, rank, world, nfw, nbk, nbm, n_las, s0.neutral_wire_count() if self.neu else 0)
            self.fwd_in, self.beam_in, self.back_in = (self.links.own.get(k) for k in ("fwd_in", "beam_in", "back_in"))
        else:
            self.fwd_in = mk(nfw) if rank > 0 else None           # from the last stage of rank-1
            self.beam_in = mk(nbm) if rank > 0 else None
            self.back_in = mk(nbk) if rank < world - 1 else None  # from the first stage of rank+1
        self.off_back = (0, nb)
        self.ev, self.pending = {}, {}
        self.w = 0
        self._ev_on = bool(os.environ.get("QPG_TRACE_EVENTS"))
        self._marks = [[] for _ in range(S)]
        self.lflags = capi.WireBuf(32 * S)            # [r]: number of the last back message stage r's sweep kernel has published
        self.nback_out, self.nback_in = [0] * S, [0] * S
        self._zeroed = [False] * S
        self._deposited = [False] * S                 # the stage's tail has already scattered the next step's beam charge (split push / deposit)
        # split beam push / deposit around the wait for the downstream stage: shortens the beam link of a multi-GPU
        # pipeline; QPG_PIPELINE_SPLIT_BEAM=0 pushes and deposits in one piece behind the message (A/B)
        self.split = bool(int(os.environ.get("QPG_PIPELINE_SPLIT_BEAM", "1")))
        if self.pgc:
            # envelope links (sim_lasers_class.f03:216-218): stage r's advance waits for the new last two slices of stage r-1; buffer and ready
            # word at the consumer, ack word at the producer; flag words [r] = ready of the link INTO stage r, [S + r] = ack of the link OUT of r
            self.lasbuf = [mk(n_las) if r > 0 else None for r in range(S)]
            self.lasflags = capi.WireBuf(32 * 2 * S)
            fl = lambda k: self.lasflags.ptr + 32 * k
            for r, sim in enumerate(self.sims):
                link_in = link_out = (None, None, None)
                if r > 0:
                    link_in = (self.lasbuf[r].data_ptr(), fl(r), fl(S + r - 1))
                elif rank > 0:
                    link_in = (self.links.own["las_in"].ptr, self.links.flag("own", "ready_las"), self.links.flag("up", "ack_las"))
                if r < S - 1:
                    link_out = (self.lasbuf[r + 1].data_ptr(), fl(r + 1), fl(S + r))
                elif rank < world - 1:
                    link_out = (self.links.down["las_in"].ptr, self.links.flag("down", "ready_las"), self.links.flag("own", "ack_las"))
                sim.laser.set_handoff(*link_in, *link_out)

    # events: recorded on the producer's stream, waited on by the consumer's stream; host order = a valid schedule
    def _rec(self, name, r):
        e = self.ev.get((name, r))
        if e is None:
            e = self.ev[(name, r)] = self.torch.cuda.Event()
        e.record(self.streams[self._cur])
        return e

    def _wait(self, name, r):
        e = self.ev.get((name, r))
        if e is not None:
            self.streams[self._cur].wait_event(e)

    # NCCL links to the neighbouring ranks (torch.distributed orders an op after the CURRENT torch stream)
    def _nccl_wait(self, name, r):
        w = self.pending.pop(name, None)
        if w is not None:
            with self.torch.cuda.stream(self.streams[r]):
                w.wait()

    def _nccl_isend(self, name, t, dst, after):
        self.comm.wait_event(after)
        with self.torch.cuda.stream(self.comm):
            self.pending[name] = self.dist.isend(t, dst)

    # optional device-time trace (QPG_TRACE_EVENTS=1): CUDA events on the stage's stream at named marks of its wave
    def _mark(self, r, name):
        if not self._ev_on:
            return
        ev = self.torch.cuda.Event(enable_timing=True)
        ev.record(self.streams[r])
        self._marks[r].append((name, ev))

    def event_report(self):
        """per stage: average device time (ms) between consecutive marks, in order of first appearance"""
        self.sync()
        out = []
        for marks in self._marks:
            acc, cnt = {}, {}
            for (n0, e0), (n1, e1) in zip(marks[:-1], marks[1:]):
                k = f"{n0}>{n1}"
                acc[k] = acc.get(k, 0.0) + e0.elapsed_time(e1)
                cnt[k] = cnt.get(k, 0) + 1
            out.append({k: round(acc[k] / cnt[k], 3) for k in acc})
        return out

    def trace_reset(self):
        self._marks = [[] for _ in range(self.S)]

    # peer-memory links: stream-ordered flag waits / writes on the stage's stream
    def _pwait(self, r, name, n):
        if n > 0:
            capi.stream_wait(self.streams[r].cuda_stream, self.links.flag("own", name), n)

    def _psignal(self, r, where, name, n):
        capi.stream_signal(self.streams[r].cuda_stream, self.links.flag(where, name), n)

    def _head(self, r, upload=None):
        s, S = self.sims[r], self.S
        self._cur = r
        remote_up = r == 0 and self.rank > 0
        remote_down = r == S - 1 and self.rank < self.world - 1
        p2p_up, p2p_down = remote_up and self.p2p, remote_down and self.p2p
        if upload is not None:
            s.species.upload(*upload)                                   # the host re-injects the plasma (species%renew)
        self._mark(r, "head")
        # everything that does not need the upstream hand-off first: the stage does it while it waits
        if not self._zeroed[r]:
            s.beam_qdp_begin()
            s.begin_step_zero()
        self._zeroed[r] = False
        if not self._deposited[r]:                                      # (the first step of a stage: no tail has run before)
            s.beam_qdp_raw()
        self._deposited[r] = False
        src = None
        self._mark(r, "w_fwd")
        if p2p_up:
            n_in = self.links.next("fwd_in")
            self._pwait(r, "ready_fwd", n_in)
            src = self.fwd_in
        elif remote_up:
            with self.torch.cuda.stream(self.streams[r]):
                self.dist.recv(self.fwd_in[:self.n_fwd], self.rank - 1)
            src = self.fwd_in
        elif r > 0:
            self._wait("fwd_ready", r - 1)
            src = self.fwd[r - 1]
        self._mark(r, "got_fwd")
        if src is not None:
            fin = lambda k: src.data_ptr() + 8 * self.off_fwd[k]
            s.field("beam_q").unpack(1, fin(0), add=True)
        s.beam_qdp_fix()
        self._mark(r, "qdp")
        s.begin_step_add()
        if src is not None:
            s.species.unpack(fin(3))
            s.field("cu").unpack(0, fin(1))
            s.field("b_spe").unpack(0, fin(2))
            if self.neu:                                                # neut%precv (after the renewal of the stage's tail)
                s.neutral_unpack(self.links.own["neu_in"].ptr if p2p_up else self.neub[r - 1].data_ptr())
            if p2p_up:
                self._psignal(r, "up", "ack_fwd", n_in)
            elif not remote_up:
                self._rec("fwd_free", r - 1)
        self._mark(r, "begun")
        ev_back = None
        if src is not None and not (remote_up and not self.p2p):
            # b and e of the first slice go back to the upstream stage FROM INSIDE the sweep kernel (it writes the wire
            # buffer, possibly in the upstream GPU's memory, and raises the flag the upstream stream waits on): one launch
            # per slab, the guard slice leaves one slice into the sweep
            if p2p_up:
                n_b = self.links.next("back_out")
                self._pwait(r, "ack_back", n_b - 1)
                bdst, flag = self.links.up["back_in"].ptr, self.links.flag("up", "ready_back")
            else:
                self.nback_out[r] += 1
                n_b = self.nback_out[r]
                self._wait("back_free", r)
                bdst, flag = self.back[r].data_ptr(), self.lflags.ptr + 32 * r
            s.set_back_handoff(bdst + 8 * self.off_back[0], bdst + 8 * self.off_back[1], flag, n_b)
            s.run_slices(1, s.nzp)
        elif src is not None:
            # NCCL transport: separate first-slice launch, pack kernels, isend
            s.run_slices(1, 1)
            self._nccl_wait("back", r)
            s.field("b").pack(1, self.back[r].data_ptr() + 8 * self.off_back[0])
            s.field("e").pack(1, self.back[r].data_ptr() + 8 * self.off_back[1])
            ev_back = self._rec("back_ready", r)
            if s.nzp > 1:
                s.run_slices(2, s.nzp)
        else:
            s.run_slices(1, s.nzp)
        self._mark(r, "swept")
        if remote_up and not self.p2p:
            # after the sweep is enqueued: posting NCCL operations can block the host for milliseconds
            self._nccl_isend("back", self.back[r], self.rank - 1, ev_back)
            with self.torch.cuda.stream(self.comm):
                self.pending["beam_in"] = self.dist.irecv(self.beam_in, self.rank - 1)
        if r < S - 1 or remote_down:
            fdst = self.fwd[r].data_ptr()
            if p2p_down:
                n_f = self.links.next("fwd_out")
                self._pwait(r, "ack_fwd", n_f - 1)
                fdst = self.links.down["fwd_in"].ptr                    # the downstream GPU's buffer
            elif remote_down:
                self._nccl_wait("fwd", r)
            else:
                self._wait("fwd_free", r)
            fout = lambda k: fdst + 8 * self.off_fwd[k]
            s.field("beam_q").pack(s.nzp + 1, fout(0))
            s.field("cu").pack(0, fout(1))
            s.field("b_spe").pack(0, fout(2))
            s.species.pack(fout(3))
            if self.neu:                                                # neut%psend
                s.neutral_pack(self.links.down["neu_in"].ptr if p2p_down else self.neub[r].data_ptr())
            if p2p_down:
                self._psignal(r, "down", "ready_fwd", n_f)
            else:
                self._rec("fwd_ready", r)
        self._mark(r, "packed")
        if self.pgc:
            s.laser_advance()                                           # simulation_class.f03:486 (+ the envelope hand-off, capi.Laser.set_handoff)
            self._mark(r, "advanced")

    def _tail(self, r, renew=True):
        s, S = self.sims[r], self.S
        self._cur = r
        remote_up = r == 0 and self.rank > 0
        remote_down = r == S - 1 and self.rank < self.world - 1
        p2p_up, p2p_down = remote_up and self.p2p, remote_down and self.p2p
        self._mark(r, "tail")
        if renew:
            s.renew()                                                   # needs nothing from the neighbours: before the wait
        s.beam_qdp_begin()                                              # the next step's zero fills too (the beam push reads
        s.begin_step_zero()                                             # the e / b VOLUMES, the fills clear slice images)
        self._zeroed[r] = True
        # the beam particles whose gather does not touch the guard slice nzp + 1 (all but those in the slab's last slice) are pushed
        # NOW, before the wait for the downstream stage's first slice: only the rest of the push stays on the backward link
        if self.split:
            s.beam_push_interior()
        self._mark(r, "pre")                                            # tail>pre is work, pre>got_back the wait for the downstream stage
        # The backward message (e, b of the downstream stage's first slice) feeds the beam push only: a stage without beam particles
        # does not wait for it (qpg_stream_wait_unless_empty looks at the device-side particle count when the stream gets there), so
        # the skew "every stage starts after the first slice of the next one" builds up over the beam-carrying stages only.  The
        # unpack below then copies a possibly stale record into the guard slice nobody reads; the message counters stay in step.
        if p2p_down:
            n_b = self.links.next("back_in")
            capi.stream_wait_unless_empty(self.streams[r].cuda_stream, s.beam.count_ptr(), self.links.flag("own", "ready_back"), n_b)
            bsrc = self.back_in
        elif remote_down:
            self._nccl_isend("fwd", self.fwd[r][:self.n_fwd], self.rank + 1, self.ev[("fwd_ready", r)])
            with self.torch.cuda.stream(self.streams[r]):
                self.dist.recv(self.back_in, self.rank + 1)
            bsrc = self.back_in
        elif r < S - 1:
            self.nback_in[r] += 1
            capi.stream_wait_unless_empty(self.streams[r].cuda_stream, s.beam.count_ptr(), self.lflags.ptr + 32 * (r + 1), self.nback_in[r])   # raised by stage r+1's sweep kernel
            bsrc = self.back[r + 1]
        else:
            bsrc = None
        self._mark(r, "got_back")
        if bsrc is not None:
            s.field("b").unpack(s.nzp + 1, bsrc.data_ptr() + 8 * self.off_back[0])
            s.field("e").unpack(s.nzp + 1, bsrc.data_ptr() + 8 * self.off_back[1])
            if p2p_down:
                self._psignal(r, "down", "ack_back", n_b)
            elif not remote_down:
                self._rec("back_free", r + 1)
        if self.split:
            s.beam_push_edge()
        else:
            s.beam_push()
        self._mark(r, "pushed")
        if p2p_up:
            n_m = self.links.next("beam_in")
            self._pwait(r, "ready_beam", n_m)
            s.beam.unpack(self.beam_in.data_ptr())
            self._psignal(r, "up", "ack_beam", n_m)
        elif remote_up:
            self._nccl_wait("beam_in", r)
            s.beam.unpack(self.beam_in.data_ptr())
        elif r > 0:
            self._wait("beam_ready", r - 1)
            s.beam.unpack(self.beamb[r - 1].data_ptr())
            self._rec("beam_free", r - 1)
        if self.split:
            if self.base + r > 0:
                s.beam_qdp_part(3)                                      # the arrivals' charge (before pack_forward compacts the set)
            self._deposited[r] = True                                   # the next head does not scatter the beam charge again
        if p2p_down:
            n_m = self.links.next("beam_out")
            self._pwait(r, "ack_beam", n_m - 1)
            s.beam.pack_forward(self.links.down["beam_in"].ptr)
            self._psignal(r, "down", "ready_beam", n_m)
        elif remote_down:
            self._nccl_wait("beam", r)
            s.beam.pack_forward(self.beamb[r].data_ptr())
            self._nccl_isend("beam", self.beamb[r], self.rank + 1, self._rec("beam_ready", r))
        elif r < S - 1:
            self._wait("beam_free", r)
            s.beam.pack_forward(self.beamb[r].data_ptr())
            self._rec("beam_ready", r)
        self._mark(r, "moved")

    def wave(self, upload=None):
        """global stage g: tail of step w-g-1, then head of step w-g.  Descending order: a stage's tail needs the first
        slice of the downstream stage's head of the same step, which this order has just enqueued."""
        w = self.w
        for r in reversed(range(self.S)):
            g = self.base + r
            if w - g - 1 >= 0:
                self._tail(r, renew=not (g == 0 and upload is not None))
            if w - g >= 0:
                self._head(r, upload if g == 0 else None)
        self.w += 1

    def fill(self):
        while self.w < self.G - 1:
            self.wave()

    def drain(self):
        """finish the steps in flight (every stage ends after the same 3D step)"""
        last = self.w - 1            # newest step the first stage has started
        for w in range(self.w, self.w + self.G):
            for r in reversed(range(self.S)):
                g = self.base + r
                n_tail, n_head = w - g - 1, w - g
                if 0 <= n_tail <= last:
                    self._tail(r)
                if 0 <= n_head <= last:
                    self._head(r)
        self.w += self.G
        for k in list(self.pending):
            self.pending.pop(k).wait()
        self.sync()

    def restart(self):
        """after drain(): start a new run of the pipeline (fill again) on the state the stages hold -- every stage has finished the
        same 3D step, nothing is in flight; the message counters of the links keep counting"""
        if self.pending:
            raise RuntimeError("restart() needs a drained pipeline")
        self.w = 0

    def sync(self):
        if self.pgc:
            for s in self.sims:
                s.laser.sync()                        # overlapped envelope advances run on side streams
        for st in self.streams:
            st.synchronize()
        if self.comm is not None:
            self.comm.synchronize()

    def stats(self):
        tot = [0, 0, 0]
        for s in self.sims:
            for k, v in enumerate(s.stats()):
                tot[k] += v
        return tuple(tot)

    def launch_count(self):
        return sum(s.ctx.launch_count() for s in self.sims)

    def close(self):
        self.sync()
        if self.links is not None:
            self.links.close()
        self.lflags.close()
        for s in self.sims:
            s.close()
        if self.pgc:
            self.lasflags.close()


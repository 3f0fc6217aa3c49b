#!/usr/bin/env python
"""bench.py -- plasma particle-slice updates/s of the quasi-static slice loop (BASELINE.json metric).

One "step" = one 3D step of the deck: the sweep over all xi slices of the slab(s) (per slice: qdeposit, the field
solves, n_it x amjdeposit + B-perp predictor-corrector, push, bound check) plus the once-per-step beam deposit / push.
Workload at N=1: C2 = blowout deck scaled to nr=1024, nz=2048, max_mode=1, plasma ppc [4,4] x 16 theta
(262 144 particles per slice, 5.37e8 particle-slice updates per step).  N>1: the same deck cut into N xi slabs, one
rank per GPU, hand-offs by NCCL send/recv (strong scaling of the 3D-step pipeline).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU restatement of the reference algorithm (oracle/, the
reference itself cannot be built in this image) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

METRIC = "plasma particle-slice updates/s"
UNIT = "updates/s"
DATA = "synthetic (lattice plasma per fdist2d rule; PCG64(10) tri-Gaussian beam thinned to one particle per radial cell width and slice, charge-scaled to deposit the deck's density)"


def workload_string(name, cfg, npp0):
    """identical in the B200 arm and the reference arm (the driver compares them)"""
    return f"{name}: nr={cfg['nr']} nz={cfg['nz']} max_mode={cfg['max_mode']} Np/slice={npp0} iter_max={cfg['iter_max']}"


def deck_config(name):
    from qpad_b200 import decks
    cfg = dict(decks.CONFIGS[name])
    beam = cfg.pop("beam", None)            # one beam block, a list of them (C3), or none (C4)
    beam = [dict(b) for b in beam] if isinstance(beam, (list, tuple)) else (dict(beam) if beam is not None else None)
    cfg.pop("nstep3d", None)
    return cfg, beam


LEGACY_BEAM = bool(int(os.environ.get("QPAD_BENCH_LEGACY_BEAM", "0")))


def make_inputs(cfg, beam, beam_lattice=(256, 512), xi_cells=None):
    """Synthetic inputs of the deck's shape: lattice plasma (deterministic) + tri-Gaussian beam (PCG64 seed 10).
    The reference initialises a beam on the simulation grid (ppc per cell); at C2 that would be ~7e7 particles, so on grids finer
    than the C1-class 256 x 512 lattice the beam is thinned to ONE particle per radial cell width and per slice -- radial lattice of
    nr / ppc_r cells (particle spacing = dr), the grid's own xi cells with one layer each -- and the charge of a particle is scaled
    by (nr / lattice cells)^2 (its weight is r / dr_lattice and the radial spacing grows with the lattice cell, both linear), so that
    the DEPOSITED beam density is the deck's in every node of every slice.  C2: 1.8e7 beam particles.
    xi_cells = (k0, k1): only the beam of those xi cells (the CPU arm's bounded sample).
    [Until the end of round 1 the lattice was capped at 256 x 512 WITHOUT the charge scaling: at C2 the beam density was 64x
    below the deck's on average and present in every 4th slice only -- a weakly driven wake (1.1 predictor-corrector iterations
    per slice instead of ~4).  QPAD_BENCH_LEGACY_BEAM=1 reproduces those inputs.]"""
    from qpad_b200 import decks
    pl = decks.plasma_uniform(cfg["nr"], cfg["rmax"], cfg["ppc1"], cfg["ppc2"], cfg["num_theta"])
    if beam is None:                      # laser-driven deck (C4): no beam particles
        return pl, (np.zeros((0, 3)), np.zeros((0, 3)), np.zeros(0))
    blocks = beam if isinstance(beam, list) else [beam]
    if LEGACY_BEAM:
        bnr, bnz, qscale = min(cfg["nr"], beam_lattice[0]), min(cfg["nz"], beam_lattice[1]), 1.0
    else:
        ppc_r = max(int(b["ppc"][0]) for b in blocks)
        bnr = cfg["nr"] if cfg["nr"] <= beam_lattice[0] else max(beam_lattice[0], -(-cfg["nr"] // ppc_r))
        bnz, qscale = cfg["nz"], (cfg["nr"] / bnr) ** 2
        if cfg["nz"] > beam_lattice[1]:   # keep the particle count: one layer per slice instead of ppc3 layers per (coarser) lattice cell
            blocks = [dict(b, ppc=(b["ppc"][0], b["ppc"][1], 1)) for b in blocks]
    if xi_cells is not None:
        xi_cells = (xi_cells[0] * bnz // cfg["nz"], -(-xi_cells[1] * bnz // cfg["nz"]))
    parts = [decks.beam_std(bnr, bnz, cfg["rmax"], cfg["zmin"], cfg["zmax"], **dict(b, seed=10 + k, xi_cells=xi_cells)) for k, b in enumerate(blocks)]
    bm = tuple(np.concatenate([p[a] for p in parts]) for a in range(3))      # beams of equal q/m share one particle set
    return pl, (bm[0], bm[1], bm[2] * qscale)


class ClockSampler:
    QUERY = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for s in self.samples:
            f = [t.strip() for t in s.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle restatement of the reference algorithm
# ------------------------------------------------------------------------------------------------
def cpu_sample(cfg, plasma, beam_arrays, nslices, fast=True, nstages=1):
    from oracle import oracle as O
    kw = {k: cfg[k] for k in ("nr", "nz", "max_mode", "rmax", "zmin", "zmax", "dt", "iter_max", "iter_reltol", "iter_abstol", "ppc1", "ppc2", "num_theta")}
    las = cfg.get("laser")
    if las:                               # C4: robust_pgc plasma + one laser envelope (launched on the host, decks.laser_gaussian)
        from qpad_b200 import decks
        kw.update(sp_push_type=5, laser_on=1, laser_iter=las["iteration"], laser_k0=las["k0"], beam_evol=0)
    neu = cfg.get("neutral")
    if neu:                               # C5: nspecies 0, one ADK neutral species (input_file/ionization)
        kw.update(sp_density=0.0, neut_on=1, neut_elem=neu["element"], neut_ion_max=neu["ion_max"], neut_ppc1=cfg["ppc1"], neut_ppc2=cfg["ppc2"],
                  neut_num_theta=cfg["num_theta"], neut_density=neu.get("density", 1.0), n0=cfg.get("n0", 1.0e17))
    sim = O.Sim(fast=fast, nstages=nstages, **kw)
    if las:
        sim.set_laser(*decks.laser_gaussian(cfg["nr"], cfg["nz"], cfg["rmax"], cfg["zmin"], cfg["zmax"], **las))
    sim.set_beam(*beam_arrays)
    t0 = time.perf_counter()
    upd = sim.run_slices(nslices)
    dt = time.perf_counter() - t0
    return upd, dt, sim.total_iters()


_cpu_barrier = None
LAST_CPU_NIT = None
LAST_CPU_INFO = {}


def _cpu_init(barrier):
    global _cpu_barrier
    _cpu_barrier = barrier


def _cpu_worker(args):
    """one host core = one stage of the reference's `mpirun -np k` xi-pipeline in steady state (parallel_module.f03:221-239): stage i
    owns slab i of the deck (equal-length slabs, options_class.f03:103-106) and all stages sweep their slabs AT THE SAME TIME, each
    on its own 3D step.  A stage's slab needs the plasma state the upstream stages hand down; here every worker produces it itself
    in an UNTIMED preparation sweep over the slices ahead of its slab (the quasi-static recurrence has no shortcut), snapshots it,
    and then times `reps` sweeps of its slab from that state -- together the k timed slabs are exactly one 3D step of the deck,
    with the deck's own predictor-corrector iteration counts."""
    name, j0, n, reps, nwarm = args
    cfg, beam = deck_config(name)
    plasma, bm = make_inputs(cfg, beam, xi_cells=(0, j0 + n + 2))       # the beam charge the slices up to the slab's end see
    from oracle import oracle as O
    kw = {k: cfg[k] for k in ("nr", "nz", "max_mode", "rmax", "zmin", "zmax", "dt", "iter_max", "iter_reltol", "iter_abstol", "ppc1", "ppc2", "num_theta")}
    las = cfg.get("laser")
    if las:
        from qpad_b200 import decks
        kw.update(sp_push_type=5, laser_on=1, laser_iter=las["iteration"], laser_k0=las["k0"], beam_evol=0)
    neu = cfg.get("neutral")
    if neu:
        kw.update(sp_density=0.0, neut_on=1, neut_elem=neu["element"], neut_ion_max=neu["ion_max"], neut_ppc1=cfg["ppc1"], neut_ppc2=cfg["ppc2"],
                  neut_num_theta=cfg["num_theta"], neut_density=neu.get("density", 1.0), n0=cfg.get("n0", 1.0e17))
    sim = O.Sim(fast=True, **kw)
    if las:
        sim.set_laser(*decks.laser_gaussian(cfg["nr"], cfg["nz"], cfg["rmax"], cfg["zmin"], cfg["zmax"], **las))
    sim.set_beam(*bm)
    tp = time.time()
    sim.run_slices(j0)                                   # untimed: the state at the slab's first slice (stage_begin + slices 1..j0)
    can_snap = not (las or neu)
    if can_snap:
        sim.snapshot()
    prep = time.time() - tp
    out = []
    for r in range(nwarm + reps):
        if r > 0:
            if not can_snap:
                break
            sim.restore()
        if _cpu_barrier is not None:
            _cpu_barrier.wait()                          # all stages start their slab together, every repetition
        i0 = sim.total_iters()
        t0 = time.time()
        upd = sim.run_range(j0 + 1, j0 + n)
        t1 = time.time()
        if r >= nwarm or not can_snap:
            out.append((upd, t0, t1, sim.total_iters() - i0))
    return out, prep, n


def cpu_parallel(name, nslices=None, ncores=None, reps=1, nwarm=0):
    """All host cores as the k stages of the reference's xi-pipeline in steady state over ONE 3D step of the deck (see _cpu_worker).
    Returns (updates of the timed steps, wall seconds = sum over steps of the slowest stage's slab time, cores, per-step list).
    `nslices` caps the slab length (a bounded sample for very large k * slab products); None = the whole deck."""
    import multiprocessing as mp
    from oracle import oracle as O
    from qpad_b200.pipeline import slab_partition
    O.build(fast=True, force=True)                      # -O3 -march=native of THIS box, once, before the workers start
    cfg, _ = deck_config(name)
    k = min(ncores or os.cpu_count() or 1, cfg["nz"] // 2)
    parts = slab_partition(cfg["nz"], k)
    if nslices:
        parts = [(a, min(n, nslices)) for a, n in parts]
    if cfg.get("laser") or cfg.get("neutral"):
        reps, nwarm = 1, 0
    ctx = mp.get_context("spawn")
    with ctx.Pool(k, initializer=_cpu_init, initargs=(ctx.Barrier(k),)) as pool:
        res = pool.map(_cpu_worker, [(name, a, n, reps, nwarm) for a, n in parts], chunksize=1)
    global LAST_CPU_NIT, LAST_CPU_INFO
    steps = []
    for r in range(len(res[0][0])):
        upd = sum(w[0][r][0] for w in res)
        wall = max(w[0][r][2] for w in res) - min(w[0][r][1] for w in res)
        steps.append((upd, wall))
    iters = sum(w[0][0][3] for w in res)
    nsl = sum(w[2] for w in res)
    LAST_CPU_NIT = iters / max(nsl, 1)
    slab_s = [w[0][-1][2] - w[0][-1][1] for w in res]
    LAST_CPU_INFO = {"slices_timed_per_step": nsl, "stages": k, "prep_s_max": max(w[1] for w in res), "slab_s_min": min(slab_s), "slab_s_max": max(slab_s)}
    return sum(u for u, _ in steps), sum(t for _, t in steps), k, steps


def cpu_sample_text(name, k):
    i = LAST_CPU_INFO
    return (f"the reference's xi-pipeline in steady state on {k} host cores: {k} concurrent single-threaded stage processes, stage i sweeps slab i of the "
            f"{name} deck (equal-length slabs, options_class.f03:103-106) -- together {i['slices_timed_per_step']} slices = "
            f"{'one whole 3D step' if i['slices_timed_per_step'] >= deck_config(name)[0]['nz'] else 'a bounded sample of the 3D step'} per timed step, "
            f"{LAST_CPU_NIT:.3f} predictor-corrector iterations per slice; a step costs the slowest slab ({i['slab_s_max']:.1f} s, fastest {i['slab_s_min']:.1f} s); "
            f"each stage first produced the plasma state at its slab's first slice in an untimed sweep (up to {i['prep_s_max']:.0f} s); "
            "oracle restatement of the reference algorithm (-O3 -march=native; the Fortran/MPI/HYPRE reference cannot be built here)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg, beam = deck_config(args.config)
    plasma, bm = make_inputs(cfg, beam) if args.config not in ("C4", "C5") else (make_inputs(cfg, None)[0], None)
    upd, dt, k, steps = cpu_parallel(args.config, args.ref_slices or None, reps=args.steps, nwarm=min(args.warmup, 1))
    nst = len(steps)
    value = upd / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": nst, "warmup": min(args.warmup, 1),
            "ms_per_step": 1e3 * dt / nst, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": DATA,
            "config": {"workload": workload_string(args.config, cfg, len(plasma[4])), "parallelism": f"cpu: {k} stage processes (one per host core)",
                       "pc_iters_per_slice": LAST_CPU_NIT},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": k, "kind": "port", "sample": cpu_sample_text(args.config, k), "pc_iters_per_slice": LAST_CPU_NIT},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from qpad_b200 import capi
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        # one NCCL kernel can be resident beside the cooperative sweep kernel (the receive of the crossing beam particles
        # spins until the upstream stage has pushed its beam): its CTAs (one per channel) must fit in the SMs
        # PipelineStage leaves free, otherwise the sweep launch waits for the message (measured: 3-4 ms per step)
        os.environ.setdefault("NCCL_MAX_P2P_NCHANNELS", "4")
        dist.init_process_group("nccl")     # lazy init: every stage pair gets its own p2p communicator / stream
    cfg, beam = deck_config(args.config)
    plasma, bm = make_inputs(cfg, beam)
    npp0 = len(plasma[4])
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        if world == 1:
            from qpad_b200.pipeline import SingleStage as Runner
        else:
            from qpad_b200.pipeline import PipelineStage as Runner
        runner = Runner(cfg, plasma, bm, stream=stream, rank=rank, world=world, device=local, use_graph=1)
        sim = runner.sim

        def sync_all():
            if world > 1:
                dist.barrier(device_ids=[local])
            torch.cuda.synchronize()
            if os.environ.get("QPG_TRACE"):
                print(f"rank {rank}: sync_all done", file=sys.stderr, flush=True)

        if args.no_sweep:
            sim.set_sweep(0)
        for _ in range(args.warmup):
            runner.step()
        if world > 1:
            runner.prime()      # fill the pipeline: stage r runs world-1-r steps ahead (untimed), see PipelineStage.prime
        step = runner.step if world == 1 else runner.step_primed
        sync_all()
        u0, i0, s0 = sim.stats()
        l0 = sim.ctx.launch_count()
        sweep_on = not args.no_sweep
        if sweep_on:
            sim.sweep_profile(reset=True)
            sim.ctx.tprof_reset(); sim.ctx.tprof_enable(True)      # CUDA events around every sweep-kernel launch
        clk = ClockSampler(local); clk.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        ev0.record(stream)
        for _ in range(args.steps):
            step()
        ev1.record(stream)
        sync_all()
        ms = ev0.elapsed_time(ev1)
        clocks = clk.stop()
        u1, i1, s1 = sim.stats()
        upd, iters, slices = u1 - u0, i1 - i0, s1 - s0
        prof, sweep_ms, sweep_n = None, None, 0
        if sweep_on:
            sweep_ms, sweep_n = sim.ctx.tprof_get("kernel sweep")
            sim.ctx.tprof_enable(False)
            prof = sim.sweep_profile()
        launches = sim.ctx.launch_count() - l0 if (sweep_on or not sim.prm.use_graph) else slices * 5 + 2 * iters + args.steps * 12
        t = torch.tensor([ms, float(upd), float(launches), float(iters), float(slices)], dtype=torch.float64, device="cuda")
        if world > 1:
            tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            ms, upd, launches, iters, slices = tmax[0].item(), tsum[1].item(), tsum[2].item(), tsum[3].item(), tsum[4].item()
        value = upd / (ms * 1e-3)

        # ---- end-to-end leg: host buffers in, host line-outs out, every step (single GPU public API) -----------
        e2e = None
        if world == 1:
            sync_all()
            ue0 = sim.stats()[0]
            t0 = time.perf_counter()
            h2d = d2h = 0
            for _ in range(args.steps):
                hb, db = runner.step_e2e()
                h2d, d2h = hb, db
            torch.cuda.synchronize()
            te = time.perf_counter() - t0
            ue = sim.stats()[0] - ue0
            e2e = {"value": ue / te, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                   "what": "per step: plasma lattice host->device through qpg_part2d_upload, full 3D step, E_z and psi on-axis line-outs + counters device->host"}

        # ---- roofline ------------------------------------------------------------------------------------------
        roof = roof_hbm = None
        if rank == 0 and sweep_on:
            # dominant kernel = the persistent sweep kernel: CUDA events around each of its launches in the timed region.
            # algorithmic bytes (fused push): 24 (qdeposit) + 64 per amjdeposit pass + 88 (push_u+push_x) per update
            peak, peak_src = hbm_peak()
            u_r, i_r, s_r = (u1 - u0), (i1 - i0), (s1 - s0)
            nit = i_r / max(s_r, 1)
            bytes_total = u_r * (112.0 + 64.0 * nit)
            ach = bytes_total / (sweep_ms * 1e-3) / 1e9
            nspc = prof["ns_total"] / max(prof["cyc_total"], 1.0)
            ph = {}
            for key, cnt, bpp in (("A", prof["slices"], None), ("amj", prof["amj_phases"], 64.0), ("C", prof["amj_phases"], None), ("push", prof["slices"], 112.0)):
                us = prof["cyc_" + key] * nspc * 1e-3 / max(cnt, 1.0)
                ph[key] = {"us_per_phase": us, "cta0_work_us": prof["work_" + key] * nspc * 1e-3 / max(cnt, 1.0)}
                if bpp:
                    ph[key]["GBs"] = bpp * (u_r / max(s_r, 1)) / (us * 1e-6) / 1e9
            traffic, traffic_src = None, None
            tpath = os.path.join(ROOT, "profiles", "r01_sweep_traffic.json")
            if world == 1 and args.config == "C2" and os.path.exists(tpath):   # ncu capture of this very launch shape (committed)
                tj = json.load(open(tpath))
                traffic = float(tj["dram_bytes_read"] + tj["dram_bytes_write"])
                traffic_src = "profiles/r01_sweep_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum of one 2048-slice launch; the particle planes stay in L2, the DRAM writes are the field-volume slice stores)"
            roof = {"bound": "hbm", "kernel": "k_sweep<%d> (persistent: all slices of a slab in one launch)" % cfg["max_mode"], "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": bytes_total / max(sweep_n, 1), "peak_source": peak_src,
                    "bytes_per_update": 112.0 + 64.0 * nit, "updates_per_launch": u_r / max(sweep_n, 1), "avg_launch_ms": sweep_ms / max(sweep_n, 1),
                    "launches_timed": int(sweep_n), "us_per_slice": sweep_ms * 1e3 / max(s_r, 1),
                    "phases": {"A||update_bound": ph["A"], "amjdeposit (64 B/particle)": ph["amj"], "C": ph["C"], "push_u+push_x+qdeposit||D (112 B/particle)": ph["push"],
                               "note": "SM-clock stamps of CTA 0 between grid-barrier releases, scaled by globaltimer; each phase includes its barrier"},
                    "note": "particle planes (16.8 MB) stay L2-resident between phases at this size; the kernel is bound by fp64 issue + barrier latency, not HBM (DESIGN.md 4)"}
            if world == 1 and not args.no_micro:
                runner.prepare_step()
                sim.run_slices(1, max(1, int(0.55 * cfg["nz"])))   # fields of a slice inside the wake
                roof_hbm = runner.kernel_microbench(peak)
                runner.finish_step()
        elif rank == 0 and world == 1:
            kern = {}
            peak, peak_src = hbm_peak()
            j0 = max(1, int(0.55 * cfg["nz"]))
            j1 = min(cfg["nz"], j0 + args.roof_slices - 1)
            runner.prepare_step()
            sim.run_slices(1, j0 - 1)          # get into the wake (clustered particles), graph mode
            sim.set_graph(0)
            sim.ctx.tprof_reset(); sim.ctx.tprof_enable(True)
            n_before = sim.species.npp()
            sim.run_slices(j0, j1)
            for ev in ("kernel amjdeposit", "kernel qdeposit", "kernel push", "kernel compact", "fused field program"):
                kern[ev] = sim.ctx.tprof_get(ev)
            sim.ctx.tprof_enable(False)
            sim.set_graph(1)
            ms_amj, n_amj = kern["kernel amjdeposit"]
            per = ms_amj / max(n_amj, 1)
            ach = 64.0 * n_before / (per * 1e-3) / 1e9
            roof = {"bound": "hbm", "kernel": "k_amjdeposit<1>", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                    "peak_source": peak_src, "bytes_per_particle": 64, "particles_per_launch": int(n_before), "avg_launch_us": per * 1e3,
                    "note": "per-slice launch path (--no-sweep): events on the launch stream, slices %d-%d" % (j0, j1),
                    "other_kernels_us": {k: (v[0] / max(v[1], 1)) * 1e3 for k, v in kern.items()}}
            roof_hbm = runner.kernel_microbench(peak)
            runner.finish_step()

        # ---- CPU baseline beside it (rank 0, bounded sample) ----------------------------------------------------
        cpu = None
        if rank == 0 and world == 1 and not args.no_cpu:
            try:
                upd_c, wall_c, k_c, _steps = cpu_parallel(args.config, args.ref_slices or None)
                cpu = {"value": upd_c / wall_c, "unit": UNIT, "cores": k_c, "kind": "port", "sample": cpu_sample_text(args.config, k_c), "pc_iters_per_slice": LAST_CPU_NIT}
            except Exception as exc:  # the GPU result must not be lost to a CPU-side problem
                cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {exc}"}

        if rank == 0:
            line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                    "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                    "data": DATA,
                    "config": {"workload": workload_string(args.config, cfg, npp0),
                               "parallelism": "single" if world == 1 else f"xi-pipeline x{world}: one xi slab per GPU, NCCL send/recv hand-offs; pipeline filled before the timed region (stage r runs {world}-1-r steps ahead), every stage then times {args.steps} steady-state steps",
                               "l2": "step working set (field volumes ~0.9 GB + beam) exceeds the 126 MB L2",
                               "pc_iters_per_slice": iters / max(slices, 1)},
                    "clocks": clocks, "gpu_launches": int(launches)}
            if e2e: line["e2e"] = e2e
            if roof: line["roofline"] = roof
            if roof_hbm: line["roofline_hbm_stream"] = roof_hbm
            if cpu: line["cpu_baseline"] = cpu
            print(json.dumps(line))
        if world > 1:
            if os.environ.get("QPG_TRACE_EVENTS"):
                print(f"rank {rank} step trace (ms): {runner.event_report()}", file=sys.stderr, flush=True)
            runner.unwind()
            sync_all()
        runner.close()
    if world > 1:
        dist.destroy_process_group()



def beam_sums(x, p, q):
    """additive moments of a beam particle set (so that shards on several ranks can be summed)"""
    v = [q.sum()]
    for k in (0, 1):
        v += [np.sum(q * x[:, k]), np.sum(q * p[:, k]), np.sum(q * x[:, k] ** 2), np.sum(q * p[:, k] ** 2), np.sum(q * x[:, k] * p[:, k])]
    v.append(np.sum(q * p[:, 2]))
    return np.array(v, dtype=np.float64)


def beam_moments_from_sums(v):
    out = {}
    for k, ax in enumerate("xy"):
        sx, sp, sxx, spp, sxp = (v[1 + 5 * k + i] / v[0] for i in range(5))
        vxx, vpp, vxp = sxx - sx * sx, spp - sp * sp, sxp - sx * sp
        out[ax] = (sx, np.sqrt(max(vxx, 0.0)), np.sqrt(max(vxx * vpp - vxp * vxp, 0.0)))
    out["pz"] = v[11] / v[0]
    return out


def parity_check(cfg, plasma, bm, lp, nsteps, device, rank, world, dist, tol=1e-6):
    """Correctness of the pipelined run carried by the bench line: after lp.drain() every stage holds the fields of 3D step
    `nsteps` - 1 on its slab; the same `nsteps` steps are re-run on ONE stage (a single sweep kernel over the whole box, the
    path the full-size oracle parity tests validate: tests/test_gpu_fullsize.py) and the E_z / psi on-axis line-outs of every slab
    and the beam centroid / rms size / emittance are compared (north star: <= 1e-6 relative).  Every rank checks its own slabs."""
    import torch
    from qpad_b200.pipeline import _make_sim
    tol = float(os.environ.get("QPG_PARITY_TOL", tol))             # (tests force the conditioning probe below with a tiny tolerance)
    st = torch.cuda.Stream(device=device)
    sim = _make_sim(cfg, len(plasma[4]), len(bm[2]), st, device, 1)
    sim.init_species(*plasma)
    sim.beam.upload(*bm)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for k in range(nsteps):
        if k == nsteps - 1:
            ev[0].record(st)
        sim.step3d()
    ev[1].record(st)
    st.synchronize()
    single_ms = ev[0].elapsed_time(ev[1])
    ez1, ps1 = sim.field("e").lineout(3, 0, 1), sim.field("psi").lineout(1, 0, 1)
    upd1, it1, sl1 = sim.stats()
    dez, dps = np.zeros(len(ez1)), np.zeros(len(ps1))              # per-slice deviations on this rank's slabs
    for (off, n), s in zip(lp.parts[lp.base:lp.base + lp.S], lp.sims):
        ez, ps = s.field("e").lineout(3, 0, 1), s.field("psi").lineout(1, 0, 1)
        dez[off:off + n] = np.abs(ez[:n] - ez1[off:off + n]) / float(np.max(np.abs(ez1)))
        dps[off:off + n] = np.abs(ps[:n] - ps1[off:off + n]) / float(np.max(np.abs(ps1)))
    err_ez, err_ps = float(dez.max()), float(dps.max())
    # A deck can be ill-conditioned in a few slices (the hosing deck where its bubble closes: rounding differences grow by 1e10 within three
    # slices, tests/test_gpu_fullsize.py): if the line-outs deviate, the conditioning of the ONE-STAGE run itself is probed -- the same steps
    # with the beam charge scaled by (1 + 1e-14) -- and a slice whose own response to that perturbation exceeds 1e-9 is held to 30x its
    # response instead of `tol` (at most 2 % of the slices may be such)
    ill, fields_ok = 0, bool(err_ez < tol and err_ps < tol)
    if not fields_ok:
        sim2 = _make_sim(cfg, len(plasma[4]), len(bm[2]), st, device, 1)
        sim2.init_species(*plasma)
        sim2.beam.upload(bm[0], bm[1], bm[2] * (1.0 + 1e-14))
        for k in range(nsteps):
            sim2.step3d()
        st.synchronize()
        rez = np.abs(sim2.field("e").lineout(3, 0, 1) - ez1) / float(np.max(np.abs(ez1)))
        rps = np.abs(sim2.field("psi").lineout(1, 0, 1) - ps1) / float(np.max(np.abs(ps1)))
        sim2.close()
        bad = (rez > 1e-9) | (rps > 1e-9)
        ill = int(bad.sum())
        fields_ok = bool(np.all(dez < np.where(bad, np.maximum(tol, 30.0 * rez), tol)) and np.all(dps < np.where(bad, np.maximum(tol, 30.0 * rps), tol))
                         and ill <= max(8, len(ez1) // 50))
        err_ez_well = float(dez[~bad].max()) if (~bad).any() else 0.0
    else:
        err_ez_well = err_ez
    sums = np.zeros(12)
    nb = 0
    for s in lp.sims:
        bx, bp, bq = s.beam.download()
        nb += len(bq)
        if len(bq):
            sums += beam_sums(bx, bp, bq)
    t = torch.tensor(list(sums) + [float(nb)], dtype=torch.float64, device="cuda")
    e = torch.tensor([err_ez, err_ps, 0.0 if fields_ok else 1.0, float(ill), err_ez_well], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dist.all_reduce(e, op=dist.ReduceOp.MAX)
    sums, nb = t[:12].cpu().numpy(), int(t[12].item())
    err_ez, err_ps, fields_ok, ill, err_ez_well = float(e[0].item()), float(e[1].item()), e[2].item() == 0.0, int(e[3].item()), float(e[4].item())
    bx, bp, bq = sim.beam.download()
    m1, mp_ = beam_moments_from_sums(beam_sums(bx, bp, bq)), beam_moments_from_sums(sums)
    sim.close()
    berr = 0.0
    for ax in "xy":
        c1, s1, e1 = m1[ax]; c2, s2, e2 = mp_[ax]
        berr = max(berr, abs(c1 - c2) / s1, abs(s1 - s2) / s1, abs(e1 - e2) / e1)
    berr = max(berr, abs(m1["pz"] - mp_["pz"]) / abs(m1["pz"]))
    ok = bool(fields_ok and berr < tol and nb == len(bq))
    return {"ok": ok, "tol": tol, "steps_compared": nsteps, "ez_lineout_rel_err": err_ez, "psi_lineout_rel_err": err_ps,
            "ill_conditioned_slices": ill, "ez_lineout_rel_err_well_conditioned_slices": err_ez_well,
            "beam_centroid_size_emittance_rel_err": berr, "beam_particles": nb, "beam_particles_single_stage": int(len(bq)),
            "against": "the same 3D steps on ONE stage (one sweep kernel over the whole box) on this GPU; that path is checked against the CPU oracle at full size in tests/test_gpu_fullsize.py"}, \
        {"single_step_ms": single_ms, "what": "latency of ONE 3D step on one GPU without the SM-partitioned pipeline (--stages 1: one sweep kernel on all SMs + beam deposit / push)",
         "updates_per_s": upd1 / max(nsteps, 1) / (single_ms * 1e-3)}


def run_c4(args):
    """config 4 (input_file/lwfa): laser-driven wake, robust_pgc plasma, envelope advanced every 3D step -- one xi stage on one
    GPU through the per-slice launch path (the laser hooks are not in the persistent sweep kernel yet).  A step = one 3D
    step = nz slices + the envelope advance."""
    import torch
    from qpad_b200 import capi, decks
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference for the CPU arm)")
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        raise SystemExit("bench.py --config C4: the laser path runs on one xi stage (one GPU)")
    cfg, _ = deck_config("C4")
    las = cfg["laser"]
    plasma, _bm = make_inputs(cfg, None)
    x, p, g, psi, q = plasma
    npp0 = len(q)
    stream = torch.cuda.Stream()
    simkw = {k: cfg[k] for k in ("nr", "nz", "max_mode", "rmax", "zmin", "zmax", "dt", "iter_max", "iter_reltol", "iter_abstol")}
    sim = capi.Sim(sp_npmax=2 * npp0, beam_npmax=64, beam_evol=0, sp_push_pgc=1, laser_iter=las["iteration"], laser_k0=las["k0"], sp_ppc_r=cfg["ppc1"],
                   use_graph=0 if args.no_graph else 1, stream=stream.cuda_stream, **simkw)
    sweep_on = not args.no_sweep
    sim.set_sweep(1 if sweep_on else 0)      # default: the laser hooks inside the persistent sweep kernel (k_sweep<M, PGC = true>)
    sim.init_species(x, p, g, psi, q)
    a_r, a_i = decks.laser_gaussian(cfg["nr"], cfg["nz"], cfg["rmax"], cfg["zmin"], cfg["zmax"], **las)
    sim.laser.upload(a_r, a_i)
    for _ in range(args.warmup):
        sim.step3d()
    torch.cuda.synchronize()
    u0, i0, s0 = sim.stats()
    l0 = sim.ctx.launch_count()
    clk = ClockSampler(0); clk.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record(stream)
    for _ in range(args.steps):
        sim.step3d()
    ev1.record(stream)
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    clocks = clk.stop()
    u1, i1, s1 = sim.stats()
    upd, iters, slices = u1 - u0, i1 - i0, s1 - s0
    launches = sim.ctx.launch_count() - l0 if (args.no_graph or sweep_on) else slices * 9 + 2 * iters + 2 * args.steps   # graph replay: head 2, tail 7, 2 per PC iteration
    # end to end: plasma lattice host -> device every step, wake + envelope line-outs device -> host
    t0 = time.perf_counter()
    ue0 = sim.stats()[0]
    d2h = 0
    for _ in range(args.steps):
        sim.species.upload(x, p, g, psi, q)
        sim.beam_qdp_begin(); sim.beam_qdp_end(); sim.begin_step()
        sim.run_slices(1, sim.nzp)
        sim.laser_advance()
        ez = sim.field("e").lineout(3, 0, 1); ps = sim.field("psi").lineout(1, 0, 1)
        d2h = 8 * (len(ez) + len(ps)) + 24
        sim.stats()
    torch.cuda.synchronize()
    te = time.perf_counter() - t0
    e2e = {"value": (sim.stats()[0] - ue0) / te, "unit": UNIT, "h2d_bytes_per_step": int(64 * npp0), "d2h_bytes_per_step": int(d2h),
           "what": "per step: plasma lattice host->device (qpg_part2d_upload), nz slices + envelope advance, E_z and psi on-axis line-outs + counters device->host"}
    peak, peak_src = hbm_peak()
    nit = iters / max(slices, 1)
    # algorithmic bytes per update: qdeposit 24 + amjdeposit_pgc 72 per pass (reads psi too) + push_u_pgc 80 + push_x 64 + deposit_chi 32
    bpu = 24.0 + 72.0 * nit + 80.0 + 64.0 + 32.0
    ach = upd * bpu / (ms * 1e-3) / 1e9
    roof = {"bound": "hbm", "kernel": ("k_sweep<0, PGC> (persistent: all slices of the step in one launch; laser slice images, pgc pushers and the susceptibility deposit inside it) + one persistent envelope-solve CTA per step; 65 536 particles per slice: barrier and field-program latency bound it, not the particle bytes"
                                       if sweep_on else "per-slice launch path: k_amjdeposit_pgc / k_push_u_pgc / k_push / k_qdeposit / k_deposit_chi + field programs (65 536 particles per slice: the latency of the ~12 dependent kernels of a slice bounds it, not the particle bytes)"),
            "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None, "peak_source": peak_src, "bytes_per_update": bpu,
            "us_per_slice": ms * 1e3 / max(slices, 1)}
    cpu = None
    if not args.no_cpu:
        try:
            upd_c, wall_c, k_c, _t = cpu_parallel("C4", args.ref_slices or None)
            cpu = {"value": upd_c / wall_c, "unit": UNIT, "cores": k_c, "kind": "port", "sample": cpu_sample_text("C4", k_c), "pc_iters_per_slice": LAST_CPU_NIT}
        except Exception as exc:
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {exc}"}
    line = {"metric": METRIC, "value": upd / (ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (lattice plasma per fdist2d rule, Gaussian x sin^2 laser pulse of the lwfa deck)",
            "config": {"workload": f"C4: nr={cfg['nr']} nz={cfg['nz']} max_mode=0 Np/slice={npp0} robust_pgc, laser a0={las['a0']} k0={las['k0']} iteration {las['iteration']}",
                       "parallelism": ("single stage: one persistent sweep kernel per 3D step + one persistent envelope-solve CTA" if sweep_on else
                                       "single stage, slice body " + ("as plain stream launches" if args.no_graph else "replayed from a CUDA graph (device-side WHILE node for the predictor-corrector loop)") + " + one persistent envelope-solve CTA per step"), "l2": "field volumes + envelope volumes ~60 MB, particle planes 4 MB: L2 resident",
                       "pc_iters_per_slice": nit},
            "clocks": clocks, "gpu_launches": int(launches), "roofline": roof, "e2e": e2e}
    if cpu: line["cpu_baseline"] = cpu
    print(json.dumps(line))
    sim.close()


def run_c4_pipeline(args):
    """config 4 on the xi-pipeline (the deck is `nodes [1,4]`; sim_lasers_class.f03:197-222): --stages S sweep kernels per GPU on SM
    partitions, every stage with its slab of the envelope, advanced after its sweep from the upstream stage's new last two slices
    (capi.Laser.set_handoff; one SM per stage is left to the envelope solves, which run beside the next sweeps).  A timed step = one wave =
    every stage sweeps and advances its slab once, in steady state; after the timed region the pipeline is drained and the SAME number of
    3D steps is re-run on one stage: envelope and wake line-outs must agree (parity_check)."""
    import torch
    import torch.distributed as dist
    from qpad_b200 import capi, decks
    from qpad_b200.pipeline import LocalPipeline
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl")
    cfg, _ = deck_config("C4")
    las = cfg["laser"]
    plasma, _bm = make_inputs(cfg, None)
    npp0 = len(plasma[4])
    empty = (np.zeros((0, 3)), np.zeros((0, 3)), np.zeros(0))
    a_r, a_i = decks.laser_gaussian(cfg["nr"], cfg["nz"], cfg["rmax"], cfg["zmin"], cfg["zmax"], **las)
    S = args.stages
    mk = lambda pt: LocalPipeline(cfg, plasma, empty, S, device=local, rank=rank, world=world, dist=dist if world > 1 else None, transport="p2p" if world > 1 else None,
                                  partition=pt, laser=(a_r, a_i))
    parts, balance_log = None, []
    if getattr(args, "balance", 0):
        # slabs of equal measured cost (the pulse and the bubble behind it take more predictor-corrector passes than the quiet plasma ahead):
        # closed loop on the running pipeline over the 3D step numbers the timed region will see (pipeline.measured_partition), all untimed
        from qpad_b200.pipeline import measured_partition
        for rnd in range(max(getattr(args, "rebalance", 0), 0)):
            lp = mk(parts)
            new_parts, stage_ms, spread = measured_partition(lp, cfg, None, nwaves=args.steps, nwarm=args.warmup)
            balance_log.append({"round": rnd, "spread": round(spread, 4), "busy_ms_by_stage": [round(v, 3) for v in stage_ms], "slab_slices": [n for _, n in lp.parts]})
            lp.drain(); lp.close()
            if new_parts is None or [tuple(q) for q in new_parts] == [tuple(q) for q in lp.parts]:
                break
            parts = new_parts
    lp = mk(parts)
    main = torch.cuda.current_stream()

    def sync_all():
        lp.sync()
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    lp.fill()
    for _ in range(args.warmup):
        lp.wave()
    sync_all()
    lp.trace_reset()
    u0, i0, s0 = lp.stats()
    l0 = lp.launch_count()
    for sim in lp.sims:
        sim.sweep_profile(reset=True)
    clk = ClockSampler(local); clk.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    ev0.record(main)
    for st in lp.streams:
        st.wait_event(ev0)
    th0 = time.perf_counter()
    for _ in range(args.steps):
        lp.wave()
    host_ms = 1e3 * (time.perf_counter() - th0) / args.steps      # host time to ENQUEUE a wave
    for st in lp.streams:     # the envelope advances on the side streams are not waited for: a stage's advance of step n overlaps its sweep of n+1
        e = torch.cuda.Event(); e.record(st); main.wait_event(e)
    ev1.record(main)
    sync_all()
    ms = ev0.elapsed_time(ev1)
    clocks = clk.stop()
    if os.environ.get("QPG_TRACE_EVENTS"):
        for r_, rep in enumerate(lp.event_report()):
            waits = sum(v for k, v in rep.items() if k in ("w_fwd>got_fwd", "pre>got_back"))
            print(f"rank {rank} stage {r_} trace (ms): busy {sum(rep.values()) - waits:.2f} waits {waits:.2f} {rep}", file=sys.stderr, flush=True)
        lp._ev_on = False
    u1, i1, s1 = lp.stats()
    upd, iters, slices = u1 - u0, i1 - i0, s1 - s0
    launches = lp.launch_count() - l0
    profs = [sim.sweep_profile() for sim in lp.sims]
    ph = {}
    for key, cntk in (("A", "slices"), ("amj", "amj_phases"), ("C", "amj_phases"), ("push", "slices")):
        ph[key] = [round(p["cyc_" + key] * (p["ns_total"] / max(p["cyc_total"], 1.0)) * 1e-3 / max(p[cntk], 1.0), 2) for p in profs]
    if world > 1:
        t = torch.tensor([ms, float(upd), float(launches), float(iters), float(slices)], dtype=torch.float64, device="cuda")
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, upd, launches, iters, slices = tmax[0].item(), tsum[1].item(), tsum[2].item(), tsum[3].item(), tsum[4].item()
    # end to end (N = 1): plasma lattice host -> device into stage 0 every wave, line-outs of every slab + counters back
    e2e = None
    if world == 1:
        ue0 = lp.stats()[0]
        t0 = time.perf_counter()
        d2h = 0
        for _ in range(args.steps):
            lp.wave(upload=plasma)
            d2h = 24 * S
            for sim in lp.sims:
                ez = sim.field("e").lineout(3, 0, 1); ps = sim.field("psi").lineout(1, 0, 1)
                d2h += 8 * (len(ez) + len(ps))
            lp.stats()
        sync_all()
        te = time.perf_counter() - t0
        e2e = {"value": (lp.stats()[0] - ue0) / te, "unit": UNIT, "h2d_bytes_per_step": int(64 * npp0), "d2h_bytes_per_step": int(d2h),
               "what": "per step (wave): plasma lattice host->device (qpg_part2d_upload) into stage 0, every stage sweeps + advances its slab, E_z and psi on-axis line-outs of every slab + counters device->host"}
    # parity of the pipelined run: the same number of 3D steps on one stage
    check = None
    if args.check:
        lp.drain()
        nsteps = lp.sims[0].stats()[2] // lp.sims[0].nzp
        simkw = {k: cfg[k] for k in ("nr", "nz", "max_mode", "rmax", "zmin", "zmax", "dt", "iter_max", "iter_reltol", "iter_abstol")}
        st1 = torch.cuda.Stream(device=local)
        one = capi.Sim(sp_npmax=2 * npp0, beam_npmax=64, beam_evol=0, sp_push_pgc=1, laser_iter=las["iteration"], laser_k0=las["k0"], sp_ppc_r=cfg["ppc1"], use_graph=1,
                       device=local, stream=st1.cuda_stream, **simkw)
        one.init_species(*plasma)
        one.laser.upload(a_r, a_i)
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for k in range(nsteps):
            if k == nsteps - 1:
                evs[0].record(st1)
            one.step3d()
        evs[1].record(st1)
        st1.synchronize()
        r1, i1_ = one.laser.download()
        ez1, ps1 = one.field("e").lineout(3, 0, 1), one.field("psi").lineout(1, 0, 1)
        err = {"ez": 0.0, "psi": 0.0, "a": 0.0}
        for (off, n), sim in zip(lp.parts[lp.base:lp.base + S], lp.sims):
            gr, gi = sim.laser.download()
            err["a"] = max(err["a"], float(np.max(np.abs(gr[:, 2:2 + n] - r1[:, off + 2:off + 2 + n]))), float(np.max(np.abs(gi[:, 2:2 + n] - i1_[:, off + 2:off + 2 + n]))))
            ez, ps = sim.field("e").lineout(3, 0, 1), sim.field("psi").lineout(1, 0, 1)
            err["ez"] = max(err["ez"], float(np.max(np.abs(ez[:n] - ez1[off:off + n]))))
            err["psi"] = max(err["psi"], float(np.max(np.abs(ps[:n] - ps1[off:off + n]))))
        err["a"] /= float(max(np.max(np.abs(r1)), np.max(np.abs(i1_)))); err["ez"] /= float(np.max(np.abs(ez1))); err["psi"] /= float(np.max(np.abs(ps1)))
        if world > 1:
            e = torch.tensor([err["a"], err["ez"], err["psi"]], dtype=torch.float64, device="cuda")
            dist.all_reduce(e, op=dist.ReduceOp.MAX)
            err["a"], err["ez"], err["psi"] = (float(v) for v in e.tolist())
        moved = float(np.max(np.abs(r1 - a_r)) / np.max(np.abs(a_r)))
        check = {"ok": bool(max(err.values()) < 1e-6 and moved > 1e-3), "tol": 1e-6, "steps_compared": int(nsteps), "envelope_rel_err": err["a"], "ez_lineout_rel_err": err["ez"],
                 "psi_lineout_rel_err": err["psi"], "envelope_change_since_launch": moved, "single_step_ms": evs[0].elapsed_time(evs[1]),
                 "against": "one xi stage (a single k_sweep<0, PGC> over the whole box + one envelope solve per step: the path tests/test_gpu_laser.py holds against the oracle)"}
        one.close()
    peak, peak_src = hbm_peak()
    nit = iters / max(slices, 1)
    bpu = 24.0 + 72.0 * nit + 80.0 + 64.0 + 32.0
    ach = upd * bpu / (ms * 1e-3) / 1e9 / world
    roof = {"bound": "hbm", "kernel": f"k_sweep<0, PGC> x {S} concurrent per GPU (persistent, one xi slab each on {(148 - S) // S} SMs; laser slice images, pgc pushers and the susceptibility deposit inside) "
                                       f"+ {S} envelope-solve CTAs on SMs of their own; 65 536 particles per slice: barrier and field-program latency bound a stage, the other stages fill it",
            "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None, "peak_source": peak_src, "bytes_per_update": bpu,
            "us_per_slice_effective": ms * 1e3 / max(slices, 1), "slab_slices_by_stage": [n for _, n in lp.parts], "slab_balance_rounds": balance_log, "host_enqueue_ms_per_step_rank0": host_ms,
            "sweep_ms_per_step_by_stage_rank0": [round(p["ns_total"] * 1e-6 / args.steps, 3) for p in profs],
            "us_per_phase_by_stage_rank0": {"A (laser slice images beside it)": ph["A"], "amjdeposit_pgc per pass": ph["amj"], "C per pass": ph["C"], "push_u_pgc+push_x+qdeposit+chi || D": ph["push"]}}
    if rank == 0:
        line = {"metric": METRIC, "value": upd / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic (lattice plasma per fdist2d rule, Gaussian x sin^2 laser pulse of the lwfa deck)",
                "config": {"workload": f"C4: nr={cfg['nr']} nz={cfg['nz']} max_mode=0 Np/slice={npp0} robust_pgc, laser a0={las['a0']} k0={las['k0']} iteration {las['iteration']}",
                           "parallelism": f"xi-pipeline: {S} stages per GPU on SM partitions x {world} GPU(s), envelope slabs with guard hand-off, steady state (filled before the timed region)",
                           "l2": "field volumes + envelope volumes ~60 MB, particle planes 4 MB: L2 resident", "pc_iters_per_slice": nit},
                "clocks": clocks, "gpu_launches": int(launches), "roofline": roof}
        if e2e: line["e2e"] = e2e
        if check: line["parity_check"] = check
        if not args.no_cpu:
            try:
                upd_c, wall_c, k_c, _t = cpu_parallel("C4", args.ref_slices or None)
                line["cpu_baseline"] = {"value": upd_c / wall_c, "unit": UNIT, "cores": k_c, "kind": "port", "sample": cpu_sample_text("C4", k_c), "pc_iters_per_slice": LAST_CPU_NIT}
            except Exception as exc:
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {exc}"}
        print(json.dumps(line))
    lp.close()
    if world > 1:
        dist.destroy_process_group()


def run_c5_pipeline(args):
    """config 5 on the xi-pipeline (the deck is `nodes [1,2]`): --stages S slabs on one GPU, every stage with the neutral species attached to
    its sim (per-slice launch path, CUDA-graph replay) on its own stream -- a C5 slice is a chain of ~16 small dependent kernels that leaves
    the GPU almost empty, so S slabs in flight overlap nearly perfectly; the neutral's state (released electrons, ion buffer, rho_ion,
    levels) travels forward with the plasma hand-off (neutral_class.f03:1025-1101).  A timed step = one wave = every stage runs its slab
    once, in steady state; afterwards the pipeline is drained and the same number of 3D steps re-run on one stage (parity_check)."""
    import torch
    import torch.distributed as dist
    from qpad_b200 import capi
    from qpad_b200.pipeline import LocalPipeline
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:          # the stages continue across GPUs: the neutral's record travels over peer memory with the forward message
        dist.init_process_group("nccl")
    cfg, beam = deck_config("C5")
    neu = cfg["neutral"]
    _pl, bm = make_inputs(cfg, beam)
    empty = (np.zeros((0, 2)), np.zeros((0, 3)), np.zeros(0), np.zeros(0), np.zeros(0))
    S = args.stages
    lp = LocalPipeline(cfg, empty, bm, S, device=local, rank=rank, world=world, dist=dist if world > 1 else None, transport="p2p" if world > 1 else None)
    main = torch.cuda.current_stream()

    def sync_all():
        lp.sync()
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    lp.fill()
    for _ in range(args.warmup):
        lp.wave()
    sync_all()
    u0, i0, s0 = lp.stats()
    l0 = lp.launch_count()
    clk = ClockSampler(local); clk.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    ev0.record(main)
    for st in lp.streams:
        st.wait_event(ev0)
    th0 = time.perf_counter()
    for _ in range(args.steps):
        lp.wave()
    host_ms = 1e3 * (time.perf_counter() - th0) / args.steps
    for st in lp.streams:
        e = torch.cuda.Event(); e.record(st); main.wait_event(e)
    ev1.record(main)
    sync_all()
    ms = ev0.elapsed_time(ev1)
    clocks = clk.stop()
    u1, i1, s1 = lp.stats()
    upd, iters, slices = u1 - u0, i1 - i0, s1 - s0
    if world > 1:
        t = torch.tensor([ms, float(upd), float(iters), float(slices)], dtype=torch.float64, device="cuda")
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, upd, iters, slices = tmax[0].item(), tsum[1].item(), tsum[2].item(), tsum[3].item()
    launches = slices * 13 + 3 * iters + 30 * S * args.steps if not args.no_graph else lp.launch_count() - l0    # graph replay: head 1, 3 per PC iteration, tail 12; ~30 hand-off / beam launches per stage and wave
    check = None
    if args.check:
        lp.drain()
        nsteps = lp.sims[0].stats()[2] // lp.sims[0].nzp
        simkw = {k: cfg[k] for k in ("nr", "nz", "max_mode", "rmax", "zmin", "zmax", "dt", "iter_max", "iter_reltol", "iter_abstol")}
        st1 = torch.cuda.Stream(device=local)
        one = capi.Sim(sp_npmax=64, beam_npmax=len(bm[2]) + 1024, use_graph=1, device=local, stream=st1.cuda_stream, **simkw)
        one.init_species(*empty)
        one.attach_neutral(neu["element"], neu["ion_max"], (cfg["ppc1"], cfg["ppc2"]), cfg["num_theta"], neu.get("q", -1.0), neu.get("m", 1.0), neu.get("density", 1.0), cfg.get("n0", 1.0e17))
        one.beam.upload(*bm)
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for k in range(nsteps):
            if k == nsteps - 1:
                evs[0].record(st1)
            one.step3d()
        evs[1].record(st1)
        st1.synchronize()
        ez1, ps1 = one.field("e").lineout(3, 0, 1), one.field("psi").lineout(1, 0, 1)
        err_ez = err_ps = 0.0
        sums, nb = np.zeros(12), 0
        for (off, n), sim in zip(lp.parts[lp.base:lp.base + S], lp.sims):
            ez, ps = sim.field("e").lineout(3, 0, 1), sim.field("psi").lineout(1, 0, 1)
            err_ez = max(err_ez, float(np.max(np.abs(ez[:n] - ez1[off:off + n]))))
            err_ps = max(err_ps, float(np.max(np.abs(ps[:n] - ps1[off:off + n]))))
            bx, bp, bq = sim.beam.download()
            nb += len(bq)
            if len(bq):
                sums += beam_sums(bx, bp, bq)
        err_ez /= float(np.max(np.abs(ez1))); err_ps /= float(np.max(np.abs(ps1)))
        if world > 1:
            tt = torch.tensor(list(sums) + [float(nb)], dtype=torch.float64, device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.SUM)
            ee = torch.tensor([err_ez, err_ps], dtype=torch.float64, device="cuda"); dist.all_reduce(ee, op=dist.ReduceOp.MAX)
            sums, nb = tt[:12].cpu().numpy(), int(tt[12].item())
            err_ez, err_ps = float(ee[0].item()), float(ee[1].item())
        bx, bp, bq = one.beam.download()
        m1, m2 = beam_moments_from_sums(beam_sums(bx, bp, bq)), beam_moments_from_sums(sums)
        berr = 0.0
        for ax in "xy":
            c1, s1_, e1 = m1[ax]; c2, s2_, e2 = m2[ax]
            berr = max(berr, abs(c1 - c2) / s1_, abs(s1_ - s2_) / s1_, abs(e1 - e2) / e1)
        upd_one = one.stats()[0]
        check = {"ok": bool(err_ez < 1e-6 and err_ps < 1e-6 and berr < 1e-6 and nb == len(bq)), "tol": 1e-6, "steps_compared": int(nsteps), "ez_lineout_rel_err": err_ez,
                 "psi_lineout_rel_err": err_ps, "beam_moments_rel_err": berr, "beam_particles": nb, "beam_particles_single_stage": int(len(bq)),
                 "updates_per_step_single_stage": upd_one / nsteps, "single_step_ms": evs[0].elapsed_time(evs[1]),
                 "against": "one xi stage with the neutral attached (qpg_sim_attach_neutral, CUDA-graph replay: the path tests/test_gpu_neutral.py holds against the oracle)"}
        one.close()
    # end to end (after the comparison above: the uploads below restart the beam of stage 0 every wave, as run_c5's one-stage loop does): the
    # beam particles of the first slab host -> device each wave, line-outs of every slab + counters back
    if args.check:
        lp.restart()
        lp.fill()
    from qpad_b200.pipeline import split_beam
    dxi = (cfg["zmax"] - cfg["zmin"]) / cfg["nz"]
    b0 = split_beam(*bm, cfg["nz"], dxi, lp.G, parts=lp.parts)[lp.base]           # the beam of this rank's first slab
    sync_all()
    ue0 = lp.stats()[0]
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(args.steps):
        lp.sims[0].beam.upload(*b0)
        lp.wave()
        d2h = 24 * S
        for sim in lp.sims:
            ez = sim.field("e").lineout(3, 0, 1); ps = sim.field("psi").lineout(1, 0, 1)
            d2h += 8 * (len(ez) + len(ps))
        lp.stats()
    sync_all()
    te = time.perf_counter() - t0
    ue = float(lp.stats()[0] - ue0)
    if world > 1:
        tt = torch.tensor([ue], dtype=torch.float64, device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.SUM); ue = tt[0].item()
    e2e = {"value": ue / te, "unit": UNIT, "h2d_bytes_per_step": int(56 * len(b0[2])), "d2h_bytes_per_step": int(d2h),
           "what": "per step (wave): beam particles of the first slab host->device (qpg_part3d_upload), every stage runs its slab with ionisation, E_z and psi on-axis line-outs of every "
                   "slab + counters device->host"}
    peak, peak_src = hbm_peak()
    nit = iters / max(slices, 1)
    bpu = 112.0 + 64.0 * nit
    ach = upd * bpu / (ms * 1e-3) / 1e9
    roof = {"bound": "hbm", "kernel": f"per-slice launch path x {S} concurrent slabs (k_amjdeposit / k_push / k_qdeposit on the released electrons + k_neutral_ionize / _scan / _add + field programs; "
                                       "the ~16 dependent launches of a slice bound one slab, the slabs overlap)",
            "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None, "peak_source": peak_src, "bytes_per_update": bpu,
            "us_per_slice_effective": ms * 1e3 / max(slices, 1), "slab_slices_by_stage": [n for _, n in lp.parts], "host_enqueue_ms_per_step": host_ms}
    line = {"metric": METRIC, "value": upd / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (tri-Gaussian beam per the ionization deck, lithium gas ionised on the device)",
            "config": {"workload": f"C5: nr={cfg['nr']} nz={cfg['nz']} max_mode={cfg['max_mode']} neutral Li ion_max={neu['ion_max']} ppc {cfg['ppc1']}x{cfg['ppc2']} num_theta {cfg['num_theta']}, updates/step={upd / args.steps:.0f}",
                       "parallelism": f"xi-pipeline: {S} stages per GPU x {world} GPU(s) as concurrent streams (slice body replayed from a CUDA graph), neutral state in the forward hand-off, steady state",
                       "l2": "field volumes ~100 MB + electron planes: larger than L2 late in the step", "pc_iters_per_slice": nit},
            "clocks": clocks, "gpu_launches": int(launches), "roofline": roof, "e2e": e2e}
    if check: line["parity_check"] = check
    if rank != 0:
        lp.close()
        dist.destroy_process_group()
        return
    if not args.no_cpu:
        try:
            upd_c, wall_c, k_c, _t = cpu_parallel("C5", args.ref_slices or None)
            line["cpu_baseline"] = {"value": upd_c / wall_c, "unit": UNIT, "cores": k_c, "kind": "port", "sample": cpu_sample_text("C5", k_c), "pc_iters_per_slice": LAST_CPU_NIT}
        except Exception as exc:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {exc}"}
    print(json.dumps(line))
    lp.close()
    if world > 1:
        dist.destroy_process_group()


def run_b200_local(args):
    """the xi-pipeline mapped onto SM partitions (pipeline.LocalPipeline): --stages S sweep kernels per GPU run
    concurrently, global stage g on 3D step n-g; with N > 1 GPUs the stages continue across ranks over NCCL.  A timed
    step = one wave = every stage sweeps its slab once = one deck's worth of slices, measured in steady state (the
    pipeline is filled before the timed region)."""
    import torch
    import torch.distributed as dist
    from qpad_b200.pipeline import LocalPipeline
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        # one NCCL kernel can be resident beside the cooperative sweep kernels (e.g. the receive of the crossing beam
        # particles spins until the upstream stage has pushed its beam): its CTAs (one per channel) must fit in the SMs
        # LocalPipeline leaves free, otherwise a sweep launch waits for the message
        os.environ.setdefault("NCCL_MAX_P2P_NCHANNELS", "4")
        dist.init_process_group("nccl")     # lazy init: every rank pair gets its own p2p communicator / stream
    cfg, beam = deck_config(args.config)
    plasma, bm = make_inputs(cfg, beam)
    npp0 = len(plasma[4])
    S = args.stages
    parts = None
    balance_log = []
    mk = lambda pt: LocalPipeline(cfg, plasma, bm, S, device=local, rank=rank, world=world, dist=dist if world > 1 else None, transport=args.transport, partition=pt)
    if args.balance and world * S > 1:
        # slabs of equal measured cost instead of equal length (pipeline.balanced_partition).  Open loop first: an untimed calibration
        # sweep of one stage alone (probe_partition); then closed loop: the running pipeline measures its own stages side by side and the
        # slabs are re-cut until the stages' sweep times agree within 3 % (measured_partition; at most --rebalance rounds, all untimed)
        from qpad_b200.pipeline import probe_partition, measured_partition
        free = 4 if (world > 1 and (args.transport or os.environ.get("QPG_PIPELINE_TRANSPORT", "p2p")) == "nccl") else 0
        parts, beam_ns = probe_partition(cfg, plasma, bm, world * S, S, device=local, rank=rank, world=world, dist=dist if world > 1 else None, free_sms=free, with_beam_cost=True)
        lp = mk(parts)
        for rnd in range(args.rebalance):
            # measured over the same 3D step numbers the timed region will see (the wake deepens as the beam focuses: a slab's cost drifts)
            new_parts, stage_ms, spread = measured_partition(lp, cfg, beam_ns, nwaves=args.steps, nwarm=args.warmup)
            balance_log.append({"round": rnd, "spread": round(spread, 4), "busy_ms_by_stage": [round(v, 3) for v in stage_ms], "slab_slices": [n for _, n in parts]})
            done = new_parts is None or [tuple(p) for p in new_parts] == [tuple(p) for p in parts]
            lp.drain(); lp.close()
            parts = parts if done else new_parts
            lp = mk(parts)             # a fresh pipeline either way: the timed region then sees the 3D steps the partition was measured on
            if done:
                break
    else:
        lp = mk(parts)
    main = torch.cuda.current_stream()

    def sync_all():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    def join():   # the timing stream waits for everything the stage streams have been given so far
        for st in lp.streams + ([lp.comm] if lp.comm is not None else []):
            e = torch.cuda.Event(); e.record(st); main.wait_event(e)

    lp.fill()
    for _ in range(args.warmup):
        lp.wave()
    sync_all()
    lp.trace_reset()
    u0, i0, s0 = lp.stats()
    l0 = lp.launch_count()
    for sim in lp.sims:
        sim.sweep_profile(reset=True)      # in-kernel clocks only: no event pairs around the library calls of the timed region
    clk = ClockSampler(local); clk.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    ev0.record(main)
    for st in lp.streams:
        st.wait_event(ev0)
    th0 = time.perf_counter()
    for _ in range(args.steps):
        lp.wave()
    host_ms_per_wave = 1e3 * (time.perf_counter() - th0) / args.steps      # host time to ENQUEUE a wave (not device time)
    join()
    ev1.record(main)
    sync_all()
    ms = ev0.elapsed_time(ev1)
    if os.environ.get("QPG_TRACE_EVENTS"):
        for r_, rep in enumerate(lp.event_report()):
            waits = sum(v for k, v in rep.items() if k in ("w_fwd>got_fwd", "pre>got_back"))
            print(f"rank {rank} stage {r_} trace (ms): busy {sum(rep.values()) - waits:.2f} waits {waits:.2f} {rep}", file=sys.stderr, flush=True)
        lp._ev_on = False
    if os.environ.get("QPG_TRACE"):
        print(f"rank {rank}: host enqueue {host_ms_per_wave:.2f} ms per wave, device {ms / args.steps:.2f} ms per wave", file=sys.stderr, flush=True)
    clocks = clk.stop()
    u1, i1, s1 = lp.stats()
    upd, iters, slices = u1 - u0, i1 - i0, s1 - s0
    sweep_ms = sweep_n = 0
    profs = []
    for sim in lp.sims:
        profs.append(sim.sweep_profile())
    launches = lp.launch_count() - l0
    sweep_ms = sum(p["ns_total"] for p in profs) * 1e-6           # %globaltimer inside the kernels
    sweep_n = args.steps * (S + (1 if (lp.transport == "nccl" and rank > 0) else 0))   # one launch per slab (NCCL transport: first slice of a rank's first stage separately)
    # device time each stage spent inside its sweep kernel per step (all ranks): the pipeline runs at the pace of the slowest
    mine = [round(p["ns_total"] * 1e-6 / args.steps, 3) for p in profs]
    sweep_by_stage = mine
    if world > 1:
        every = [None] * world
        dist.all_gather_object(every, mine)
        sweep_by_stage = [v for m in every for v in m]
    if world > 1:
        t = torch.tensor([ms, float(upd), float(launches), float(iters), float(slices)], dtype=torch.float64, device="cuda")
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, upd, launches, iters, slices = tmax[0].item(), tsum[1].item(), tsum[2].item(), tsum[3].item(), tsum[4].item()
    value = upd / (ms * 1e-3)

    # ---- end to end (N = 1): the plasma lattice goes host -> device every step (stage 0), line-outs of every slab come back
    e2e = None
    if world == 1:
        torch.cuda.synchronize()
        ue0 = lp.stats()[0]
        t0 = time.perf_counter()
        d2h = 0
        for _ in range(args.steps):
            lp.wave(upload=plasma)
            d2h = 0
            for sim in lp.sims:                       # each stage's slab of the E_z and psi on-axis line-outs (the step it has just swept)
                ez = sim.field("e").lineout(3, 0, 1); ps = sim.field("psi").lineout(1, 0, 1)
                d2h += 8 * (len(ez) + len(ps))
            d2h += 24 * S
            lp.stats()
        torch.cuda.synchronize()
        te = time.perf_counter() - t0
        ue = lp.stats()[0] - ue0
        e2e = {"value": ue / te, "unit": UNIT, "h2d_bytes_per_step": int(8 * 8 * npp0), "d2h_bytes_per_step": int(d2h),
               "what": "per step (wave): plasma lattice host->device through qpg_part2d_upload into stage 0, every stage sweeps its slab, E_z and psi on-axis line-outs of every slab + counters device->host"}

    peak, peak_src = hbm_peak()
    nit = iters / max(slices, 1)
    bytes_total = upd * (112.0 + 64.0 * nit)
    ach = bytes_total / (ms * 1e-3) / 1e9 / world      # per GPU
    ph = {}
    for key, cntk in (("A", "slices"), ("amj", "amj_phases"), ("C", "amj_phases"), ("push", "slices")):
        us = [p["cyc_" + key] * (p["ns_total"] / max(p["cyc_total"], 1.0)) * 1e-3 / max(p[cntk], 1.0) for p in profs]
        ph[key] = {"us_per_phase_by_stage": [round(u, 2) for u in us]}
    roof = {"bound": "hbm", "kernel": "k_sweep<%d> x %d concurrent per GPU (persistent; each on 1/%d of the SMs, one xi slab each)" % (cfg["max_mode"], S, S),
            "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None, "peak_source": peak_src,
            "bytes_per_update": 112.0 + 64.0 * nit, "algorithmic_bytes_timed": bytes_total, "launches_timed_rank0": int(sweep_n),
            "avg_launch_ms_rank0": sweep_ms / max(sweep_n, 1), "us_per_slice_effective": ms * 1e3 / max(slices, 1), "host_enqueue_ms_per_step_rank0": host_ms_per_wave,
            "us_per_slice_per_stage_rank0": [round(p["ns_total"] * 1e-3 / max(p["slices"], 1.0), 2) for p in profs],
            "sweep_ms_per_step_by_stage": sweep_by_stage,
            "slab_slices_by_stage": [n for _, n in lp.parts],
            "slab_partition": "cost-balanced: untimed calibration sweep (pipeline.probe_partition) + closed loop on the running pipeline (pipeline.measured_partition)" if parts is not None else "equal length (options_class.f03:103-106)",
            "slab_balance_rounds": balance_log,
            "stage_time_spread": (max(sweep_by_stage) / (sum(sweep_by_stage) / len(sweep_by_stage)) - 1.0) if sweep_by_stage else None,
            "phases_rank0": {"A||update_bound": ph["A"], "amjdeposit (64 B/particle)": ph["amj"], "C": ph["C"], "push_u+push_x+qdeposit||D (112 B/particle)": ph["push"]},
            "note": "achieved = algorithmic bytes of ALL sweep launches in the timed region / its duration, per GPU (the S kernels of a GPU overlap: a stage's latency-bound field phases and barriers hide behind the other stages' particle phases); particle planes stay L2-resident"}
    # the other roofline of this path: fp64 issue.  SASS of this build (cuobjdump, max_mode 1): amjdeposit 252 fp64 instructions per
    # thread + 8 DMMA per warp (= 64 DFMA-equivalents), push 206, qdeposit 45 + 8 DMMA -> per warp of 32 particles 316 per
    # amjdeposit pass and 315 for push + qdeposit; B200 issues 2 fp64 warp-instructions per clock and SM (37 TFLOP/s measured)
    try:
        if cfg["max_mode"] == 1:
            wi = 316.0 * nit + 315.0
            mhz = clocks.get("sm_max_mhz") or 1965.0
            bound = 148 * mhz * 1e6 * 64.0 / wi
            roof["fp64_pipe"] = {"fp64_warp_instr_per_32_updates": wi, "bound_updates_per_s_per_gpu": bound, "achieved_frac": value / world / bound,
                                 "note": "arithmetic intensity ~10 flop/B against a machine balance of 5.6 flop/B: the fp64 pipe, not HBM, is the nearer roof"}
    except Exception:   # an explanatory extra must never cost the bench line
        pass
    try:   # DRAM traffic: ncu --set full of ONE sweep launch in this very launch shape (37 CTAs = one SM partition), per slice, x the slices of a launch
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_sweep_traffic.json")))
        if args.config == "C2":
            per_launch = slices / max(sweep_n * world, 1)
            roof["traffic"] = float(tj["dram_bytes_per_slice"]) * per_launch
            roof["traffic_source"] = (f"{tj['source']}: {tj['dram_bytes_per_slice'] / 1e3:.0f} KB of DRAM traffic per slice x {per_launch:.0f} slices per launch; algorithmic particle bytes "
                                      f"per slice {npp0 * (112.0 + 64.0 * nit) / 1e6:.1f} MB: the particle planes stay in L2, DRAM sees the field-volume slice stores")
            roof["algorithmic_bytes_per_launch"] = bytes_total / max(sweep_n * world, 1)
    except Exception:
        pass
    roof_hbm = None
    if world == 1 and not args.no_micro:
        # the push and deposit kernels individually, streaming from HBM (north star: >= 60 % of the HBM roofline): fields of
        # the slab that holds the wake (the stage's current slice images)
        from qpad_b200.pipeline import kernel_microbench
        lp.sync()
        roof_hbm = kernel_microbench(lp.sims[min(1, S - 1)], cfg, peak)
    # ---- correctness of THIS run + what a user gets outside the steady state --------------------------------------------
    nsteps_done = lp.w                 # after the drain every stage has finished 3D steps 0 .. lp.w - 1
    lp.drain()
    sync_all()
    parity = single = None
    if args.check:
        parity, single = parity_check(cfg, plasma, bm, lp, nsteps_done, local, rank, world, dist if world > 1 else None)
    fill_incl = None
    if args.fill_steps:
        fill_incl = {"what": "updates/s of a run of K 3D steps INCLUDING pipeline fill and drain (K + stages - 1 waves), host wall clock around wave() x K + drain(), max over ranks",
                     "stages": lp.G}
        for K in args.fill_steps:
            lp.restart()
            sync_all()
            u_a = lp.stats()[0]
            t_a = time.perf_counter()
            for _ in range(K):
                lp.wave()
            lp.drain()
            sync_all()
            dt_f = time.perf_counter() - t_a
            u_f = lp.stats()[0] - u_a
            tt = torch.tensor([dt_f, float(u_f)], dtype=torch.float64, device="cuda")
            if world > 1:
                tm = tt.clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
                ts = tt.clone(); dist.all_reduce(ts, op=dist.ReduceOp.SUM)
                dt_f, u_f = tm[0].item(), ts[1].item()
            fill_incl[f"K={K}"] = {"updates_per_s": u_f / dt_f, "seconds": dt_f, "model_K_over_K_plus_S_minus_1": K / (K + lp.G - 1.0)}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            upd_c, wall_c, k_c, _steps = cpu_parallel(args.config, args.ref_slices or None)
            cpu = {"value": upd_c / wall_c, "unit": UNIT, "cores": k_c, "kind": "port", "sample": cpu_sample_text(args.config, k_c), "pc_iters_per_slice": LAST_CPU_NIT}
        except Exception as exc:
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {exc}"}
    if rank == 0:
        how = {"p2p": "hand-offs between GPUs = pack kernels writing into the next GPU's memory over NVLink (CUDA IPC mapping) + flag words awaited by stream memory operations",
               "nccl": "NCCL send/recv between GPUs"}.get(lp.transport)
        where = "one GPU" if world == 1 else f"{world} GPUs x {S} stages, {how}"
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": DATA,
                "config": {"workload": workload_string(args.config, cfg, npp0),
                           "parallelism": f"{where}: xi-pipeline over {world * S} stages, each an SM partition running one persistent sweep kernel; stage g sweeps slab g of 3D step n-g (the reference's pipeline, parallel_module.f03:221-239); filled before the timed region, a timed step = every stage sweeps its slab once = {cfg['nz']} slices",
                           "l2": "step working set (field volumes ~0.9 GB + beam) exceeds the 126 MB L2",
                           "pc_iters_per_slice": nit},
                "clocks": clocks, "gpu_launches": int(launches), "roofline": roof}
        if e2e: line["e2e"] = e2e
        if roof_hbm: line["roofline_hbm_stream"] = roof_hbm
        if cpu: line["cpu_baseline"] = cpu
        if parity: line["parity_check"] = parity
        if single: line["single_step"] = single
        if fill_incl: line["fill_inclusive"] = fill_incl
        print(json.dumps(line))
    lp.close()
    if world > 1:
        dist.destroy_process_group()


def run_c5(args):
    """config 5 (input_file/ionization): nspecies 0, one lithium neutral species ionised by the beam (ADK), electrons created on the
    device -- one xi stage on one GPU, the neutral attached to qpg_sim (qpg_sim_attach_neutral: per-slice launch path, CUDA-graph
    replay of the slice body).  A step = one 3D step = nz slices + beam push + renewal; the particle count of a slice grows from 0
    inside the step, so `value` counts the updates the device counters report (released electrons only)."""
    import torch
    from qpad_b200 import capi, decks
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference for the CPU arm)")
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        raise SystemExit("bench.py --config C5: the neutral species runs on one xi stage (one GPU)")
    cfg, beam = deck_config("C5")
    neu = cfg["neutral"]
    _pl, bm = make_inputs(cfg, beam)
    stream = torch.cuda.Stream()
    simkw = {k: cfg[k] for k in ("nr", "nz", "max_mode", "rmax", "zmin", "zmax", "dt", "iter_max", "iter_reltol", "iter_abstol")}
    sim = capi.Sim(sp_npmax=64, beam_npmax=len(bm[2]) + 1024, use_graph=0 if args.no_graph else 1, stream=stream.cuda_stream, **simkw)
    empty = (np.zeros((0, 2)), np.zeros((0, 3)), np.zeros(0), np.zeros(0), np.zeros(0))
    sim.init_species(*empty)
    ne = sim.attach_neutral(neu["element"], neu["ion_max"], (cfg["ppc1"], cfg["ppc2"]), cfg["num_theta"], neu.get("q", -1.0), neu.get("m", 1.0),
                            neu.get("density", 1.0), cfg.get("n0", 1.0e17))
    sim.beam.upload(*bm)
    for _ in range(args.warmup):
        sim.step3d()
    torch.cuda.synchronize()
    u0, i0, s0 = sim.stats()
    l0 = sim.ctx.launch_count()
    clk = ClockSampler(0); clk.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record(stream)
    for _ in range(args.steps):
        sim.step3d()
    ev1.record(stream)
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    clocks = clk.stop()
    u1, i1, s1 = sim.stats()
    upd, iters, slices = u1 - u0, i1 - i0, s1 - s0
    # graph replay: head 1, PC iteration 3 (two amjdeposits + program C), tail 12 (program D, push, compact, qdeposit, neutral update 4, push, compact, 2 qdeposits)
    launches = sim.ctx.launch_count() - l0 if args.no_graph else slices * 13 + 3 * iters + 8 * args.steps
    t0 = time.perf_counter()
    ue0 = sim.stats()[0]
    d2h = 0
    for _ in range(args.steps):
        sim.beam.upload(*bm)
        sim.beam_qdp_begin(); sim.beam_qdp_end(); sim.begin_step()
        sim.run_slices(1, sim.nzp)
        ez = sim.field("e").lineout(3, 0, 1); ps = sim.field("psi").lineout(1, 0, 1)
        d2h = 8 * (len(ez) + len(ps)) + 24
        sim.stats()
        sim.beam_push(); sim.renew()
    torch.cuda.synchronize()
    te = time.perf_counter() - t0
    e2e = {"value": (sim.stats()[0] - ue0) / te, "unit": UNIT, "h2d_bytes_per_step": int(56 * len(bm[2])), "d2h_bytes_per_step": int(d2h),
           "what": "per step: beam particles host->device (qpg_part3d_upload), nz slices with ionisation, E_z and psi on-axis line-outs + counters device->host, beam push, renewal"}
    peak, peak_src = hbm_peak()
    nit = iters / max(slices, 1)
    bpu = 112.0 + 64.0 * nit
    ach = upd * bpu / (ms * 1e-3) / 1e9
    roof = {"bound": "hbm", "kernel": "per-slice launch path: k_amjdeposit / k_push / k_qdeposit on the released electrons + k_neutral_ionize / _scan / _add + field programs (the electron count of a slice grows from 0 to its final value inside the step: the ~16 dependent launches of a slice bound it, not the particle bytes)",
            "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None, "peak_source": peak_src, "bytes_per_update": bpu,
            "us_per_slice": ms * 1e3 / max(slices, 1)}
    cpu = None
    if not args.no_cpu:
        try:
            upd_c, wall_c, k_c, _t = cpu_parallel("C5", args.ref_slices or None)
            cpu = {"value": upd_c / wall_c, "unit": UNIT, "cores": k_c, "kind": "port", "sample": cpu_sample_text("C5", k_c), "pc_iters_per_slice": LAST_CPU_NIT}
        except Exception as exc:
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {exc}"}
    line = {"metric": METRIC, "value": upd / (ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (no pre-ionised plasma, lithium neutral gas, bi-Gaussian beam of the ionization deck with a fixed-seed NumPy generator)",
            "config": {"workload": f"C5: nr={cfg['nr']} nz={cfg['nz']} max_mode={cfg['max_mode']} neutral Li ion_max={neu['ion_max']} ppc {cfg['ppc1']}x{cfg['ppc2']} num_theta {cfg['num_theta']}, updates/step={upd // max(args.steps, 1)}",
                       "parallelism": "single stage, slice body " + ("as plain stream launches" if args.no_graph else "replayed from a CUDA graph (device-side WHILE node for the predictor-corrector loop)"),
                       "l2": "field volumes ~100 MB + up to 65 MB of electron planes: larger than L2 late in the step", "pc_iters_per_slice": nit},
            "clocks": clocks, "gpu_launches": int(launches), "roofline": roof, "e2e": e2e}
    if cpu: line["cpu_baseline"] = cpu
    print(json.dumps(line))
    sim.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C2")
    ap.add_argument("--ref-slices", type=int, default=0, help="CPU arm: cap of the slab length each host core times (0 = the whole deck: k cores x their equal-length slabs = one 3D step)")
    ap.add_argument("--roof-slices", type=int, default=64)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--check", type=int, default=1, help="1 (default): after the timed region compare the pipelined run's E_z / psi line-outs and beam moments with a one-stage run of the same 3D steps (parity_check in the JSON line)")
    ap.add_argument("--fill-steps", type=int, nargs="*", default=[28, 128], help="report updates/s of runs of K 3D steps including pipeline fill and drain (fill_inclusive); empty = skip")
    ap.add_argument("--no-graph", action="store_true", help="C4: plain stream launches instead of CUDA-graph replay of the slice body")
    ap.add_argument("--no-sweep", action="store_true", help="per-slice CUDA-graph launches instead of the persistent sweep kernel")
    ap.add_argument("--no-micro", action="store_true", help="skip the stream-from-HBM kernel microbenchmark")
    ap.add_argument("--legacy-pipeline", action="store_true", help="N>1: one stage per GPU through pipeline.PipelineStage")
    ap.add_argument("--rebalance", type=int, default=3, help="closed-loop rounds of the slab balancing (pipeline.measured_partition); 0 = open-loop calibration only")
    ap.add_argument("--balance", type=int, default=1, help="1 = cost-balanced xi slabs from an untimed calibration sweep (default), 0 = the reference's equal-length slabs")
    ap.add_argument("--transport", default=None, choices=["p2p", "nccl"], help="N>1: how the stage hand-offs cross GPUs (default p2p = peer-memory writes + flags, csrc/p2p.cu)")
    ap.add_argument("--stages", type=int, default=0, help="xi-pipeline stages mapped onto SM partitions of ONE GPU (LocalPipeline); 0 = auto (up to 4), 1 = a single sweep kernel on all SMs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "C4":
        if args.stages == 0:
            args.stages = 5 if not (args.no_sweep or args.no_graph) else 1      # measured on one B200: 1 stage 2.4e9, 4 stages 5.1e9, 5 stages 5.5e9, 6 stages 4.6e9 updates/s
        if args.stages > 3:
            os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")     # S stage streams + S envelope side streams > the default 8 hardware queues (+2 %)
        (run_c4_pipeline if args.stages > 1 else run_c4)(args)
    elif args.config == "C5":
        if args.stages == 0:
            args.stages = 1 if args.no_graph else 24      # measured on one B200: 1 stage 1.5e8, 4 stages 3.6e8; unrolled slice graph: 8 stages 4.5e8, 16 5.6e8, 24 6.4e8, 32 6.6e8
        if args.stages > 8:
            # more than 8 stage streams: the default of 8 hardware queues would serialise them in pairs (8, 12, 16, 24 stages then all run at the pace of 8);
            # must be set before the CUDA context exists (torch is imported inside the run functions)
            os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
        (run_c5_pipeline if args.stages > 1 else run_c5)(args)
    else:
        if args.stages == 0:       # auto: as many stages as the field team (one CTA per 32 radial nodes) and the slab length allow, at most 4
            cfg, _ = deck_config(args.config)
            nteam, stages = (cfg["nr"] + 31) // 32, 1
            world = int(os.environ.get("WORLD_SIZE", "1"))
            # (measured on one B200: C2 -- team of 32 CTAs -- 4 stages; C1 / C3 -- team of 8 -- 8 stages: 2.0e9 / 1.07e9 updates/s against
            # 1.4e9 / 7.1e8 with 4: the smaller the deck, the more of a slice is barrier and field-program latency that other stages can fill)
            for cand in (2, 3, 4, 5, 6, 7, 8):
                if 148 // cand > nteam + 1 and cfg["nz"] // (cand * world) >= 48 and cfg["max_mode"] <= 2:
                    stages = cand
            args.stages = stages
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if not args.no_sweep and (args.stages > 1 or (world > 1 and not args.legacy_pipeline)):
            run_b200_local(args)
        else:
            run_b200(args)


if __name__ == "__main__":
    main()

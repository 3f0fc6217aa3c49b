"""torchrun worker of tests/test_gpu_pipeline.py: a 2-stage xi-pipeline on 2 GPUs, results saved per rank."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from qpad_b200 import decks  # noqa: E402
from qpad_b200.pipeline import LocalPipeline, PipelineStage  # noqa: E402


def main():
    out, nsteps = sys.argv[1], int(sys.argv[2])
    stages = int(sys.argv[3]) if len(sys.argv) > 3 else 0      # 0: one PipelineStage per rank; S > 0: LocalPipeline with S stages per rank
    transport = sys.argv[4] if len(sys.argv) > 4 else None     # LocalPipeline transport between ranks: p2p | nccl
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    one_device = bool(os.environ.get("QPG_TEST_ONE_DEVICE"))   # every rank on GPU 0 (peer-memory links through IPC work inside one GPU too)
    if one_device:
        local = 0
    torch.cuda.set_device(local)
    dist.init_process_group("gloo" if one_device else "nccl")
    barrier = (lambda: dist.barrier()) if one_device else (lambda: dist.barrier(device_ids=[local]))
    cfg = dict(nr=64, nz=32, max_mode=1, rmax=5.0, zmin=-5.0, zmax=5.0, dt=10.0, ppc1=2, ppc2=2, num_theta=8, iter_max=2,
               iter_reltol=1e-3, iter_abstol=1e-3)
    beam = dict(decks.CONFIGS["C1"]["beam"])
    bm = decks.beam_std(cfg["nr"], cfg["nz"], cfg["rmax"], cfg["zmin"], cfg["zmax"], **beam)
    plasma = decks.plasma_uniform(cfg["nr"], cfg["rmax"], cfg["ppc1"], cfg["ppc2"], cfg["num_theta"])
    lwfa = len(sys.argv) > 5 and sys.argv[5] == "lwfa"          # laser-driven deck: the envelope slabs and their guard hand-off cross the ranks too
    laser = None
    if lwfa:
        z = np.load(os.path.join(out, "lwfa_inputs.npz"))
        cfg = dict(nr=int(z["nr"]), nz=int(z["nz"]), max_mode=0, rmax=12.0, zmin=-3.0, zmax=6.0, dt=2.0, iter_max=6, iter_reltol=1e-3, iter_abstol=1e-6, ppc1=4, ppc2=2,
                   num_theta=8, laser=dict(k0=20.0, iteration=3))
        plasma = tuple(z[k] for k in ("x", "p", "g", "psi", "q"))
        bm = (np.zeros((0, 3)), np.zeros((0, 3)), np.zeros(0))
        laser = (z["ar"], z["ai"])
    if len(sys.argv) > 5 and sys.argv[5] == "neutral":           # ionisation deck: the neutral's record crosses the ranks with the forward message
        cfg = dict(nr=64, nz=36, max_mode=1, rmax=6.0, zmin=0.0, zmax=8.0, dt=10.0, iter_max=3, iter_reltol=1e-3, iter_abstol=1e-3, ppc1=2, ppc2=2, num_theta=8, n0=1.0e17,
                   neutral=dict(element=3, ion_max=2))
        bm = decks.beam_std(cfg["nr"], cfg["nz"], cfg["rmax"], cfg["zmin"], cfg["zmax"], **dict(decks.CONFIGS["C5"]["beam"]))
        plasma = (np.zeros((0, 2)), np.zeros((0, 3)), np.zeros(0), np.zeros(0), np.zeros(0))
    if stages > 0:
        lp = LocalPipeline(cfg, plasma, bm, stages, device=local, rank=rank, world=world, dist=dist, transport=transport, laser=laser)
        lp.fill()
        for _ in range(nsteps):
            lp.wave()
        lp.drain()                                           # every stage has finished 3D steps 0 .. G-2+nsteps
        torch.cuda.synchronize()
        for r, s in enumerate(lp.sims):
            bx, bp, bq = s.beam.download()
            extra = {}
            if lwfa:
                extra["ar"], extra["ai"] = s.laser.download()
            np.savez(os.path.join(out, f"stage{rank * stages + r}.npz"), psi=s.field("psi").download_f2(), e=s.field("e").download_f2(), bx=bx, bp=bp,
                     bq=bq, stats=np.array(s.stats()), noff2=s.noff2, nzp=s.nzp, **extra)
        barrier()
        lp.close()
        dist.destroy_process_group()
        return
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        st = PipelineStage(cfg, plasma, bm, stream=stream, rank=rank, world=world, device=local)
        # the steady-state driver of bench.py: complete steps, then prime / primed steps / unwind
        st.step()
        st.prime()
        for _ in range(nsteps - 1):
            st.step_primed()
        st.unwind()
        st.drain()
        torch.cuda.synchronize()
        s = st.sim
        bx, bp, bq = s.beam.download()
        np.savez(os.path.join(out, f"rank{rank}.npz"), psi=s.field("psi").download_f2(), e=s.field("e").download_f2(), bx=bx, bp=bp, bq=bq,
                 stats=np.array(s.stats()), noff2=st.noff2, nzp=st.nzp)
        dist.barrier(device_ids=[local])
        st.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

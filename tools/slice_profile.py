"""per-slice cost profile of a deck along xi (qpg_sim_slice_trace): python tools/slice_profile.py [config] [ctas] -> gpurun_out/slice_profile_<config>_<ctas>.npz"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from bench import deck_config, make_inputs  # noqa: E402
from qpad_b200.pipeline import probe_slice_costs, balanced_partition, slab_partition  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
ctas = int(sys.argv[2]) if len(sys.argv) > 2 else 0
cfg, beam = deck_config(name)
plasma, bm = make_inputs(cfg, beam)
ns, it, beam_ns = probe_slice_costs(cfg, plasma, bm, 0, ctas)
print(f'beam deposit+push+move: {beam_ns.sum() * 1e-6:.3f} ms on the whole GPU')
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez(os.path.join(ROOT, "gpurun_out", f"slice_profile_{name}_{ctas}.npz"), ns=ns, it=it, beam_ns=beam_ns)
blk = max(1, len(ns) // 32)
print(f"{name} ctas={ctas}: total {ns.sum() * 1e-6:.2f} ms, mean {ns.mean() * 1e-3:.1f} us/slice, iterations/slice {it.mean():.3f}")
print("us/slice by 1/32 of the box:", [round(float(ns[k:k + blk].mean()) * 1e-3, 1) for k in range(0, len(ns), blk)])
print("iters/slice by 1/32        :", [round(float(it[k:k + blk].mean()), 2) for k in range(0, len(ns), blk)])
for G in (4, 8, 16, 32):
    if G * 2 > len(ns):
        continue
    u = max(ns[a:a + n].sum() for a, n in slab_partition(len(ns), G)); b = max(ns[a:a + n].sum() for a, n in balanced_partition(ns, G))
    print(f"G={G}: slowest slab uniform {u * 1e-6:.2f} ms, balanced {b * 1e-6:.2f} ms, ideal {ns.sum() / G * 1e-6:.2f} ms")

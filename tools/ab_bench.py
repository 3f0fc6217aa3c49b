#!/usr/bin/env python
"""A/B of compile-time variants of libqpadb200.so on the GPU box.

  python tools/ab_bench.py build  name1:"-DFLAG ..." name2:"..."     (here, no GPU: nvcc -> qpad_b200/variants/lib_<name>.so)
  python tools/ab_bench.py run [--args "<bench.py args>"] [names...]   (on the box: bench.py per variant via QPG_LIB, one summary line each)

The variants travel to the box with the snapshot (*.so is git-ignored, not gpurun-ignored)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VDIR = os.path.join(ROOT, "qpad_b200", "variants")
sys.path.insert(0, ROOT)


def build(specs):
    import __graft_entry__ as g
    os.makedirs(VDIR, exist_ok=True)
    procs = []
    for spec in specs:
        name, _, flags = spec.partition(":")
        out = os.path.join(VDIR, f"lib_{name}.so")
        cmd = [os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")] + g.NVCC_FLAGS + flags.split() + [os.path.join(g.CSRC, "lib.cu"), "-o", out]
        procs.append((name, subprocess.Popen(cmd)))
    for name, p in procs:
        assert p.wait() == 0, name
        print("built", name)


def run(names, bench_args):
    names = names or sorted(f[4:-3] for f in os.listdir(VDIR) if f.startswith("lib_") and f.endswith(".so"))
    for name in names:
        env = dict(os.environ, QPG_LIB=os.path.join(VDIR, f"lib_{name}.so"))
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + bench_args, env=env, capture_output=True, text=True)
        try:
            j = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception:
            print(f"{name}: FAILED rc={r.returncode} {r.stderr[-600:]}")
            continue
        rf, st = j.get("roofline", {}), j.get("roofline_hbm_stream", {})
        print(f"{name}: value {j['value']:.4g} ms/step {j['ms_per_step']:.2f} frac {rf.get('frac', 0):.3f} nit {j['config'].get('pc_iters_per_slice', 0):.3f} "
              f"stage_ms {rf.get('sweep_ms_per_step_by_stage')} us/slice {rf.get('us_per_slice')} "
              f"phases {json.dumps(rf.get('phases_rank0') or rf.get('phases'))[:400]} "
              f"stream amj {st.get('amjdeposit_frac', 0):.3f} qdep {st.get('qdeposit_frac', 0):.3f} push {st.get('push_u_plus_push_x_frac', 0):.3f} "
              f"parity {j.get('parity_check', {}).get('ok')}", flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    else:
        rest = sys.argv[2:]
        bargs = ["--no-cpu", "--check", "0", "--fill-steps", "--steps", "6"]
        if rest and rest[0] == "--args":
            bargs = rest[1].split(); rest = rest[2:]
        run(rest, bargs)

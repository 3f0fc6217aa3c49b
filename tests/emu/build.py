"""TEST INFRASTRUCTURE ONLY: builds tests/emu/_build/libqpademu.so -- the per-routine part of qpad_b200/csrc (everything except the
persistent sweep kernel, the cluster kernels, the CUDA-graph simulation object and the peer-memory transport) compiled for the HOST through
tests/emu/cuda_runtime.h (fibers for CTA threads, exact barriers and warp collectives).  The product sources are not touched;
the textual transformations applied to the copies under _build/ are:
  * `kernel<<<grid, block, smem, stream>>>(args)`  ->  `emu::launch(grid, block, smem, [&] { kernel(args); })`
  * `extern __shared__ T name[];`                  ->  `T *name = (T *)emu::dyn_smem;`
  * the five inline-PTX statements of particles.cu ->  their C++ meaning (PTX table below; each must match exactly once)"""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "qpad_b200", "csrc")
OUT = os.path.join(HERE, "_build")
COMPILE_ONLY = ["fused.cu"]        # needed to link sim.cu; its kernels need live thread-block clusters and are not run
SOURCES = ["fields.cu", "particles.cu", "beam.cu", "laser.cu", "fused.cu", "sweep.cu", "sim.cu", "neutral.cu", "subcyc.cu", "vpot.cu", "diag.cu"]
# by default the emulation runs the plain per-slice launches of qpg_sim: a request for CUDA graphs is ignored (same launches,
# replayed or not), the cluster programs refuse to be switched on, the persistent sweep kernel is off unless qpg_sim_set_sweep(s, 1)
PATCH = {
    "sim.cu": [
        ("s->prm = *prm;", "s->prm = *prm; s->prm.use_graph = 0;"),
        ("s->prm.use_graph = use_graph != 0; return 0; }", "(void)use_graph; return 0; }"),
        ("s->use_fused = (prm->nr <= FT * FC && prm->max_mode <= 2);", "s->use_fused = false;"),
        ("s->use_sweep = sweep_supported(*prm);", "s->use_sweep = getenv(\"QPAD_EMU_SWEEP\") ? sweep_supported(*prm) : false;"),
        ("const bool can = s->prm.nr <= FT * FC && s->prm.max_mode <= 2;", "const bool can = false;"),
        # the persistent sweep kernel runs as a cooperative launch of the emulation: all CTAs alive at once (emu::launch_coop)
        ("void *args[] = {(void *)&a};\n    return cudaLaunchCooperativeKernel((const void *)k_sweep<M, PGC>, dim3(grid), dim3(SW_T), args, sweep_smem<M>(), st);",
         "emu::launch_coop(dim3(grid), dim3(SW_T), sweep_smem<M>(), [&] { k_sweep<M, PGC>(a); });\n    return cudaSuccess;"),
    ],
    "sweep.cu": [
        ("__shared__ int sm_i[48];", "int *sm_i = (int *)emu::cta_static(48 * sizeof(int), 1);"),     # per CTA: the CTAs of this kernel are alive together
    ],
}
PTX = {
    # sweep.cu: the loads that poll for another CTA (grid / team barrier counters, flagged exchange words) yield to the scheduler
    "sweep.cu": [
        ('asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");', "v = *(const volatile unsigned *)p; emu::poll_yield();"),
        ('asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");', "(void)id; (void)nthreads; emu_unsupported_ptx();"),
        ('asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((unsigned)__double2loint(v)), "r"(seq), "r"((unsigned)__double2hiint(v)), "r"(seq)\n                 : "memory");',
         "*p = uint4{(unsigned)__double2loint(v), seq, (unsigned)__double2hiint(v), seq};"),
        ('asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p) : "memory");',
         "{ const uint4 w = *p; a = w.x; b = w.y; c = w.z; d = w.w; } emu::poll_yield();"),
        ('asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");', "*p = v;"),
        ('asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(a.las_progress) : "memory");', "v = *(const volatile unsigned *)a.las_progress; emu::poll_yield();"),
        ('asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(nstart));', "nstart = clock64();"),
        ('asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));', "now = clock64();"),
        ('asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(nend));', "nend = clock64();"),
    ],
    "laser.cu": [
        ('asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(progress), "r"(progress_base + (unsigned)j) : "memory");', "*progress = progress_base + (unsigned)j;"),
    ],
    "particles.cu": [
        ('asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(y));', "r = 1.0 / y;"),
        ('asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));', "y = 1.0 / std::sqrt(x);"),
        ('asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");', "*p += v;"),
        ('asm volatile("red.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");', "*p += v;"),
        ('asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));',
         "emu_dmma_m8n8k4(c0, c1, a, b);"),
    ],
}
LIB = os.path.join(OUT, "libqpademu.so")


def _match(text, i, open_ch, close_ch):
    """index just past the bracket that closes the one at text[i]"""
    depth = 0
    while True:
        ch = text[i]
        if ch == open_ch:
            depth += 1
        elif ch == close_ch:
            depth -= 1
            if depth == 0:
                return i + 1
        i += 1


def _split_top(s):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip()); cur = ""
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


def transform(text):
    out, pos = "", 0
    for m in re.finditer(r"([A-Za-z_]\w*(?:<[^<>;(){}]*>)?)\s*<<<", text):
        if m.start() < pos:
            continue
        cfg_end = text.index(">>>", m.end())
        cfg = _split_top(text[m.end():cfg_end])
        a0 = text.index("(", cfg_end)
        a1 = _match(text, a0, "(", ")")
        smem = cfg[2] if len(cfg) > 2 else "0"
        out += text[pos:m.start()] + f"emu::launch({cfg[0]}, {cfg[1]}, {smem}, [&] {{ {m.group(1)}{text[a0:a1]}; }})"
        pos = a1
    out += text[pos:]
    return re.sub(r"extern\s+__shared__\s+(\w+)\s+(\w+)\[\];", r"\1 *\2 = (\1 *)emu::dyn_smem;", out)


def strip_ptx(text):
    """compile-only files: every inline-PTX statement becomes a call that aborts if it is ever reached"""
    out, pos = "", 0
    for m in re.finditer(r"\basm\s+volatile\s*\(", text):
        end = _match(text, m.end() - 1, "(", ")")
        out += text[pos:m.start()] + "emu_unsupported_ptx()"
        pos = end
    return out + text[pos:]


def transform_file(name, text):
    if name in COMPILE_ONLY:
        text = strip_ptx(text)
    for ptx, cpp in PTX.get(name, []) + PATCH.get(name, []):
        assert text.count(ptx) == 1, (name, ptx)
        text = text.replace(ptx, cpp)
    assert not re.search(r"\basm\b", re.sub(r"//.*", "", text)), f"{name}: unhandled inline assembly"
    return transform(text)


def build(force=False):
    """QPAD_EMU_ASAN=1: the AddressSanitizer build (run python with LD_PRELOAD=$(gcc -print-file-name=libasan.so) and
    ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0): every access of a kernel to 'device' memory (the host heap) is
    bounds-checked -- a memcheck of the device sources without a GPU."""
    global LIB
    asan = bool(int(os.environ.get("QPAD_EMU_ASAN", "0")))
    extra = os.environ.get("QPAD_EMU_DEFS", "").split()          # e.g. -DQPG_GATHER_FOLDED: an experimental variant of the device sources
    tag = ("_asan" if asan else "") + "".join("_" + re.sub(r"\W", "", d) for d in extra)
    LIB = os.path.join(OUT, f"libqpademu{tag}.so")
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(CSRC, f) for f in SOURCES if os.path.exists(os.path.join(CSRC, f))]
    deps = srcs + [os.path.join(HERE, f) for f in ("cuda_runtime.h", "emu_lib.cpp", "build.py")] + [os.path.join(CSRC, "common.cuh"),
                                                                                                      os.path.join(ROOT, "include", "qpad_b200.h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    for s in srcs:
        with open(s) as f:
            t = transform_file(os.path.basename(s), f.read())
        with open(os.path.join(OUT, os.path.basename(s) + ".cpp"), "w") as f:
            f.write(t)
    defs = [f"-DEMU_HAVE_{os.path.basename(s).split('.')[0].upper()}" for s in srcs]
    san = ["-fsanitize=address", "-fno-omit-frame-pointer"] if asan else []
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-ffp-contract=off", "-fPIC", "-shared"] + san + ["-I", HERE, "-I", OUT, "-I", CSRC] + defs + extra +
                          [os.path.join(HERE, "emu_lib.cpp"), "-o", LIB])
    return LIB


if __name__ == "__main__":
    print(build(force=True))

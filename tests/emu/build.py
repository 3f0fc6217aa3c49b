"""TEST INFRASTRUCTURE ONLY: builds tests/emu/_build/libqpademu.so -- the shuffle-free kernels of qpad_b200/csrc compiled for the
HOST through tests/emu/cuda_runtime.h (fibers for CTA threads).  The only source transformation is the launch syntax:
`kernel<<<grid, block, smem, stream>>>(args)` becomes `emu::launch(grid, block, [&] { kernel(args); })`."""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "qpad_b200", "csrc")
OUT = os.path.join(HERE, "_build")
SOURCES = ["neutral.cu", "subcyc.cu", "vpot.cu", "diag.cu"]      # device code that has no warp-level primitives / PTX
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
        out += text[pos:m.start()] + f"emu::launch({cfg[0]}, {cfg[1]}, [&] {{ {m.group(1)}{text[a0:a1]}; }})"
        pos = a1
    return out + text[pos:]


def build(force=False):
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(CSRC, f) for f in SOURCES if os.path.exists(os.path.join(CSRC, f))]
    deps = srcs + [os.path.join(HERE, f) for f in ("cuda_runtime.h", "emu_lib.cpp", "build.py")] + [os.path.join(CSRC, "common.cuh"),
                                                                                                      os.path.join(ROOT, "include", "qpad_b200.h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    for s in srcs:
        with open(s) as f:
            t = transform(f.read())
        with open(os.path.join(OUT, os.path.basename(s) + ".cpp"), "w") as f:
            f.write(t)
    defs = [f"-DEMU_HAVE_{os.path.basename(s).split('.')[0].upper()}" for s in srcs]
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-ffp-contract=off", "-fPIC", "-shared", "-I", HERE, "-I", OUT, "-I", CSRC] + defs +
                          [os.path.join(HERE, "emu_lib.cpp"), "-o", LIB])
    return LIB


if __name__ == "__main__":
    print(build(force=True))

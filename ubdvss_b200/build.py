"""In-tree build of libubd.so for sm_100a (nvcc cross-compiles without a GPU).

The shared object is git-ignored but travels to the GPU box with the snapshot; nothing is JIT-built
at import time and nothing is installed into site-packages."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libubd.so")
STAMP = os.path.join(CSRC, ".libubd.stamp")
SOURCES = ["ubd_api.cu", "ubd_rect.cpp"]

# UBD_TC_TRACE=1 in the environment compiles the in-kernel cycle trace in (tools/tc_trace.py, tools/stem_trace.py)
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
] + (["-DUBD_TC_TRACE=1"] if os.environ.get("UBD_TC_TRACE") == "1" else [])


def _digest() -> str:
    h = hashlib.sha256()
    inc = os.path.join(os.path.dirname(HERE), "include", "ubd.h")
    for f in sorted(os.listdir(CSRC)) + [inc]:
        p = f if os.path.isabs(f) else os.path.join(CSRC, f)
        if p.endswith((".cu", ".cuh", ".cpp", ".h", ".inc")):
            with open(p, "rb") as fh:
                h.update(p.encode()); h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/ -> csrc/libubd.so (skipped when sources are unchanged).  Returns the path."""
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == dig:
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libubd.so (no CPU fallback exists)")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas"); cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    with open(STAMP, "w") as fh:
        fh.write(dig)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

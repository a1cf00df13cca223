"""Builds the C part of the CPU oracle / CPU baseline (test infrastructure, NOT product code):
oracle/csrc/flat_select.c -> oracle/_build/libflatselect.so (git-ignored; it travels to the GPU box with the snapshot,
and is rebuilt there with the image's gcc when missing or stale).  `__graft_entry__.build()` calls build()."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "flat_select.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libflatselect.so")
# no -march=native: the library is built in one container and may run on another host CPU; the hot loop (a running
# maximum over 64 scores) is cloned for AVX2 / AVX-512 by the compiler and dispatched at load time.
FLAGS = ["-O3", "-fopenmp", "-shared", "-fPIC", "-std=gnu11"]


def _digest() -> str:
    with open(SRC, "rb") as fh:
        return hashlib.sha256(fh.read() + " ".join(FLAGS).encode()).hexdigest()


def build(force: bool = False) -> str:
    """Returns the path of the built library; raises RuntimeError when there is no gcc or the compile fails."""
    stamp = os.path.join(OUT_DIR, "digest.txt")
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
        return LIB
    # plain `gcc` first: $CC may name a toolchain without OpenMP support (this image: /opt/gcc has no libgomp.spec)
    compilers = [c for c in (shutil.which("gcc"), shutil.which(os.environ.get("CC", "cc"))) if c]
    if not compilers:
        raise RuntimeError("gcc not found: cannot build oracle/_build/libflatselect.so")
    os.makedirs(OUT_DIR, exist_ok=True)
    errors = []
    for cc in compilers:
        r = subprocess.run([cc, *FLAGS, "-o", LIB, SRC], capture_output=True, text=True)
        if r.returncode == 0:
            break
        errors.append(f"{cc}: {r.stderr.strip()}")
    else:
        raise RuntimeError(f"could not compile {SRC}:\n" + "\n".join(errors))
    with open(stamp, "w") as fh:
        fh.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force=True))

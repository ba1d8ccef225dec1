"""Build libtatva_b200.so in-tree with nvcc for sm_100a (B200).

    python -m tatva_b200.build [--force] [--verbose]

The shared library travels to the GPU box with the repo snapshot; nothing is JIT-compiled
at import time.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtatva_b200.so")
SOURCES = ["generic.cu", "neo_hookean.cu", "user_law.cu", "halo_nccl.cu", "host.cpp", "xla_ffi_shim.cc"]  # the shim is empty without jaxlib headers
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(HERE, "..", "include", "tatva_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fopenmp,-O3",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libtatva_b200.so")


STAMP = LIB + ".stamp"


def _source_hash() -> str:
    """Content hash of every source, header and flag that goes into the library.  File times do not survive the copy to
    the GPU box, so staleness is decided by content: the stamp written next to the library travels with it."""
    import hashlib

    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for path in [os.path.join(CSRC, s) for s in SOURCES] + HEADERS + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".inc"))):
        with open(path, "rb") as f:
            h.update(os.path.basename(path).encode() + b"\0" + f.read())
    return h.hexdigest()


def needs_build() -> bool:
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as f:
        return f.read().strip() != _source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile (if needed) and return the library path.  Safe to call from several processes at once (one rank per
    GPU importing a fresh checkout): an exclusive file lock serialises them, the late comers find the library up to
    date, and the link goes to a temporary name that is renamed into place."""
    import fcntl

    if not force and not needs_build():
        return LIB
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    with open(os.path.join(objdir, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():  # another process built it while we waited
                return LIB
            return _build_locked(verbose, objdir)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _embed_sources(objdir: str) -> None:
    """The text of the kernel-template headers as C string literals (build/embedded_sources.inc): user_law.cu hands them
    to NVRTC when it compiles a user-supplied constitutive law into the same fused kernel templates."""
    out = []
    for var, name in (("kSrcCommon", "common.cuh"), ("kSrcFused", "fused.cuh")):
        with open(os.path.join(CSRC, name)) as f:
            text = f.read()
        assert ')TATVA_SRC"' not in text
        # string literals are limited to 64 KiB by some front ends: emit adjacent raw literals of <= 16 KiB
        chunks = [text[i : i + 16000] for i in range(0, len(text), 16000)]
        out.append(f"static const char {var}[] =\n" + "\n".join(f'R"TATVA_SRC({c})TATVA_SRC"' for c in chunks) + ";\n")
    path = os.path.join(objdir, "embedded_sources.inc")
    new = "".join(out)
    if not os.path.exists(path) or open(path).read() != new:
        with open(path, "w") as f:
            f.write(new)


def _build_locked(verbose: bool, objdir: str) -> str:
    nvcc = _nvcc()
    _embed_sources(objdir)
    objs = []
    procs = []
    for s in SOURCES:
        o = os.path.join(objdir, os.path.splitext(s)[0] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-I", objdir, "-x", "cu", "-c", os.path.join(CSRC, s), "-o", o]
        if verbose:
            cmd[1:1] = ["-Xptxas", "-v"]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}")
    tmp = LIB + f".tmp{os.getpid()}"
    link = [nvcc, "-shared", "-o", tmp, *objs, "-Xcompiler", "-fopenmp", "-lgomp", "-ldl"]
    subprocess.run(link, check=True)
    os.replace(tmp, LIB)
    with open(STAMP, "w") as f:
        f.write(_source_hash())
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)

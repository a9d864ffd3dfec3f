"""Build the sm_100a shared library (``libartspeech_b200.so``) in-tree with nvcc.

The library is the C-ABI boundary declared in ``include/artspeech_b200.h``.  It is built
explicitly for ``compute_100a/sm_100a`` (tcgen05 / TMEM / TMA need the arch-specific target) with
``-lineinfo`` so ncu's source page maps to the .cu files.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libartspeech_b200.so")
_STAMP = os.path.join(HERE, "csrc", ".build_stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _fingerprint() -> str:
    h = hashlib.sha256()
    files = _sources() + sorted(
        os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    files.append(os.path.join(HERE, "..", "include", "artspeech_b200.h"))
    for f in files:
        h.update(os.path.basename(f).encode())     # content-addressed: the repo may live anywhere
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    if not (os.path.isfile(LIB_PATH) and os.path.isfile(_STAMP)):
        return False
    with open(_STAMP) as f:
        return f.read().strip() == _fingerprint()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every ``csrc/*.cu`` and link ``libartspeech_b200.so``.  Returns the library path.

    Safe under ``torchrun``: an exclusive file lock serialises concurrent builders (the ranks that lose the race find
    an up-to-date library when they get the lock) and the library is linked to a temporary name and renamed into
    place, so a process that is loading it never sees a half-written file."""
    if not force and is_current():
        return LIB_PATH
    import fcntl
    with open(os.path.join(HERE, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and is_current():
                return LIB_PATH
            return _build_locked(force, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force: bool, verbose: bool) -> str:
    nvcc = os.environ.get("NVCC", "nvcc")
    objs = []
    procs = []
    os.makedirs(os.path.join(CSRC, "obj"), exist_ok=True)
    hdr = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)) + [os.path.join("..", "..", "include", "artspeech_b200.h")]:
        if f.endswith((".cuh", ".h")):
            with open(os.path.join(CSRC, f), "rb") as fh:
                hdr.update(fh.read())
    hdr.update(" ".join(NVCC_FLAGS).encode())
    stamps = []
    for src in _sources():
        obj = os.path.join(CSRC, "obj", os.path.basename(src)[:-3] + ".o")
        # per-object stamp: a translation unit is recompiled only when it or a header changed
        with open(src, "rb") as fh:
            digest = hashlib.sha256(hdr.digest() + fh.read()).hexdigest()
        objs.append(obj)
        if not force and os.path.isfile(obj) and os.path.isfile(obj + ".stamp") and \
                open(obj + ".stamp").read().strip() == digest:
            continue
        stamps.append((obj + ".stamp", digest))
        cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for src, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out.decode(), file=sys.stderr)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{out.decode()}")
    for path, digest in stamps:
        with open(path, "w") as f:
            f.write(digest)
    tmp = LIB_PATH + f".tmp{os.getpid()}"
    link = [nvcc, "-shared", "-o", tmp, *objs, "-cudart", "static",
            "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout.decode()}")
    os.replace(tmp, LIB_PATH)
    with open(_STAMP, "w") as f:
        f.write(_fingerprint())
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

"""Build the C-ABI CUDA library in-tree with nvcc for sm_100a (no torch extension machinery:
the boundary is plain C, loaded through ctypes)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libegogen_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + \
           [os.path.join(PKG_DIR, "..", "include", "egogen_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def _obj_dir() -> str:
    d = os.path.join(PKG_DIR, "_obj")
    os.makedirs(d, exist_ok=True)
    return d


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile egogen_b200/csrc/*.cu -> egogen_b200/libegogen_b200.so. Cross-compiles without a GPU.
    Each source is compiled to its own object (in parallel, re-used when neither it nor a header changed) and linked."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build libegogen_b200.so")
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if not f.endswith(".cu")] + \
              [os.path.join(PKG_DIR, "..", "include", "egogen_b200.h")]
    t_hdr = max(os.path.getmtime(h) for h in headers)
    flags = [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else []) + \
        os.environ.get("EG_NVCC_EXTRA", "").split()      # debug builds only (e.g. -DEG_LBS_PROF=1)
    jobs, objs = [], []
    for src in sources():
        obj = os.path.join(_obj_dir(), os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), t_hdr):
            jobs.append((src, subprocess.Popen([nvcc] + flags + ["-c", "-o", obj, src], stdout=subprocess.PIPE,
                                               stderr=subprocess.STDOUT, text=True)))
    log = ""
    for src, p in jobs:
        out, _ = p.communicate()
        log += out
        if p.returncode != 0:
            for _, q in jobs:
                if q.poll() is None:
                    q.kill()
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, out))
    r = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs,
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(log)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

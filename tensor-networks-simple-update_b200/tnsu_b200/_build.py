"""In-tree build of libtnsu_b200.so for sm_100a (nvcc cross-compiles without a GPU).

The kernel variants live in `csrc/edge_inst.cu`, compiled once per (scalar type, MMAX) so the
translation units build in parallel; `csrc/engine.cu` holds the C ABI.  Objects go to `csrc/build/`
(git-ignored), the library next to this file so it travels with the source snapshot.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(PKG_DIR), "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libtnsu_b200.so")
NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC"] + os.environ.get("TNSU_NVCC_EXTRA", "").split()


def _sources_digest() -> str:
    h = hashlib.sha256()
    for name in sorted(os.listdir(CSRC)):
        if name.endswith((".cu", ".cuh")):
            with open(os.path.join(CSRC, name), "rb") as f:
                h.update(name.encode() + b"\0" + f.read())
    for header in ("tnsu_b200.h", "tnsu_b200_debug.h"):
        with open(os.path.join(os.path.dirname(os.path.dirname(PKG_DIR)), "include", header), "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the library if the sources changed; returns the library path."""
    stamp = LIB_PATH + ".sha256"
    digest = _sources_digest()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    bdir = os.path.join(CSRC, "build")
    os.makedirs(bdir, exist_ok=True)
    jobs = [(os.path.join(bdir, "engine.o"), [os.path.join(CSRC, "engine.cu")]),
            (os.path.join(bdir, "svdw_inst.o"), [os.path.join(CSRC, "svdw_inst.cu")]),
            (os.path.join(bdir, "rc_inst.o"), [os.path.join(CSRC, "rc_inst.cu")]),
            (os.path.join(bdir, "lqc_inst.o"), [os.path.join(CSRC, "lqc_inst.cu")]),
            (os.path.join(bdir, "fused_inst.o"), [os.path.join(CSRC, "fused_inst.cu")])]
    for cx in (0, 1):
        for mm in (8, 16, 32):
            jobs.append((os.path.join(bdir, f"edge_{'c' if cx else 'd'}{mm}.o"),
                         [f"-DINST_CPLX={cx}", f"-DINST_MM={mm}", os.path.join(CSRC, "edge_inst.cu")]))

    def compile_one(job):
        obj, args = job
        cmd = [nvcc, *NVCC_FLAGS, "-Xptxas", "-v", "-c", "-o", obj, *args]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed: {' '.join(cmd)}\n{res.stdout}\n{res.stderr}")
        with open(obj + ".ptxas.log", "w") as f:
            f.write(res.stderr)
        return obj

    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        objs = list(pool.map(compile_one, jobs))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH, *objs]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed: {' '.join(cmd)}\n{res.stdout}\n{res.stderr}")
    with open(stamp, "w") as f:
        f.write(digest)
    if verbose:
        print(f"built {LIB_PATH}")
    return LIB_PATH


if __name__ == "__main__":
    build(force=True, verbose=True)

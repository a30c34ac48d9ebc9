"""In-tree nvcc build of libconvofusion_b200.so (sm_100a only; cross-compiles without a GPU)."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "lib" / "libconvofusion_b200.so"
SOURCES = ["api.cu", "gemm_simt.cu", "gemm_tc.cu", "gemm_split.cu", "rowops.cu", "attention.cu", "cross_tc.cu", "sched.cu", "weg.cu", "rowblock.cu", "denoiser.cu", "vae.cu"]
NVCC_FLAGS = os.environ.get("CFB_EXTRA_NVCC", "").split() + ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "--threads", "0"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: convofusion_b200 needs the CUDA toolkit to build its kernels")


def _fingerprint() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "convofusion_b200.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    stamp = LIB.with_suffix(".stamp")
    fp = _fingerprint()
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == fp:
        return LIB
    LIB.parent.mkdir(parents=True, exist_ok=True)
    objdir = PKG / "lib" / "obj"
    objdir.mkdir(exist_ok=True)
    nvcc = _nvcc()
    procs = []
    for src in SOURCES:
        obj = objdir / (src + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for src, obj, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
        objs.append(str(obj))
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *objs]
    subprocess.run(link, check=True)
    stamp.write_text(fp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

"""Builds libnkbk.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

Usage:  python -m nkb_classification_b200.build [--force] [--verbose]
The library lands next to this file so it travels with the source tree.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libnkbk.so"
SOURCES = ["nkbk_api.cu", "k1_general.cu", "k1_fast.cu", "k1_fast_aug.cu", "k2_heads.cu", "k2_fused.cu", "k2_tc.cu", "k3_metrics.cu", "k4_comm.cu", "k4_peer.cu", "k5_auc.cu"]
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]


def nvcc_path() -> str:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(cand).exists():
        raise RuntimeError("nvcc not found; libnkbk.so cannot be built")
    return cand


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [PKG.parent / "include" / "nkbk.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    nvcc = nvcc_path()
    objs = []
    build_dir = CSRC / "_obj"
    build_dir.mkdir(exist_ok=True)
    common = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--use_fast_math=false"] + ARCH_FLAGS
    common = [f for f in common if f != "--use_fast_math=false"]  # never fast-math: parity needs IEEE ops
    if verbose:
        common += ["-Xptxas", "-v"]
    procs = []
    for src in SOURCES:
        obj = build_dir / (Path(src).stem + ".o")
        cmd = [nvcc, "-c", str(CSRC / src), "-o", str(obj)] + common
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(obj))
    failed = False
    for src, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            failed = True
            sys.stderr.write(f"--- nvcc {src} failed ---\n{out}\n")
        elif verbose and out:
            sys.stderr.write(f"--- nvcc {src} ---\n{out}\n")
    if failed:
        raise RuntimeError("nvcc compilation failed")
    tmp = LIB.with_suffix(".so.tmp")
    cmd = [nvcc, "-shared", "-o", str(tmp)] + objs + ARCH_FLAGS + ["-Xcompiler", "-fPIC", "-ldl"]
    subprocess.check_call(cmd)
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)

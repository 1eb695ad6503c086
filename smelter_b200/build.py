"""Builds smelter_b200/libsmelter_b200.so in-tree with nvcc for sm_100a (no torch, no cmake).

Usage:  python -m smelter_b200.build [--force] [--verbose]
Also builds oracle/libsmelter_oracle.so (plain C restatement of the host-side index maps) when asked by
__graft_entry__.build().
"""
from __future__ import annotations

import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libsmelter_b200.so")

SOURCES = [
    "onnx_wire.cc",
    "engine.cc",
    "converters.cc",
    "capi.cc",
    "nccl_shim.cc",
    "kernels/conv_igemm.cu",
    "kernels/conv_pair.cu",
    "kernels/elementwise.cu",
    "kernels/pool_norm.cu",
]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-std=c++17", "-O3", "-lineinfo",
    "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function",
    "-DSMELTER_BUILDING",
]
for _macro in ("SMELTER_TRYWAIT_HINT_NS", "SMELTER_MMA_LOOKAHEAD", "SMELTER_CLUSTER_RELAXED"):  # compile-time experiment switches
    if os.environ.get(_macro):
        NVCC_FLAGS.append(f"-D{_macro}={os.environ[_macro]}")
if os.environ.get("SMELTER_CONV_INSTRUMENT"):  # perf experiments: %globaltimer stamps + ablation flags in the conv kernel
    NVCC_FLAGS.append("-DSMELTER_CONV_INSTRUMENT=1")


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _all_inputs():
    out = [os.path.join(HERE, "..", "include", "smelter_b200.h")]
    for root, _, files in os.walk(CSRC):
        for f in files:
            if f.endswith((".cc", ".cu", ".h", ".cuh")):
                out.append(os.path.join(root, f))
    return [os.path.abspath(p) for p in out]


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    stamp = os.path.join(BUILD, "stamp")
    digest = _digest(_all_inputs())
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    nvcc = _nvcc()

    def compile_one(src: str) -> str:
        obj = os.path.join(BUILD, src.replace("/", "_") + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose and r.stderr.strip():
            print(r.stderr, file=sys.stderr)
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-Xlinker", "--no-undefined", "-ldl"]
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)

"""Compiles oracle/host_oracle.c (plain C, gcc) into oracle/_build/libhost_oracle.so.

oracle/_ref/ (a build of the reference's own sources) does not exist for this project: the reference is Swift
on Apple's MetalPerformanceShaders/Accelerate (no Swift toolchain here, frameworks are Apple-only) and its one
Python file imports `onnx` + `onnx.optimizer`, neither installed nor installable.  See DESIGN.md §Oracle.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libhost_oracle.so")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    subprocess.run(["gcc", "-O2", "-std=c99", "-Wall", "-Wextra", "-shared", "-fPIC", SRC, "-o", LIB], check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True))

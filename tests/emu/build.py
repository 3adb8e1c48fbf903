"""Builds tests/emu/libmft_emu.so: the product's device thread bodies compiled for the host (TEST INFRASTRUCTURE; see the
headers of the .cpp files here).  -ffp-contract=off mirrors nvcc -fmad=false."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libmft_emu.so")
CSRC = os.path.join(HERE, "..", "..", "meshfreetrixi.jl_b200", "csrc")


def build(force=False):
    srcs = [os.path.join(HERE, f) for f in sorted(os.listdir(HERE)) if f.endswith(".cpp")]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-Wall", "-o", LIB] + srcs
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed building the emulation harness:\n" + res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True))

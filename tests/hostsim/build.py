"""Builds the CPU simulation of the solver core (test infrastructure, see hostsim.cpp)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build", "libcave_hostsim.so")


def build(force=False):
    src = os.path.join(HERE, "hostsim.cpp")
    deps = [src] + [os.path.join(HERE, "..", "..", "cave_b200", "csrc", f) for f in ("solver_core.cuh", "ctx.cuh")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(p) for p in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", src, "-o", OUT])
    return OUT


if __name__ == "__main__":
    print(build(force=True))

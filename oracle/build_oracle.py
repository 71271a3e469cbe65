"""Build recipe for the C part of the oracle (gcc only; no reference sources are compiled here -- the reference's own
my_cpp/common.cpp is compiled, with its FCL/octomap boundary shimmed, by oracle/build_ref.py into oracle/_ref/)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libfilter_ref.so")


def build(force=False):
    srcs = [os.path.join(HERE, f) for f in ("filter_ref.c", "occupancy_ref.c")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in srcs):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["gcc", "-O2", "-mfma", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", "-o", LIB] + srcs + ["-lm"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force=True))

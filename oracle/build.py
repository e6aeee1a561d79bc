"""Compile the C restatement of the scoring simulator (checker only) into oracle/_build/."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "nbody_ref.c")
OUT = os.path.join(HERE, "_build", "libnbody_ref.so")


def build(force=False):
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if force or not os.path.exists(OUT) or os.path.getmtime(OUT) < os.path.getmtime(SRC):
        subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", SRC, "-o", OUT, "-lm"], check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))

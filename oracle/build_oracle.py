#!/usr/bin/env python
"""Compile oracle/ptg_oracle.c (the CPU restatement; TEST INFRASTRUCTURE ONLY) into
oracle/_build/libptg_oracle.so with gcc -O2 and FMA contraction off (the reference
build has no FMA: plain x86-64 SSE2 code from gcc -O2)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "ptg_oracle.c")
OUT = os.path.join(HERE, "_build", "libptg_oracle.so")


def build(force=False):
    if (not force and os.path.exists(OUT)
            and os.path.getmtime(OUT) >= os.path.getmtime(SRC)):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-std=c99",
                           "-Wall", "-Wextra", "-shared", "-fPIC", SRC, "-o", OUT])
    return OUT


if __name__ == "__main__":
    print(build(force=True))

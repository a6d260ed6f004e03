"""TEST INFRASTRUCTURE -- compile oracle/*.c into oracle/_build/liboracle.so (plain gcc, no reference sources)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liboracle.so")
SRCS = ["quad_oracle.c", "maze_oracle.c"]
DEPS = SRCS + ["quad_oracle_impl.h", "quad_oracle.h"]
CFLAGS = ["-O2", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-Wall", "-Wno-unused-function"]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(HERE, d)) > t for d in DEPS)


def build(force=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = ["gcc"] + CFLAGS + ["-shared", "-o", LIB] + [os.path.join(HERE, s) for s in SRCS] + ["-lm", "-lpthread"]
    subprocess.check_call(cmd, cwd=HERE)
    return LIB


if __name__ == "__main__":
    print(build(force=True))

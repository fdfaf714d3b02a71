"""Build the compiled part of the oracle (oracle/ws_ref.cpp -> oracle/_build/libwsref.so).

TEST INFRASTRUCTURE.  `__graft_entry__.build()` calls this here; the .so travels to the GPU box
with the snapshot (built files are git-ignored, not gpurun-ignored).  g++ is also on the GPU box, so
a missing or stale library is rebuilt on first use."""

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "ws_ref.cpp")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libwsref.so")


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    gxx = shutil.which("g++")
    if gxx is None:
        if os.path.exists(LIB):
            return LIB
        raise RuntimeError("g++ not found and oracle/_build/libwsref.so is not built")
    os.makedirs(OUT_DIR, exist_ok=True)
    res = subprocess.run([gxx, "-O2", "-std=c++17", "-shared", "-fPIC", "-o", LIB + ".tmp", SRC],
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + res.stdout + res.stderr)
    os.replace(LIB + ".tmp", LIB)
    return LIB


if __name__ == "__main__":
    print(build(force=True))

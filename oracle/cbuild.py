"""TEST INFRASTRUCTURE ONLY -- builds oracle/csrc/stag_ref.c into oracle/_build/libstag_ref.so
(gcc -O3 -fopenmp) and binds it with ctypes.  `oracle/_build/` is git-ignored but travels to
the GPU box with the snapshot."""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "stag_ref.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libstag_ref.so")

_lib = None


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["gcc", "-O3", "-march=x86-64-v2", "-fopenmp", "-fPIC", "-shared", "-fvisibility=hidden",
           "-o", LIB, SRC, "-lm"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("gcc failed on the oracle's C restatement:\n" + r.stdout)
    return LIB


def load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.ref_num_threads.restype = ctypes.c_int
    return _lib

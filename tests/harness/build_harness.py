"""Builds the host-side interpreter harness (g++) used by the CPU tests."""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_interp_host.so")
SRC = os.path.join(HERE, "interp_host.cpp")
HDR = os.path.join(HERE, "..", "..", "gaussianprocessfundamentals_b200", "csrc", "program.cuh")


def load():
    stale = (not os.path.exists(OUT)) or any(os.path.getmtime(f) > os.path.getmtime(OUT) for f in (SRC, HDR))
    if stale:
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", "-ffp-contract=off", SRC, "-o", OUT])
    lib = ctypes.CDLL(OUT)
    lib.h_matrix.restype = ctypes.c_int
    lib.h_matrix.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                             ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                             ctypes.c_void_p]
    return lib

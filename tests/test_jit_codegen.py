"""CPU: the run-time specialisation of the assembly / gradient kernels (csrc/jit.cu).  Source generation is host-only
and NVRTC compiles for sm_100a without a GPU, so the generated kernels of every tree the parity tests use - and of the
deepest tree bench.py's C3 grammar can draw - are checked here: they compile, and their SASS contains no local-memory
access (LDL / STL) at all (the interpreter's dynamically indexed stack / tape / adjoints lived in local
memory).  The numerical parity of these kernels is tested on the GPU (tests/test_gpu_*.py run on them by default)."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

from gaussianprocessfundamentals_b200 import _lib
from gaussianprocessfundamentals_b200.program import compile_spec

COMPOSITE = ("MUL", [("ADD", [("SE",), ("PER",)]), ("LIN",)])
DEEP = ("ADD", [("MUL", [("SE",), ("PER",), ("LIN",)]), ("MUL", [("SE",), ("ADD", [("LIN",), ("PER",)])])])
CP3 = ("CP", [("SE",), ("PER",), ("ADD", [("SE",), ("LIN",)])])
TERNARY_PER = ("MUL", [("MUL", [("MUL", [("PER",)] * 3)] * 3)] * 3)      # depth 3, 3-ary, 27 PER leaves: tape 106 > 64


def _lib_or_skip():
    try:
        return _lib.load()
    except _lib.GpbError as exc:
        pytest.skip(str(exc))


def _source(lib, cp, cp_mode=1):
    code = np.ascontiguousarray(cp.code, dtype=np.int32)
    need = ctypes.c_size_t()
    rc = lib.gpb_jit_source(code.ctypes.data_as(_lib.c_int32_p), cp.n_ops, cp.dim, cp_mode, None, 0, ctypes.byref(need))
    assert rc == 0, lib.gpb_last_error()
    buf = ctypes.create_string_buffer(need.value)
    assert lib.gpb_jit_source(code.ctypes.data_as(_lib.c_int32_p), cp.n_ops, cp.dim, cp_mode, buf, need.value,
                              ctypes.byref(need)) == 0
    return buf.value.decode()


def _cubin(lib, cp, cp_mode=1):
    code = np.ascontiguousarray(cp.code, dtype=np.int32)
    need = ctypes.c_size_t()
    rc = lib.gpb_jit_cubin(code.ctypes.data_as(_lib.c_int32_p), cp.n_ops, cp.dim, cp_mode, b"sm_100a", None, 0,
                           ctypes.byref(need))
    if rc == 3000 and b"cannot load libnvrtc" in lib.gpb_last_error():
        pytest.skip("libnvrtc is not available on this host")
    assert rc == 0, lib.gpb_last_error().decode()
    buf = ctypes.create_string_buffer(need.value)
    assert lib.gpb_jit_cubin(code.ctypes.data_as(_lib.c_int32_p), cp.n_ops, cp.dim, cp_mode, b"sm_100a", buf, need.value,
                             ctypes.byref(need)) == 0
    return buf.raw[:need.value]


def test_generated_source_is_straight_line():
    lib = _lib_or_skip()
    src = _source(lib, compile_spec(COMPOSITE, 1, False))
    body = src[src.index("struct Prog"):]
    assert "N_HP = 4, DIM = 1" in body
    # one named scalar per node, hyper-parameter offsets as literals, the reverse sweep as single assignments
    for needle in ("const double v4 = v2 * v3;", "g[0] += a0 * d0_0;", "g[2] += a1 * d1_1;", "g[3] += a3 * d3_0;",
                   "const double a2 = a4 * v3;", "gpb_sincos(u1"):
        assert needle in body, needle
    assert "for (" not in body.split("extern \"C\"")[0].split("static __device__")[1]      # value(): no loops, no dispatch
    # indicator change points carry no gradient (Operators.py:396-400); the smooth modes do
    ind = _source(lib, compile_spec(CP3, 1, False), cp_mode=1)
    smooth = _source(lib, compile_spec(CP3, 1, False), cp_mode=2)
    assert "dp" not in ind[ind.index("struct Prog"):].replace("gpb_spec", "") and "const double dp" in smooth


@pytest.mark.parametrize("name,spec,dim,cp_mode", [
    ("c2", COMPOSITE, 1, 1), ("deep", DEEP, 1, 1), ("cp_indicator", CP3, 1, 1), ("cp_sigmoid", CP3, 1, 0),
    ("cp_approx", CP3, 1, 2), ("matern_wn", ("ADD", [("MAT32",), ("MAT52",), ("WN",)]), 1, 1),
    ("se_ard_d8", ("SE_ARD",), 8, 1), ("mixed_d3", ("ADD", [("SE_ARD",), ("LIN",), ("SE",)]), 3, 1)])
def test_specialised_kernels_compile_without_local_memory(name, spec, dim, cp_mode):
    lib = _lib_or_skip()
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    cubin = _cubin(lib, compile_spec(spec, dim, False), cp_mode)
    path = "/tmp/gpb_jit_test_%s_%d.cubin" % (name, os.getpid())
    with open(path, "wb") as f:
        f.write(cubin)
    try:
        sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
        usage = subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True, check=True).stdout
    finally:
        os.remove(path)
    assert "gpb_spec_assemble" in sass and "gpb_spec_grad" in sass
    assert sass.count("LDL") == 0 and sass.count("STL") == 0, name
    assert "STACK:0" in usage and "STACK:8" not in usage, usage
    assert "DFMA" in sass


def test_c3_grammar_worst_case_compiles():
    """a depth-3, 3-ary MUL of PER leaves - legal in bench.py's C3 grammar - needs a gradient tape of 106
    entries and was refused by the interpreter (GPB_MAX_TAPE 64).  The specialised kernels have no tape."""
    lib = _lib_or_skip()
    cp = compile_spec(TERNARY_PER, 1, False)
    assert cp.tape == 106 and cp.n_hp == 54 and cp.n_ops == 53
    cubin = _cubin(lib, cp)
    assert len(cubin) > 10000

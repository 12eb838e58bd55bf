"""Dumps the kernels libgpb generates for a kernel program (csrc/jit.cu): CUDA source, SASS and resource usage, all on the
CPU (NVRTC cross-compiles for sm_100a).  usage: python tools/jit_dump.py <name> <out_prefix>   (name: c2 | c3big | ard8)"""
import ctypes
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gaussianprocessfundamentals_b200 import _lib  # noqa: E402
from gaussianprocessfundamentals_b200.program import compile_spec  # noqa: E402

COMPOSITE = ("MUL", [("ADD", [("SE",), ("PER",)]), ("LIN",)])


def program(name):
    if name == "c2":
        return compile_spec(COMPOSITE, 1, False)
    if name == "ard8":
        return compile_spec(("SE_ARD",), 8, False)
    if name == "c3big":
        import bench
        trees, _ = bench.candidate_trees(256)
        return max((compile_spec(t, 1, False) for t in trees), key=lambda c: c.n_ops)
    raise SystemExit("unknown program " + name)


def main():
    name, prefix = sys.argv[1], sys.argv[2]
    lib = _lib.load()
    cp = program(name)
    code = np.ascontiguousarray(cp.code, dtype=np.int32)
    ptr = code.ctypes.data_as(_lib.c_int32_p)
    need = ctypes.c_size_t()
    _lib.check(lib.gpb_jit_source(ptr, cp.n_ops, cp.dim, 1, None, 0, ctypes.byref(need)), "gpb_jit_source")
    buf = ctypes.create_string_buffer(need.value)
    _lib.check(lib.gpb_jit_source(ptr, cp.n_ops, cp.dim, 1, buf, need.value, ctypes.byref(need)), "gpb_jit_source")
    src = buf.value.decode()
    with open(prefix + "_generated.cu", "w") as f:
        f.write(src[src.index("namespace gpb {\nstruct Prog"):])       # the generated part (the rest is csrc/*.cuh verbatim)
    _lib.check(lib.gpb_jit_cubin(ptr, cp.n_ops, cp.dim, 1, b"sm_100a", None, 0, ctypes.byref(need)), "gpb_jit_cubin")
    cub = ctypes.create_string_buffer(need.value)
    _lib.check(lib.gpb_jit_cubin(ptr, cp.n_ops, cp.dim, 1, b"sm_100a", cub, need.value, ctypes.byref(need)), "gpb_jit_cubin")
    path = prefix + ".cubin"
    with open(path, "wb") as f:
        f.write(cub.raw[:need.value])
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    usage = subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True, check=True).stdout
    os.remove(path)
    lines = [l for l in sass.splitlines() if "/*" in l and not l.strip().startswith("/* 0x")]
    mnem = {}
    for l in lines:
        parts = l.split("*/", 1)[1].split()
        if not parts:
            continue
        op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
        op = op.split(".")[0].rstrip(";")
        mnem[op] = mnem.get(op, 0) + 1
    with open(prefix + "_sass_summary.txt", "w") as f:
        f.write("program %s: %d ops, %d hyper-parameters, dim %d\n" % (name, cp.n_ops, cp.n_hp, cp.dim))
        f.write(usage)
        f.write("\nlocal-memory instructions: LDL %d, STL %d\n" % (sass.count("LDL"), sass.count("STL")))
        f.write("instruction mix (both kernels): " + ", ".join("%s %d" % kv for kv in sorted(mnem.items(), key=lambda kv: -kv[1])) + "\n")
    with open(prefix + ".sass", "w") as f:
        f.write(sass)
    print(open(prefix + "_sass_summary.txt").read())


if __name__ == "__main__":
    main()

"""Developer probe for a B200 box: FP64 pipe rates, cuBLAS/cuSOLVER library rates, stage timings of the plan.
Writes one JSON object per line to gpurun_out/probe.jsonl.  Not part of the product or of bench.py."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussianprocessfundamentals_b200 import _lib, engine as eng  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "probe.jsonl")
os.makedirs(os.path.dirname(OUT), exist_ok=True)


def emit(**kw):
    print(json.dumps(kw), flush=True)
    with open(OUT, "a") as f:
        f.write(json.dumps(kw) + "\n")


def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))


def main():
    lib = _lib.load()
    dev = torch.device("cuda")
    emit(what="device", name=torch.cuda.get_device_name(0), sms=torch.cuda.get_device_properties(0).multi_processor_count)
    sp = eng._stream_ptr
    for blocks_per_sm in (1, 2, 4):
        blocks = 148 * blocks_per_sm
        iters = 20000
        t, _ = timed(lambda: lib.gpb_microbench(0, iters, blocks, sp()))
        emit(what="dmma_peak", blocks_per_sm=blocks_per_sm, ms=t, tflops=blocks * 8 * iters * 8 * 512 / (t * 1e-3) / 1e12)
        t, _ = timed(lambda: lib.gpb_microbench(1, iters, blocks, sp()))
        emit(what="dfma_peak", blocks_per_sm=blocks_per_sm, ms=t, tflops=blocks * 8 * iters * 16 * 64 / (t * 1e-3) / 1e12)
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    c = torch.empty(n, n, dtype=torch.float64, device=dev)
    t, tm = timed(lambda: torch.matmul(a, b, out=c))
    emit(what="cublas_dgemm_8192", ms=t, ms_median=tm, tflops=2 * n ** 3 / (t * 1e-3) / 1e12)
    for cfg in (2, 4):
        for akm, bkm in ((0, 0), (0, 1), (1, 1)):
            t, tm = timed(lambda: eng.gemm(akm | cfg, bkm, a, n, b, n, c, n, n, n, n, 1.0, 0.0))
            emit(what="gpb_gemm_8192", cfg={2: "big", 4: "half"}[cfg], akm=akm, bkm=bkm, ms=t, ms_median=tm,
                 tflops=2 * n ** 3 / (t * 1e-3) / 1e12)
        t, tm = timed(lambda: eng.gemm(cfg, 0, a, n, b, n, c, n, n, n, 128, -1.0, 1.0))
        emit(what="gpb_gemm_8192_k128_rmw", cfg={2: "big", 4: "half"}[cfg], ms=t, tflops=2 * n * n * 128 / (t * 1e-3) / 1e12)
    spd = a @ a.t() + n * torch.eye(n, dtype=torch.float64, device=dev)
    t, tm = timed(lambda: torch.linalg.cholesky(spd))
    emit(what="cusolver_potrf_8192", ms=t, tflops=n ** 3 / 3 / (t * 1e-3) / 1e12)
    del a, b, c, spd
    torch.cuda.empty_cache()

    tree = ("MUL", [("ADD", [("SE",), ("PER",)]), ("LIN",)])
    hp = np.array([0.1, 0.1, 0.1, 0.01])
    sizes = [int(v) for v in os.environ.get("PROBE_SIZES", "1000,2048,4096,8192,16384").split(",")]
    for n in sizes:
        rng = np.random.default_rng(1)
        x = np.linspace(0, 1, n)[:, None]
        y = x * np.sin(40 * x) + 0.1 * rng.standard_normal((n, 1))
        prog = eng.DeviceProgram.get(tree, 1, False, 1)
        plan = eng.Plan([prog], [n], want_grad=True)
        plan.set_data(0, torch.tensor(x), torch.tensor(y))
        plan.set_hp(0, hp, 1e-2)
        rec = {"what": "plan_stages", "n": n}
        for name, st in (("assemble", 1), ("potrf", 2), ("nll", 4), ("inverse", 8), ("grad", 16)):
            if st in (1, 2):
                # potrf destroys K: re-assemble before every timed factorisation
                def run(st=st):
                    if st == 2:
                        plan.eval(1)
                    plan.eval(st)
                t_all, _ = timed(run, reps=3)
                if st == 2:
                    t_all -= rec["assemble_ms"]
                rec[name + "_ms"] = t_all
            else:
                plan.eval(1); plan.eval(2); plan.eval(4)
                if st == 16:
                    plan.eval(8)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); plan.eval(st); e1.record(); torch.cuda.synchronize()
                rec[name + "_ms"] = e0.elapsed_time(e1)
        t, tm = timed(lambda: plan.eval(eng.STAGES_LML_GRAD), reps=3)
        rec["full_ms"] = t
        rec["potrf_tflops"] = n ** 3 / 3 / (rec["potrf_ms"] * 1e-3) / 1e12
        rec["inverse_tflops"] = 2 * n ** 3 / 3 / (rec["inverse_ms"] * 1e-3) / 1e12
        l0 = eng.launch_count(); plan.eval(eng.STAGES_LML_GRAD); rec["launches"] = eng.launch_count() - l0
        nll, grads, info = plan.results()
        rec["nll"] = float(nll[0]); rec["info"] = int(info[0])
        emit(**rec)
        del plan
        torch.cuda.empty_cache()

    # batched: B x n
    for B, n in ((64, 1024), (32, 2048)):
        prog = eng.DeviceProgram.get(tree, 1, False, 1)
        plan = eng.Plan([prog] * B, [n] * B, want_grad=True)
        x = np.linspace(0, 1, n)[:, None]
        for b in range(B):
            rng = np.random.default_rng(b)
            y = x * np.sin(40 * x) + 0.1 * rng.standard_normal((n, 1))
            plan.set_data(b, torch.tensor(x), torch.tensor(y))
            plan.set_hp(b, hp, 1e-2)
        t, tm = timed(lambda: plan.eval(eng.STAGES_LML_GRAD), reps=3)
        emit(what="batched", B=B, n=n, ms=t, gp_evals_per_s=B / (t * 1e-3), tflops=B * n ** 3 / (t * 1e-3) / 1e12)
        del plan
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()

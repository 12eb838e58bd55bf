"""Benchmark of the exact-GP likelihood hot path (BASELINE.json: "LML+grad evals/sec at n=8k/32k FP64; FP64 TC % of
peak in Cholesky").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload c2|m32k|c1|c3|c4|c5|c5s|c5g|m32kd] [--grid PxQ]

A step = one full LML + gradient evaluation (assembly -> Cholesky with carried y -> NLL -> inverse -> trace gradient)
of the workload; default workload C2 = composite (SE+PER)xLIN kernel, n = 8192, d = 1.
  value  : evaluations / s with X, y, theta resident in HBM (CUDA events around K steps, max over ranks)
  e2e    : the same through the host-buffer C-ABI call gpb_plan_eval_host: X, y, theta copied H2D from pinned memory and
           NLL + gradient + info copied back every step
  N > 1  : C2 / M32k / C1 do not shard -> "replicas only": every rank evaluates its own replica, no collective
           (DESIGN.md); C3 / C4 (batched candidates / partition blocks) shard their GPs across ranks, no data-path
           collective either.  value = GP evaluations of all ranks / max-over-ranks time.
           C5 / C5s (LML) and C5g / M32kd (LML + gradient) evaluate ONE matrix on all N GPUs: distributed Cholesky,
           inverse and gradient with NCCL panel broadcasts (strong scaling; N = 1 is the single-GPU plan).
--impl reference times the CPU restatement of the reference's unfused op sequence (oracle/gp_oracle.py, torch CPU,
all host threads) on the same workload; TensorFlow - the reference's own engine - is not installable offline.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

COMPOSITE = ("MUL", [("ADD", [("SE",), ("PER",)]), ("LIN",)])
METRIC = "LML+grad evals/sec"
WORKLOADS = {
    "c1": dict(name="C1: SE, n=1000, d=1", n=1000, B=1, tree=("SE",), hp=[0.1]),
    "c2": dict(name="C2: (SE+PER)xLIN, n=8192, d=1", n=8192, B=1, tree=COMPOSITE, hp=[0.1, 0.1, 0.1, 0.01]),
    "m32k": dict(name="M32k: (SE+PER)xLIN, n=32768, d=1", n=32768, B=1, tree=COMPOSITE, hp=[0.1, 0.1, 0.1, 0.01]),
    "c3": dict(name="C3: 256 candidate kernels x n=2048", n=2048, B=256, tree=None, hp=None),
    "c4": dict(name="C4: 1024 partition blocks x n=1024 (SE)", n=1024, B=1024, tree=("SE",), hp=None),
    # one large GP, likelihood only, strong scaling: distributed Cholesky over a P x Q grid when N > 1
    "c5": dict(name="C5: SE-ARD, n=65536, d=8, LML (distributed Cholesky, NCCL panel broadcasts)", n=65536, B=1,
               tree=("SE_ARD",), hp=None, d=8),
    "c5s": dict(name="C5 at n=32768: SE-ARD, d=8, LML (distributed Cholesky, NCCL panel broadcasts)", n=32768, B=1,
                tree=("SE_ARD",), hp=None, d=8),
    # the same with the gradient stages (distributed inverse + trace gradient), strong scaling
    "c5g": dict(name="C5 + gradient: SE-ARD, n=65536, d=8, LML+grad (distributed Cholesky, inverse, trace gradient)",
                n=65536, B=1, tree=("SE_ARD",), hp=None, d=8, grad=True),
    "m32kd": dict(name="M32k distributed: SE-ARD, n=32768, d=8, LML+grad over N GPUs (strong scaling)", n=32768, B=1,
                  tree=("SE_ARD",), hp=None, d=8, grad=True),
}
DIST_WORKLOADS = ("c5", "c5s", "c5g", "m32kd")


# ---- synthetic data --------------------------------------------------------------------------------------
def make_xy(n, seed):
    rng = np.random.default_rng(seed)
    x = np.linspace(0.0, 1.0, n)[:, None]
    y = x * np.sin(40 * x) + 0.1 * rng.standard_normal((n, 1))
    return x, y


def candidate_trees(B, seed=2):
    """random compositional kernels from the grammar {SE, PER, LIN} x {ADD, MUL}, depth <= 3 (config C3)"""
    rng = np.random.default_rng(seed)

    def gen(depth):
        if depth == 0 or rng.uniform() < 0.3:
            return (["SE", "PER", "LIN"][rng.integers(3)],)
        k = 2 + int(rng.integers(2))
        return (["ADD", "MUL"][rng.integers(2)], [gen(depth - 1) for _ in range(k)])

    def hp_of(t):
        if t[0] == "SE":
            return [abs(0.1 + 0.05 * rng.standard_normal())]
        if t[0] == "PER":
            return [abs(0.3 + 0.1 * rng.standard_normal()), rng.uniform(0.05, 0.5)]
        if t[0] == "LIN":
            return [rng.uniform(-1.0, 2.0)]
        out = []
        for c in t[1]:
            out += hp_of(c)
        return out

    trees = [gen(3) for _ in range(B)]
    return trees, [np.asarray(hp_of(t), dtype=np.float64) for t in trees]


def build_workload(key, rank, world):
    """(trees, flat hps, ns, xs, ys) of the GPs this rank evaluates"""
    w = WORKLOADS[key]
    n = w["n"]
    if key == "c3":
        trees, hps = candidate_trees(w["B"])
        mine = list(range(rank, w["B"], world))
        x, _ = make_xy(n, 2)
        ys = [make_xy(n, 1000 + b)[1] for b in mine]
        return [trees[b] for b in mine], [hps[b] for b in mine], [n] * len(mine), [x] * len(mine), ys
    if key == "c4":
        mine = list(range(rank, w["B"], world))
        N = w["B"] * n
        xs, ys, hps = [], [], []
        rng = np.random.default_rng(3)
        ls = rng.uniform(0.2, 1.0, w["B"]) / w["B"]
        for b in mine:
            xb = (np.arange(b * n, (b + 1) * n) / N)[:, None]
            r = np.random.default_rng(3000 + b)
            xs.append(xb)
            ys.append(np.sin((50 + b % 7) * 40 * xb) + 0.1 * r.standard_normal((n, 1)))
            hps.append(np.asarray([ls[b]]))
        return [w["tree"]] * len(mine), hps, [n] * len(mine), xs, ys
    x, y = make_xy(n, 1 + rank)
    return [w["tree"]], [np.asarray(w["hp"], dtype=np.float64)], [n], [x], [y]


# ---- clocks --------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            p = [v.strip() for v in s.split(",")]
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
                for nm, v in zip(names, p[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---- CPU baseline (reference-equivalent unfused op sequence) ------------------------------------------------------------
def cpu_eval_seconds(tree, hp_flat, n, seed=1):
    from gaussianprocessfundamentals_b200.program import compile_spec
    from oracle import gp_oracle as orc
    x, y = make_xy(n, seed)
    flat = np.asarray(hp_flat, dtype=np.float64)
    hp = [flat[o] if s == 1 else flat[o:o + s] for o, s in compile_spec(tree, 1, False).entries]
    t0 = time.perf_counter()
    orc.nll_and_grad(tree, hp, 1e-2, x, y, reference_distance=True)
    return time.perf_counter() - t0


def cpu_baseline(key, budget_s=25.0):
    """evals/s of the CPU port on the workload, from a bounded sample: two sizes are timed and t(n) = a n^2 + b n^3 is
    extrapolated when the full size does not fit the budget"""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    w = WORKLOADS[key]
    tree = w["tree"] if w["tree"] is not None else COMPOSITE
    hp = w["hp"] if w["hp"] is not None else ([0.1, 0.1, 0.1, 0.01] if tree is COMPOSITE else [0.1])
    n = w["n"]
    cpu_eval_seconds(tree, hp, 256)  # warm-up of the thread pools
    n1 = min(n, 2048)
    t1 = cpu_eval_seconds(tree, hp, n1)
    if n1 == n:
        per_eval, sample = t1, "1 full eval at n=%d" % n
    else:
        n2 = min(n, 4096)
        t2 = cpu_eval_seconds(tree, hp, n2)
        if n2 == n:
            per_eval, sample = t2, "1 full eval at n=%d" % n
        else:
            A = np.array([[n1 ** 2, n1 ** 3], [n2 ** 2, n2 ** 3]], dtype=np.float64)
            a, b = np.linalg.solve(A, np.array([t1, t2]))
            if a < 0 or b < 0:
                a, b = 0.0, t2 / n2 ** 3
            per_eval = a * n ** 2 + b * n ** 3
            sample = "evals at n=%d (%.2fs) and n=%d (%.2fs), t = a n^2 + b n^3 extrapolated to n=%d" % (n1, t1, n2, t2, n)
    return {"value": 1.0 / per_eval, "unit": "evals/s", "cores": cores,
            "kind": "port", "sample": sample, "seconds_per_eval": per_eval}


# ---- C5: one large GP over a process grid (strong scaling) -----------------------------------------------------------------
def make_c5(n, d, seed=4):
    rng = np.random.default_rng(seed)
    x = rng.uniform(0.0, 1.0, size=(n, d))
    y = np.sum(np.sin(3 * x), axis=1, keepdims=True) + 0.1 * rng.standard_normal((n, 1))
    ell = np.linspace(0.5, 1.2, d)
    return x, y, ell


def run_c5(args, rank, world, local_rank):
    """LML (workloads with grad=True: LML + gradient) of one n-point SE-ARD GP: N = 1 single-GPU plan, N > 1 distributed
    plan.  Strong scaling: the work is fixed, value = evaluations / s of the whole job."""
    import torch
    import torch.distributed as dist
    from gaussianprocessfundamentals_b200 import _lib, engine as eng
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    w = WORKLOADS[args.workload]
    n, d = w["n"], w["d"]
    want_grad = bool(w.get("grad", False))
    STAGES = eng.STAGES_LML_GRAD if want_grad else eng.STAGES_LML
    x, y, ell = make_c5(n, d)
    prog = eng.DeviceProgram.get(w["tree"], d, False, 1)
    grid = None
    if world > 1:
        P, Q = (args.grid if args.grid else eng.ProcessGrid.default_shape(world))
        grid = eng.ProcessGrid(P, Q)
    plan = eng.Plan([prog], [n], want_grad=want_grad, grid=grid)
    plan.set_data(0, torch.tensor(x), torch.tensor(y))
    plan.set_hp(0, ell, 1e-2)

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    l0 = eng.launch_count()
    plan.eval(STAGES)
    torch.cuda.synchronize()
    launches = eng.launch_count() - l0
    stage_ms = {}
    names = ["assemble", "potrf", "nll"] + (["trtri", "lauum", "grad"] if want_grad else [])
    bits = [eng.STAGE_ASSEMBLE, eng.STAGE_POTRF, eng.STAGE_NLL] + \
        ([eng.STAGE_TRTRI, eng.STAGE_LAUUM, eng.STAGE_GRAD] if want_grad else [])
    marks = [ev() for _ in range(len(bits) + 1)]
    barrier()
    marks[0].record()
    for i, bit in enumerate(bits):
        plan.eval(bit); marks[i + 1].record()
    barrier()
    for i, name in enumerate(names):
        stage_ms[name] = marks[i].elapsed_time(marks[i + 1])
    lib = _lib.load()
    best = 1e9
    for _ in range(3):
        e0, e1 = ev(), ev()
        e0.record(); lib.gpb_microbench(0, 20000, 148 * 4, eng._stream_ptr()); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    dmma_peak = 148 * 4 * 8 * 20000 * 8 * 512 / (best * 1e-3) / 1e12
    for _ in range(args.warmup):
        plan.eval(STAGES)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(args.steps):
        plan.eval(STAGES)
    e1.record()
    barrier()
    elapsed_ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    nll, _, info = plan.results()
    # e2e: host buffers in, scalars out, every step
    hx, hy = torch.tensor(x).pin_memory().numpy(), torch.tensor(y.reshape(-1)).pin_memory().numpy()
    plan.eval_host([ell], [1e-2], [hx], [hy], stages=STAGES)
    barrier()
    t0 = time.perf_counter()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(args.steps):
        plan.eval_host([ell], [1e-2], [hx], [hy], stages=STAGES)
    e1.record()
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)
    if world > 1:
        keys = sorted(stage_ms)
        t = torch.tensor([elapsed_ms, e2e_ms] + [stage_ms[k_] for k_ in keys], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, e2e_ms = float(t[0]), float(t[1])
        for i, k_ in enumerate(keys):
            stage_ms[k_] = float(t[2 + i])
    if rank == 0:
        value = args.steps / (elapsed_ms * 1e-3)
        flops = n ** 3 / 3.0
        tf = flops / (stage_ms["potrf"] * 1e-3) / 1e12
        line = {
            "metric": METRIC if want_grad else "LML evals/sec (likelihood only; distributed Cholesky)", "value": value,
            "unit": "evals/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["name"], "noise": 1e-2,
                       "grid": [grid.P, grid.Q] if grid else [1, 1],
                       "multi_gpu": "2D block-cyclic block ownership, NCCL panel broadcasts" if grid else "single GPU",
                       "l2": "working set (%.1f GiB) exceeds the 126 MB L2; no explicit flush" % (8.0 * n * n / 2 ** 30)},
            "e2e": {"value": args.steps / (e2e_ms * 1e-3), "unit": "evals/s",
                    "h2d_bytes_per_step": int(hx.nbytes + hy.nbytes + ell.nbytes + 8),
                    "d2h_bytes_per_step": 8 + 4 + (8 * (d + 1) if want_grad else 0),
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches * args.steps), "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "Cholesky (gemm_kernel trailing updates + panels + diagonal blocks), "
                         "whole job", "achieved": tf, "peak": dmma_peak * world, "unit": "TFLOP/s",
                         "frac": tf / (dmma_peak * world), "traffic": None,
                         "peak_source": "measured live on rank 0: DMMA.8x8x4 probe x n_gpus"},
            "stages_ms": stage_ms,
            "cholesky": {"tflops": tf, "frac_of_fp64_tensor_peak": tf / (dmma_peak * world)},
            "check": {"nll0": float(nll[0]), "info_max": int(np.max(info))},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---- main ---------------------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    times = []
    base = None
    for i in range(args.warmup + args.steps):
        base = cpu_baseline(args.workload)
        if i >= args.warmup:
            times.append(base["seconds_per_eval"])
    per_eval = float(np.mean(times))
    value = 1.0 / per_eval
    line = {"metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": per_eval * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": w["name"], "note": "CPU port of the reference's unfused op sequence (oracle), torch CPU "
                       "float64; each step is a bounded sample extrapolated to the workload size"},
            "cpu_baseline": {"value": value, "unit": "evals/s", "cores": base["cores"], "kind": "port",
                             "sample": base["sample"]},
            "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--grid", type=lambda v: tuple(int(t) for t in v.split("x")), default=None,
                    help="process grid PxQ of the distributed workloads (default: ProcessGrid.default_shape)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    if args.workload in DIST_WORKLOADS:
        run_c5(args, rank, world, local_rank)
        return

    import torch
    import torch.distributed as dist
    from gaussianprocessfundamentals_b200 import _lib, engine as eng

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    w = WORKLOADS[args.workload]
    trees, hps, ns, xs, ys = build_workload(args.workload, rank, world)
    progs = [eng.DeviceProgram.get(t, 1, False, 1) for t in trees]
    plan = eng.Plan(progs, ns, want_grad=True)
    for b in range(len(ns)):
        plan.set_data(b, torch.tensor(xs[b]), torch.tensor(ys[b]))
        plan.set_hp(b, hps[b], 1e-2)
    n = w["n"]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # launches of one evaluation (counted eagerly: a graph replay does not pass through the host-side counter)
    l0 = eng.launch_count()
    plan.eval(eng.STAGES_LML_GRAD)
    torch.cuda.synchronize()
    launches_per_eval = eng.launch_count() - l0

    # ---- stage timings (one evaluation, CUDA events on the launching stream) -------------------------------------------
    def ev():
        return torch.cuda.Event(enable_timing=True)
    stage_ms = {}
    for _ in range(2):
        marks = [ev() for _ in range(7)]
        marks[0].record()
        plan.eval(eng.STAGE_ASSEMBLE); marks[1].record()
        plan.eval(eng.STAGE_POTRF); marks[2].record()
        plan.eval(eng.STAGE_NLL); marks[3].record()
        plan.eval(eng.STAGE_TRTRI); marks[4].record()
        plan.eval(eng.STAGE_LAUUM); marks[5].record()      # ONE launch: Kinv = W^T W, the roofline kernel
        plan.eval(eng.STAGE_GRAD); marks[6].record()
        torch.cuda.synchronize()
        for i, name in enumerate(["assemble", "potrf", "nll", "trtri", "lauum", "grad"]):
            stage_ms[name] = marks[i].elapsed_time(marks[i + 1])
        stage_ms["inverse"] = stage_ms["trtri"] + stage_ms["lauum"]

    # ---- FP64 peak, measured live (MEASURED_PEAKS.json carries no FP64 entry) --------------------------------------------
    lib = _lib.load()
    iters, blocks = 20000, 148 * 4
    best = 1e9
    for _ in range(3):
        e0, e1 = ev(), ev()
        e0.record(); lib.gpb_microbench(0, iters, blocks, eng._stream_ptr()); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    dmma_peak = blocks * 8 * iters * 8 * 512 / (best * 1e-3) / 1e12
    a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    c = torch.empty_like(a)
    best = 1e9
    for _ in range(3):
        e0, e1 = ev(), ev()
        e0.record(); torch.matmul(a, a, out=c); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    cublas_dgemm = 2 * 8192 ** 3 / (best * 1e-3) / 1e12
    del a, c

    # ---- value: K evaluations, inputs resident (CUDA graph of the whole evaluation when capture works) ------------------
    graph = None
    if not args.no_graph:
        try:
            g = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                with torch.cuda.graph(g, stream=s):
                    plan.eval(eng.STAGES_LML_GRAD)
            torch.cuda.current_stream().wait_stream(s)
            g.replay(); torch.cuda.synchronize()
            graph = g
        except Exception as exc:  # capture of the multi-stream look-ahead is best effort
            sys.stderr.write("graph capture unavailable (%r); timing eager launches\n" % (exc,))
            graph = None
            torch.cuda.synchronize()

    def step():
        if graph is not None:
            graph.replay()
        else:
            plan.eval(eng.STAGES_LML_GRAD)

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    elapsed_ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    nll, grads, info = plan.results()

    # ---- e2e: host buffers in pinned memory, H2D + kernels + D2H + sync inside every step --------------------------------
    pin_x = [torch.tensor(x).pin_memory() for x in xs]
    pin_y = [torch.tensor(y.reshape(-1)).pin_memory() for y in ys]
    hx, hy = [t.numpy() for t in pin_x], [t.numpy() for t in pin_y]
    noises = [1e-2] * len(ns)
    for _ in range(2):
        plan.eval_host(hps, noises, hx, hy)
    barrier()
    t0 = time.perf_counter()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(args.steps):
        nll_h, grads_h, info_h = plan.eval_host(hps, noises, hx, hy)
    e1.record()
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)
    h2d = sum(x.nbytes + y.nbytes for x, y in zip(hx, hy)) + sum(h.nbytes for h in hps) + 8 * len(ns)
    d2h = 8 * len(ns) + sum((h.size + 1) * 8 for h in hps) + 4 * len(ns)

    if world > 1:
        t = torch.tensor([elapsed_ms, e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, e2e_ms = float(t[0]), float(t[1])
        cnt = torch.tensor([len(ns)], dtype=torch.float64, device="cuda")
        dist.all_reduce(cnt)
        total_gps = int(cnt[0])
    else:
        total_gps = len(ns)

    if rank == 0:
        evals = total_gps * args.steps
        value = evals / (elapsed_ms * 1e-3)
        flops_potrf, flops_inv = n ** 3 / 3.0 * len(ns), 2.0 * n ** 3 / 3.0 * len(ns)
        tf_potrf = flops_potrf / (stage_ms["potrf"] * 1e-3) / 1e12
        tf_inv = flops_inv / (stage_ms["inverse"] * 1e-3) / 1e12
        tf_all = (flops_potrf + flops_inv) / ((stage_ms["potrf"] + stage_ms["inverse"]) * 1e-3) / 1e12
        flops_lauum = n ** 3 / 3.0 * len(ns)
        tf_lauum = flops_lauum / (stage_ms["lauum"] * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get(args.workload)
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["name"], "gps_per_step": total_gps, "noise": 1e-2,
                       "multi_gpu": "replicas only" if w["B"] == 1 else "GPs sharded across ranks, no collective",
                       "l2": "working set (K and K^-1, %.1f GiB per GP) exceeds the 126 MB L2; no explicit flush"
                             % (2 * 8.0 * n * n / 2 ** 30),
                       "cuda_graph": graph is not None},
            "e2e": {"value": evals / (e2e_ms * 1e-3), "unit": "evals/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches_per_eval * args.steps),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "gemm_kernel<CfgHalf,1,1,GeoLauum> (FP64 DMMA mainloop; the single "
                         "largest launch of an evaluation: K^-1 = W^T W, lower tiles)", "achieved": tf_lauum,
                         "peak": dmma_peak, "unit": "TFLOP/s", "frac": tf_lauum / dmma_peak, "traffic": traffic,
                         "flops_per_launch": flops_lauum, "launch_ms": stage_ms["lauum"],
                         "peak_source": "measured live: DMMA.8x8x4 register-operand probe (gpb_microbench); "
                                        "MEASURED_PEAKS.json has no FP64 entry; cuBLAS DGEMM 8192^3 = %.2f TFLOP/s"
                                        % cublas_dgemm,
                         "all_gemm_launches": {"achieved": tf_all, "frac": tf_all / dmma_peak,
                                               "flops_per_eval": flops_potrf + flops_inv}},
            "stages_ms": stage_ms,
            "cholesky": {"tflops": tf_potrf, "frac_of_fp64_tensor_peak": tf_potrf / dmma_peak,
                         "inverse_tflops": tf_inv, "inverse_frac": tf_inv / dmma_peak},
            "check": {"nll0": float(nll[0]), "info_max": int(np.max(info))},
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(args.workload)
            line["cpu_baseline"].pop("seconds_per_eval", None)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

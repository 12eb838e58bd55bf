"""Benchmark of the exact-GP likelihood hot path (BASELINE.json: "LML+grad evals/sec at n=8k/32k FP64; FP64 TC % of
peak in Cholesky").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload m32kd|c2|m32k|c1|c3|c4|c5|c5s|c5g] [--grid PxQ] [--no-sub-records]

A step = one full LML + gradient evaluation (assembly -> Cholesky with carried y -> NLL -> inverse -> trace gradient).

Default workload (every N): M32kd = ONE n = 32768 SE-ARD GP, LML + gradient - the metric's n = 32k and the path that
really shards with a data-plane exchange: N = 1 is the single-GPU plan, N > 1 the distributed plan (2D block-cyclic block
ownership, NCCL panel broadcasts; strong scaling, value = evaluations / s of the whole job).  The same JSON line carries
`sub_records`: C2 (n = 8192, the metric's n = 8k, N = 1 only), C1 (N = 1 only), and the independent-GP workloads C3
(256 candidate kernels x n = 2048) and C4 (1024 partition blocks x n = 1024), which shard their GPs across the ranks
with no data-path collective - so one driver run measures 8k and 32k, and the 1 -> 8 curves of both sharding kinds.
  value  : evaluations / s with X, y, theta resident in HBM (CUDA events around K steps, max over ranks)
  e2e    : the same through the host-buffer C-ABI call gpb_plan_eval_host: X, y, theta copied H2D from pinned memory and
           NLL + gradient + info copied back every step
--impl reference times the CPU restatement of the reference's unfused op sequence (oracle/gp_oracle.py, torch CPU,
all host threads): C2 (n = 8192) IN FULL, no extrapolation; for n = 32768 (autodiff tape ~ 20 n^2 doubles = 170 GB) each
step is a full n = 8192 evaluation of the same kernel scaled by (32768 / 8192)^3, labelled as such.  TensorFlow - the
reference's own engine - is not installable offline and /root/reference is not on the GPU box.
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
DEFAULT_WORKLOAD = "m32kd"


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
def _cpu_inputs(key, n):
    w = WORKLOADS[key]
    if key in DIST_WORKLOADS:
        x, y, ell = make_c5(n, w["d"])
        return w["tree"], [ell], x, y
    tree = w["tree"] if w["tree"] is not None else COMPOSITE
    flat = np.asarray(w["hp"] if w["hp"] is not None else ([0.1, 0.1, 0.1, 0.01] if tree is COMPOSITE else [0.1]),
                      dtype=np.float64)
    from gaussianprocessfundamentals_b200.program import compile_spec
    hp = [flat[o] if s == 1 else flat[o:o + s] for o, s in compile_spec(tree, 1, False).entries]
    x, y = make_xy(n, 1)
    return tree, hp, x, y


def cpu_eval_seconds(key, n):
    """one LML + gradient evaluation of workload `key` at size n on the host: the oracle's unfused op sequence with
    autodiff through the Cholesky (what the reference's TF path does), all host threads"""
    from oracle import gp_oracle as orc
    tree, hp, x, y = _cpu_inputs(key, n)
    t0 = time.perf_counter()
    orc.nll_and_grad(tree, hp, 1e-2, x, y, reference_distance=(x.shape[1] == 1))
    return time.perf_counter() - t0


CPU_FULL_MAX_N = 8192      # largest size the CPU port evaluates IN FULL (n = 8192: ~5 s and ~15 GB of autodiff tape)


def cpu_step_seconds(key):
    """(seconds per workload step, description of what was timed).  Workloads up to n = 8192 with one GP are timed in
    full.  Larger single matrices: one full evaluation at n = 8192 of the same kernel, scaled by (n / 8192)^3 (the n^3
    terms dominate beyond 8192; the tape of a full n = 32768 evaluation would need ~170 GB).  Batches of independent GPs
    (C3 / C4): one GP of the batch in full, times the batch size."""
    w = WORKLOADS[key]
    n, B = w["n"], w["B"]
    if n <= CPU_FULL_MAX_N:
        t = cpu_eval_seconds(key, n)
        if B == 1:
            return t, "1 full evaluation at n=%d (no extrapolation)" % n
        return t * B, "1 of the %d GPs evaluated in full at n=%d (%.3f s), times %d" % (B, n, t, B)
    t = cpu_eval_seconds(key, CPU_FULL_MAX_N)
    f = (n / CPU_FULL_MAX_N) ** 3
    return t * f, "1 full evaluation at n=%d (%.2f s) scaled by (%d/%d)^3 = %g" % (CPU_FULL_MAX_N, t, n, CPU_FULL_MAX_N, f)


def cpu_baseline(key, repeats=3):
    """evals/s of the CPU port on the workload: median of `repeats` timed steps after one warm-up of the thread pools"""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cpu_eval_seconds("c1", 256)
    samples, what = [], ""
    for _ in range(repeats):
        t, what = cpu_step_seconds(key)
        samples.append(t)
    per_step = float(np.median(samples))
    gps = WORKLOADS[key]["B"]
    return {"value": gps / per_step, "unit": "evals/s", "cores": cores, "kind": "port",
            "sample": "median of %d: %s" % (repeats, what), "seconds_per_step": per_step,
            "tensorflow_importable": _tensorflow_importable()}


def _tensorflow_importable():
    """if the real TensorFlow were on the box the reference itself would be timed.  It is not in
    the image, and the reference's sources (/root/reference) do not travel to the GPU box either - recorded, not hidden."""
    try:
        import importlib.util
        return importlib.util.find_spec("tensorflow") is not None and os.path.isdir("/root/reference/main/gpbasics")
    except Exception:
        return False


# ---- shared device-side helpers ----------------------------------------------------------------------------------------
class Ctx:
    def __init__(self, args, rank, world, local_rank):
        import torch
        self.args, self.rank, self.world, self.local_rank = args, rank, world, local_rank
        self.torch = torch
        self._peak = None
        self._cublas = None

    def ev(self):
        return self.torch.cuda.Event(enable_timing=True)

    def barrier(self):
        import torch.distributed as dist
        self.torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        import torch.distributed as dist
        if self.world == 1:
            return [float(v) for v in values]
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def sum_over_ranks(self, value):
        import torch.distributed as dist
        if self.world == 1:
            return float(value)
        t = self.torch.tensor([value], dtype=self.torch.float64, device="cuda")
        dist.all_reduce(t)
        return float(t[0])

    def dmma_peak(self):
        """FP64 tensor peak of this GPU, measured live (MEASURED_PEAKS.json carries no FP64 entry): DMMA.8x8x4 register
        probe, best of 3"""
        if self._peak is None:
            from gaussianprocessfundamentals_b200 import _lib, engine as eng
            lib = _lib.load()
            iters, blocks = 20000, 148 * 4
            best = 1e9
            for _ in range(3):
                e0, e1 = self.ev(), self.ev()
                e0.record(); lib.gpb_microbench(0, iters, blocks, eng._stream_ptr()); e1.record()
                self.torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            self._peak = blocks * 8 * iters * 8 * 512 / (best * 1e-3) / 1e12
        return self._peak

    def cublas_dgemm(self):
        if self._cublas is None:
            torch = self.torch
            a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
            c = torch.empty_like(a)
            best = 1e9
            for _ in range(3):
                e0, e1 = self.ev(), self.ev()
                e0.record(); torch.matmul(a, a, out=c); e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            self._cublas = 2 * 8192 ** 3 / (best * 1e-3) / 1e12
            del a, c
        return self._cublas


def _traffic(key):
    """ncu-measured DRAM bytes of the roofline launch (profiles/roofline_traffic.json: a STATIC number from a committed
    capture, not measured in this run), or None"""
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        return json.load(open(tpath)).get(key)
    except Exception:
        return None


# ---- one large GP over a process grid (strong scaling): C5 / M32kd ----------------------------------------------------------
def make_c5(n, d, seed=4):
    rng = np.random.default_rng(seed)
    x = rng.uniform(0.0, 1.0, size=(n, d))
    y = np.sum(np.sin(3 * x), axis=1, keepdims=True) + 0.1 * rng.standard_normal((n, 1))
    ell = np.linspace(0.5, 1.2, d)
    return x, y, ell


def dist_config(key, world, grid_shape=None):
    """`config` of a one-large-GP workload; shared by both arms so that the driver sees the same configuration"""
    w = WORKLOADS[key]
    n = w["n"]
    P, Q = grid_shape if grid_shape else (1, world)
    return {"workload": w["name"], "noise": 1e-2, "grid": [P, Q],
            "multi_gpu": "2D block-cyclic block ownership, NCCL panel broadcasts" if world > 1 else "single GPU",
            "l2": "working set (%.1f GiB) exceeds the 126 MB L2; no explicit flush" % (8.0 * n * n / 2 ** 30)}


def batch_config(key, total_gps, used_graph=True):
    w = WORKLOADS[key]
    n = w["n"]
    ws_mb = 2 * 8.0 * n * n * w["B"] / 2 ** 20
    l2 = ("working set (K and K^-1 of all GPs, %.0f MiB) exceeds the 126 MB L2; no explicit flush" % ws_mb) if ws_mb > 126 \
        else ("working set (%.0f MiB) fits the 126 MB L2: every step rebuilds K from X and y (16 n bytes), nothing is "
              "reused across steps; no explicit flush" % ws_mb)
    return {"workload": w["name"], "gps_per_step": total_gps, "noise": 1e-2,
            "multi_gpu": "replicas only" if w["B"] == 1 else "GPs sharded across ranks, no data-path collective",
            "l2": l2, "cuda_graph": used_graph}


def run_dist(ctx, key, steps, warmup):
    """LML (workloads with grad=True: LML + gradient) of one n-point SE-ARD GP: N = 1 single-GPU plan, N > 1 distributed
    plan.  Strong scaling: the work is fixed, value = evaluations / s of the whole job.  Returns the record (rank 0)."""
    import torch
    from gaussianprocessfundamentals_b200 import engine as eng
    args, rank, world = ctx.args, ctx.rank, ctx.world
    w = WORKLOADS[key]
    n, d = (args.size if args.size else w["n"]), w["d"]
    want_grad = bool(w.get("grad", False))
    STAGES = eng.STAGES_LML_GRAD if want_grad else eng.STAGES_LML
    x, y, ell = make_c5(n, d)
    prog = eng.DeviceProgram.get(w["tree"], d, False, 1)
    grid = None
    if world > 1:
        P, Q = (args.grid if args.grid else eng.ProcessGrid.default_shape(world))
        grid = eng.ProcessGrid(P, Q)
    storage = args.storage if (world > 1 and not want_grad) else "replicated"
    plan = eng.Plan([prog], [n], want_grad=want_grad, grid=grid, storage=storage)
    plan.set_data(0, torch.tensor(x), torch.tensor(y))
    plan.set_hp(0, ell, 1e-2)
    ev, barrier = ctx.ev, ctx.barrier

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
    dmma_peak = ctx.dmma_peak()
    for _ in range(warmup):
        plan.eval(STAGES)
    sampler = ClockSampler(ctx.local_rank)
    barrier()
    sampler.start()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(steps):
        plan.eval(STAGES)
    e1.record()
    barrier()
    elapsed_ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    nll, grads, info = plan.results()
    # e2e: host buffers in, scalars out, every step
    hx, hy = torch.tensor(x).pin_memory().numpy(), torch.tensor(y.reshape(-1)).pin_memory().numpy()
    plan.eval_host([ell], [1e-2], [hx], [hy], stages=STAGES)
    barrier()
    t0 = time.perf_counter()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(steps):
        plan.eval_host([ell], [1e-2], [hx], [hy], stages=STAGES)
    e1.record()
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)
    keys = sorted(stage_ms)
    red = ctx.max_over_ranks([elapsed_ms, e2e_ms] + [stage_ms[k_] for k_ in keys])
    elapsed_ms, e2e_ms = red[0], red[1]
    for i, k_ in enumerate(keys):
        stage_ms[k_] = red[2 + i]
    ws_gib = round(plan.ws_bytes / 2 ** 30, 2)
    del plan
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    value = steps / (elapsed_ms * 1e-3)
    tf = n ** 3 / 3.0 / (stage_ms["potrf"] * 1e-3) / 1e12
    peak = dmma_peak * world
    rec = {
        "metric": METRIC if want_grad else "LML evals/sec (likelihood only; distributed Cholesky)", "value": value,
        "unit": "evals/s",
        "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": elapsed_ms / steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(dist_config(key, world, (grid.P, grid.Q) if grid else None), n=n, storage=storage,
                       workspace_gib_per_rank=ws_gib),
        "e2e": {"value": steps / (e2e_ms * 1e-3), "unit": "evals/s",
                "h2d_bytes_per_step": int(hx.nbytes + hy.nbytes + ell.nbytes + 8),
                "d2h_bytes_per_step": 8 + 4 + 16 + (8 * (d + 1) if want_grad else 0),
                "ms_per_step": e2e_ms / steps},
        "gpu_launches": int(launches * steps), "clocks": clocks,
        # the DOMINANT stage of the evaluation at this size: the Cholesky factorisation (all its launches: trailing-update
        # gemm_kernel<GeoSyrk>, panel products, diagonal blocks), n^3 / 3 flops over the stage's CUDA-event time
        "roofline": {"bound": "tensor", "kernel": "Cholesky factorisation stage (gemm_kernel<GeoSyrk> trailing updates, "
                     "panel products, diagonal blocks; %d GPU%s)" % (world, "" if world == 1 else "s"),
                     "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak,
                     "traffic": _traffic(key + "_syrk_bulk_first") if world == 1 else None,
                     "traffic_note": ("static: ncu capture under profiles/ of the stage's LARGEST launch (first bulk trailing "
                                      "update, k = 1024, 26.5 ms), DRAM bytes per launch; algorithmic bytes of that launch "
                                      "8.3e9 (RED read-modify-write of the lower trailing matrix + the panel)"
                                      if world == 1 else "not measured (tensor-bound stage)"),
                     "flops_per_stage": n ** 3 / 3.0, "stage_ms": stage_ms["potrf"],
                     "peak_source": "measured live on rank 0: DMMA.8x8x4 register-operand probe (gpb_microbench) x n_gpus; "
                                    "MEASURED_PEAKS.json has no FP64 entry"},
        "stages_ms": stage_ms,
        "cholesky": {"tflops": tf, "frac_of_fp64_tensor_peak": tf / peak},
        "check": {"nll0": float(nll[0]), "info_max": int(np.max(info))},
    }
    if want_grad:
        t_inv = stage_ms["trtri"] + stage_ms["lauum"]
        tf_l = n ** 3 / 3.0 / (stage_ms["lauum"] * 1e-3) / 1e12
        rec["roofline"]["secondary"] = {"kernel": "K^-1 = W^T W (gemm_kernel<GeoLauum / GeoDistLauum>, one launch per rank)",
                                        "achieved": tf_l, "frac": tf_l / peak, "launch_ms": stage_ms["lauum"]}
        rec["cholesky"]["inverse_tflops"] = 2 * n ** 3 / 3.0 / (t_inv * 1e-3) / 1e12
        rec["cholesky"]["inverse_frac"] = rec["cholesky"]["inverse_tflops"] / peak
        rec["check"]["grad_norm"] = float(np.linalg.norm(grads[0]))
    return rec


# ---- independent GPs on every rank: C1 / C2 / M32k (replicas), C3 / C4 (GPs sharded over the ranks) -----------------------
def run_batch(ctx, key, steps, warmup, cpu_leg):
    import torch
    from gaussianprocessfundamentals_b200 import engine as eng
    args, rank, world = ctx.args, ctx.rank, ctx.world
    w = WORKLOADS[key]
    trees, hps, ns, xs, ys = build_workload(key, rank, world)
    t_jit = time.perf_counter()
    progs = eng.DeviceProgram.get_many(trees, 1, False, 1)      # kernels specialised per program: compiled in parallel
    t_jit = time.perf_counter() - t_jit
    plan = eng.Plan(progs, ns, want_grad=True)
    for b in range(len(ns)):
        plan.set_data(b, torch.tensor(xs[b]), torch.tensor(ys[b]))
        plan.set_hp(b, hps[b], 1e-2)
    n = w["n"]
    ev, barrier = ctx.ev, ctx.barrier

    # launches of one evaluation (counted eagerly: a graph replay does not pass through the host-side counter)
    l0 = eng.launch_count()
    plan.eval(eng.STAGES_LML_GRAD)
    torch.cuda.synchronize()
    launches_per_eval = eng.launch_count() - l0
    nll_first = plan.results()[0].copy()      # compared bit for bit with the values after the timed region

    # ---- stage timings (one evaluation, CUDA events on the launching stream) -------------------------------------------
    stage_ms = {}
    for _ in range(2):
        marks = [ev() for _ in range(7)]
        marks[0].record()
        plan.eval(eng.STAGE_ASSEMBLE); marks[1].record()
        plan.eval(eng.STAGE_POTRF); marks[2].record()
        plan.eval(eng.STAGE_NLL); marks[3].record()
        plan.eval(eng.STAGE_TRTRI); marks[4].record()
        plan.eval(eng.STAGE_LAUUM); marks[5].record()      # ONE launch: Kinv = W^T W
        plan.eval(eng.STAGE_GRAD); marks[6].record()
        torch.cuda.synchronize()
        for i, name in enumerate(["assemble", "potrf", "nll", "trtri", "lauum", "grad"]):
            stage_ms[name] = marks[i].elapsed_time(marks[i + 1])
        stage_ms["inverse"] = stage_ms["trtri"] + stage_ms["lauum"]
    dmma_peak = ctx.dmma_peak()
    cublas_dgemm = ctx.cublas_dgemm()

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

    for _ in range(warmup):
        step()
    sampler = ClockSampler(ctx.local_rank)
    barrier()
    sampler.start()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier()
    elapsed_ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    nll, grads, info = plan.results()

    # ---- e2e: host buffers in ONE pinned staging buffer laid out like the plan's input region (a single H2D for all
    #      GPs), H2D + kernels + D2H + sync inside every step ------------------------------------------------------------
    hx, hy = plan.host_inputs()
    for b in range(len(ns)):
        hx[b][...] = xs[b]
        hy[b][...] = ys[b].reshape(-1)
    noises = [1e-2] * len(ns)
    for _ in range(2):
        plan.eval_host(hps, noises, hx, hy)
    barrier()
    t0 = time.perf_counter()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(steps):
        nll_h, grads_h, info_h = plan.eval_host(hps, noises, hx, hy)
    e1.record()
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)
    h2d = sum(x.nbytes + y.nbytes for x, y in zip(hx, hy)) + sum(h.nbytes for h in hps) + 8 * len(ns)
    d2h = (8 + 4 + 16) * len(ns) + sum((h.size + 1) * 8 for h in hps)
    keys = sorted(stage_ms)
    red = ctx.max_over_ranks([elapsed_ms, e2e_ms] + [stage_ms[k_] for k_ in keys])
    elapsed_ms, e2e_ms = red[0], red[1]
    for i, k_ in enumerate(keys):
        stage_ms[k_] = red[2 + i]
    total_gps = int(round(ctx.sum_over_ranks(len(ns))))
    used_graph = graph is not None
    del plan, graph
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    evals = total_gps * steps
    value = evals / (elapsed_ms * 1e-3)
    peak = dmma_peak * world
    flops_potrf, flops_inv = n ** 3 / 3.0 * total_gps, 2.0 * n ** 3 / 3.0 * total_gps
    tf_potrf = flops_potrf / (stage_ms["potrf"] * 1e-3) / 1e12
    tf_inv = flops_inv / (stage_ms["inverse"] * 1e-3) / 1e12
    tf_all = (flops_potrf + flops_inv) / ((stage_ms["potrf"] + stage_ms["inverse"]) * 1e-3) / 1e12
    tf_lauum = (n ** 3 / 3.0 * total_gps) / (stage_ms["lauum"] * 1e-3) / 1e12
    rec = {
        "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": elapsed_ms / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": batch_config(key, total_gps, used_graph),
        "e2e": {"value": evals / (e2e_ms * 1e-3), "unit": "evals/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms / steps,
                "h2d_copies_per_step": 2 if len(ns) > 1 else 3},
        "gpu_launches": int(launches_per_eval * steps),
        "clocks": clocks,
        # the DOMINANT stage: the Cholesky factorisation (all its launches), n^3 / 3 flops per GP over the stage time
        "roofline": {"bound": "tensor", "kernel": "Cholesky factorisation stage (gemm_kernel<GeoSyrk> trailing updates, "
                     "panel products, diagonal blocks)", "achieved": tf_potrf, "peak": peak, "unit": "TFLOP/s",
                     "frac": tf_potrf / peak, "traffic": None,
                     "traffic_note": "not measured in this run (tensor / latency-bound stage)",
                     "flops_per_stage": flops_potrf, "stage_ms": stage_ms["potrf"],
                     "secondary": {"kernel": "gemm_kernel<GeoLauum> (K^-1 = W^T W, the single largest launch)",
                                   "achieved": tf_lauum, "frac": tf_lauum / peak, "launch_ms": stage_ms["lauum"],
                                   "traffic": _traffic(key), "traffic_note": "static: ncu capture under profiles/"},
                     "peak_source": "measured live: DMMA.8x8x4 register-operand probe (gpb_microbench) x n_gpus; "
                                    "MEASURED_PEAKS.json has no FP64 entry; cuBLAS DGEMM 8192^3 = %.2f TFLOP/s"
                                    % cublas_dgemm,
                     "all_gemm_stages": {"achieved": tf_all, "frac": tf_all / peak,
                                         "flops_per_step": flops_potrf + flops_inv}},
        "stages_ms": stage_ms,
        "cholesky": {"tflops": tf_potrf, "frac_of_fp64_tensor_peak": tf_potrf / peak,
                     "inverse_tflops": tf_inv, "inverse_frac": tf_inv / peak},
        "check": {"nll0": float(nll[0]), "info_max": int(np.max(info)),
                  "e2e_matches_resident": bool(abs(nll_h[0] - nll[0]) <= 1e-12 * abs(nll[0])),
                  "bitwise_reproducible": bool(np.array_equal(nll_first, nll))},
        # assembly and trace gradient run on kernels generated per kernel program and compiled at program creation
        # (csrc/jit.cu, NVRTC); set-up cost, outside the timed region like the plan construction
        "specialised_kernels": {"programs": len(set(id(p_) for p_ in progs)),
                                "specialised": int(sum(1 for p_ in set(progs) if p_.specialised)),
                                "create_s": round(t_jit, 3)},
    }
    if cpu_leg:
        rec["cpu_baseline"] = cpu_baseline(key)
        rec["cpu_baseline"].pop("seconds_per_step", None)
    return rec


def _brief(rec):
    """sub-record: the numbers of a workload without the boilerplate of a full line"""
    keep = ("value", "unit", "ms_per_step", "scaling", "config", "e2e", "gpu_launches", "stages_ms", "cholesky", "check",
            "cpu_baseline", "n_gpus", "steps", "specialised_kernels")
    out = {k: rec[k] for k in keep if k in rec}
    out["roofline_frac"] = rec["roofline"]["frac"]
    return out


# ---- main ---------------------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    key = args.workload
    w = WORKLOADS[key]
    cpu_eval_seconds("c1", 256)      # thread-pool warm-up
    times, what = [], ""
    for i in range(args.warmup + args.steps):
        t, what = cpu_step_seconds(key)
        if i >= args.warmup:
            times.append(t)
    per_step = float(np.mean(times))
    gps = w["B"]
    value = gps / per_step
    scaling = "strong" if key in DIST_WORKLOADS else "weak"
    line = {"metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            # the SAME configuration as our arm prints for this workload and N (what differs is the implementation)
            "config": dist_config(key, args.gpus, args.grid) if key in DIST_WORKLOADS else
                      batch_config(key, gps * (args.gpus if gps == 1 else 1)),
            "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": "port",
                             "sample": "CPU port of the reference's unfused op sequence with autodiff through the Cholesky "
                                       "(oracle/gp_oracle.py), torch CPU float64, all host threads; mean of %d steps; "
                                       "each step: %s" % (args.steps, what),
                             "tensorflow_importable": _tensorflow_importable()},
            "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if key == DEFAULT_WORKLOAD and not args.no_sub_records:
        # the metric's n = 8k on the same cores, IN FULL (median of 3 evaluations, no extrapolation)
        base = cpu_baseline("c2", repeats=3)
        line["sub_records"] = {"c2": {"value": base["value"], "unit": "evals/s", "ms_per_step": base["seconds_per_step"] * 1e3,
                                      "config": {"workload": WORKLOADS["c2"]["name"], "noise": 1e-2},
                                      "cpu_baseline": {k: v for k, v in base.items() if k != "seconds_per_step"}}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub-records", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--storage", default="replicated", choices=["replicated", "columns"],
                    help="likelihood-only distributed workloads (c5, c5s) on N > 1 GPUs: 'columns' keeps only the own block "
                         "columns on every rank (gpb_plan_create_dist_columns)")
    ap.add_argument("--size", type=int, default=0, help="override the size of a one-large-GP workload (c5, c5s, c5g, m32kd)")
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

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = Ctx(args, rank, world, local_rank)
    key = args.workload
    cpu_leg = (not args.no_cpu_baseline) and world == 1
    if key in DIST_WORKLOADS:
        line = run_dist(ctx, key, args.steps, args.warmup)
    else:
        line = run_batch(ctx, key, args.steps, args.warmup, cpu_leg)
    if key == DEFAULT_WORKLOAD and not args.no_sub_records:
        # the other configurations of BASELINE.json in the same driver run: C3 / C4 shard their GPs over the ranks (no
        # data-path collective); C2 (the metric's n = 8k, with the CPU port timed IN FULL beside it) and C1 do not shard
        # and are measured at N = 1 only
        sub_steps = max(3, min(args.steps, 10))
        subs = {}
        for sk in (["c2", "c1"] if world == 1 else []) + ["c3", "c4"]:
            rec = run_batch(ctx, sk, sub_steps, 3, cpu_leg and sk == "c2")
            if rank == 0:
                subs[sk] = _brief(rec)
        if rank == 0:
            line["sub_records"] = subs
            if "c2" in subs and "cpu_baseline" in subs["c2"]:
                # the bounded CPU sample of the contract: the metric's n = 8k evaluated in full on the host cores
                line["cpu_baseline"] = dict(subs["c2"]["cpu_baseline"], workload=WORKLOADS["c2"]["name"],
                                            gpu_value_same_workload=subs["c2"]["value"])
    elif rank == 0 and cpu_leg and "cpu_baseline" not in line:
        line["cpu_baseline"] = cpu_baseline(key)
        line["cpu_baseline"].pop("seconds_per_step", None)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

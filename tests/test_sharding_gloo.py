"""world_size = 2 coverage of the sharded block / candidate evaluation (first row) on CPU with `gloo`.

The device plan cannot run here, so the ranks evaluate their own blocks with the CPU oracle through the same
SegmentedCovarianceMatrix.block_nll_and_grad / sharding.Sharding code the GPU path uses; what is tested is the host
logic: the unit -> rank map, that each rank touches only its own blocks, the all-gather, and that every rank returns
the single-process totals.  The same scenario runs on two B200s in tests/test_gpu_multi.py."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gaussianprocessfundamentals_b200 import sharding  # noqa: E402


def free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


# ---- the unit -> rank map ----------------------------------------------------------------------------------------
def test_assign_is_a_partition_and_deterministic():
    rng = np.random.default_rng(0)
    for world in (1, 2, 3, 4, 8):
        costs = list(rng.integers(1, 50, size=37).astype(float) ** 3)
        m = sharding.assign(costs, world)
        assert sorted(i for part in m for i in part) == list(range(37))
        assert all(part == sorted(part) for part in m)
        assert m == sharding.assign(costs, world)
        loads = [sum(costs[i] for i in part) for part in m]
        # LPT guarantee: makespan <= 4/3 optimum; optimum >= mean load and >= the largest unit
        assert max(loads) <= 4.0 / 3.0 * max(sum(costs) / world, max(costs)) + 1e-9


def test_assign_equal_costs_is_round_robin():
    m = sharding.assign([1.0] * 10, 4)
    assert m == [[0, 4, 8], [1, 5, 9], [2, 6], [3, 7]]
    assert sharding.assign([], 3) == [[], [], []]
    assert sharding.assign([5.0], 2) == [[0], []]


def test_single_process_sharding_is_identity():
    sh = sharding.Sharding([8.0, 1.0, 27.0], rank=0, world=1)
    assert sh.mine == [0, 1, 2]
    nll, grads, info = sh.combine([1.0, 2.0, 3.0], [np.array([1.0, 2.0]), np.array([3.0]), np.array([4.0, 5.0, 6.0])],
                                  [0, 0, 0], [2, 1, 3])
    assert nll.tolist() == [1.0, 2.0, 3.0] and info.tolist() == [0, 0, 0]
    assert [g.tolist() for g in grads] == [[1.0, 2.0], [3.0], [4.0, 5.0, 6.0]]


# ---- two ranks ---------------------------------------------------------------------------------------------------
SPECS = [("SE",), ("ADD", [("SE",), ("LIN",)]), ("PER",), ("MUL", [("SE",), ("PER",)]), ("SE",)]
SIZES = [40, 90, 0, 60, 25]      # block 2 is empty: its hyper-parameters are skipped but still consumed


def _make_problem():
    from gaussianprocessfundamentals_b200.DataHandling import DataInput as di
    from gaussianprocessfundamentals_b200.KernelBasics import BaseKernels as bk, Operators as op
    from gaussianprocessfundamentals_b200.KernelBasics import PartitioningModel as pm, PartitionOperator as po
    from gaussianprocessfundamentals_b200.MeanFunctionBasics import BaseMeanFunctions as bmf
    from gaussianprocessfundamentals_b200.Statistics import GaussianProcess as gproc

    def build(spec):
        if spec[0] == "SE":
            return bk.SquaredExponentialKernel(1)
        if spec[0] == "PER":
            return bk.PeriodicKernel(1)
        if spec[0] == "LIN":
            return bk.LinearKernel(1)
        cls = op.AdditionOperator if spec[0] == "ADD" else op.MultiplicationOperator
        return cls(1, [build(c) for c in spec[1]])

    rng = np.random.default_rng(11)
    edges = np.concatenate([[0.0], np.cumsum([max(s, 5) for s in SIZES])]).astype(np.float64)
    xs = []
    for i, s in enumerate(SIZES):
        xs.append(np.sort(rng.uniform(edges[i] + 0.01, edges[i + 1] - 0.01, size=s)))
    x = np.concatenate(xs)[:, None]
    y = np.sin(0.7 * x) + 0.1 * rng.standard_normal(x.shape)
    model = pm.PartitioningModel(pm.PartitioningClass.SELF_SUFFICIENT, [])
    model.init_partitioning([pm.IntervalCriterion(edges[i], edges[i + 1]) for i in range(len(SIZES))])
    kern = po.PartitionOperator(1, [build(s) for s in SPECS], model)
    hp = []
    for d in kern.get_hyper_parameter_dimensionalities():
        size = 1 if len(d) == 0 else d[0]
        hp.append(torch.tensor(rng.uniform(0.5, 2.0, size=size)).reshape(d))
    pdi = model.partition_data_input(di.DataInput(x, y, x, y))
    pdi.set_mean_function(bmf.ZeroMeanFunction(1))
    pgp = gproc.PartitionedGaussianProcess(kern, bmf.ZeroMeanFunction(1))
    return pgp, pdi, hp


class OracleBlocks:
    """stand-in for Statistics._device.DeviceBlocks with the same evaluate() contract, computing with the CPU oracle
    (test infrastructure).  Records which blocks it was built for."""

    built_for = []

    def __init__(self, kernels, xs, ys):
        from gaussianprocessfundamentals_b200.program import compile_spec
        self.kernels, self.xs, self.ys = kernels, xs, ys
        self.entries = [compile_spec(k.to_spec(), 1, False).entries for k in kernels]
        self.last = None
        OracleBlocks.built_for.append([int(x.shape[0]) for x in xs])

    def matches(self, kernels, y_source):
        return len(kernels) == len(self.kernels)

    def evaluate(self, hp_lists, noises, grad, check=True):
        from oracle import gp_oracle as orc
        nll, grads = [], []
        for k, x, y, hp, s2 in zip(self.kernels, self.xs, self.ys, hp_lists, noises):
            hpv = [np.asarray(torch.as_tensor(h)) for h in hp]
            v, g, gn = orc.nll_and_grad(k.to_spec(), hpv, s2, np.asarray(x), np.asarray(y), reference_distance=False)
            nll.append(v)
            grads.append(np.concatenate([np.asarray(t, dtype=np.float64).reshape(-1) for t in g] + [[gn]]))
        info = np.zeros(len(nll), dtype=np.int32)
        self.last = (np.asarray(nll), grads, info)
        return np.asarray(nll), grads

    def grads_as_lists(self, grads, hp_lists):
        from gaussianprocessfundamentals_b200.program import unflatten_grad
        return [unflatten_grad(e, g[:-1], like=hp) for e, g, hp in zip(self.entries, grads, hp_lists)], \
            [float(g[-1]) for g in grads]


def _evaluate(shard: bool):
    from gaussianprocessfundamentals_b200.Metrics import Auxiliary as met_aux, Metrics as met
    pgp, pdi, hp = _make_problem()
    pgp.set_data_input(pdi)
    cov = pgp.covariance_matrix
    cov._make_blocks = lambda kernels, xs, ys: OracleBlocks(kernels, xs, ys)
    OracleBlocks.built_for = []
    if shard:
        cov.shard()
    metric = met_aux.get_metric_by_type(met.MetricType.blockwise_LL, pgp)
    noise = torch.tensor(1e-2, dtype=torch.float64)
    val = float(metric.get_metric(hp, noise, None))
    per_block = list(metric.last_block_values)
    grads, gnoise = metric.get_gradients(hp, noise, with_noise=True)
    flat = np.concatenate([np.asarray(g).reshape(-1) for g in grads] + [[float(gnoise)]])
    return val, per_block, flat, OracleBlocks.built_for


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        val, per_block, flat, built = _evaluate(shard=True)
        out.put((rank, val, per_block, flat, built))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_blockwise_nll_two_ranks_gloo():
    want_val, want_blocks, want_flat, built_single = _evaluate(shard=False)
    assert built_single[0] == [s for s in SIZES if s > 0]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([out.get(timeout=240) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    active_sizes = [s for s in SIZES if s > 0]
    m = sharding.assign([sharding.estimated_cost(s) for s in active_sizes], 2)
    for rank, val, per_block, flat, built in got:
        # each rank built (and therefore evaluated) only its own blocks
        assert built[0] == [active_sizes[p] for p in m[rank]]
        # ... and still returns the totals of the single-process evaluation (same oracle arithmetic per block:
        # bit-identical block values; the sums differ at most by reassociation)
        assert per_block == want_blocks
        assert val == pytest.approx(want_val, rel=1e-14)
        assert np.allclose(flat, want_flat, rtol=1e-13, atol=0)
    assert sorted(got[0][4][0] + got[1][4][0]) == sorted(active_sizes)


# ---- bench.py's own sharding of the batched workloads -----------------------------------------------------------
def test_bench_workloads_partition_their_gps_across_ranks():
    """C3 / C4 of bench.py: every GP is evaluated by exactly one rank, whatever the number of ranks, and a rank's share is
    the same data it would hold in the single-process run (build_workload is pure host code)."""
    import bench
    for key, B in (("c3", 256), ("c4", 1024)):
        t1, h1, n1, x1, y1 = bench.build_workload(key, 0, 1)
        assert len(t1) == B and all(n == bench.WORKLOADS[key]["n"] for n in n1)
        for world in (2, 8):
            seen = []
            for rank in range(world):
                t, h, ns, xs, ys = bench.build_workload(key, rank, world)
                mine = list(range(rank, B, world))
                assert len(t) == len(mine)
                seen += mine
                for pos in (0, len(mine) - 1):                      # spot-check: same tree, hyper-parameters and targets
                    b = mine[pos]
                    assert t[pos] == t1[b]
                    assert np.array_equal(h[pos], h1[b])
                    assert np.array_equal(ys[pos], y1[b]) and np.array_equal(xs[pos], x1[b])
            assert sorted(seen) == list(range(B))


# ---- batched kernel search (search.CandidateBatch) on two ranks ---------------------------------------------------------
CANDIDATES = [("SE",), ("PER",), ("ADD", [("SE",), ("LIN",)]), ("MUL", [("SE",), ("PER",)]),
              ("ADD", [("MUL", [("SE",), ("LIN",)]), ("PER",)]), ("LIN",), ("MUL", [("PER",), ("LIN",)])]


def _candidate_problem():
    from gaussianprocessfundamentals_b200.DataHandling import DataInput as di
    from gaussianprocessfundamentals_b200.KernelBasics import BaseKernels as bk, Operators as op
    from gaussianprocessfundamentals_b200.MeanFunctionBasics import BaseMeanFunctions as bmf

    def build(spec):
        if spec[0] == "SE":
            return bk.SquaredExponentialKernel(1)
        if spec[0] == "PER":
            return bk.PeriodicKernel(1)
        if spec[0] == "LIN":
            return bk.LinearKernel(1)
        cls = op.AdditionOperator if spec[0] == "ADD" else op.MultiplicationOperator
        return cls(1, [build(c) for c in spec[1]])

    rng = np.random.default_rng(21)
    n = 70
    x = np.sort(rng.uniform(0, 1, (n, 1)), axis=0)
    y = np.sin(6 * x) + 0.5 * x + 0.1 * rng.standard_normal((n, 1))
    din = di.DataInput(x, y, x, y)
    din.set_mean_function(bmf.ZeroMeanFunction(1))
    kernels = [build(s) for s in CANDIDATES]
    hps = []
    for k in kernels:
        hp = []
        for d in k.get_hyper_parameter_dimensionalities():
            size = 1 if len(d) == 0 else d[0]
            hp.append(torch.tensor(rng.uniform(0.3, 1.2, size=size)).reshape(d))
        hps.append(hp)
    return kernels, din, hps


def _candidate_eval():
    from gaussianprocessfundamentals_b200 import search
    kernels, din, hps = _candidate_problem()
    batch = search.CandidateBatch(kernels, din)
    batch._make_blocks = lambda ks, xs, ys: OracleBlocks(ks, xs, ys)
    OracleBlocks.built_for = []
    nll, grads, gnoise = batch.evaluate(hps, torch.tensor(1e-2, dtype=torch.float64))
    flat = [np.concatenate([np.asarray(t).reshape(-1) for t in g]) for g in grads]
    return nll, flat, gnoise, batch.sharding.mine, batch.best(hps, 1e-2)


def _candidate_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out.put((rank,) + _candidate_eval())
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_candidate_batch_two_ranks_gloo():
    nll1, flat1, gn1, mine1, best1 = _candidate_eval()
    assert mine1 == list(range(len(CANDIDATES)))
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_candidate_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([out.get(timeout=240) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    shares = [g[4] for g in got]
    assert sorted(shares[0] + shares[1]) == list(range(len(CANDIDATES))) and shares[0] and shares[1]
    for rank, nll, flat, gn, mine, best in got:
        assert np.array_equal(nll, nll1) and best == best1          # same oracle arithmetic per candidate
        assert all(np.array_equal(a, b) for a, b in zip(flat, flat1))
        assert np.array_equal(gn, gn1)

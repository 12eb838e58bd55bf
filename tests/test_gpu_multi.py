"""Multi-GPU parity (needs >= 2 B200s: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`).

  * distributed Cholesky / LML of one GP over a P x Q process grid (2D block-cyclic block ownership, NCCL panel
    broadcasts) against the single-GPU plan on the same inputs and against the CPU oracle;
  * sharded blockwise likelihood (independent blocks, one all-gather of scalars) against the single-process value.
Both run one process per GPU with torch.distributed (nccl) on 127.0.0.1."""
import os
import socket
import sys
import traceback

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

LL_RTOL = 1e-10     # north-star tolerance on the log-likelihood
GRAD_RTOL = 1e-8    # ... and on the gradient


def _ngpu():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem(n, d, seed):
    rng = np.random.default_rng(seed)
    if d == 1:
        x = np.sort(rng.uniform(0, 1, size=(n, 1)), axis=0)
        y = x * np.sin(40 * x) + 0.1 * rng.standard_normal((n, 1))
        tree, hp = ("MUL", [("ADD", [("SE",), ("PER",)]), ("LIN",)]), np.array([0.1, 0.1, 0.1, 0.01])
    else:
        x = rng.uniform(0, 1, size=(n, d))
        y = np.sum(np.sin(3 * x), axis=1, keepdims=True) + 0.1 * rng.standard_normal((n, 1))
        tree, hp = ("SE_ARD",), rng.uniform(0.3, 1.0, size=d)
    return tree, hp, x, y


def _dist_worker(rank, world, port, cases, out):
    try:
        import torch.distributed as dist
        from gaussianprocessfundamentals_b200 import engine as eng
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        res = []
        grids = {}
        for (n, d, P, Q) in cases:
            if (P, Q) not in grids:
                grids[(P, Q)] = eng.ProcessGrid(P, Q)
            grid = grids[(P, Q)]
            tree, hp, x, y = _problem(n, d, 7 + n)
            prog = eng.DeviceProgram.get(tree, d, False, 1)
            # single-GPU plan on this rank: the reference value of the comparison
            ref = eng.Plan([prog], [n], want_grad=False)
            ref.set_data(0, torch.tensor(x), torch.tensor(y)); ref.set_hp(0, hp, 1e-2)
            ref.eval(eng.STAGES_LML)
            nll_ref, _, info_ref = ref.results()
            L_ref = ref.lower_matrix(0).clone()
            z_ref = ref.buffer(0, eng.BUF_Z).clone()
            # distributed plan (collective)
            dp = eng.Plan([prog], [n], want_grad=False, grid=grid)
            dp.set_data(0, torch.tensor(x), torch.tensor(y)); dp.set_hp(0, hp, 1e-2)
            dp.eval(eng.STAGES_LML)
            torch.cuda.synchronize()
            nll_d, _, info_d = dp.results()
            L_d = dp.lower_matrix(0)
            tril = torch.tril(torch.ones(n, n, dtype=torch.bool, device="cuda"))
            dL = float((L_d - L_ref)[tril].abs().max())
            dz = float((dp.buffer(0, eng.BUF_Z) - z_ref).abs().max())
            # a second evaluation must reproduce the first (staging buffers / events are reused)
            dp.eval(eng.STAGES_LML)
            torch.cuda.synchronize()
            nll_d2 = dp.results()[0]
            # column storage (1 x Q grids): every rank keeps only its own block columns, panels travel in place over NCCL
            nll_cs, info_cs, ws_ratio = None, -1, 0.0
            if P == 1:
                cs = eng.Plan([prog], [n], want_grad=False, grid=grid, storage="columns")
                cs.set_data(0, torch.tensor(x), torch.tensor(y)); cs.set_hp(0, hp, 1e-2)
                for _ in range(2):
                    cs.eval(eng.STAGES_LML)
                    torch.cuda.synchronize()
                nll_c, _, info_c = cs.results()
                nll_cs, info_cs, ws_ratio = float(nll_c[0]), int(info_c[0]), cs.ws_bytes / dp.ws_bytes
                del cs
            # host-buffer call
            nll_h, _, info_h = dp.eval_host([hp], [1e-2], [x], [y.reshape(-1)], stages=eng.STAGES_LML)
            # gradient stages: W = inv(L) and inv(K) split by block column, one all-reduce of the gradient
            refg = eng.Plan([prog], [n], want_grad=True)
            refg.set_data(0, torch.tensor(x), torch.tensor(y)); refg.set_hp(0, hp, 1e-2)
            refg.eval(eng.STAGES_LML_GRAD)
            nll_rg, g_ref, _ = refg.results()
            dpg = eng.Plan([prog], [n], want_grad=True, grid=grid)
            dpg.set_data(0, torch.tensor(x), torch.tensor(y)); dpg.set_hp(0, hp, 1e-2)
            dpg.eval(eng.STAGES_LML_GRAD)
            torch.cuda.synchronize()
            nll_g, g_d, info_g = dpg.results()
            dW = float((dpg.lower_matrix(0) - refg.lower_matrix(0))[tril].abs().max())
            Wmax = float(refg.lower_matrix(0)[tril].abs().max())
            dalpha = float((dpg.buffer(0, eng.BUF_ALPHA) - refg.buffer(0, eng.BUF_ALPHA)).abs().max())
            amax = float(refg.buffer(0, eng.BUF_ALPHA).abs().max())
            dpg.eval(eng.STAGES_LML_GRAD)
            torch.cuda.synchronize()
            g_d2 = dpg.results()[1]
            res.append(dict(n=n, d=d, P=P, Q=Q, nll_cs=nll_cs, info_cs=info_cs, ws_ratio=ws_ratio,
                            nll_ref=float(nll_ref[0]), nll=float(nll_d[0]), nll2=float(nll_d2[0]),
                            nll_host=float(nll_h[0]), info=int(info_d[0]), info_ref=int(info_ref[0]), dL=dL, dz=dz,
                            Lmax=float(L_ref[tril].abs().max()), nll_grad_run=float(nll_g[0]), info_g=int(info_g[0]),
                            grad=[float(v) for v in g_d[0]], grad_ref=[float(v) for v in g_ref[0]],
                            grad2=[float(v) for v in g_d2[0]], dW=dW, Wmax=Wmax, dalpha=dalpha, amax=amax))
            del ref, dp, refg, dpg
            torch.cuda.empty_cache()
        # non positive definite input: every rank must report the same pivot
        n = 700
        tree, hp, x, y = _problem(n, 1, 3)
        prog = eng.DeviceProgram.get(tree, 1, False, 1)
        dp = eng.Plan([prog], [n], want_grad=False, grid=grids[tuple(cases[0][2:])])
        dp.set_data(0, torch.tensor(x), torch.tensor(y)); dp.set_hp(0, hp, -5.0)
        dp.eval(eng.STAGES_LML)
        nll_bad, _, info_bad = dp.results()
        res.append(dict(bad_info=int(info_bad[0]), bad_nll_isnan=bool(np.isnan(nll_bad[0]))))
        dist.barrier()
        out.put((rank, res, None))
        dist.destroy_process_group()
    except Exception:
        out.put((rank, None, traceback.format_exc()))


def _run(world, target, *args):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=target, args=(r, world, port) + args + (out,)) for r in range(world)]
    for p in procs:
        p.start()
    got = [out.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
    for rank, res, err in got:
        assert err is None, "rank %d failed:\n%s" % (rank, err)
    return sorted(got, key=lambda t: t[0])


@pytest.mark.timeout(900)
@pytest.mark.parametrize("outer_blocks", ["", "2"])
def test_distributed_cholesky_matches_single_gpu(outer_blocks, monkeypatch):
    """outer_blocks = "2" forces the 256-wide outer panels (two panels / two block rows per far update) that the drivers
    only choose for n >= 6144, so that the test sizes exercise them too (GPB_POTRF_KB is read once per process; the
    spawned ranks inherit it)."""
    world = min(_ngpu(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    if outer_blocks:
        monkeypatch.setenv("GPB_POTRF_KB", outer_blocks)
    else:
        monkeypatch.delenv("GPB_POTRF_KB", raising=False)
    world = 4 if world >= 4 else 2
    shapes = [(1, 2), (2, 1)] if world == 2 else [(1, 4), (2, 2), (4, 1)]
    cases = []
    for (P, Q) in shapes:
        cases += [(1000, 1, P, Q), (1024, 1, P, Q), (2304 + 17, 1, P, Q), (127, 1, P, Q), (1536, 8, P, Q)]
    got = _run(world, _dist_worker, cases)
    from oracle import gp_oracle as orc
    for rank, res, _ in got:
        for r in res[:-1]:
            tag = "rank %d case %r" % (rank, r)
            assert r["info"] == 0 and r["info_ref"] == 0, tag
            assert abs(r["nll"] - r["nll_ref"]) <= LL_RTOL * abs(r["nll_ref"]), tag
            assert r["nll2"] == r["nll"], tag
            assert r["nll_host"] == r["nll"], tag
            if r["P"] == 1:    # column storage: same likelihood from a fraction of the workspace
                assert r["info_cs"] == 0 and abs(r["nll_cs"] - r["nll_ref"]) <= LL_RTOL * abs(r["nll_ref"]), tag
                if r["n"] >= 1000:
                    assert r["ws_ratio"] < 0.8, tag
            assert r["dL"] <= 1e-12 * r["Lmax"], tag
            assert r["dz"] <= 1e-9, tag
            assert r["info_g"] == 0 and abs(r["nll_grad_run"] - r["nll_ref"]) <= LL_RTOL * abs(r["nll_ref"]), tag
            g, gr = np.asarray(r["grad"]), np.asarray(r["grad_ref"])
            assert np.max(np.abs(g - gr)) <= GRAD_RTOL * np.max(np.abs(gr)), tag
            assert r["grad2"] == r["grad"], tag
            assert r["dW"] <= 1e-10 * r["Wmax"], tag
            assert r["dalpha"] <= 1e-9 * r["amax"], tag
        assert res[-1]["bad_info"] > 0 and res[-1]["bad_nll_isnan"]
    # every rank reports the same numbers
    for r0, r1 in zip(got[0][1], got[-1][1]):
        r0, r1 = dict(r0), dict(r1)
        r0.pop("ws_ratio", None); r1.pop("ws_ratio", None)     # the ranks own different numbers of block columns
        assert r0 == r1
    # against the CPU oracle (small case)
    r = got[0][1][0]
    tree, hp, x, y = _problem(r["n"], r["d"], 7 + r["n"])
    hpl = [np.asarray(v) for v in hp]
    want, _, _ = orc.nll_and_grad(tree, hpl, 1e-2, x, y, reference_distance=False)
    # The strict 1e-10 comparison above is with the single-GPU plan (itself held to 1e-10 against oracle values frozen in
    # the build container, tests/test_gpu_parity_sizes.py).  The LIVE oracle runs on the CPU of whatever box the test
    # lands on, and its LAPACK result for this ill-conditioned composite kernel moves in the 10th digit from box to box
    # (seen: 1.6e-10 on one pool box, 3e-12 on others), so it is held to 1e-8 here.
    assert abs(r["nll"] - want) <= 1e-8 * abs(want)


def _shard_worker(rank, world, port, out):
    try:
        import torch.distributed as dist
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        from tests.test_sharding_gloo import _make_problem
        from gaussianprocessfundamentals_b200.Metrics import Auxiliary as met_aux, Metrics as met
        vals = {}
        for shard in (False, True):
            pgp, pdi, hp = _make_problem()
            pgp.set_data_input(pdi)
            if shard:
                pgp.covariance_matrix.shard()
            metric = met_aux.get_metric_by_type(met.MetricType.blockwise_LL, pgp)
            noise = torch.tensor(1e-2, dtype=torch.float64)
            v = float(metric.get_metric(hp, noise, None))
            grads, gn = metric.get_gradients(hp, noise, with_noise=True)
            flat = np.concatenate([np.asarray(g).reshape(-1) for g in grads] + [[float(gn)]])
            vals[shard] = (v, flat, list(metric.last_block_values))
        dist.barrier()
        out.put((rank, vals, None))
        dist.destroy_process_group()
    except Exception:
        out.put((rank, None, traceback.format_exc()))


@pytest.mark.timeout(600)
def test_sharded_blockwise_likelihood_two_gpus():
    if _ngpu() < 2:
        pytest.skip("needs >= 2 GPUs")
    got = _run(2, _shard_worker)
    for rank, vals, _ in got:
        v0, g0, b0 = vals[False]
        v1, g1, b1 = vals[True]
        assert b0 == b1                       # per-block values: same kernels on the same inputs
        assert abs(v0 - v1) <= 1e-13 * abs(v0)
        assert np.allclose(g0, g1, rtol=1e-12, atol=0)
    assert got[0][1][True][0] == got[1][1][True][0]


def _surface_worker(rank, world, port, out):
    """the reference-shaped surface on a process grid: GaussianProcess -> LogLikelihood with covariance_matrix.distribute"""
    try:
        import torch.distributed as dist
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        from gaussianprocessfundamentals_b200 import engine as eng
        from gaussianprocessfundamentals_b200.DataHandling import DataInput as di
        from gaussianprocessfundamentals_b200.KernelBasics import BaseKernels as bk, Operators as op
        from gaussianprocessfundamentals_b200.MeanFunctionBasics import BaseMeanFunctions as bmf
        from gaussianprocessfundamentals_b200.Metrics import Auxiliary as met_aux, Metrics as met
        from gaussianprocessfundamentals_b200.Statistics import GaussianProcess as gproc
        n = 1500
        rng = np.random.default_rng(5)
        x = np.sort(rng.uniform(0, 1, size=(n, 1)), axis=0)
        y = np.sin(11 * x) + 0.1 * rng.standard_normal((n, 1))
        hp = [torch.tensor(0.15, dtype=torch.float64), torch.tensor(0.4, dtype=torch.float64),
              torch.tensor(0.3, dtype=torch.float64)]
        noise = torch.tensor(1e-2, dtype=torch.float64)
        vals = {}
        grid = eng.ProcessGrid(1, world)
        for distributed in (False, True):
            kern = op.AdditionOperator(1, [bk.SquaredExponentialKernel(1), bk.PeriodicKernel(1)])
            din = di.DataInput(x, y, x[:50], y[:50])
            din.set_mean_function(bmf.ZeroMeanFunction(1))
            gp = gproc.GaussianProcess(kern, bmf.ZeroMeanFunction(1))
            gp.set_data_input(din)
            if distributed:
                gp.covariance_matrix.distribute(grid)
            metric = met_aux.get_metric_by_type(met.MetricType.LL, gp)
            v = float(metric.get_metric(hp, noise, None))
            g = np.concatenate([np.asarray(t).reshape(-1) for t in metric.get_gradients(hp, noise)])
            gp.covariance_matrix.reset()
            alpha = gp.covariance_matrix.get_L_alpha(hp, noise).cpu().numpy().reshape(-1)
            vals[distributed] = (v, g, alpha)
        dist.barrier()
        out.put((rank, vals, None))
        dist.destroy_process_group()
    except Exception:
        out.put((rank, None, traceback.format_exc()))


@pytest.mark.timeout(600)
def test_distributed_gp_through_the_reference_surface():
    if _ngpu() < 2:
        pytest.skip("needs >= 2 GPUs")
    got = _run(2, _surface_worker)
    for rank, vals, _ in got:
        v0, g0, a0 = vals[False]
        v1, g1, a1 = vals[True]
        assert abs(v0 - v1) <= LL_RTOL * abs(v0)
        assert np.max(np.abs(g0 - g1)) <= GRAD_RTOL * np.max(np.abs(g0))
        assert np.max(np.abs(a0 - a1)) <= 1e-9 * np.max(np.abs(a0))
    assert got[0][1][True][0] == got[1][1][True][0]

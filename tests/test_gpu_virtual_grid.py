"""ONE-GPU tests of the DISTRIBUTED path (the "virtual grid"): P x Q virtual ranks inside one
process on one device run the distributed Cholesky, triangular inverse, K^-1 and trace gradient through the SAME code
the multi-GPU path runs - block ownership, panel packing (`panel_copy_kernel`), `GeoDistPanel`, `GeoDistSyrk`, the
look-ahead schedule with 128- and 256-wide outer panels, `run_trtri_dist`, `GeoDistLauum`, the gradient split by block
column - with the NCCL collectives replaced by the loop-back transport (stream-ordered device copies;
gpb_dist_loopback_create).  Every virtual rank is compared with the single-GPU plan and with the CPU oracle.
Tolerances (north_star): relative <= 1e-10 on the likelihood, <= 1e-8 on gradients."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import gp_oracle as orc

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LL_RTOL, GRAD_RTOL = 1e-10, 1e-8


def _eng():
    from gaussianprocessfundamentals_b200 import engine
    return engine


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


def run_case(eng, n, d, P, Q, check_oracle=True):
    """the distributed LML + gradient of one GP on a P x Q virtual grid against the single-GPU plan (and the oracle)"""
    tree, hp, x, y = _problem(n, d, 7 + n)
    prog = eng.DeviceProgram.get(tree, d, False, 1)
    ref = eng.Plan([prog], [n], want_grad=True)
    ref.set_data(0, torch.tensor(x), torch.tensor(y)); ref.set_hp(0, hp, 1e-2)
    ref.eval(eng.STAGES_LML)
    torch.cuda.synchronize()
    nll_ref = float(ref.results()[0][0])
    L_ref = ref.lower_matrix(0).clone()
    z_ref = ref.buffer(0, eng.BUF_Z).clone()
    ref.eval(eng.STAGES_LML_GRAD)
    torch.cuda.synchronize()
    _, g_ref, _ = ref.results()
    W_ref = ref.lower_matrix(0).clone()
    a_ref = ref.buffer(0, eng.BUF_ALPHA).clone()
    tril = torch.tril(torch.ones(n, n, dtype=torch.bool, device="cuda"))
    vg = eng.VirtualGrid(P, Q)
    plans = [eng.Plan([prog], [n], want_grad=True, grid=vg.ranks[r]) for r in range(vg.world)]

    def work(r, grid):
        dp = plans[r]
        dp.set_data(0, torch.tensor(x), torch.tensor(y)); dp.set_hp(0, hp, 1e-2)
        out = {}
        dp.eval(eng.STAGES_LML)
        torch.cuda.current_stream().synchronize()
        nll, _, info = dp.results()
        out["nll"], out["info"] = float(nll[0]), int(info[0])
        out["dL"] = float((dp.lower_matrix(0) - L_ref)[tril].abs().max())
        out["dz"] = float((dp.buffer(0, eng.BUF_Z) - z_ref).abs().max())
        # gradient stages on top of the factor that is already there (stage-by-stage calls are collective too)
        dp.eval(eng.STAGE_TRTRI | eng.STAGE_LAUUM | eng.STAGE_GRAD)
        torch.cuda.current_stream().synchronize()
        out["grad_staged"] = dp.results()[1][0].copy()
        # one call, twice: staging buffers and events are reused
        for key in ("grad", "grad2"):
            dp.eval(eng.STAGES_LML_GRAD)
            torch.cuda.current_stream().synchronize()
            nll_g, g, info_g = dp.results()
            out[key], out["nll_" + key], out["info_" + key] = g[0].copy(), float(nll_g[0]), int(info_g[0])
        out["dW"] = float((dp.lower_matrix(0) - W_ref)[tril].abs().max())
        out["dalpha"] = float((dp.buffer(0, eng.BUF_ALPHA) - a_ref).abs().max())
        # host-buffer call
        nll_h, g_h, info_h = dp.eval_host([hp], [1e-2], [x], [y.reshape(-1)])
        out["nll_host"], out["grad_host"], out["info_host"] = float(nll_h[0]), g_h[0].copy(), int(info_h[0])
        return out

    res = vg.run(work)
    Lmax, Wmax, amax = float(L_ref[tril].abs().max()), float(W_ref[tril].abs().max()), float(a_ref.abs().max())
    gmax = float(np.max(np.abs(g_ref[0])))
    for r, o in enumerate(res):
        tag = (n, d, P, Q, r)
        assert o["info"] == 0 and o["info_grad"] == 0 and o["info_host"] == 0, tag
        for key in ("nll", "nll_grad", "nll_grad2", "nll_host"):
            assert abs(o[key] - nll_ref) <= LL_RTOL * abs(nll_ref), (tag, key, o[key], nll_ref)
        assert o["dL"] <= 1e-12 * Lmax and o["dz"] <= 1e-11 * float(z_ref.abs().max()), (tag, o["dL"], o["dz"])
        assert o["dW"] <= 1e-11 * Wmax and o["dalpha"] <= 1e-10 * amax, (tag, o["dW"], o["dalpha"])
        for key in ("grad", "grad2", "grad_staged", "grad_host"):
            assert np.max(np.abs(o[key] - g_ref[0])) <= GRAD_RTOL * gmax, (tag, key, o[key], g_ref[0])
        # all ranks hold the same numbers bit for bit (every rank sums the gradient in rank order)
        assert np.array_equal(o["grad"], res[0]["grad"]) and o["nll"] == res[0]["nll"], tag
    if check_oracle:
        hp_list = [hp[0], hp[1], hp[2], hp[3:4]] if d == 1 else [hp]
        want, gw, gn = orc.nll_and_grad(tree, hp_list, 1e-2, x, y, reference_distance=False)
        gflat = np.concatenate([np.asarray(t).reshape(-1) for t in gw] + [[gn]])
        assert abs(res[0]["nll"] - want) <= LL_RTOL * abs(want)
        assert np.max(np.abs(res[0]["grad"] - gflat)) <= GRAD_RTOL * np.max(np.abs(gflat))
    return res


def run_store_case(eng, n, d, Q, noise=1e-2):
    """COLUMN STORAGE (gpb_plan_create_dist_columns): every virtual rank keeps only its own block columns; the
    likelihood must equal the single-GPU plan's, on every rank bit for bit, with a smaller workspace"""
    tree, hp, x, y = _problem(n, d, 11 + n)
    prog = eng.DeviceProgram.get(tree, d, False, 1)
    ref = eng.Plan([prog], [n], want_grad=False)
    ref.set_data(0, torch.tensor(x), torch.tensor(y)); ref.set_hp(0, hp, noise)
    ref.eval(eng.STAGES_LML)
    torch.cuda.synchronize()
    nll_ref, _, info_ref = ref.results()
    vg = eng.VirtualGrid(1, Q)
    plans = [eng.Plan([prog], [n], want_grad=False, grid=vg.ranks[r], storage="columns") for r in range(Q)]
    rep_bytes = eng.Plan([prog], [n], want_grad=False, grid=vg.ranks[0]).ws_bytes

    def work(r, grid):
        dp = plans[r]
        dp.set_data(0, torch.tensor(x), torch.tensor(y)); dp.set_hp(0, hp, noise)
        out = []
        for _ in range(3):                       # ring slots, events and the reduction scratch are reused
            dp.eval(eng.STAGES_LML)
            torch.cuda.current_stream().synchronize()
            nll, _, info = dp.results()
            out.append((float(nll[0]), int(info[0])))
        nll_h, _, info_h = dp.eval_host([hp], [noise], [x], [y.reshape(-1)], stages=eng.STAGES_LML)
        out.append((float(nll_h[0]), int(info_h[0])))
        return out
    res = vg.run(work)
    for r, o in enumerate(res):
        for nll, info in o:
            assert info == int(info_ref[0]), (n, Q, r, info, info_ref)
            if info == 0:
                assert abs(nll - float(nll_ref[0])) <= LL_RTOL * abs(float(nll_ref[0])), (n, Q, r, nll, nll_ref)
                assert nll == res[0][0][0], (n, Q, r)             # every rank, every repetition: the same bits
            else:
                assert np.isnan(nll)
    if n >= 1000 and Q >= 2:
        assert max(pl.ws_bytes for pl in plans) < rep_bytes, (plans[0].ws_bytes, rep_bytes)
    return res


@pytest.mark.parametrize("Q", [2, 3, 4])
def test_column_storage_likelihood_matches_single_gpu_plan(Q):
    eng = _eng()
    run_store_case(eng, 1000, 1, Q)          # ragged last block
    run_store_case(eng, 1024, 8, Q)          # the carried row opens a block column of its own
    run_store_case(eng, 1409, 1, Q)          # odd number of blocks: ragged last group
    run_store_case(eng, 385, 1, Q)           # fewer groups than ranks
    run_store_case(eng, 700, 1, Q, noise=-5.0)   # not positive definite: the same info on every rank, NaN likelihood


def test_column_storage_refuses_gradient_stages():
    eng = _eng()
    from gaussianprocessfundamentals_b200 import _lib
    tree, hp, x, y = _problem(300, 1, 5)
    prog = eng.DeviceProgram.get(tree, 1, False, 1)
    vg = eng.VirtualGrid(1, 2)
    with pytest.raises(_lib.GpbError):
        eng.Plan([prog], [300], want_grad=True, grid=vg.ranks[0], storage="columns")
    plans = [eng.Plan([prog], [300], want_grad=False, grid=vg.ranks[r], storage="columns") for r in range(2)]

    def work(r, grid):
        dp = plans[r]
        dp.set_data(0, torch.tensor(x), torch.tensor(y)); dp.set_hp(0, hp, 1e-2)
        dp.eval(eng.STAGES_LML)
        refused = 0
        for bit in (eng.STAGE_TRTRI, eng.STAGE_BACKSOLVE, eng.STAGE_GRAD):
            try:
                dp.eval(bit)
            except _lib.GpbError:
                refused += 1
        return refused
    assert vg.run(work) == [3, 3]
    vg2 = eng.VirtualGrid(2, 1)
    with pytest.raises(_lib.GpbError):
        eng.Plan([prog], [300], want_grad=False, grid=vg2.ranks[0], storage="columns")   # needs a 1 x Q grid


@pytest.mark.parametrize("P,Q", [(1, 2), (2, 1), (2, 2), (1, 3), (3, 2)])
def test_virtual_grid_matches_single_gpu_plan(P, Q):
    eng = _eng()
    run_case(eng, 1000, 1, P, Q)            # ragged last block, carried row in its own block row (1000 = 7 * 128 + 104)
    run_case(eng, 1024, 8, P, Q)            # n multiple of 128: the carried y row opens block row 8
    run_case(eng, 385, 1, P, Q, check_oracle=False)   # fewer blocks than some grids have ranks


def test_virtual_grid_not_positive_definite_same_info_on_every_rank():
    eng = _eng()
    n = 700
    tree, hp, x, y = _problem(n, 1, 3)
    prog = eng.DeviceProgram.get(tree, 1, False, 1)
    vg = eng.VirtualGrid(2, 2)
    plans = [eng.Plan([prog], [n], want_grad=True, grid=vg.ranks[r]) for r in range(4)]

    def work(r, grid):
        dp = plans[r]
        dp.set_data(0, torch.tensor(x), torch.tensor(y)); dp.set_hp(0, hp, -5.0)
        dp.eval(eng.STAGES_LML_GRAD)
        torch.cuda.current_stream().synchronize()
        nll, g, info = dp.results()
        return int(info[0]), bool(np.isnan(nll[0])), bool(np.all(np.isnan(g[0])))
    res = vg.run(work)
    assert res[0][0] > 0 and all(r == res[0] for r in res) and res[0][1] and res[0][2]


def test_stage_order_is_enforced():
    """after the distributed inverse the buffer no longer holds L (nor z): BACKSOLVE / NLL must be refused,
    not read garbage; the single-GPU plan refuses BACKSOLVE after the inverse and POTRF without a fresh assembly"""
    eng = _eng()
    from gaussianprocessfundamentals_b200 import _lib
    n = 300
    tree, hp, x, y = _problem(n, 1, 5)
    prog = eng.DeviceProgram.get(tree, 1, False, 1)
    plan = eng.Plan([prog], [n], want_grad=True)
    plan.set_data(0, torch.tensor(x), torch.tensor(y)); plan.set_hp(0, hp, 1e-2)
    with pytest.raises(_lib.GpbError):
        plan.eval(eng.STAGE_POTRF)                     # nothing assembled yet
    plan.eval(eng.STAGES_LML)
    with pytest.raises(_lib.GpbError):
        plan.eval(eng.STAGE_POTRF)                     # would factorise L a second time
    with pytest.raises(_lib.GpbError):
        plan.eval(eng.STAGE_GRAD)                      # no K^-1 yet
    plan.eval(eng.STAGE_BACKSOLVE)                     # legal: L is there
    plan.eval(eng.STAGE_INVERSE)
    plan.eval(eng.STAGE_NLL)                           # single GPU: z^T survives the inverse
    with pytest.raises(_lib.GpbError):
        plan.eval(eng.STAGE_BACKSOLVE)                 # L is gone
    plan.eval(eng.STAGE_GRAD)
    vg = eng.VirtualGrid(1, 2)
    plans = [eng.Plan([prog], [n], want_grad=True, grid=vg.ranks[r]) for r in range(2)]

    def work(r, grid):
        dp = plans[r]
        dp.set_data(0, torch.tensor(x), torch.tensor(y)); dp.set_hp(0, hp, 1e-2)
        dp.eval(eng.STAGES_LML_GRAD)
        refused = 0
        for bit in (eng.STAGE_NLL, eng.STAGE_BACKSOLVE):
            try:
                dp.eval(bit)
            except _lib.GpbError:
                refused += 1
        return refused
    assert vg.run(work) == [2, 2]


def test_virtual_grid_wide_outer_panels_subprocess():
    """GPB_POTRF_KB=2 (read once per process) forces the 256-wide outer panels of the distributed factorisation and of the
    distributed forward substitution at a size a test can afford"""
    script = r"""
import sys
sys.path.insert(0, %r)
from gaussianprocessfundamentals_b200 import engine as eng
from tests.test_gpu_virtual_grid import run_case
for (P, Q) in [(1, 2), (2, 2), (2, 1), (1, 3)]:
    run_case(eng, 1100, 1, P, Q)
    run_case(eng, 1280, 8, P, Q, check_oracle=False)
    run_case(eng, 1409, 1, P, Q, check_oracle=False)
print("ok")
""" % ROOT
    env = dict(os.environ, GPB_POTRF_KB="2")
    out = subprocess.run([sys.executable, "-c", script], env=env, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0 and "ok" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


@pytest.mark.parametrize("ow", ["1", "4"])
def test_virtual_grid_column_group_widths_subprocess(ow):
    """GPB_DIST_OW (read at grid creation): block columns owned one by one / in groups of four on 1 x Q grids - the
    default (groups of two) is what the tests above run.  Sizes with a ragged last group and fewer groups than ranks."""
    script = r"""
import sys
sys.path.insert(0, %r)
from gaussianprocessfundamentals_b200 import engine as eng
from tests.test_gpu_virtual_grid import run_case, run_store_case
for (P, Q) in [(1, 2), (1, 3), (1, 4)]:
    run_case(eng, 1100, 1, P, Q)
    run_case(eng, 1280, 8, P, Q, check_oracle=False)
    run_case(eng, 385, 1, P, Q, check_oracle=False)
    run_store_case(eng, 1100, 1, Q)
    run_store_case(eng, 1280, 8, Q)
print("ok")
""" % ROOT
    # ow = 1 also runs the un-overlapped exchange of W (one exchange after the inverse, W^T W as one launch)
    env = dict(os.environ, GPB_DIST_OW=ow, GPB_DIST_OVERLAP="0" if ow == "1" else "1")
    out = subprocess.run([sys.executable, "-c", script], env=env, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0 and "ok" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]

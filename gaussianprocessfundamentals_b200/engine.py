"""Host-side driver of libgpb: device programs, plans (workspace + launch sequence) and their torch views.

PyTorch is used for plumbing only (device memory, streams, CUDA graphs); every FLOP of the path runs in libgpb's
kernels.  There is no CPU fallback: constructing a DeviceProgram or a Plan without CUDA raises.
"""
import ctypes
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .program import CompiledProgram, compile_spec, flatten_hp, unflatten_grad

STAGE_ASSEMBLE, STAGE_POTRF, STAGE_NLL, STAGE_INVERSE, STAGE_GRAD, STAGE_BACKSOLVE = 1, 2, 4, 8, 16, 32
STAGE_TRTRI, STAGE_LAUUM = 64, 128   # the two halves of STAGE_INVERSE
STAGES_LML, STAGES_LML_GRAD = 7, 31
BUF_A, BUF_KINV, BUF_ALPHA, BUF_Z, BUF_X, BUF_Y, BUF_HP, BUF_NOISE, BUF_NLL, BUF_GRAD, BUF_INFO, BUF_TERMS = range(12)


def require_cuda():
    if not torch.cuda.is_available():
        raise _lib.GpbError("gaussianprocessfundamentals_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def _stream_ptr() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class NotPositiveDefinite(ArithmeticError):
    """Cholesky met a non-positive pivot (the reference surfaces TF's InvalidArgumentError here)."""

    def __init__(self, info):
        super().__init__("Cholesky decomposition was not successful: non-positive pivot at index %s (1-based)" % (info,))
        self.info = info


class DeviceProgram:
    """A compiled kernel tree resident on the device (gpb_program_t)."""

    _cache = {}

    def __init__(self, compiled: CompiledProgram, cp_mode: int):
        require_cuda()
        lib = _lib.load()
        self.compiled = compiled
        self.cp_mode = int(cp_mode)
        code = np.ascontiguousarray(compiled.code, dtype=np.int32)
        h = ctypes.c_void_p()
        _lib.check(lib.gpb_program_create(code.ctypes.data_as(_lib.c_int32_p), compiled.n_ops, compiled.dim,
                                          self.cp_mode, ctypes.byref(h)), "gpb_program_create")
        self.handle = h
        assert lib.gpb_program_num_hp(h) == compiled.n_hp

    @classmethod
    def get(cls, spec, dim: int, scaled: bool, cp_mode: int) -> "DeviceProgram":
        compiled = compile_spec(spec, dim, scaled)
        key = (compiled.signature(), int(cp_mode), torch.cuda.current_device())
        prog = cls._cache.get(key)
        if prog is None:
            prog = cls(compiled, cp_mode)
            cls._cache[key] = prog
        return prog

    @classmethod
    def get_many(cls, specs: Sequence, dim: int, scaled: bool, cp_mode: int) -> List["DeviceProgram"]:
        """One program per spec; programs that do not exist yet are created on a thread pool: gpb_program_create compiles
        the kernels specialised for the program with NVRTC (0.4 - 1.5 s each) and releases the GIL while it does."""
        import os
        from concurrent.futures import ThreadPoolExecutor
        compiled = [compile_spec(spec, dim, scaled) for spec in specs]
        dev = torch.cuda.current_device()
        keys = [(c.signature(), int(cp_mode), dev) for c in compiled]
        missing = {}
        for k, c in zip(keys, compiled):
            if k not in cls._cache and k not in missing:
                missing[k] = c
        if len(missing) > 1:
            def make(c):
                torch.cuda.set_device(dev)
                return cls(c, cp_mode)
            with ThreadPoolExecutor(max_workers=min(len(missing), os.cpu_count() or 4, 32)) as pool:
                for k, prog in zip(missing, pool.map(make, missing.values())):
                    cls._cache[k] = prog
        elif missing:
            (k, c), = missing.items()
            cls._cache[k] = cls(c, cp_mode)
        return [cls._cache[k] for k in keys]

    @property
    def n_hp(self) -> int:
        return self.compiled.n_hp

    @property
    def specialised(self) -> bool:
        """the program runs on kernels generated and compiled for it (csrc/jit.cu), not on the interpreter"""
        return bool(_lib.load().gpb_program_is_specialised(self.handle))

    @property
    def jit_note(self) -> str:
        return _lib.load().gpb_program_jit_note(self.handle).decode("utf-8", "replace")

    def __del__(self):
        try:
            if getattr(self, "handle", None) is not None and _lib._lib is not None:
                _lib._lib.gpb_program_destroy(self.handle)
        except Exception:
            pass


def assemble(prog: DeviceProgram, X: torch.Tensor, X2: Optional[torch.Tensor], hp_flat: torch.Tensor,
             noise: Optional[torch.Tensor], lower_only: bool = False) -> torch.Tensor:
    """K[n, m] (logical, row-major semantics) as a transposed view of the column-major buffer the kernel writes."""
    require_cuda()
    lib = _lib.load()
    n = X.shape[0]
    m = n if X2 is None else X2.shape[0]
    ld = max(n, 1)
    buf = torch.empty((m, ld), dtype=torch.float64, device=X.device)
    if lower_only:
        buf.zero_()
    _lib.check(lib.gpb_assemble(prog.handle, X.data_ptr(), None if X2 is None else X2.data_ptr(), n, m,
                                hp_flat.data_ptr() if hp_flat.numel() else None,
                                None if noise is None else noise.data_ptr(), buf.data_ptr(), ld,
                                1 if lower_only else 0, _stream_ptr()), "gpb_assemble")
    return buf.t()[:n, :m]


class Plan:
    """B independent GPs evaluated together: assembly -> Cholesky (+ carried y) -> NLL -> inverse -> gradient."""

    def __init__(self, programs: Sequence[DeviceProgram], ns: Sequence[int], want_grad: bool = True,
                 device: Optional[torch.device] = None, grid: Optional["ProcessGrid"] = None,
                 storage: str = "replicated"):
        """grid: a ProcessGrid makes this the distributed plan of ONE GP (2D block-cyclic block ownership, NCCL panel
        broadcasts; inverse and gradient split by block column).  Every rank of the grid must construct it and call
        eval collectively.  storage = "columns" (1 x Q grids, want_grad = False): every rank keeps only the block
        columns it owns - the likelihood of matrices that do not fit one GPU (STAGES_LML only)."""
        require_cuda()
        lib = _lib.load()
        self.lib = lib
        self.programs = list(programs)
        self.ns = [int(v) for v in ns]
        self.B = len(self.programs)
        assert self.B == len(self.ns) and self.B > 0
        self.want_grad = bool(want_grad)
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        handles = (ctypes.c_void_p * self.B)(*[p.handle for p in self.programs])
        n_arr = (ctypes.c_int64 * self.B)(*self.ns)
        h = ctypes.c_void_p()
        self.grid = grid
        self.storage = storage
        if storage not in ("replicated", "columns"):
            raise ValueError("storage must be 'replicated' or 'columns'")
        if storage == "columns":
            if grid is None or self.B != 1 or want_grad:
                raise _lib.GpbError("column storage: one GP on a process grid, likelihood only (want_grad=False)")
            _lib.check(lib.gpb_plan_create_dist_columns(self.programs[0].handle, self.ns[0], grid.handle, ctypes.byref(h)),
                       "gpb_plan_create_dist_columns")
        elif grid is not None:
            if self.B != 1:
                raise _lib.GpbError("a distributed plan holds one GP")
            _lib.check(lib.gpb_plan_create_dist(self.programs[0].handle, self.ns[0], 1 if want_grad else 0, grid.handle,
                                                ctypes.byref(h)),
                       "gpb_plan_create_dist")
        else:
            _lib.check(lib.gpb_plan_create(self.B, handles, n_arr, 1 if want_grad else 0, ctypes.byref(h)),
                       "gpb_plan_create")
        self.handle = h
        self.ws_bytes = int(lib.gpb_plan_workspace_bytes(h))
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.device)
        _lib.check(lib.gpb_plan_bind(h, ctypes.c_void_p(self.ws.data_ptr())), "gpb_plan_bind")
        self.hp_sizes = [p.n_hp for p in self.programs]
        self.grad_offsets = np.concatenate([[0], np.cumsum([s + 1 for s in self.hp_sizes])]).astype(np.int64)
        # host staging for eval_host
        self._nll_h = np.zeros(self.B, dtype=np.float64)
        self._grad_h = np.zeros(int(self.grad_offsets[-1]), dtype=np.float64)
        self._info_h = np.zeros(self.B, dtype=np.int32)

    # -- buffers -----------------------------------------------------------------------------------------------
    def buffer(self, b: int, which: int) -> torch.Tensor:
        ptr, nbytes, ld = ctypes.c_void_p(), ctypes.c_size_t(), ctypes.c_int64()
        _lib.check(self.lib.gpb_plan_buffer(self.handle, b, which, ctypes.byref(ptr), ctypes.byref(nbytes),
                                            ctypes.byref(ld)), "gpb_plan_buffer")
        off = ptr.value - self.ws.data_ptr()
        raw = self.ws[off:off + nbytes.value]
        if which == BUF_INFO:
            return raw.view(torch.int32)
        t = raw.view(torch.float64)
        n = self.ns[b]
        if which == BUF_A:
            if self.storage == "columns":
                return t.view(-1, ld.value)      # the packed own block columns (row r = local column r)
            return t.view(n + 1, ld.value)       # row r of this view = column r of the column-major matrix
        if which == BUF_KINV:
            return t.view(n, ld.value)
        if which == BUF_X:
            return t.view(n, -1)
        return t

    def lower_matrix(self, b: int, which: int = BUF_A) -> torch.Tensor:
        """logical n x n matrix view M[i, j] (lower triangle valid) of BUF_A / BUF_KINV"""
        n = self.ns[b]
        return self.buffer(b, which)[:n, :n].t()

    def set_data(self, b: int, X: torch.Tensor, y: torch.Tensor):
        self.buffer(b, BUF_X).copy_(X.reshape(self.ns[b], -1).to(self.device, torch.float64))
        self.buffer(b, BUF_Y).copy_(y.reshape(-1).to(self.device, torch.float64))

    def set_hp(self, b: int, hp_flat, noise: float):
        if self.hp_sizes[b]:
            self.buffer(b, BUF_HP).copy_(torch.as_tensor(np.asarray(hp_flat, dtype=np.float64)))
        self.buffer(b, BUF_NOISE).copy_(torch.as_tensor(np.asarray([noise], dtype=np.float64)))

    # -- evaluation --------------------------------------------------------------------------------------------
    def eval(self, stages: int = STAGES_LML_GRAD):
        _lib.check(self.lib.gpb_plan_eval(self.handle, int(stages), _stream_ptr()), "gpb_plan_eval")

    def results(self):
        """(nll[B], [grad_b], info[B]) read back from the device (synchronises)."""
        nll = torch.stack([self.buffer(b, BUF_NLL)[0] for b in range(self.B)]).cpu().numpy()
        info = torch.stack([self.buffer(b, BUF_INFO)[0] for b in range(self.B)]).cpu().numpy()
        grads = [self.buffer(b, BUF_GRAD).cpu().numpy().copy() for b in range(self.B)]
        return nll, grads, info

    def _ptr_array(self, key: str, arrs):
        """ctypes array of the host pointers of `arrs` (one float64 C-contiguous array per GP).  Rebuilt only when the
        caller passes different array objects: a fit loop hands over the same X / y every step, and for a thousand small
        GPs the conversion was a fifth of the end-to-end time."""
        if arrs is None:
            return None
        cached = self._ptr_cache.get(key)
        if cached is not None and len(cached[0]) == len(arrs) and all(a is b for a, b in zip(cached[0], arrs)):
            return cached[2]
        keep = [np.ascontiguousarray(a, dtype=np.float64) for a in arrs]
        out = (ctypes.c_void_p * self.B)(*[a.ctypes.data for a in keep])
        if all(k is a for k, a in zip(keep, arrs)):       # no conversion copies: the pointers follow in-place updates
            self._ptr_cache[key] = (list(arrs), keep, out)
        else:
            self._ptr_cache.pop(key, None)
            self._ptr_keep = keep                          # keep the converted copies alive for this call
        return out

    def eval_host(self, hp_flats: Sequence[np.ndarray], noises: Sequence[float], Xs=None, ys=None,
                  stages: int = STAGES_LML_GRAD):
        """The end-to-end call: host buffers in, host results out (H2D + kernels + D2H + sync inside)."""
        B = self.B
        if not hasattr(self, "_ptr_cache"):
            self._ptr_cache = {}
        # hyper-parameters change every call: one flat staging array, pointers into it
        sizes = [max(1, int(np.size(h))) for h in hp_flats]
        flat = np.zeros(int(np.sum(sizes)), dtype=np.float64)
        hp_ptrs = (ctypes.c_void_p * B)()
        pos = 0
        for b, h in enumerate(hp_flats):
            k = int(np.size(h))
            if k:
                flat[pos:pos + k] = np.asarray(h, dtype=np.float64).reshape(-1)
            hp_ptrs[b] = flat.ctypes.data + 8 * pos
            pos += sizes[b]
        x_ptrs = self._ptr_array("x", Xs)
        y_ptrs = self._ptr_array("y", ys)
        nz = np.ascontiguousarray(noises, dtype=np.float64)
        _lib.check(self.lib.gpb_plan_eval_host(self.handle, int(stages), x_ptrs, y_ptrs, hp_ptrs,
                                               nz.ctypes.data, self._nll_h.ctypes.data, self._grad_h.ctypes.data,
                                               self._info_h.ctypes.data, _stream_ptr()), "gpb_plan_eval_host")
        gh = self._grad_h.copy()
        off = self.grad_offsets
        grads = [gh[off[b]:off[b + 1]] for b in range(B)]
        return self._nll_h.copy(), grads, self._info_h.copy()

    def set_grad_weights(self, b: int, w_quad: float, w_logdet: float):
        """gradient of GP b = d/dtheta [w_quad * 1/2 y^T K^-1 y + w_logdet * sum(log diag L)] (default 1, 1 = the NLL)"""
        _lib.check(self.lib.gpb_plan_set_grad_weights(self.handle, int(b), float(w_quad), float(w_logdet)),
                   "gpb_plan_set_grad_weights")

    def last_terms(self):
        """(y^T K^-1 y [B], sum(log diag L) [B]) as copied back by the latest eval_host"""
        quad, logdet = np.zeros(self.B), np.zeros(self.B)
        _lib.check(self.lib.gpb_plan_last_terms(self.handle, quad.ctypes.data, logdet.ctypes.data), "gpb_plan_last_terms")
        return quad, logdet

    def host_inputs(self):
        """([X_b views], [y_b views]) into ONE pinned host buffer laid out like the plan's input region: passing these
        to eval_host moves the inputs of all GPs in a single host-to-device copy."""
        if getattr(self, "_host_in", None) is None:
            xo, yo, tot = ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t()
            offs = []
            for b in range(self.B):
                _lib.check(self.lib.gpb_plan_input_layout(self.handle, b, ctypes.byref(xo), ctypes.byref(yo),
                                                          ctypes.byref(tot)), "gpb_plan_input_layout")
                offs.append((xo.value, yo.value))
            pinned = torch.zeros(tot.value, dtype=torch.uint8).pin_memory()
            raw = pinned.numpy()
            xs, ys = [], []
            for b, (ox, oy) in enumerate(offs):
                n, d = self.ns[b], self.programs[b].compiled.dim
                xs.append(raw[ox:ox + 8 * n * d].view(np.float64).reshape(n, d))
                ys.append(raw[oy:oy + 8 * n].view(np.float64))
            self._host_in = (pinned, xs, ys)
        return self._host_in[1], self._host_in[2]

    def __del__(self):
        try:
            if getattr(self, "handle", None) is not None and _lib._lib is not None:
                _lib._lib.gpb_plan_destroy(self.handle)
        except Exception:
            pass


class ProcessGrid:
    """P x Q grid of the processes of a torch.distributed group (one process per GPU) with its own NCCL communicator
    inside libgpb (gpb_dist_init).  rank = p * Q + q.  torch.distributed is used once, to ship the NCCL unique id."""

    def __init__(self, P: int, Q: int, group=None):
        import torch.distributed as dist
        require_cuda()
        lib = _lib.load()
        if not (dist.is_available() and dist.is_initialized()):
            raise _lib.GpbError("ProcessGrid needs an initialised torch.distributed process group")
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if P * Q != self.world:
            raise ValueError("P x Q = %d x %d does not match the group size %d" % (P, Q, self.world))
        self.P, self.Q, self.p, self.q = int(P), int(Q), self.rank // Q, self.rank % Q
        ident = (ctypes.c_ubyte * 128)()
        if self.rank == 0:
            _lib.check(lib.gpb_dist_unique_id(ident), "gpb_dist_unique_id")
        box = [bytes(ident)]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        ident = (ctypes.c_ubyte * 128).from_buffer_copy(box[0])
        h = ctypes.c_void_p()
        _lib.check(lib.gpb_dist_init(ident, self.rank, self.world, self.P, self.Q, ctypes.byref(h)), "gpb_dist_init")
        self.handle = h

    @staticmethod
    def default_shape(world: int):
        """(P, Q) = (1, world): block-column-cyclic ownership.  The NVSwitch fabric is uniform, so the grid shape only
        balances the load, and P = 1 saves the diagonal-block round trip of every step (measured at n = 65536 on
        8 x B200: 1x8 660 ms, 2x4 728 ms; 1x4 1212 ms, 2x2 1230 ms).  Any P x Q = world with P <= 8 is accepted."""
        return 1, world

    def owner(self, I: int, J: int) -> int:
        return _lib.load().gpb_dist_owner_w(I, J, self.P, self.Q, _lib.load().gpb_dist_col_width(self.P))

    def __del__(self):
        try:
            if getattr(self, "handle", None) is not None and _lib._lib is not None:
                _lib._lib.gpb_dist_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class VirtualGrid:
    """A P x Q grid of VIRTUAL ranks on the current device (gpb_dist_loopback_create): every rank is a ProcessGrid-like
    handle for Plan(..., grid=...).  `run(fn)` calls fn(rank, grid_handle) on one host thread per rank, each inside its
    own CUDA stream, and returns the results in rank order - the calling convention of the distributed plan (every rank
    makes the same calls in the same order) on a single GPU.  Used by the one-GPU tests of the distributed path."""

    class _Rank:
        def __init__(self, handle, rank, world, P, Q):
            self.handle, self.rank, self.world, self.P, self.Q = handle, rank, world, P, Q
            self.p, self.q = rank // Q, rank % Q

        def owner(self, I: int, J: int) -> int:
            return _lib.load().gpb_dist_owner_w(I, J, self.P, self.Q, _lib.load().gpb_dist_col_width(self.P))

    def __init__(self, P: int, Q: int):
        require_cuda()
        lib = _lib.load()
        self.P, self.Q, self.world = int(P), int(Q), int(P) * int(Q)
        handles = (ctypes.c_void_p * self.world)()
        _lib.check(lib.gpb_dist_loopback_create(self.world, self.P, self.Q, handles), "gpb_dist_loopback_create")
        self.ranks = [self._Rank(ctypes.c_void_p(handles[r]), r, self.world, self.P, self.Q) for r in range(self.world)]
        self.streams = [torch.cuda.Stream() for _ in range(self.world)]

    def run(self, fn):
        import threading
        out, err = [None] * self.world, [None] * self.world
        dev = torch.cuda.current_device()

        def work(r):
            try:
                torch.cuda.set_device(dev)
                with torch.cuda.stream(self.streams[r]):
                    out[r] = fn(r, self.ranks[r])
                    self.streams[r].synchronize()
            except BaseException as exc:   # surfaced on the calling thread
                err[r] = exc
        torch.cuda.synchronize()
        threads = [threading.Thread(target=work, args=(r,)) for r in range(self.world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for e in err:
            if e is not None:
                raise e
        return out

    def __del__(self):
        try:
            if _lib._lib is not None:
                for rk in getattr(self, "ranks", []):
                    if rk.handle is not None:
                        _lib._lib.gpb_dist_destroy(rk.handle)
                        rk.handle = None
        except Exception:
            pass


def gemm(a_kmajor: bool, b_kmajor: bool, A: torch.Tensor, lda: int, Bm: torch.Tensor, ldb: int, C: torch.Tensor,
         ldc: int, M: int, N: int, K: int, alpha: float, beta: float):
    lib = _lib.load()
    _lib.check(lib.gpb_gemm(int(a_kmajor), int(b_kmajor), A.data_ptr(), lda, Bm.data_ptr(), ldb, C.data_ptr(), ldc,
                            M, N, K, float(alpha), float(beta), _stream_ptr()), "gpb_gemm")


def launch_count() -> int:
    return int(_lib.load().gpb_launch_count())


class trace:
    """with engine.trace() as t: plan.eval(...)  ->  t.spans = [(tag, a, b, stream, start_us, end_us)] of every launch
    (developer tool, gpb_trace_begin / gpb_trace_end; not usable around eval_host's graph replay)"""

    def __enter__(self):
        self.spans = []
        _lib.check(_lib.load().gpb_trace_begin(_stream_ptr()), "gpb_trace_begin")
        return self

    def __exit__(self, *exc):
        lib = _lib.load()
        need = ctypes.c_size_t()
        _lib.check(lib.gpb_trace_end(None, 0, ctypes.byref(need)), "gpb_trace_end")
        buf = ctypes.create_string_buffer(max(1, need.value))
        _lib.check(lib.gpb_trace_end(buf, len(buf), None), "gpb_trace_end")
        for line in buf.value.decode().splitlines():
            tag, a, b, st, t0, t1 = line.split()
            self.spans.append((tag, int(a), int(b), int(st), float(t0), float(t1)))
        return False


def _padded(t: torch.Tensor) -> torch.Tensor:
    """row-major 2-d tensor with an even number of columns and 16-byte aligned storage (copy only if needed)"""
    t = t.contiguous()
    if t.shape[1] % 2 == 0 and t.data_ptr() % 16 == 0:
        return t
    out = torch.zeros((t.shape[0], t.shape[1] + (t.shape[1] % 2)), dtype=torch.float64, device=t.device)
    out[:, :t.shape[1]].copy_(t)
    return out


def matmul(A: torch.Tensor, Bm: torch.Tensor, trans_a: bool = False, trans_b: bool = False) -> torch.Tensor:
    """C = op(A) op(B) for row-major float64 CUDA tensors on the FP64 tensor-core GEMM (gpb_gemm).

    A row-major [r, c] tensor is the column-major c x r matrix with leading dimension c, so C (row-major [m, n]) is
    computed as the column-major n x m product  C^T = op(B)^T op(A)^T."""
    require_cuda()
    m, kk = (A.shape[1], A.shape[0]) if trans_a else (A.shape[0], A.shape[1])
    k2, n = (Bm.shape[1], Bm.shape[0]) if trans_b else (Bm.shape[0], Bm.shape[1])
    assert kk == k2, "inner dimensions differ"
    Ap, Bp = _padded(A), _padded(Bm)
    C = torch.empty((m, n), dtype=torch.float64, device=A.device)
    # gemm operand 1 = op(B)^T as an (n x k) matrix: element (j, k) = op(B)[k, j]
    #   no trans_b: B[k, j] at k*ldB + j  -> MN-major (a_kmajor = 0);  trans_b: B[j, k] at j*ldB + k -> K-major
    # gemm operand 2 = op(A) as an (m x k) matrix: element (i, k) = op(A)[i, k]
    #   no trans_a: A[i, k] at i*ldA + k  -> K-major (b_kmajor = 1);   trans_a: A[k, i] at k*ldA + i -> MN-major
    gemm(bool(trans_b), not trans_a, Bp, Bp.shape[1], Ap, Ap.shape[1], C, n, n, m, kk, 1.0, 0.0)
    return C

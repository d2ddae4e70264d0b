"""Whole-sweep execution: CUDA-graph capture of one gradient evaluation and batch-sharded gradients.

``CapturedGradient``  — forward pass + reverse sweep of ``∇(f)`` recorded ONCE as a CUDA graph
    (SURVEY.md §8(f) rank 1).  The tape is built by the same host code as the eager path
    (core.py / rules.py, mirroring src/core.jl:214-229); only the launches are replayed, so the
    ~100 launches of the small MLP/VAE configurations cost one ``cudaGraphLaunch``.
``ShardedGradient``   — batch (column) sharding over G ranks: every rank evaluates
    the objective on its shard with the batch-independent prior scaled by 1/G, then ONE grouped
    NCCL allreduce(sum) of the objective and every parameter sensitivity runs on the sweep stream.
    The reference has no distributed layer; single-rank results are the parity target."""
import ctypes as C

import numpy as np

from ._lib import check
from .core import Leaf, Node, Tape, nabla, oneslike, unbox, unthunk, zeroslike
from .device import get_context
from .scalar import CATALOGUE, compile_scalar

_ADD_PROG = compile_scalar(CATALOGUE["add"], 2)


class CapturedGradient:
    """``∇(f; get_output=true)(params...)`` captured as a CUDA graph.

    ``params`` are DevArrays; ``f`` may close over further DevArrays (the data batch).  Replays read
    whatever those buffers hold at launch time, so a new batch is fed by copying into the same
    buffers (``DevArray.copy_from_host`` / ``Context.h2d_async``).  ``f`` must not read device values
    on the host while tracing (that would synchronise inside the capture)."""

    def __init__(self, f, params, ctx=None, warmup=True):
        self.ctx = ctx or (params[0].ctx if params else get_context())
        self.params = list(params)
        self._graph = None
        L, h = self.ctx.L, self.ctx.h
        if warmup:
            # one eager evaluation first: compiles the scalar programs and sizes the pool, so the
            # capture itself makes no cudaMalloc
            self._evaluate(f)
            self.ctx.sync()
        self.ctx.flush_deferred()
        check(L.nb_graph_begin(h), h)
        try:
            self.y, self.grads, self._tape = self._evaluate(f)
            self.ctx.flush_deferred(only_side_outputs=True)   # nothing the graph owes may launch after the capture
        except BaseException:
            g = C.c_void_p()
            L.nb_graph_end(h, C.byref(g))
            if g:
                L.nb_graph_destroy(g)
            raise
        g = C.c_void_p()
        check(L.nb_graph_end(h, C.byref(g)), h)
        self._graph = g
        before = self.ctx.stats()["kernel_launches"]
        self()                      # first replay materialises the outputs
        self.launches_per_replay = self.ctx.stats()["kernel_launches"] - before

    def _evaluate(self, f):
        t = Tape()
        leaves = [Leaf(t, p) for p in self.params]
        y = f(*leaves)
        if not isinstance(y, Node):
            raise TypeError("CapturedGradient: the objective does not depend on the parameters")
        df = nabla(y)
        grads = [unthunk(df.raw(l)) if df.isassigned(l) else zeroslike(l) for l in leaves]
        # a sensitivity that aliases another slot (a rule returned ȳ itself) gets its own buffer so
        # the in-place allreduce of ShardedGradient never sums one buffer twice
        grads = [g.copy() if g.shared else g for g in grads]
        # deferred products (fusion.py) must be launched now — inside the capture when there is one
        for g in grads:
            g.ptr
        unbox(y).ptr
        # keep the forward tape and the reverse tape alive: the graph refers to their buffers
        return unbox(y), grads, (t, df)

    def __call__(self):
        check(self.ctx.L.nb_graph_launch(self.ctx.h, self._graph), self.ctx.h)
        return self.y, self.grads

    def close(self):
        if self._graph is not None and self.ctx.h:
            self.ctx.L.nb_graph_destroy(self._graph)
        self._graph = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Communicator:
    """NCCL communicator bound to a context (include/nabla_cuda.h ``nb_comm_*``).  The 128-byte
    unique id is created on rank 0 and handed to ``exchange(id_bytes_or_None) -> id_bytes`` — any
    out-of-band broadcast (bench.py uses ``torch.distributed``)."""

    def __init__(self, ctx, rank, nranks, exchange):
        self.ctx, self.rank, self.nranks = ctx, rank, nranks
        if nranks > 1:
            uid = (C.c_uint8 * 128)()
            if rank == 0:
                check(ctx.L.nb_comm_unique_id(uid))
            raw = exchange(bytes(uid) if rank == 0 else None)
            uid = (C.c_uint8 * 128)(*raw)
            check(ctx.L.nb_comm_init(ctx.h, nranks, rank, uid), ctx.h)
            ctx.nranks, ctx.rank = nranks, rank

    def allreduce_sum(self, arrays):
        """In-place sum over ranks of every array (one grouped launch per dtype on the ctx stream)."""
        if self.nranks == 1 or not arrays:
            return
        ctx = self.ctx
        for dt in (np.dtype(np.float32), np.dtype(np.float64)):
            group = [a for a in arrays if a.dtype == dt]
            if not group:
                continue
            n = len(group)
            ptrs = (C.c_void_p * n)(*[a.ptr for a in group])
            counts = (C.c_int64 * n)(*[a.size for a in group])
            check(ctx.L.nb_allreduce_sum(ctx.h, n, ptrs, counts, group[0].nb_dtype), ctx.h)

    def allreduce_sum_async(self, arrays):
        """Start the in-place sum on the communicator's own stream, ordered after the work enqueued so
        far on the sweep stream; ``join()`` makes the sweep stream wait for all of them."""
        if self.nranks == 1 or not arrays:
            return
        ctx = self.ctx
        for dt in (np.dtype(np.float32), np.dtype(np.float64)):
            group = [a for a in arrays if a.dtype == dt]
            if not group:
                continue
            n = len(group)
            ptrs = (C.c_void_p * n)(*[a.ptr for a in group])
            counts = (C.c_int64 * n)(*[a.size for a in group])
            check(ctx.L.nb_allreduce_sum_async(ctx.h, n, ptrs, counts, group[0].nb_dtype), ctx.h)

    def join(self):
        if self.nranks > 1:
            check(self.ctx.L.nb_comm_join(self.ctx.h), self.ctx.h)

    def close(self):
        if self.nranks > 1 and self.ctx.h:
            self.ctx.L.nb_comm_destroy(self.ctx.h)


def shard_bounds(B, rank, nranks):
    """Contiguous, near-equal column range of rank ``rank`` (matches workloads.shard_columns)."""
    return (B * rank) // nranks, (B * (rank + 1)) // nranks


class OverlappedShardedGradient:
    """Batch-sharded gradient of ``data_objective + shared_objective`` with the exchange OVERLAPPED
    with the reverse sweep.

    ``data_objective`` is this rank's shard of the batch-dependent term (its gradients are summed over
    ranks); ``shared_objective`` is the batch-independent term (the prior), evaluated identically on
    every rank.  Per step:

    1. the shared term is swept FIRST with the seed ȳ = 1/G: its sensitivities (−λ/G·W for the prior of
       examples/mlp.jl:61-65; exact for G a power of two) pre-fill the reverse tape of the data sweep — the reference's
       own pre-filled-slot path (``compute_Dv_update``, src/finite_differencing.jl:57-85);
    2. the data term is swept: every rule accumulates into the pre-filled slot (``add!!``, core.jl:203-205 — β = 1 in the
       GEMM epilogue, ``+=`` in the row-sum), and the moment a parameter's sensitivity is final
       (``propagate(..., on_final)``) its allreduce(sum) is started on the communicator's stream, so the weights of the
       last layers travel over NVLink while the GEMMs of the layers below still run;
    3. the sweep stream joins the communicator stream.  Nothing runs after the last exchange: the sum over ranks of
       (data shard + shared/G) is the full gradient.

    With ``capture=True`` the whole step — both sweeps, every NCCL call on the side stream and the join — is recorded once
    as ONE CUDA graph and replayed (the tape signature of a fixed model is constant; NCCL collectives are capturable),
    which removes the host launch path (~90 launches + 11 collectives per step) from the critical path."""

    def __init__(self, data_objective, shared_objective, params, comm, capture=False):
        self.comm, self.ctx = comm, comm.ctx
        self.params = list(params)
        self.f_data, self.f_shared = data_objective, shared_objective
        self._graph = None
        self._keep = None
        if capture:
            self._capture()

    # -- one step, enqueued on the context (and communicator) streams ---------------------------------------------
    def _enqueue(self, keep_tapes=False):
        from .core import UNDEF, own, propagate, reverse_tape
        from .kernels import _timed, bcast_fwd
        comm, ctx = self.comm, self.ctx
        G = comm.nranks
        # ---- sweep A: shared term, seeded with 1/G
        pre = [None] * len(self.params)
        y2v = None
        tapes = []
        if self.f_shared is not None:
            with _timed(ctx, "shared_sweep", 0, "ms"):      # (CUDA-event bracket while bench.py profiles the step)
                t2 = Tape()
                leaves2 = [Leaf(t2, p) for p in self.params]
                y2 = self.f_shared(*leaves2)
                if isinstance(y2, Node):
                    y2v = unbox(y2)
                    rt2 = propagate(t2, reverse_tape(y2, ctx.full(y2v.shape, 1.0 / G, y2v.dtype)))
                    for i, l in enumerate(leaves2):
                        if rt2.isassigned(l):
                            pre[i] = own(unthunk(rt2.raw(l)))
                    tapes.append((t2, rt2))
                    if not keep_tapes:
                        t2.tape.clear()
        # ---- sweep B: data term accumulates into the pre-filled slots; allreduce of every final leaf slot overlapped
        t = Tape()
        leaves = [Leaf(t, p) for p in self.params]
        y = self.f_data(*leaves)
        rt = reverse_tape(y, oneslike(y))
        for l, g in zip(leaves, pre):
            if g is not None:
                rt.tape[l.pos - 1] = g
        grads = {}

        def on_final(leaf_pos, rvs):
            g = unthunk(rvs.raw(leaf_pos))
            if g.shared:
                g = g.copy()
            rvs.tape[leaf_pos - 1] = g
            grads[leaf_pos] = g
            comm.allreduce_sum_async([g])

        # the objective's exchange goes first (a copy: the forward value itself may still be read by a pullback), so
        # that it is not one more latency-bound collective behind the last weight sensitivity
        yd = unbox(y).copy() if G > 1 else unbox(y)
        comm.allreduce_sum_async([yd])
        propagate(t, rt, on_final=on_final)
        rest = []
        for l in leaves:            # parameters only the shared term touches: G copies of g/G
            if l.pos not in grads and rt.tape[l.pos - 1] is not UNDEF:
                grads[l.pos] = unthunk(rt.raw(l.pos))
                rest.append(grads[l.pos])
        comm.allreduce_sum_async(rest)
        out = [grads[l.pos] if l.pos in grads else zeroslike(l) for l in leaves]
        with _timed(ctx, "join_wait", 0, "ms"):           # what of the exchange is NOT hidden behind the sweep
            comm.join()
        if y2v is not None:
            ytot = ctx.empty((), yd.dtype)
            bcast_fwd(_ADD_PROG, [yd, y2v], ytot)
        else:
            ytot = yd
        tapes.append((t, rt))
        if not keep_tapes:
            t.tape.clear()
        return ytot, out, tapes

    def _capture(self):
        ctx = self.ctx
        L, h = ctx.L, ctx.h
        self._enqueue()          # eager once: compiles programs, sizes the pool (no cudaMalloc inside the capture)
        ctx.sync()
        ctx.flush_deferred()
        check(L.nb_graph_begin(h), h)
        try:
            y, grads, tapes = self._enqueue(keep_tapes=True)
            for g in grads:
                g.ptr               # deferred launches belong inside the capture
            y.ptr
            ctx.flush_deferred(only_side_outputs=True)
        except BaseException:
            g = C.c_void_p()
            L.nb_graph_end(h, C.byref(g))
            if g:
                L.nb_graph_destroy(g)
            raise
        g = C.c_void_p()
        check(L.nb_graph_end(h, C.byref(g)), h)
        self._graph = g
        self._keep = (y, grads, tapes)   # the graph refers to these buffers by address

    def step(self):
        if self._graph is not None:
            check(self.ctx.L.nb_graph_launch(self.ctx.h, self._graph), self.ctx.h)
            return self._keep[0], self._keep[1]
        y, grads, _ = self._enqueue()
        return y, grads

    def close(self):
        if self._graph is not None and self.ctx.h:
            self.ctx.L.nb_graph_destroy(self._graph)
        self._graph = None
        self._keep = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ShardedGradient:
    """Batch-sharded gradient evaluation.

    ``make_objective(prior_scale)`` returns the objective closed over THIS rank's data shard, with
    batch-independent terms multiplied by ``prior_scale`` = 1/G.  ``step()`` runs the (optionally
    graph-captured) sweep and the allreduce; results stay on the device."""

    def __init__(self, make_objective, params, comm, capture=True):
        self.comm = comm
        self.ctx = comm.ctx
        self.params = list(params)
        self.f = make_objective(1.0 / comm.nranks)
        self.captured = CapturedGradient(self.f, self.params, self.ctx) if capture else None

    def step(self):
        if self.captured is not None:
            y, grads = self.captured()
        else:
            y, grads = nabla(self.f, get_output=True)(*self.params)
            y, grads = unbox(y), [unthunk(g) for g in grads]
            for i, g in enumerate(grads):   # a sensitivity that aliases another slot gets its own buffer
                if g.shared:
                    grads[i] = g.copy()
        self.comm.allreduce_sum([y] + list(grads))
        return y, grads

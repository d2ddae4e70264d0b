"""Debug: who keeps never-launched deferred products alive after a sharded step."""
import os, sys, gc, types
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nabla_jl_b200 as nb
from nabla_jl_b200 import workloads as wl
ctx = nb.get_context()
dims = (784, 2048, 2048, 2048, 2048, 10)
B = 2048
W, b, X, Y = wl.mlp_inputs(dims, B, np.float32, seed=4)
params = [ctx.asarray(p) for p in (*W, *b)]
Xd, Yd = ctx.asarray(X), ctx.asarray(Y)
comm = nb.Communicator(ctx, 0, 1, None)
sg = nb.OverlappedShardedGradient(wl.mlp_data_objective(nb, Xd, Yd, 5), wl.mlp_prior_objective(nb, 1e-3, 5), params, comm)
gc.disable()
for i in range(3):
    sg.step(); ctx.sync()
    print("step", i, "live without gc:", sum(1 for r in ctx._deferred if r() is not None and not r().done), flush=True)
live = [r() for r in ctx._deferred if r() is not None and not r().done]
print("live", len(live))
skip = {id(live)}
def name(o):
    t = type(o).__name__
    if isinstance(o, types.FunctionType):
        return f"function {o.__qualname__}"
    if isinstance(o, types.FrameType):
        return f"frame {o.f_code.co_name}"
    if isinstance(o, dict):
        return "dict keys=" + str(list(o.keys())[:5])
    if isinstance(o, (list, tuple)):
        return f"{t} len={len(o)}"
    if t == "cell":
        return "cell"
    if t in ("Branch", "Leaf"):
        return f"{t} pos={getattr(o, 'pos', None)} f={getattr(getattr(o, 'f', None), '__name__', getattr(o, 'f', None))}"
    if t == "DevArray":
        return f"DevArray {o.shape} pending={type(o.pending).__name__ if o.pending is not None else None}"
    if t == "Deferred":
        return f"Deferred kind={o.kind} stage={o.stage} {o.m}x{o.n}x{o.k} out_alive={o.out_ref() is not None if o.out_ref else None}"
    return t
def walk(o, depth, seen):
    if depth > 7 or id(o) in seen:
        return
    seen.add(id(o))
    for ref in gc.get_referrers(o):
        if id(ref) in skip or isinstance(ref, types.FrameType) or type(ref).__name__ in ("list_iterator",):
            continue
        if isinstance(ref, dict) and "__name__" in ref and ref.get("__name__") == "__main__":
            continue
        print("  " * depth + "<- " + name(ref), flush=True)
        walk(ref, depth + 1, seen)
for d in live[:2]:
    print("==", name(d))
    walk(d, 1, set())

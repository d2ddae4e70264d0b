"""Debug driver: Float64 MLP gradient (data term + shared term, get_output) at a given shape; prints ok or the error."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nabla_jl_b200 as nb
from nabla_jl_b200 import workloads as wl

width, B = int(sys.argv[1]), int(sys.argv[2])
dt = np.float64 if len(sys.argv) < 4 or sys.argv[3] == "f64" else np.float32
ctx = nb.get_context()
dims = (784, width, width, 10)
W, b, X, Y = wl.mlp_inputs(dims, B, dt, seed=4)
Xd, Yd = ctx.asarray(X), ctx.asarray(Y)
params = [ctx.asarray(p) for p in (*W, *b)]
y, g = nb.nabla(wl.mlp_data_objective(nb, Xd, Yd, 3), get_output=True)(*params)
print("data objective", float(nb.to_host(y)), flush=True)
y, g = nb.nabla(wl.mlp_prior_objective(nb, 1e-3, 3), get_output=True)(*params)
print("prior objective", float(nb.to_host(y)), flush=True)
y, g = nb.nabla(wl.mlp_objective(nb, Xd, Yd, 1e-3, 3), get_output=True)(*params)
print("full objective", float(nb.to_host(y)), [float(np.linalg.norm(x.numpy())) for x in g][:3], flush=True)
print("ok")

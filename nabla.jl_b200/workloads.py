"""The BASELINE configurations, written ONCE against an ops namespace ``nb`` so the same definition
runs on the device backend (``nb = nabla_jl_b200``) and on the CPU oracle (``nb = nabla_oracle``).

C1/C5  mlp_log_joint         examples/mlp.jl:56-82
C4     vae_log_joint         examples/vae.jl:32-39, 67-90
C2     quadratic forms       examples/symmetric_quadratic.jl:15, asymmetric_quadratic.jl:15
C3     bcast_sum             sum(tanh.(A .+ b))  

Synthetic inputs follow BASELINE.md (NumPy ``default_rng(seed)``)."""
import numpy as np


def logistic_of(nb):
    """``logistic(x) = 1 / (1 + exp(-x))`` (examples/vae.jl:32, mlp.jl:42) as a user lambda: the
    reference differentiates it with ForwardDiff duals; the device traces it to a program."""
    return lambda x: 1.0 / (1.0 + nb.exp(-x))


def apply_transforms(nb, X, W, b, f):
    """examples/mlp.jl:79-82 — one Branch per `*`, `.+` and `f.` (no dot fusion)."""
    h = X
    for Wn, bn in zip(W, b):
        h = nb.broadcast(f, nb.broadcast(nb.add, nb.matmul(Wn, h), bn))
    return h


def mlp_log_joint(nb, X, Y, W, b, lam, eps=1e-15):
    """examples/mlp.jl:56-77.  ``1 - Y`` (line 73) is read as ``1 .- Y``; ``Y`` is data, not a Node, so
    that broadcast is a plain forward computation re-done every evaluation, as in the reference."""
    log_prior = 0.0
    for Wn, bn in zip(W, b):
        log_prior = log_prior + (nb.sum(nb.abs2, Wn) + nb.sum(nb.abs2, bn))
    log_prior = log_prior * (-0.5 * lam)
    f = nb.broadcast(logistic_of(nb), apply_transforms(nb, X, W, b, nb.tanh))
    f = nb.broadcast(nb.div, f, nb.sum(f, dims=1))
    one_minus_Y = nb.broadcast(nb.sub, 1.0, Y)
    t1 = nb.broadcast(nb.mul, Y, nb.broadcast(nb.log, nb.broadcast(nb.add, f, eps)))
    t2 = nb.broadcast(nb.mul, one_minus_Y, nb.broadcast(nb.log, nb.broadcast(nb.sub, 1.0 + eps, f)))
    log_lik = nb.sum(nb.broadcast(nb.add, t1, t2))
    return log_lik + log_prior, f


def mlp_data_objective(nb, X, Y, nlayers, eps=1e-15):
    """The batch-dependent part of ``mlp_log_joint``: the log-likelihood (examples/mlp.jl:67-73)."""
    def obj(*params):
        W, b = params[:nlayers], params[nlayers:]
        f = nb.broadcast(logistic_of(nb), apply_transforms(nb, X, W, b, nb.tanh))
        f = nb.broadcast(nb.div, f, nb.sum(f, dims=1))
        one_minus_Y = nb.broadcast(nb.sub, 1.0, Y)
        t1 = nb.broadcast(nb.mul, Y, nb.broadcast(nb.log, nb.broadcast(nb.add, f, eps)))
        t2 = nb.broadcast(nb.mul, one_minus_Y, nb.broadcast(nb.log, nb.broadcast(nb.sub, 1.0 + eps, f)))
        return nb.sum(nb.broadcast(nb.add, t1, t2))
    return obj


def mlp_prior_objective(nb, lam, nlayers):
    """The batch-independent part of ``mlp_log_joint``: the log-prior (examples/mlp.jl:61-65)."""
    def obj(*params):
        W, b = params[:nlayers], params[nlayers:]
        log_prior = 0.0
        for Wn, bn in zip(W, b):
            log_prior = log_prior + (nb.sum(nb.abs2, Wn) + nb.sum(nb.abs2, bn))
        return log_prior * (-0.5 * lam)
    return obj


def mlp_objective(nb, X, Y, lam, nlayers, prior_scale=1.0, eps=1e-15):
    """Scalar objective of the parameters (W1..Wn, b1..bn) for ``nabla``.  ``prior_scale`` = 1/G
    when the batch is sharded over G ranks so the allreduce-sum counts the prior once
    (SURVEY.md §7 hard part 7)."""
    def obj(*params):
        W, b = params[:nlayers], params[nlayers:]
        return mlp_log_joint(nb, X, Y, W, b, lam * prior_scale, eps)[0]
    return obj


def vae_log_joint(nb, X, noise, Wh, Wmu, Wsig, Vh, Vf, lam, scal, dz, eps=1e-12, prior_scale=1.0):
    """examples/vae.jl:67-90 with L = 1.  ``scal`` uses the GLOBAL batch size; ``prior_scale`` = 1/G
    under batch sharding (the allreduce-sum then counts the prior once)."""
    logistic = logistic_of(nb)
    logprior = (-0.5 * lam * prior_scale) * (nb.sum(nb.abs2, Wh) + nb.sum(nb.abs2, Wmu)
                                               + nb.sum(nb.abs2, Wsig) + nb.sum(nb.abs2, Vh)
                                               + nb.sum(nb.abs2, Vf))
    h_enc = nb.broadcast(nb.tanh, nb.matmul(Wh, X))                       # encode, vae.jl:34-37
    mu = nb.matmul(Wmu, h_enc)
    sig = nb.broadcast(nb.exp, nb.matmul(Wsig, h_enc))
    z = nb.broadcast(nb.add, mu, nb.broadcast(nb.mul, sig, noise))       # vae.jl:77
    f = nb.broadcast(logistic, nb.matmul(Vf, nb.broadcast(nb.tanh, nb.matmul(Vh, z))))  # decode :38
    one_minus_X = nb.broadcast(nb.sub, 1.0, X)
    t1 = nb.broadcast(nb.mul, X, nb.broadcast(nb.log, nb.broadcast(nb.add, f, eps)))
    t2 = nb.broadcast(nb.mul, one_minus_X, nb.broadcast(nb.log, nb.broadcast(nb.sub, 1.0 + eps, f)))
    loglik = 0.0 + nb.sum(nb.broadcast(nb.add, t1, t2))
    loglik = loglik / 1.0
    sig2 = nb.broadcast(nb.mul, sig, sig)                                  # σz .* σz, vae.jl:85
    # klvae, vae.jl:39; the constant D is batch-independent, so a shard carries 1/G of it
    klqp = 0.5 * (nb.sum(sig2) + nb.sum(nb.abs2, mu) - float(dz) * prior_scale
                  - nb.sum(nb.broadcast(nb.log, sig2)))
    elbo = loglik - klqp
    return scal * elbo + logprior


def vae_data_objective(nb, X, noise, scal, dz, eps=1e-12):
    """The batch-dependent part of ``vae_log_joint`` for THIS shard's columns of X and of the noise:
    ``scal * (loglik - klqp_data)`` with ``scal`` from the GLOBAL batch (examples/vae.jl:59, 77-90)."""
    logistic = logistic_of(nb)

    def obj(Wh, Wmu, Wsig, Vh, Vf):
        h_enc = nb.broadcast(nb.tanh, nb.matmul(Wh, X))
        mu = nb.matmul(Wmu, h_enc)
        sig = nb.broadcast(nb.exp, nb.matmul(Wsig, h_enc))
        z = nb.broadcast(nb.add, mu, nb.broadcast(nb.mul, sig, noise))
        f = nb.broadcast(logistic, nb.matmul(Vf, nb.broadcast(nb.tanh, nb.matmul(Vh, z))))
        one_minus_X = nb.broadcast(nb.sub, 1.0, X)
        t1 = nb.broadcast(nb.mul, X, nb.broadcast(nb.log, nb.broadcast(nb.add, f, eps)))
        t2 = nb.broadcast(nb.mul, one_minus_X, nb.broadcast(nb.log, nb.broadcast(nb.sub, 1.0 + eps, f)))
        loglik = nb.sum(nb.broadcast(nb.add, t1, t2))
        sig2 = nb.broadcast(nb.mul, sig, sig)
        kl = 0.5 * (nb.sum(sig2) + nb.sum(nb.abs2, mu) - nb.sum(nb.broadcast(nb.log, sig2)))
        return scal * (loglik - kl)
    return obj


def vae_shared_objective(nb, lam, scal, dz):
    """The batch-independent part: the log-prior (examples/vae.jl:70-71) and the constant ``D`` of klvae
    (examples/vae.jl:39), which enters the objective as ``+ scal * D / 2``."""
    def obj(Wh, Wmu, Wsig, Vh, Vf):
        logprior = (-0.5 * lam) * (nb.sum(nb.abs2, Wh) + nb.sum(nb.abs2, Wmu) + nb.sum(nb.abs2, Wsig)
                                   + nb.sum(nb.abs2, Vh) + nb.sum(nb.abs2, Vf))
        return logprior + (0.5 * scal * float(dz))
    return obj


def vae_objective(nb, X, noise, lam, scal, dz, prior_scale=1.0):
    def obj(*params):
        return vae_log_joint(nb, X, noise, *params, lam=lam, scal=scal, dz=dz, prior_scale=prior_scale)
    return obj


def shard_columns(a, rank, nranks):
    """Batch (column) shard of a (d x B) array for rank ``rank`` of ``nranks``:
    contiguous, near-equal column ranges in rank order."""
    B = a.shape[-1]
    lo, hi = (B * rank) // nranks, (B * (rank + 1)) // nranks
    return a[..., lo:hi]


def quad_symmetric(nb, A):
    """``f(x) = x' * (A * x)`` (examples/symmetric_quadratic.jl:15); ∇ = (A + A')x."""
    return lambda x: nb.matmul(nb.adjoint(x), nb.matmul(A, x))


def quad_asymmetric(nb, A):
    """``f(x, y) = x' * (A * y)`` (examples/asymmetric_quadratic.jl:15); ∇ = (A y, A' x)."""
    return lambda x, y: nb.matmul(nb.adjoint(x), nb.matmul(A, y))


def bcast_sum(nb):
    """C3: ``s = sum(tanh.(A .+ b))``."""
    return lambda A, b: nb.sum(nb.broadcast(nb.tanh, nb.broadcast(nb.add, A, b)))


# ---- synthetic inputs ---------------------------------------------------------------
def mlp_inputs(dims, B, dtype=np.float64, seed=0, w_scale=0.1):
    rng = np.random.default_rng(seed)
    W = [(w_scale * rng.standard_normal((dims[i + 1], dims[i]))).astype(dtype) for i in range(len(dims) - 1)]
    b = [(w_scale * rng.standard_normal(dims[i + 1])).astype(dtype) for i in range(len(dims) - 1)]
    X = rng.random((dims[0], B)).astype(dtype)
    labels = rng.integers(0, dims[-1], B)
    Y = np.zeros((dims[-1], B), dtype=dtype)
    Y[labels, np.arange(B)] = 1.0
    return W, b, X, Y


def vae_inputs(dx=784, dz=10, dh=500, B=4096, dtype=np.float64, seed=3):
    rng = np.random.default_rng(seed)
    Wh = (0.1 * rng.standard_normal((dh, dx))).astype(dtype)
    Wmu = (0.1 * rng.standard_normal((dz, dh))).astype(dtype)
    Wsig = (0.1 * rng.standard_normal((dz, dh))).astype(dtype)
    Vh = (0.1 * rng.standard_normal((dh, dz))).astype(dtype)
    Vf = (0.1 * rng.standard_normal((dx, dh))).astype(dtype)
    X = rng.random((dx, B)).astype(dtype)
    noise = rng.standard_normal((dz, B)).astype(dtype)
    return (Wh, Wmu, Wsig, Vh, Vf), X, noise


def quad_inputs(N, dtype=np.float64, seed=1):
    rng = np.random.default_rng(seed)
    Bm = rng.standard_normal((N, N))
    A = (Bm.T @ Bm + 1e-6 * np.eye(N)).astype(dtype)
    return A, rng.standard_normal(N).astype(dtype), rng.standard_normal(N).astype(dtype)


def c3_inputs(m, n, dtype=np.float64, seed=2, row_vector=False):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((m, n)).astype(dtype)
    b = rng.standard_normal((1, n) if row_vector else (m,)).astype(dtype)
    return A, b

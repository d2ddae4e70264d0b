"""Error of the Float64-on-int8 split for different slicings, in NumPy (no GPU): 7-bit slices x 6 / x 7 (csrc/ozaki.cu)
against signed 8-bit slices x 6 with a carry pass (candidate: 21 passes at close to the 28-pass accuracy).
Prints the norm-wise relative error of A @ B for Gaussian operands."""
import numpy as np


def scales(X, axis, extra):
    mx = np.max(np.abs(X), axis=axis, keepdims=True)
    return np.ldexp(1.0, np.frexp(mx)[1] + 1 + extra)      # |x| / scale < 1/2 (extra = 0) or < 1/4 (extra = 1)


def slices(X, S, bits):
    out = []
    f = float(1 << bits)
    for _ in range(S):
        t = X * f
        q = np.rint(t)
        out.append(q.astype(np.int64))
        X = t - q
    if bits == 8:                                           # carry pass: bring every slice into [-128, 127]
        for p in range(S - 1, 0, -1):
            over = out[p] >= 128
            out[p] = np.where(over, out[p] - 256, out[p])
            out[p - 1] = out[p - 1] + over
        assert all(q.min() >= -128 and q.max() <= 127 for q in out)
    return out


def product(A, B, S, bits):
    extra = 1 if bits == 8 else 0                           # 8-bit: |x| < 1/4 keeps slice 0 (plus a carry) in range
    ra, cb = scales(A, 1, extra), scales(B, 0, extra)
    sa, sb = slices(A / ra, S, bits), slices(B / cb, S, bits)
    C = np.zeros((A.shape[0], B.shape[1]))
    for l in range(S):
        acc = sum(sa[p] @ sb[l - p] for p in range(l + 1))
        C += np.ldexp(acc.astype(np.float64), -bits * (l + 2))
    return C * ra * cb


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    m = n = 192
    for k in (512, 4096):
        A, B = rng.standard_normal((m, k)), rng.standard_normal((k, n))
        ref = np.asarray(A.astype(np.longdouble) @ B.astype(np.longdouble), dtype=np.float64)
        for S, bits in ((6, 7), (7, 7), (6, 8), (7, 8)):
            C = product(A, B, S, bits)
            print(f"k={k} S={S} bits={bits} passes={S * (S + 1) // 2}: rel err {np.linalg.norm(C - ref) / np.linalg.norm(ref):.2e}")

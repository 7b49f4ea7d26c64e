"""CPU ORACLE (test infrastructure, NOT product code) for the empirical-Bayes callers of the
likelihood matrix: the alternative L builder and the regulariser
gradients.  Plain numpy restatement; every function cites the reference file:line it follows
(paths relative to the reference repo root).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import this.

Unlike the forward solve, the arithmetic of these functions is IN the reference tree (pure Julia,
BLAS-level), so the restatement follows the source operation by operation (same binomial formula,
same loops).  The reference's tests hold no numeric vectors for them (test/priorestimation.jl:17-29
only calls them); what pins this oracle are the reference's own internal consistency checks,
restated in tests/test_eb_oracle.py:
  * dhztest (eb/regularizers.jl:157-170): dhzmatrix == dhzloop == dhzloop2 == gradient of hz;
  * projectsimplex_heap! == projectsimplex_sort! (eb/projectsimplex.jl:11-46);
  * likelihoodmat_nanfast(renormalize=false) against the element-wise pdf form it replaced
    (eb/likelihoodmat.jl:9, utils.jl:89-101) up to the documented normalisation;
  * emiteration!(w, L) == w .* dlogl / M (em.jl:27-41 vs regularizers.jl:7-10).
"""
from __future__ import annotations

import numpy as np


# ------------------------------------------------------------------------------ likelihood matrix
def sqeucdist_nan(A, B):
    """eb/likelihoodmat.jl:61-95 — squared Euclidean distances between the columns of A (D x I) and
    B (D x J), NaN entries ignored; the reference's binomial formula with the NaN corrections."""
    A = np.array(A, dtype=np.float64)
    B = np.array(B, dtype=np.float64)
    An, Bn = np.isnan(A), np.isnan(B)
    A[An] = 0.0                                                   # :68-69
    B[Bn] = 0.0
    sa2 = (A * A).sum(axis=0)                                     # :70-71
    sb2 = (B * B).sum(axis=0)
    ca = An.astype(np.float64).T @ (B * B)                        # :75-78   ca[i,:] += abs2(B[k,:]) for NaN A[k,i]
    ca += (A * A).T @ Bn.astype(np.float64)                       # :80-83
    r = A.T @ B                                                   # :85
    return sa2[:, None] + sb2[None, :] - 2.0 * r - ca             # :87-92


def sqeucdist_nan_direct(A, B):
    """The same quantity by its definition (sum over jointly observed entries of (a-b)^2) — used to
    bound the cancellation error of the binomial form when comparing with the device kernel."""
    A = np.asarray(A, dtype=np.float64)
    B = np.asarray(B, dtype=np.float64)
    out = np.empty((A.shape[1], B.shape[1]))
    for j in range(B.shape[1]):
        d = A - B[:, j:j + 1]
        out[:, j] = np.nansum(d * d, axis=0)
    return out


def likelihoodmat_nanfast(xs, ys, sigmas, renormalize=True):
    """eb/likelihoodmat.jl:34-58 for d::MatrixNormalCentered(sigmas).  xs, ys: sequences of equally
    shaped matrices (NaN = missing).  Returns the len(xs) x len(ys) matrix.  Note :47 — the
    normalisation is sqrt(prod(sigma[mask]) * (2 pi)^count), with sigma (not sigma^2) under the root,
    exactly as in the source."""
    sig = np.asarray(sigmas, dtype=np.float64)
    svec = sig.ravel(order="F")
    A = np.stack([(np.asarray(x, dtype=np.float64) / sig).ravel(order="F") for x in xs], axis=1)   # :35
    B = np.stack([(np.asarray(y, dtype=np.float64) / sig).ravel(order="F") for y in ys], axis=1)   # :36
    sq = sqeucdist_nan(A, B)                                                                        # :37
    norm = np.ones_like(sq)                                                                         # :40
    if renormalize:
        Am, Bm = ~np.isnan(A), ~np.isnan(B)
        for j in range(B.shape[1]):
            for i in range(A.shape[1]):
                m = Am[:, i] & Bm[:, j]
                norm[i, j] = np.sqrt(np.prod(svec[m]) * (2 * np.pi) ** int(m.sum()))                # :47-50
        sq = sq - sq.min()                                                                          # :53
        norm = norm / norm.max()                                                                    # :54
    return np.exp(-sq / 2) / norm                                                                   # :57


def matrixnormal_logpdf(sigmas, x):
    """utils.jl:93-101 — logpdf(MatrixNormalCentered(sigmas), x), NaN entries skipped."""
    s = np.asarray(sigmas, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    ok = ~np.isnan(x)
    return float(np.sum(-0.5 * (x[ok] / s[ok]) ** 2 - np.log(s[ok]) - 0.5 * np.log(2 * np.pi)))


# ------------------------------------------------------------------------------ regularisers
def logl(w, L):
    """eb/regularizers.jl:5 — logl(wx, Ldx) = sum(log(Ldx * wx)); Ldx is (data x samples)."""
    return float(np.sum(np.log(L @ w)))


def dlogl(w, L):
    """eb/regularizers.jl:7-10 — sum(L ./ (L*w), 1); L is (data x samples)."""
    return (L / (L @ w)[:, None]).sum(axis=0)


def repeatweights(w, nz):
    """eb/regularizers.jl:16-19."""
    zmult = nz // len(w)
    return np.tile(w, zmult) / zmult


def hz(wx, Lzx, wz=None):
    """eb/regularizers.jl:30-45 — -sum_z wz * log(rhoz), rhoz = Lzx*wx, zero rows skipped."""
    wz = wx if wz is None else wz
    rhoz = Lzx @ wx
    l = 0.0
    for r, w in zip(rhoz, wz):
        if r == 0:
            continue
        l -= np.log(r) * w
    return l


def dhzdiscr(w, L, zmult=None):
    """eb/regularizers.jl:59-79 given L = likelihoodmat(zs, ys) (Z x K, Z = K*zmult)."""
    Z, K = L.shape
    zmult = Z // K if zmult is None else zmult
    wz = repeatweights(w, Z)
    rhoz = L @ w                                                    # :65
    factors = wz / rhoz                                             # :66
    d = np.zeros(K)
    for k in range(K):
        d[k] -= float(np.dot(L[:, k], factors))                     # :70-72
        for m in range(zmult):
            d[k] -= np.log(rhoz[k + K * m]) / zmult                 # :73-75
    if np.any(np.isnan(d)):
        raise FloatingPointError("encountered NaN in dhz")          # :77
    return d


def dhzmatrix(L, w, wz):
    """eb/regularizers.jl:132-135."""
    rhoz = L @ w
    return -(((wz * np.log(rhoz) / rhoz)[:, None] * L).sum(axis=0) + 1.0)


def dhzloop(L, w, wz):
    """eb/regularizers.jl:137-147."""
    rhoz = L @ w
    d = -np.ones_like(w)
    for x in range(len(w)):
        for z in range(len(wz)):
            d[x] -= L[z, x] / rhoz[z] * wz[z] * np.log(rhoz[z])
    return d


def dhzloop2(L, w, wz):
    """eb/regularizers.jl:149-160 (the body of dhzcont, :50-57)."""
    rhoz = L @ w
    d = -np.ones_like(w)
    for z in range(len(wz)):
        fact = wz[z] * np.log(rhoz[z]) / rhoz[z]
        d -= fact * L[z, :]
    return d


def HKL(w, L, wz=None):
    """eb/regularizers.jl:194-204."""
    wz = w if wz is None else wz
    ev = L @ w
    h = 0.0
    for k in range(L.shape[1]):
        for l in range(L.shape[0]):
            h += wz[l] * L[l, k] * w[k] / ev[l] * np.log(L[l, k] / ev[l])
    return h / L.shape[0] / L.shape[1]


# ------------------------------------------------------------------------------ simplex projection / ascent
def projectsimplex_sort(y):
    """eb/projectsimplex.jl:29-46 (algorithm 1, sort based)."""
    y = np.array(y, dtype=np.float64)
    temp = np.sort(y)[::-1]
    cumsum, t = 0.0, 0.0
    for k in range(1, len(y) + 1):
        uk = temp[k - 1]
        cumsum += uk
        normalized = (cumsum - 1.0) / k
        if normalized >= uk:
            break
        t = normalized
    return np.maximum(y - t, 0.0)


def projectsimplex_heap(y):
    """eb/projectsimplex.jl:11-26 (algorithm 2, heap based; the default, :8)."""
    import heapq
    y = np.array(y, dtype=np.float64)
    heap = [-v for v in y]
    heapq.heapify(heap)
    cumsum, t = 0.0, 0.0
    for k in range(1, len(y) + 1):
        uk = -heapq.heappop(heap)
        cumsum += uk
        normalized = (cumsum - 1.0) / k
        if normalized >= uk:
            break
        t = normalized
    return np.maximum(y - t, 0.0)


projectsimplex = projectsimplex_heap


def psi_S(w, gradient):
    """eb/projectsimplex.jl:48-53."""
    return w + w * gradient - w * np.dot(w, gradient)


def movefromboundary(wold, wnew, relstep=0.5):
    """eb/gradientascent.jl:19-25."""
    if np.any(wnew == 0):
        return (1 - relstep) * wold + relstep * wnew
    return wnew


def gradientascent(df, w0, n, h, projection=projectsimplex):
    """eb/gradientascent.jl:1-5 with autodiff=false: [w0, iter(w0), ...] (n entries)."""
    out = []
    w = np.array(w0, dtype=np.float64)
    for _ in range(n):
        out.append(w)
        w = movefromboundary(w, projection(w + h * df(w)))
    return out


def dmple_obj(Ldata_t, Lz, reg):
    """eb/likelihoodmodel.jl:56-64 — gradient of reg*hz + (1-reg)*logl.  Ldata_t is (data x samples)."""
    if reg == 0:
        return lambda w: dlogl(w, Ldata_t)
    if reg == 1:
        return lambda w: dhzdiscr(w, Lz)
    return lambda w: reg * dhzdiscr(w, Lz) + (1 - reg) * dlogl(w, Ldata_t)


def mple_stepsize(h, reg, ndata):
    """eb/likelihoodmodel.jl:33-39 (autoh)."""
    c1, c2 = 1 / 100, 1 / 1000
    return h / ((1 - reg) * (ndata / c1) + reg / c2)

"""Host-side mirror of the reference's empirical-Bayes callers of the likelihood matrix (src/eb), over the C ABI.

Names, argument meaning and error behaviour follow the reference (paths relative to its repo):
  MatrixNormalCentered                 src/gync/utils.jl:81-105
  likelihoodmat / likelihoodmat_nanfast  src/eb/likelihoodmat.jl:6-58 (sqeucdist_nan :61-95)
  LikelihoodModel, em, mple, uniformweights   src/eb/likelihoodmodel.jl:3-64
  logl, dlogl, repeatweights, hz, dhz   src/eb/regularizers.jl:3-79
  projectsimplex(!), psi_S             src/eb/projectsimplex.jl:5-53
  gradientascent, movefromboundary     src/eb/gradientascent.jl:1-25
  gyncmodel, phi                       src/eb/gync.jl:3-25, regularizers.jl:215
All numerics run in the CUDA library (csrc/gync_eb.cuh); nothing here falls back to the CPU.  The two likelihood
matrices of a LikelihoodModel are built once on the device and stay in HBM (the reference memoises them, :8).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import api as _api
from . import lib as _L

model_measerrors = _api.model_measmagnitude / 10.0            # model.jl:8


class MatrixNormalCentered:
    """utils.jl:81-105 — independent centred normals with a matrix of standard deviations."""

    def __init__(self, sigmas):
        self.sigmas = np.array(sigmas, dtype=np.float64)

    def rand(self, rng):
        return rng.standard_normal(self.sigmas.shape) * self.sigmas

    def __eq__(self, other):
        return isinstance(other, MatrixNormalCentered) and np.array_equal(self.sigmas, other.sigmas)

    def __hash__(self):
        return hash(self.sigmas.tobytes())


def _cols(mats, shape=None):
    """hcat([vec(x) for x in xs]...) (likelihoodmat.jl:35): D x n, column-major."""
    if isinstance(mats, np.ndarray) and mats.ndim == 2 and shape is None:
        return _L.fcol(mats)
    out = np.empty((int(np.prod(np.shape(mats[0]))), len(mats)), order="F")
    for i, x in enumerate(mats):
        out[:, i] = np.asarray(x, dtype=np.float64).ravel(order="F")
    return out


def likelihoodmat_nanfast(xs, ys, d: MatrixNormalCentered, renormalize=True):
    """likelihoodmat.jl:34-58: len(xs) x len(ys) matrix of N(x - y; 0, sigmas) over the jointly observed entries."""
    lib = _L.load()
    A, B = _cols(xs), _cols(ys)
    sig = np.ascontiguousarray(d.sigmas.ravel(order="F"))
    if A.shape[0] != sig.size or B.shape[0] != sig.size:
        raise ValueError("size mismatch between the samples and d.sigmas (utils.jl:94)")
    Lm = np.empty((A.shape[1], B.shape[1]), order="F")
    _L.check(lib.gync_likelihoodmat(_L.dptr(A), A.shape[1], _L.dptr(B), B.shape[1], sig.size, _L.dptr(sig),
                                    1 if renormalize else 0, _L.dptr(Lm)))
    return Lm


def likelihoodmat(zs, ys, d):
    """likelihoodmat.jl:6-14 (the memoisation lives in LikelihoodModel here)."""
    if not isinstance(d, MatrixNormalCentered):
        raise NotImplementedError("only MatrixNormalCentered measurement errors run on the device "
                                  "(the reference's fallback, likelihoodmat.jl:29-32, is element-wise pdf calls)")
    return likelihoodmat_nanfast(zs, ys, d)


def uniformweights(n):
    """likelihoodmodel.jl:22-24."""
    n = n if isinstance(n, (int, np.integer)) else (len(n.xs) if isinstance(n, LikelihoodModel) else len(n))
    return np.full(n, 1.0 / n)


def repeatweights(w, zs):
    """regularizers.jl:16-19."""
    zmult = len(zs) // len(w)
    return np.tile(w, zmult) / zmult


class LikelihoodModel:
    """likelihoodmodel.jl:3-14.  xs: samples, ys = phi.(xs), zs: ys + measurement noise (zmult copies), datas."""

    def __init__(self, xs, ys, zs, datas, measerr, zsampledistr=None):
        self.xs, self.ys, self.zs, self.datas = xs, ys, zs, datas
        self.measerr = measerr
        self.zsampledistr = measerr if zsampledistr is None else zsampledistr
        self._h = None

    @classmethod
    def from_matrices(cls, Ld, Lz=None):
        """A model over precomputed matrices: Ld = likelihoodmat(ys, datas) (K x M), Lz = likelihoodmat(zs, ys) (Z x K)."""
        m = cls.__new__(cls)
        Ld = None if Ld is None else _L.fcol(Ld)
        Lz = None if Lz is None else _L.fcol(Lz)
        K = Ld.shape[0] if Ld is not None else Lz.shape[1]
        m.xs = m.ys = list(range(K))
        m.zs = list(range(0 if Lz is None else Lz.shape[0]))
        m.datas = list(range(0 if Ld is None else Ld.shape[1]))
        m.measerr = m.zsampledistr = None
        h = C.c_void_p()
        _L.check(_L.load().gync_ebmodel_from_matrices(_L.dptr(Ld), K, len(m.datas), _L.dptr(Lz), len(m.zs), C.byref(h)))
        m._h = h
        return m

    def handle(self):
        if self._h is None:
            lib = _L.load()
            if not isinstance(self.measerr, MatrixNormalCentered):
                raise NotImplementedError("only MatrixNormalCentered measurement errors run on the device")
            Y, Dm = _cols(self.ys), _cols(self.datas)
            Zm = _cols(self.zs) if len(self.zs) else None
            sm = np.ascontiguousarray(self.measerr.sigmas.ravel(order="F"))
            sz = np.ascontiguousarray(self.zsampledistr.sigmas.ravel(order="F"))
            h = C.c_void_p()
            _L.check(lib.gync_ebmodel_create(_L.dptr(Y), Y.shape[1], _L.dptr(Dm), Dm.shape[1], _L.dptr(Zm),
                                             0 if Zm is None else Zm.shape[1], sm.size, _L.dptr(sm), _L.dptr(sz), C.byref(h)))
            self._h = h
        return self._h

    def matrices(self):
        """(Ld, Lz) copied back from the device."""
        lib = _L.load()
        h = self.handle()
        K, M, Z = len(self.ys), len(self.datas), len(self.zs)
        Ld = np.empty((K, M), order="F") if M else None
        Lz = np.empty((Z, K), order="F") if Z else None
        _L.check(lib.gync_ebmodel_matrices(h, _L.dptr(Ld), _L.dptr(Lz)))
        return Ld, Lz

    def close(self):
        if self._h is not None:
            _L.load().gync_ebmodel_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _w(w):
    return np.ascontiguousarray(w, dtype=np.float64)


def logl(m: LikelihoodModel, w):
    """logl(m, w) = sum(log(L*w)), L = likelihoodmat(datas, ys) (regularizers.jl:3-5, likelihoodmodel.jl:72)."""
    out = C.c_double()
    _L.check(_L.load().gync_ebmodel_logl(m.handle(), _L.dptr(_w(w)), C.byref(out)))
    return out.value


def dlogl(m: LikelihoodModel, w):
    """regularizers.jl:7-10, likelihoodmodel.jl:73."""
    d = np.empty(len(m.ys))
    _L.check(_L.load().gync_ebmodel_dlogl(m.handle(), _L.dptr(_w(w)), _L.dptr(d)))
    return d


def hz(m: LikelihoodModel, w, wz=None):
    """hz(m, w) (likelihoodmodel.jl:69) = hz(w, L, repeatweights(w, zs)) (regularizers.jl:24-45)."""
    w = _w(w)
    if not np.all(np.isfinite(w)):
        raise AssertionError("all(isfinite(wx))")                   # regularizers.jl:33
    out = C.c_double()
    wz = None if wz is None else _w(wz)
    if wz is not None and wz.size != len(m.zs):
        raise AssertionError("size(Lzx, 1) == length(wz)")          # :31
    _L.check(_L.load().gync_ebmodel_hz(m.handle(), _L.dptr(w), _L.dptr(wz), C.byref(out)))
    return out.value


def dhz(m: LikelihoodModel, w, cont=False):
    """dhz(m, w) (likelihoodmodel.jl:70): dhzdiscr (regularizers.jl:59-79; DHZCONT = false, :47) or dhzcont (:50-57).
    Raises on NaN like the reference (:77)."""
    d = np.empty(len(m.ys))
    _L.check(_L.load().gync_ebmodel_dhz(m.handle(), _L.dptr(_w(w)), 1 if cont else 0, _L.dptr(d)))
    return d


def em(m: LikelihoodModel, w0, niter):
    """em(m, w0, niter) (likelihoodmodel.jl:17-20) = emiterations(w0, L, niter): [w0, ..., em^(niter-1)(w0)]."""
    w0 = np.array(w0, dtype=np.float64)
    if niter <= 1:
        return [w0][:max(niter, 0)]
    hist = np.empty((w0.size, niter - 1), order="F")
    w = w0.copy()
    _L.check(_L.load().gync_ebmodel_em(m.handle(), _L.dptr(w), niter - 1, _L.dptr(hist)))
    return [w0] + [hist[:, i].copy() for i in range(niter - 1)]


def projectsimplex_bang(y):
    """projectsimplex!(y) (projectsimplex.jl:8): Euclidean projection onto the unit simplex, in place."""
    if not (isinstance(y, np.ndarray) and y.dtype == np.float64 and y.flags.c_contiguous):
        raise TypeError("y must be a contiguous float64 vector (it is updated in place)")
    _L.check(_L.load().gync_projectsimplex(_L.dptr(y), y.size))
    return y


def projectsimplex(y):
    """projectsimplex.jl:5."""
    return projectsimplex_bang(np.array(y, dtype=np.float64))


def psi_S(w, gradient):
    """projectsimplex.jl:48-53 (three vector operations; host)."""
    w, gradient = np.asarray(w, float), np.asarray(gradient, float)
    return w + w * gradient - w * np.dot(w, gradient)


def movefromboundary(wold, wnew, relstep=0.5):
    """gradientascent.jl:19-25."""
    return (1 - relstep) * wold + relstep * wnew if np.any(wnew == 0) else wnew


def gradientascent(f, w0, n, h, projection=projectsimplex, autodiff=False):
    """gradientascent.jl:1-5 for an arbitrary gradient callable (host loop, projection on the device).
    mple() below runs the whole loop on the device for the LikelihoodModel objectives."""
    if autodiff:
        raise NotImplementedError("ForwardDiff gradients (autodiff=true) are out of scope; pass the gradient (autodiff=false)")
    out, w = [], np.array(w0, dtype=np.float64)
    for _ in range(n):
        out.append(w)
        w = movefromboundary(w, projection(w + h * f(w)))
    return out


def dmple_obj(m: LikelihoodModel, reg):
    """likelihoodmodel.jl:56-64."""
    if reg == 0:
        return lambda w: dlogl(m, w)
    if reg == 1:
        return lambda w: dhz(m, w)
    return lambda w: reg * dhz(m, w) + (1 - reg) * dlogl(m, w)


def mple(m: LikelihoodModel, reg, h, niter, autoh=True, w0=None, cont=False):
    """mple(m, reg, h, niter; autoh, w0) (likelihoodmodel.jl:33-41): projected gradient ascent on
    reg*hz + (1-reg)*logl, returning the niter iterates [w0, ...].  Device-resident loop."""
    w0 = uniformweights(len(m.xs)) if w0 is None else np.array(w0, dtype=np.float64)
    if autoh:
        c1, c2 = 1 / 100, 1 / 1000
        h = h / ((1 - reg) * (len(m.datas) / c1) + reg / c2)        # :34-38
    if niter <= 1:
        return [w0][:max(niter, 0)]
    hist = np.empty((w0.size, niter - 1), order="F")
    w = w0.copy()
    _L.check(_L.load().gync_ebmodel_mple(m.handle(), _L.dptr(w), float(reg), float(h), niter - 1, 1 if cont else 0,
                                         _L.dptr(hist)))
    return [w0] + [hist[:, i].copy() for i in range(niter - 1)]


# ------------------------------------------------------------------------------ gync glue (eb/gync.jl)
def phi(x, rtol=None, atol=None):
    """phi(x) = forwardsol(x)[:, measuredinds] on t = 0:30 (eb/gync.jl:11): (116,) -> (31,4); (N,116) -> list of (31,4)."""
    lib = _L.load()
    x = np.asarray(x, dtype=np.float64)
    X = _L.fcol(np.atleast_2d(x))
    N = X.shape[0]
    Y = np.empty((124, N), order="F")
    st = np.zeros(N, dtype=np.int32)
    o = _L.opts(rtol or _api.DEFAULT_RTOL, atol or _api.DEFAULT_ATOL)
    _L.check(lib.gync_phi_batch(_L.dptr(X), N, C.byref(o), _L.dptr(Y), st.ctypes.data_as(_L.c_int32_p)))
    ys = [Y[:, n].reshape((31, 4), order="F").copy() for n in range(N)]
    return ys[0] if x.ndim == 1 else ys


def gyncmodel(xs, datas, zmult=0, sigma=0.1, seed=0, rtol=None, atol=None):
    """gyncmodel(xs, datas; zmult, sigma) (eb/gync.jl:10-25): ys = phi.(xs), samples with NaN results removed,
    err = MatrixNormalCentered(repmat(sigma * model_measerrors' * 10, 31)), zs = repmat(ys, zmult) + rand(err)."""
    xs = [np.asarray(x, dtype=np.float64) for x in xs]
    ys = phi(np.stack(xs), rtol, atol)
    keep = [i for i, y in enumerate(ys) if not np.any(np.isnan(y))]                   # :14-18
    xs, ys = [xs[i] for i in keep], [ys[i] for i in keep]
    err = MatrixNormalCentered(np.tile(sigma * model_measerrors[None, :] * 10, (31, 1)))   # :20
    rng = np.random.default_rng(seed)
    zs = [y + err.rand(rng) for _ in range(zmult) for y in ys]                         # :22  (repmat order)
    return LikelihoodModel(xs, ys, zs, datas, err)

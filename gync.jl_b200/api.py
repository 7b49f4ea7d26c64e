"""Host-side mirror of the reference's Julia API for the hot path, over the C-ABI library.

Names, argument meaning and error behaviour follow the reference (paths relative to its repo):
  Patient / Config            src/gync/model.jl:27-53          Lausanne / Pfizer   src/gync/data.jl:8,38
  gync / forwardsol           src/gync/gyncycle.jl:8-28, model.jl:90
  llh / logprior              src/gync/model.jl:111-163, sampling.jl:32
  Sampling / sample / sample! src/gync/sampling.jl:1-79 (sample! is `sample_bang`)
  WeightedChain               src/gync/weightedchain.jl:4-71
  emiteration(!) / emiterations   src/eb/em.jl:5-41
Julia is not in this image, so this Python layer stands where src/gync + src/eb would `ccall`
(julia/GynCB200.jl holds those stubs).  All numerics run in the CUDA library; nothing here
falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
import json
import os
from dataclasses import dataclass, field

import numpy as np

from . import lib as _L

_HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(_HERE, "data", "reference_constants.json")) as _f:
    _C = json.load(_f)

refy0 = np.array(_C["refy0"])                       # gyncycle.jl:1
refallparms = np.array(_C["refallparms"])           # gyncycle.jl:2
measuredinds = np.array([2, 7, 24, 25])             # model.jl:1 (1-based, as in the reference)
hillinds = np.array([4, 6, 10, 18, 20, 22, 26, 33, 36, 39, 43, 47, 49, 52, 55, 59, 65, 95, 98, 101, 103])
sampledinds = np.array([i for i in range(1, 104) if i not in set(hillinds.tolist())])    # model.jl:4
refparms = refallparms[sampledinds - 1]             # model.jl:6
model_measmagnitude = np.array([120.0, 10.0, 400.0, 15.0])   # model.jl:10
speciesnames = _C["speciesnames"]
parameternames = _C["parameternames"]

#: Solver tolerances used when the caller gives none.  The reference's gync() defaults to
#: reltol=abstol=1e-4 (gyncycle.jl:8), at which CVODE is ~3e-2 off the converged trajectory;
#: parity is stated against the converged solution, so the default here is the
#: tolerance at which log-likelihoods are within 1e-8 relative of it for every case of the
#: validation sets (18 432 near-set cases, profiles/r2_validation_cpu.json; Rodas5P).
DEFAULT_RTOL = 7e-10
DEFAULT_ATOL = 1e-11


def uniformpropvar(relprop):
    """model.jl:14 — eye(116) * log(1 + relprop^2)."""
    return np.eye(116) * np.log(1.0 + relprop ** 2)


defaultpropvar = uniformpropvar(0.1)                # model.jl:16


def allparms(parms):
    """model.jl:20-24."""
    p = refallparms.copy()
    p[sampledinds - 1] = parms
    return p


# ------------------------------------------------------------------------------ data
@dataclass
class Patient:
    """model.jl:27-30."""
    data: np.ndarray
    id: object = None


_PAT = None


def _patients():
    global _PAT
    if _PAT is None:
        z = np.load(os.path.join(_HERE, "data", "patients.npz"))
        _PAT = (z["lausanne"], z["pfizer"])
    return _PAT


def Lausanne(i: int) -> Patient:
    """data.jl:8 (1-based id)."""
    return Patient(_patients()[0][i - 1].copy(), f"l{i}")


def Pfizer(i: int) -> Patient:
    """data.jl:38."""
    return Patient(_patients()[1][i - 1].copy(), f"p{i}")


def alldatas(minmeas=20):
    """data.jl:3-6."""
    la, pf = _patients()
    return [d for d in list(la) + list(pf) if d.size - np.isnan(d).sum() >= minmeas]


# ------------------------------------------------------------------------------ forward solve
def gync(y0=None, p=None, t=None, reltol=None, abstol=None):
    """gync(y0,p,t;reltol,abstol) (gyncycle.jl:8-28): nT x 33 trajectory, t[0] = initial time;
    all-NaN on solver failure."""
    y0 = refy0 if y0 is None else y0
    p = refallparms if p is None else p
    t = np.arange(0.0, 31.0) if t is None else np.asarray(t, dtype=np.float64)
    Y, _ = gync_batch(np.asarray(y0, float)[None, :], np.asarray(p, float)[None, :], t, reltol, abstol)
    return Y[0]


def gync_batch(Y0, P, t, reltol=None, abstol=None):
    """Batched gync: Y0 (N,33), P (N,114) -> (N,nT,33), status (N,)."""
    lib = _L.load()
    Y0 = _L.fcol(Y0)
    P = _L.fcol(P)
    t = np.ascontiguousarray(t, dtype=np.float64)
    N = Y0.shape[0]
    Y = np.empty((N, t.size, 33), order="F")
    st = np.zeros(N, dtype=np.int32)
    o = _L.opts(reltol or DEFAULT_RTOL, abstol or DEFAULT_ATOL)
    _L.check(lib.gync_solve(_L.dptr(Y0), _L.dptr(P), N, _L.dptr(t), t.size, C.byref(o), _L.dptr(Y),
                            st.ctypes.data_as(_L.c_int32_p)))
    return Y, st


def forwardsol(x, tspan=None, reltol=None, abstol=None):
    """forwardsol(x,tspan=0:30) = gync(y0(x), allparms(parms(x)), tspan) (model.jl:90).
    x: (116,) -> (nT,33);  x: (N,116) -> (N,nT,33)."""
    lib = _L.load()
    x = np.asarray(x, dtype=np.float64)
    single = x.ndim == 1
    X = _L.fcol(x[None, :] if single else x)
    t = np.arange(0.0, 31.0) if tspan is None else np.ascontiguousarray(tspan, dtype=np.float64)
    N = X.shape[0]
    Y = np.empty((N, t.size, 33), order="F")
    st = np.zeros(N, dtype=np.int32)
    o = _L.opts(reltol or DEFAULT_RTOL, abstol or DEFAULT_ATOL)
    _L.check(lib.gync_forwardsol(_L.dptr(X), N, _L.dptr(t), t.size, C.byref(o), _L.dptr(Y),
                                 st.ctypes.data_as(_L.c_int32_p)))
    return Y[0] if single else Y


_REFSOL = None


def referencesolution(resolution=1):
    """gyncycle.jl:30-38 — trajectory of the reference point on 0:resolution:30 with non-positive
    entries replaced by the column's smallest positive one.  (The reference integrates this at
    reltol=abstol=1e-4 with CVODE; here it is the converged trajectory — DESIGN.md §3.)"""
    global _REFSOL
    if resolution == 1 and _REFSOL is not None:
        return _REFSOL.copy()
    sol = gync(refy0, allparms(refparms), np.arange(0.0, 30.0 + 1e-9, resolution))
    for j in range(sol.shape[1]):
        small = sol[:, j] <= 0
        if small.any():
            sol[small, j] = sol[~small, j].min()
    if resolution == 1:
        _REFSOL = sol.copy()
    return sol


# ------------------------------------------------------------------------------ priors / config
@dataclass
class GaussianMixture:
    """gaussianmixture(y, stdfactor) (model.jl:73-77): uniform mixture of diagonal normals centred
    on the rows of y with the column standard deviations."""
    means: np.ndarray
    stds: np.ndarray


def priory0(sigma=1.0, refsol=None):
    """model.jl:69."""
    y = referencesolution() if refsol is None else np.asarray(refsol, dtype=np.float64)
    return GaussianMixture(y.copy(), y.std(axis=0, ddof=1) * sigma)


def priorparms(alphas):
    """model.jl:70-71 — Truncated(Flat(),0,alpha_i): represented by the upper bounds."""
    return np.asarray(alphas, dtype=np.float64).copy()


@dataclass
class Config:
    """model.jl:35-53 (keyword constructor defaults of :51)."""
    patient: Patient = None
    sigma_rho: float = 0.1
    propvar: np.ndarray = None
    adapt: bool = False
    thin: int = 1
    initparms: np.ndarray = None
    inity0: np.ndarray = None
    priorparms: np.ndarray = None
    priory0: GaussianMixture = None
    rtol: float = None          # solver tolerances of this port (None -> DEFAULT_*)
    atol: float = None

    def __post_init__(self):
        if self.patient is None:
            self.patient = Lausanne(1)
        if self.propvar is None:
            self.propvar = defaultpropvar
        if self.initparms is None:
            self.initparms = refparms.copy()
        if self.inity0 is None:
            self.inity0 = refy0.copy()
        if self.priorparms is None:
            self.priorparms = priorparms(5 * self.initparms)

    @property
    def data(self):
        return self.patient.data

    def prior(self):
        if self.priory0 is None:
            self.priory0 = priory0(1.0)
        pr = _L.Prior()
        pr.gmm_means[:] = np.ascontiguousarray(self.priory0.means, dtype=np.float64).ravel().tolist()
        pr.gmm_stds[:] = self.priory0.stds.tolist()
        pr.upper[:] = self.priorparms.tolist()
        pr.mu_T, pr.sd_T = 28.9, 3.4          # model.jl:116
        return pr

    def filename(self):
        """model.jl:48."""
        return f"p{self.patient.id}s{self.sigma_rho}r{np.trace(self.propvar)}t{self.thin}a{str(self.adapt).lower()}.jld"


def init(c: Config):
    """model.jl:94."""
    return np.concatenate([c.initparms, c.inity0, [28.0]])


# ------------------------------------------------------------------------------ densities
def loglik_matrix(X, datas, sigmas, rtol=None, atol=None, want_traj=False, want_steps=False):
    """K x M log-likelihoods of samples X (K,116) against M patient tables — the llh block of
    WeightedChain(...) (weightedchain.jl:53-56); one forward solve per sample serves all patients."""
    lib = _L.load()
    X = _L.fcol(np.atleast_2d(X))
    K = X.shape[0]
    datas = [np.asarray(d, dtype=np.float64) for d in datas]
    M = len(datas)
    D = np.empty((31, 4, M), order="F")
    for m, d in enumerate(datas):
        if d.shape != (31, 4):
            raise ValueError("patient data must be 31x4 (data.jl:15)")
        D[:, :, m] = d
    sig = np.ascontiguousarray(np.broadcast_to(np.asarray(sigmas, dtype=np.float64), (M,)))
    LL = np.empty((K, M), order="F")
    st = np.zeros(K, dtype=np.int32)
    Ym = np.empty((K, 62, 4), order="F") if want_traj else None
    ns = np.zeros((K, 2), dtype=np.int32, order="F") if want_steps else None
    o = _L.opts(rtol or DEFAULT_RTOL, atol or DEFAULT_ATOL)
    _L.check(lib.gync_loglik_batch(_L.dptr(X), K, _L.dptr(D), M, _L.dptr(sig), C.byref(o), _L.dptr(LL),
                                   _L.dptr(Ym), st.ctypes.data_as(_L.c_int32_p),
                                   None if ns is None else ns.ctypes.data_as(_L.c_int32_p)))
    out = [LL, st]
    if want_traj:
        out.append(Ym)
    if want_steps:
        out.append(ns)
    return tuple(out)


def llh(c, x=None):
    """llh(c::Config, x::Vector) (model.jl:128) -> float;  llh(c, X (N,116)) -> (N,);
    llh(s::Sampling) (sampling.jl:32) -> (N,)."""
    if isinstance(c, Sampling):
        c, x = c.config, c.samples
    x = np.asarray(x, dtype=np.float64)
    LL, _ = loglik_matrix(np.atleast_2d(x), [c.data], [c.sigma_rho], c.rtol, c.atol)
    return float(LL[0, 0]) if x.ndim == 1 else LL[:, 0].copy()


def logprior(c: Config, x):
    """logprior(c,x) (model.jl:111-117) up to the additive constant of Truncated(Flat()) which
    cancels in every Metropolis ratio and in weightedchain.jl:59."""
    lib = _L.load()
    x = np.asarray(x, dtype=np.float64)
    X = _L.fcol(np.atleast_2d(x))
    LP = np.empty(X.shape[0])
    pr = c.prior()
    _L.check(lib.gync_logprior_batch(_L.dptr(X), X.shape[0], C.byref(pr), _L.dptr(LP)))
    return float(LP[0]) if x.ndim == 1 else LP


def logpost(c: Config, x):
    """model.jl:119-124."""
    l = logprior(c, x)
    return l if l == -np.inf else l + llh(c, x)


# ------------------------------------------------------------------------------ sampling
@dataclass
class Variate:
    """Stands for Mamba's AMMVariate held by Sampling.variate (sampling.jl:4): log-space state,
    current log-target, proposal factor, and the counter-based RNG position."""
    state: np.ndarray
    logf: float
    SigmaL: np.ndarray
    seed: int
    iteration: int = 0
    naccept: int = 0
    # Philox stream of THIS chain (the draws of a chain are addressed by (seed, chain_id, iteration)): it stays with the
    # variate through sample_chains groups of any composition, save() / load() and batch(), so a chain resumed alone
    # continues exactly the stream it would have drawn from inside its group.
    chain_id: int = None
    # AMMTune adaptation state (Mamba): counter, running mean / second moment of the log-space
    # states, factor of the adapted proposal.  m < 0: adaptation has not started.
    m: int = -1
    Mv: np.ndarray = None
    Mvv: np.ndarray = None
    SigmaLm: np.ndarray = None
    beta: float = 0.05          # model.jl:105
    scale: float = 2.38         # model.jl:106


@dataclass
class Sampling:
    """sampling.jl:1-5."""
    samples: np.ndarray
    config: Config
    variate: Variate

    def __len__(self):
        return self.samples.shape[0]


_process_seed = []
_next_chain_id = [0]


def process_seed() -> int:
    """The seed of chains created without one: drawn once per process from the OS (the reference's chains use Julia's
    global RNG, seeded at start-up).  Such chains are told apart by a process-wide chain counter, so chains launched
    separately — the per-patient workflow of batch.jl — never share their proposal or accept noise."""
    if not _process_seed:
        _process_seed.append(int(np.random.SeedSequence().generate_state(1, np.uint64)[0]))
    return _process_seed[0]


def SamplerVariate(c: Config, seed=None, chain_id=None) -> Variate:
    """model.jl:97-107 — state = log(init(c)); SigmaL = chol(propvar).L.
    The draws of a chain are addressed by (seed, chain_id, iteration).  seed=None: the process seed and the next value of
    the process-wide chain counter.  Explicit seed: reproducible; chain_id defaults to the chain's position in the first
    sample_chains call that advances it (0 for a single chain) and stays with the variate from then on."""
    if seed is None:
        seed = process_seed()
        if chain_id is None:
            chain_id = _next_chain_id[0]
            _next_chain_id[0] += 1
    return Variate(np.log(init(c)), np.nan, np.linalg.cholesky(c.propvar), int(seed),
                   chain_id=None if chain_id is None else int(chain_id))


def new_sampling(c: Config, seed=None, chain_id=None) -> Sampling:
    """Sampling(c) (sampling.jl:53-58)."""
    return Sampling(np.empty((0, 116)), c, SamplerVariate(c, seed, chain_id))


def propinit(s: "Sampling"):
    """sampling.jl:17-20 — covariance of the initial proposal."""
    return s.variate.SigmaL @ s.variate.SigmaL.T


def propadapt(s: "Sampling"):
    """sampling.jl:26-29 — covariance of the adaptive proposal (the initial one before the first full-rank estimate)."""
    l = s.variate.SigmaL if s.variate.SigmaLm is None else s.variate.SigmaLm
    return l @ l.T


def sample_chains(samplings, iters: int):
    """Advance many chains by `iters` Metropolis steps in ONE launch (the batched form of
    sample!, sampling.jl:62-79; replaces the Slurm fan-out of batch.jl:3-22).  Chains may have
    different patients / sigma_rho; thin, propvar and tolerances are taken per group."""
    lib = _L.load()
    groups = {}
    for i, s in enumerate(samplings):
        if s.variate.chain_id is None:
            s.variate.chain_id = i
        key = (s.config.thin, s.config.adapt, s.variate.SigmaL.tobytes(), s.config.rtol, s.config.atol, s.variate.seed,
               s.variate.iteration, id(s.config.priory0) if s.config.priory0 is not None else 0,
               s.config.priorparms.tobytes())
        groups.setdefault(key, []).append(i)
    for idx in groups.values():
        ss = [samplings[i] for i in idx]
        c0, v0 = ss[0].config, ss[0].variate
        Cn = len(ss)
        thin = c0.thin
        n_out = iters // thin
        state = _L.fcol(np.stack([s.variate.state for s in ss]))
        logf = np.array([s.variate.logf for s in ss], dtype=np.float64)
        D = np.empty((31, 4, Cn), order="F")
        for m, s in enumerate(ss):
            D[:, :, m] = s.config.data
        sig = np.array([s.config.sigma_rho for s in ss], dtype=np.float64)
        pat = np.arange(Cn, dtype=np.int32)
        SL = v0.SigmaL
        diag = np.allclose(SL, np.diag(np.diag(SL))) and np.allclose(np.diag(SL), SL[0, 0])
        chol = None if diag else _L.fcol(SL)
        out = np.empty((n_out, 116, Cn), order="F")
        nacc = np.zeros(Cn, dtype=np.int64)
        pr = c0.prior()
        o = _L.opts(c0.rtol or DEFAULT_RTOL, c0.atol or DEFAULT_ATOL)
        ids = np.array([s.variate.chain_id for s in ss], dtype=np.uint64)      # each chain's own Philox stream
        idp = ids.ctypes.data_as(C.POINTER(C.c_uint64))
        if c0.adapt:                                        # Mamba.sample!(variate, adapt=true) (sampling.jl:72)
            am = np.array([s.variate.m for s in ss], dtype=np.int64)
            aMv = np.zeros((Cn, 116))
            aMvv = np.zeros((Cn, 116, 116))
            aLm = np.zeros((Cn, 116, 116))
            for m, s in enumerate(ss):
                if s.variate.m >= 0:
                    aMv[m], aMvv[m], aLm[m] = s.variate.Mv, s.variate.Mvv, s.variate.SigmaLm
            tune = _L.AmmTune(v0.beta, v0.scale, am.ctypes.data_as(_L.c_int64_p), _L.dptr(aMv), _L.dptr(aMvv), _L.dptr(aLm))
            _L.check(lib.gync_mcmc_run_amm_ids(_L.dptr(state), _L.dptr(logf), Cn, _L.dptr(chol), float(SL[0, 0]) if diag else 0.0,
                                               _L.dptr(D), Cn, _L.dptr(sig), pat.ctypes.data_as(_L.c_int32_p), C.byref(pr),
                                               C.byref(o), iters, thin, v0.seed, v0.iteration, idp, None, None, None,
                                               _L.dptr(out), nacc.ctypes.data_as(_L.c_int64_p), C.byref(tune)))
            for m, s in enumerate(ss):
                s.variate.m, s.variate.Mv, s.variate.Mvv, s.variate.SigmaLm = int(am[m]), aMv[m].copy(), aMvv[m].copy(), aLm[m].copy()
        else:
            _L.check(lib.gync_mcmc_run_ids(_L.dptr(state), _L.dptr(logf), Cn, _L.dptr(chol), float(SL[0, 0]) if diag else 0.0,
                                           _L.dptr(D), Cn, _L.dptr(sig), pat.ctypes.data_as(_L.c_int32_p), C.byref(pr),
                                           C.byref(o), iters, thin, v0.seed, v0.iteration, idp, None, None, _L.dptr(out),
                                           nacc.ctypes.data_as(_L.c_int64_p)))
        for m, s in enumerate(ss):
            s.samples = np.vstack([s.samples, out[:, :, m]])            # sampling.jl:66-68,74
            s.variate.state = state[m].copy()
            s.variate.logf = float(logf[m])
            s.variate.iteration += iters
            s.variate.naccept += int(nacc[m])
    return samplings


def sample_bang(s: Sampling, iters: int) -> Sampling:
    """sample!(s::Sampling, iters) (sampling.jl:62-79)."""
    return sample_chains([s], iters)[0]


def sample(obj, n, seed=None):
    """sample(c::Config, iters) (sampling.jl:60)  |  sample(w::WeightedChain, n) (weightedchain.jl:18-26)."""
    if isinstance(obj, Config):
        return sample_bang(new_sampling(obj, seed), n)
    if isinstance(obj, WeightedChain):
        return obj.sample(n, seed)
    raise TypeError(type(obj))


# ------------------------------------------------------------------------------ checkpoint / resume
def save(path, s: Sampling):
    """save(path, s::Sampling) (utils.jl:49).  The reference writes JLD (HDF5) with the Mamba variate as a Julia
    `Base.serialize` blob (utils.jl:31-44) — neither is readable outside Julia, so the container here is .npz with the
    same content: samples, the Config fields, and the variate (state, current log-target, proposal factor, RNG position
    and the AMM tuning state), enough for sample!(load(path), n) to continue the chain bit-identically."""
    c, v = s.config, s.variate
    if c.priory0 is None:
        c.prior()
    z = {"samples": s.samples, "patient_data": c.patient.data, "patient_id": np.array(str(c.patient.id)),
         "sigma_rho": c.sigma_rho, "propvar": c.propvar, "adapt": c.adapt, "thin": c.thin, "initparms": c.initparms,
         "inity0": c.inity0, "priorparms": c.priorparms, "priory0_means": c.priory0.means, "priory0_stds": c.priory0.stds,
         "rtol": np.nan if c.rtol is None else c.rtol, "atol": np.nan if c.atol is None else c.atol,
         "v_state": v.state, "v_logf": v.logf, "v_SigmaL": v.SigmaL, "v_seed": v.seed, "v_iteration": v.iteration,
         "v_naccept": v.naccept, "v_m": v.m, "v_beta": v.beta, "v_scale": v.scale, "v_chain_id": -1 if v.chain_id is None else v.chain_id}
    if v.m >= 0:
        z.update(v_Mv=v.Mv, v_Mvv=v.Mvv, v_SigmaLm=v.SigmaLm)
    with open(path, "wb") as f:                      # np.savez would append ".npz" to the reference's ".jld" file names
        np.savez(f, **z)


def load(path) -> Sampling:
    """load(path) (utils.jl:48)."""
    z = np.load(path, allow_pickle=False)
    nz = lambda k: None if np.isnan(z[k]) else float(z[k])
    c = Config(patient=Patient(z["patient_data"], str(z["patient_id"])), sigma_rho=float(z["sigma_rho"]), propvar=z["propvar"],
               adapt=bool(z["adapt"]), thin=int(z["thin"]), initparms=z["initparms"], inity0=z["inity0"],
               priorparms=z["priorparms"], priory0=GaussianMixture(z["priory0_means"], z["priory0_stds"]),
               rtol=nz("rtol"), atol=nz("atol"))
    v = Variate(z["v_state"], float(z["v_logf"]), z["v_SigmaL"], int(z["v_seed"]), int(z["v_iteration"]), int(z["v_naccept"]),
                (int(z["v_chain_id"]) if int(z["v_chain_id"]) >= 0 else None) if "v_chain_id" in z.files else 0, int(z["v_m"]), beta=float(z["v_beta"]),
                scale=float(z["v_scale"]))
    if v.m >= 0:
        v.Mv, v.Mvv, v.SigmaLm = z["v_Mv"], z["v_Mvv"], z["v_SigmaLm"]
    return Sampling(z["samples"], c, v)


def batch(c: Config, iters, path, overwrite=False, seed=None):
    """batch(c, iters::Vector{Int}, path) (batch.jl:26-46): sample up to each iteration count in `iters`, saving after
    each; an existing file is continued unless `overwrite`.  (The Slurm fan-out of batch.jl:3-22 is replaced by
    sample_chains: many chains in one launch.)"""
    if not overwrite and os.path.isfile(path):
        s = load(path)
        c = s.config
    else:
        s = new_sampling(c, seed)
    thin = c.thin
    for i in iters:
        n = s.samples.shape[0] * thin
        i = i - n
        if i < thin:
            continue
        s = sample_bang(s, i)
        save(path, s)
    return s


# ------------------------------------------------------------------------------ WeightedChain / EM
class WeightedChain:
    """weightedchain.jl:4-16: samples (K x 116), likelihoods (K x M), weights (K), upd = density ./ weights."""

    def __init__(self, samples, likelihoods=None, weights=None, density=None, burnin=0):
        if isinstance(samples, (list, tuple)) and samples and isinstance(samples[0], Sampling):
            samples, likelihoods, weights, density = _weightedchain_from_samplings(samples, burnin)
        K = np.asarray(samples).shape[0]
        self.samples = np.asarray(samples, dtype=np.float64)
        self.likelihoods = _L.fcol(likelihoods)
        self.weights = np.ones(K) if weights is None else np.array(weights, dtype=np.float64)
        d = np.ones(K) if density is None else np.asarray(density, dtype=np.float64)
        with np.errstate(all="ignore"):
            self.upd = d / self.weights                      # weightedchain.jl:13

    def density(self):
        """weightedchain.jl:74."""
        return self.upd * self.weights

    # The likelihood matrix stays in HBM between emiteration!(w) calls (a host loop that calls it once per iteration would
    # otherwise upload K x M doubles every time): one library handle per WeightedChain, re-created when `likelihoods` is
    # rebound to another array.  In-place edits of the matrix are not noticed: call forget_device_matrix() after one.
    _em = None
    _em_key = None

    def _em_handle(self):
        L = self.likelihoods
        key = (L.ctypes.data, L.shape)
        if self._em is None or self._em_key != key:
            self.forget_device_matrix()
            h = C.c_void_p()
            _L.check(_L.load().gync_em_create(_L.dptr(L), L.shape[0], L.shape[1], C.byref(h)))
            self._em, self._em_key = h, key
        return self._em

    def forget_device_matrix(self):
        if self._em is not None:
            try:
                _L.load().gync_em_destroy(self._em)
            except Exception:
                pass
        self._em, self._em_key = None, None

    def __del__(self):
        self.forget_device_matrix()

    def sample(self, n=1, seed=0):
        """weightedchain.jl:18-26 — weighted resampling by inverse CDF."""
        cum = np.cumsum(self.weights)
        tgt = np.random.default_rng(seed).random(n) * cum[-1]
        return self.samples[np.searchsorted(cum, tgt, side="left")]


def collapse_duplicates(samplings, burnin=0):
    """weightedchain.jl:31-46 — collapse consecutive duplicate rows over the concatenated chains."""
    S = np.vstack([s.samples[burnin:] for s in samplings])
    if S.shape[0] == 0:
        return S, np.zeros(0, dtype=np.int64)
    new = np.ones(S.shape[0], dtype=bool)
    new[1:] = np.any(S[1:] != S[:-1], axis=1)
    starts = np.flatnonzero(new)
    counts = np.diff(np.append(starts, S.shape[0])).astype(np.int64)
    return S[starts], counts


def collapse_duplicates_device(samplings, burnin=0):
    """weightedchain.jl:31-46 on the device (gync_collapse_duplicates): same result as collapse_duplicates."""
    lib = _L.load()
    S = _L.fcol(np.vstack([s.samples[burnin:] for s in samplings]))
    K = S.shape[0]
    if K == 0:
        return S, np.zeros(0, dtype=np.int64)
    rows = np.empty((K, 116), order="F")
    counts = np.zeros(K, dtype=np.int64)
    n = C.c_int64()
    _L.check(lib.gync_collapse_duplicates(_L.dptr(S), K, _L.dptr(rows), counts.ctypes.data_as(_L.c_int64_p), C.byref(n)))
    return np.ascontiguousarray(rows[:n.value]), counts[:n.value].copy()


def _weightedchain_from_samplings(samplings, burnin=0):
    """WeightedChain(samplings::Vector{Sampling}, burnin) (weightedchain.jl:29-71)."""
    lib = _L.load()
    samplevec, counts = collapse_duplicates_device(samplings, burnin)
    c0 = samplings[0].config
    prior = logprior(c0, samplevec)                                             # :49-51
    LL, _ = loglik_matrix(samplevec, [s.config.data for s in samplings],
                          [s.config.sigma_rho for s in samplings], c0.rtol, c0.atol)   # :53-56
    K, M = LL.shape
    Lm = np.empty((K, M), order="F")
    w = np.empty(K)
    d = np.empty(K)
    _L.check(lib.gync_wc_normalize(_L.dptr(LL), _L.dptr(np.ascontiguousarray(prior)),
                                   counts.ctypes.data_as(_L.c_int64_p), K, M, _L.dptr(Lm), _L.dptr(w), _L.dptr(d)))   # :59-67
    return samplevec, Lm, w, d


def emiteration_bang(w, L=None, niter=1):
    """emiteration!(w,L) (em.jl:27-41) / emiteration!(c::WeightedChain) (em.jl:7): in place."""
    lib = _L.load()
    if isinstance(w, WeightedChain):
        if not (w.weights.dtype == np.float64 and w.weights.flags.c_contiguous):
            w.weights = np.ascontiguousarray(w.weights, dtype=np.float64)
        _L.check(lib.gync_em_iterate_handle(w._em_handle(), _L.dptr(w.weights), niter, None))
        return w
    if not (isinstance(w, np.ndarray) and w.dtype == np.float64 and w.flags.c_contiguous):
        raise TypeError("w must be a contiguous float64 vector (it is updated in place)")
    L = _L.fcol(L)
    K, M = L.shape
    if w.shape != (K,):
        raise ValueError("size mismatch between w and L")
    _L.check(lib.gync_em_iterate(_L.dptr(w), _L.dptr(L), K, M, niter, None))
    return w


def emiteration(w, L):
    """em.jl:23."""
    return emiteration_bang(np.array(w, dtype=np.float64), L)


def emiterations(w0, L, niter):
    """em.jl:5 — [w0, em(w0), ..., em^(niter-1)(w0)]  (w0 first; niter entries)."""
    lib = _L.load()
    w0 = np.array(w0, dtype=np.float64)
    if niter <= 1:
        return [w0][:max(niter, 0)]
    L = _L.fcol(L)
    K, M = L.shape
    hist = np.empty((K, niter - 1), order="F")
    w = w0.copy()
    _L.check(lib.gync_em_iterate(_L.dptr(w), _L.dptr(L), K, M, niter - 1, _L.dptr(hist)))
    return [w0] + [hist[:, i].copy() for i in range(niter - 1)]

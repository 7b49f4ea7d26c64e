# GynCB200.jl — the `ccall` stubs a GynC.jl maintainer adds so that the hot path runs in
# libgync_b200.so (C ABI: include/gync_b200.h).  Signatures and return conventions of the
# overridden methods are the reference's own; only the bodies change.
#
# NOTE: Julia is not present in the build image, so this file has been reviewed but never
# executed; the Python host layer (gync.jl_b200/api.py + lib.py) exercises the identical ABI
# (column-major arrays, Int64/Cint widths, in-place mutation) in the test-suite.
#
# Usage (inside module GynC, after the includes of src/GynC.jl):   include("GynCB200.jl")

const LIBGYNC = get(ENV, "GYNC_B200_LIB", "libgync_b200.so")

# Julia is one process: drive all GPUs of the node from it (samples / EM rows are sharded inside the library).
function b200init(ndev::Int=1)
  rc = ndev > 1 ? ccall((:gync_init_multi, LIBGYNC), Cint, (Cint, Ptr{Cint}), ndev, C_NULL) :
                  ccall((:gync_init, LIBGYNC), Cint, (Cint,), 0)
  rc == 0 || error("gync_b200: ", bytestring(ccall((:gync_last_error, LIBGYNC), Ptr{UInt8}, ())))
end

immutable SolverOpts          # gync_solver_opts
  rtol::Cdouble
  atol::Cdouble
  h0::Cdouble
  hmax::Cdouble
  max_steps::Cint
  reserved::Cint
end
SolverOpts(rtol, atol) = SolverOpts(rtol, atol, 0.0, 0.0, 0, 0)

# gync_prior: 31x33 row-major means | 33 stds | 82 upper bounds | mu_T | sd_T  (1140 doubles)
function priorvector(c::Config)
  gmm    = c.priory0                                   # MixtureModel of MvNormal (model.jl:73-77)
  means  = hcat([mean(comp) for comp in gmm.components]...)'      # 31 x 33
  stds   = sqrt(diag(cov(gmm.components[1])))
  upper  = [p.upper for p in c.priorparms]             # Truncated(Flat(),0,alpha) (model.jl:70-71)
  vcat(vec(means'), stds, upper, 28.9, 3.4)            # model.jl:116
end

function check(rc::Cint)
  rc == 0 || error("gync_b200: ", bytestring(ccall((:gync_last_error, LIBGYNC), Ptr{UInt8}, ())))
end

# --- gync(y0,p,t;reltol,abstol)  (src/gync/gyncycle.jl:8-26) -------------------------------
function gync(y0::Vector, p::Vector, t::Vector; reltol=7e-10, abstol=1e-11)
  Y  = Array(Float64, length(t), length(y0))           # N = 1: exactly the reference's nT x 33 layout
  st = Cint[0]
  o  = Ref(SolverOpts(reltol, abstol))
  check(ccall((:gync_solve, LIBGYNC), Cint,
              (Ptr{Cdouble}, Ptr{Cdouble}, Int64, Ptr{Cdouble}, Cint, Ptr{SolverOpts}, Ptr{Cdouble}, Ptr{Cint}),
              y0, p, 1, convert(Vector{Float64}, t), length(t), o, Y, st))
  Y                                                    # all-NaN when st[1] != 0 (gyncycle.jl:18-19)
end

# --- llh(c,x) (model.jl:128-155), llh(s::Sampling) (sampling.jl:32) --------------------------
function loglikmatrix(X::Matrix{Float64}, datas::Vector{Matrix{Float64}}, sigmas::Vector{Float64}; reltol=7e-10, abstol=1e-11)
  N, M = size(X, 1), length(datas)
  D  = cat(3, datas...)                                # 31 x 4 x M, NaN = missing
  LL = Array(Float64, N, M)
  st = Array(Cint, N)
  o  = Ref(SolverOpts(reltol, abstol))
  check(ccall((:gync_loglik_batch, LIBGYNC), Cint,
              (Ptr{Cdouble}, Int64, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Ptr{SolverOpts}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cint}, Ptr{Cint}),
              X, N, D, M, sigmas, o, LL, C_NULL, st, C_NULL))
  LL
end
llh(c::Config, x::Vector, periods::Int=2) = loglikmatrix(reshape(x, 1, 116), [data(c)], [c.sigma_rho])[1]
llh(s::Sampling) = vec(loglikmatrix(s.samples, [data(s)], [s.config.sigma_rho]))

# --- logprior(c,x) (model.jl:111-117) -------------------------------------------------------
function logprior(c::Config, x::Vector)
  lp = Cdouble[0]
  check(ccall((:gync_logprior_batch, LIBGYNC), Cint, (Ptr{Cdouble}, Int64, Ptr{Cdouble}, Ptr{Cdouble}),
              reshape(x, 1, 116), 1, priorvector(c), lp))
  lp[1]
end

# --- sample!(s::Sampling, iters) (sampling.jl:62-79); the sampler state replaces Mamba's variate ---
type B200Variate
  state::Vector{Float64}        # log-space (model.jl:98)
  logf::Float64
  SigmaL::Matrix{Float64}       # chol(propvar).L (model.jl:104)
  seed::UInt64
  iteration::UInt64
  chain_id::UInt64              # Philox stream of this chain: draws are addressed by (seed, chain_id, iteration), so a chain
                                # keeps its stream through batch()/save/load and whatever group it is advanced in
  # AMMTune adaptation state (used when c.adapt; m < 0: adaptation not started)
  m::Vector{Int64}              # length 1
  Mv::Vector{Float64}           # 116
  Mvv::Matrix{Float64}          # 116 x 116 (symmetric)
  SigmaLm::Matrix{Float64}      # 116 x 116, the library stores the lower factor ROW-major == its transpose column-major
end

# SamplerVariate(c) (model.jl:97-107): state = log(init(c)), SigmaL = chol(propvar)'.  Chains created without a seed share a
# per-process random seed and take the next value of a process-wide counter as their stream id, so separately launched
# chains never share their noise (the reference draws from Julia's global RNG).
const PROCESS_SEED  = UInt64[rand(RandomDevice(), UInt64)]
const NEXT_CHAIN_ID = UInt64[0]
function SamplerVariate(c::Config; seed::UInt64=PROCESS_SEED[1], chain_id::UInt64=(NEXT_CHAIN_ID[1] += 1) - 1)
  B200Variate(log(init(c)), NaN, full(chol(c.propvar)'), seed, UInt64(0), chain_id, Int64[-1], zeros(116), zeros(116, 116), zeros(116, 116))
end

immutable AmmTune               # gync_amm_tune
  beta::Cdouble
  scale::Cdouble
  m::Ptr{Int64}
  Mv::Ptr{Cdouble}
  Mvv::Ptr{Cdouble}
  SigmaLm::Ptr{Cdouble}
end

function sample!(s::Sampling, iters::Int)
  c, v  = s.config, s.variate::B200Variate
  thin  = c.thin
  n     = round(Int, iters / thin, RoundDown)
  out   = Array(Float64, n, 116)
  nacc  = Int64[0]
  state = reshape(copy(v.state), 1, 116)
  lf    = [v.logf]
  o     = Ref(SolverOpts(7e-10, 1e-11))
  if c.adapt                                            # Mamba.sample!(variate, adapt=true) (sampling.jl:72, model.jl:104-106)
    tune = Ref(AmmTune(0.05, 2.38, pointer(v.m), pointer(v.Mv), pointer(v.Mvv), pointer(v.SigmaLm)))
    check(ccall((:gync_mcmc_run_amm_ids, LIBGYNC), Cint,
                (Ptr{Cdouble}, Ptr{Cdouble}, Int64, Ptr{Cdouble}, Cdouble, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Ptr{Cint},
                 Ptr{Cdouble}, Ptr{SolverOpts}, Int64, Cint, UInt64, UInt64, Ptr{UInt64}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble},
                 Ptr{Cdouble}, Ptr{Int64}, Ptr{AmmTune}),
                state, lf, 1, v.SigmaL, 0.0, data(c), 1, [c.sigma_rho], Cint[0], priorvector(c), o,
                iters, thin, v.seed, v.iteration, UInt64[v.chain_id], C_NULL, C_NULL, C_NULL, out, nacc, tune))
  else
  check(ccall((:gync_mcmc_run_ids, LIBGYNC), Cint,           # one persistent launch; chains shard over gync_init_multi's devices
              (Ptr{Cdouble}, Ptr{Cdouble}, Int64, Ptr{Cdouble}, Cdouble, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Ptr{Cint},
               Ptr{Cdouble}, Ptr{SolverOpts}, Int64, Cint, UInt64, UInt64, Ptr{UInt64}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Int64}),
              state, lf, 1, v.SigmaL, 0.0, data(c), 1, [c.sigma_rho], Cint[0], priorvector(c), o,
              iters, thin, v.seed, v.iteration, UInt64[v.chain_id], C_NULL, C_NULL, out, nacc))
  end
  v.state, v.logf, v.iteration = vec(state), lf[1], v.iteration + iters
  s.samples = vcat(s.samples, out)                      # sampling.jl:66-68,74
  s
end

# --- emiteration!(w,L) (src/eb/em.jl:27-41) -------------------------------------------------
function emiteration!(w::DenseVector{Float64}, L::DenseMatrix{Float64})
  K, M = size(L)
  check(ccall((:gync_em_iterate, LIBGYNC), Cint, (Ptr{Cdouble}, Ptr{Cdouble}, Int64, Cint, Cint, Ptr{Cdouble}),
              w, L, K, M, 1, C_NULL))
  w
end

# emiteration!(c::WeightedChain) (em.jl:7), typically called once per iteration from a script loop: the likelihood matrix
# stays in HBM between the calls (one handle per WeightedChain, freed by its finalizer); only the K weights travel.
const EM_HANDLES = ObjectIdDict()
function emhandle(c::WeightedChain)
  get!(EM_HANDLES, c) do
    K, M = size(c.likelihoods)
    h = Ref{Ptr{Void}}(C_NULL)
    check(ccall((:gync_em_create, LIBGYNC), Cint, (Ptr{Cdouble}, Int64, Cint, Ptr{Ptr{Void}}), c.likelihoods, K, M, h))
    finalizer(c, x -> (ccall((:gync_em_destroy, LIBGYNC), Void, (Ptr{Void},), EM_HANDLES[x]); delete!(EM_HANDLES, x)))
    h[]
  end
end
function emiteration!(c::WeightedChain)
  check(ccall((:gync_em_iterate_handle, LIBGYNC), Cint, (Ptr{Void}, Ptr{Cdouble}, Cint, Ptr{Cdouble}), emhandle(c), c.weights, 1, C_NULL))
  c
end

# emiterations(w0,L,niter) (em.jl:5): all iterates in one launch
function emiterations(w0, L, niter)
  K, M = size(L)
  w, H = copy(w0), Array(Float64, K, max(niter - 1, 0))
  niter > 1 && check(ccall((:gync_em_iterate, LIBGYNC), Cint, (Ptr{Cdouble}, Ptr{Cdouble}, Int64, Cint, Cint, Ptr{Cdouble}),
                           w, L, K, M, niter - 1, H))
  vcat(Vector[w0], [H[:, i] for i in 1:size(H, 2)])
end

# --- WeightedChain(samplings, burnin) tail (weightedchain.jl:49-67) ---------------------------
function wcnormalize(LL::Matrix{Float64}, prior::Vector{Float64}, counts::Vector{Int})
  K, M = size(LL)
  L, w, d = similar(LL), Array(Float64, K), Array(Float64, K)
  check(ccall((:gync_wc_normalize, LIBGYNC), Cint,
              (Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Int64}, Int64, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
              LL, prior, counts, K, M, L, w, d))
  L, w, d
end

# propadapt(s) (sampling.jl:26-29): SigmaLm is held row-major by the library, i.e. transposed in Julia's view
propadapt(s::Sampling) = (l = s.variate.SigmaLm'; l * l')

# === src/eb: likelihood matrix, LikelihoodModel, regularisers, projected gradient ascent =================

cols(ms) = hcat([vec(m) for m in ms]...)               # likelihoodmat.jl:35: D x n, NaN = missing

# --- likelihoodmat_nanfast(xs, ys, d::MatrixNormalCentered; renormalize) (eb/likelihoodmat.jl:34-58) ---
function likelihoodmat_nanfast(xs, ys, d::MatrixNormalCentered; renormalize=true)
  A, B, s = cols(xs), cols(ys), vec(d.sigmas)
  L = Array(Float64, size(A, 2), size(B, 2))
  check(ccall((:gync_likelihoodmat, LIBGYNC), Cint,
              (Ptr{Cdouble}, Int64, Ptr{Cdouble}, Int64, Cint, Ptr{Cdouble}, Cint, Ptr{Cdouble}),
              A, size(A, 2), B, size(B, 2), length(s), s, renormalize ? 1 : 0, L))
  L
end

# --- phi(x) = forwardsol(x)[:, measuredinds] (eb/gync.jl:11): all samples in one launch ---
function phi(xs::Vector{Vector{Float64}})
  X  = vcat([x' for x in xs]...)                        # N x 116
  Y  = Array(Float64, 124, length(xs))
  st = Array(Cint, length(xs))
  o  = Ref(SolverOpts(7e-10, 1e-11))
  check(ccall((:gync_phi_batch, LIBGYNC), Cint, (Ptr{Cdouble}, Int64, Ptr{SolverOpts}, Ptr{Cdouble}, Ptr{Cint}),
              X, length(xs), o, Y, st))
  [reshape(Y[:, n], 31, 4) for n in 1:length(xs)]       # NaN matrix on solver failure (gyncycle.jl:18-19)
end

# --- LikelihoodModel (eb/likelihoodmodel.jl:3-14): one handle per model, both likelihood matrices stay in HBM.
#     Keep it in a field / WeakKeyDict next to the model and free it with a finalizer.
function b200model(m::LikelihoodModel)
  Y, D, Zs = cols(m.ys), cols(m.datas), cols(m.zs)
  h = Ref{Ptr{Void}}(C_NULL)
  check(ccall((:gync_ebmodel_create, LIBGYNC), Cint,
              (Ptr{Cdouble}, Int64, Ptr{Cdouble}, Int64, Ptr{Cdouble}, Int64, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Ptr{Void}}),
              Y, size(Y, 2), D, size(D, 2), Zs, size(Zs, 2), size(Y, 1), vec(m.measerr.sigmas), vec(m.zsampledistr.sigmas), h))
  h[]
end
b200free(h::Ptr{Void}) = ccall((:gync_ebmodel_destroy, LIBGYNC), Void, (Ptr{Void},), h)

function logl(h::Ptr{Void}, w::Vector{Float64})         # logl(m, w) (regularizers.jl:5, likelihoodmodel.jl:72)
  r = Cdouble[0]
  check(ccall((:gync_ebmodel_logl, LIBGYNC), Cint, (Ptr{Void}, Ptr{Cdouble}, Ptr{Cdouble}), h, w, r)); r[1]
end
function dlogl(h::Ptr{Void}, w::Vector{Float64})        # regularizers.jl:7-10
  d = similar(w)
  check(ccall((:gync_ebmodel_dlogl, LIBGYNC), Cint, (Ptr{Void}, Ptr{Cdouble}, Ptr{Cdouble}), h, w, d)); d
end
function hz(h::Ptr{Void}, w::Vector{Float64})           # hz(m, w) (regularizers.jl:24-45); wz = repeatweights(w, zs)
  r = Cdouble[0]
  check(ccall((:gync_ebmodel_hz, LIBGYNC), Cint, (Ptr{Void}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}), h, w, C_NULL, r)); r[1]
end
function dhz(h::Ptr{Void}, w::Vector{Float64})          # dhzdiscr (regularizers.jl:59-79); errors on NaN like :77
  d = similar(w)
  check(ccall((:gync_ebmodel_dhz, LIBGYNC), Cint, (Ptr{Void}, Ptr{Cdouble}, Cint, Ptr{Cdouble}), h, w, DHZCONT ? 1 : 0, d)); d
end

# em(m, w0, niter) (likelihoodmodel.jl:17-20)
function em(h::Ptr{Void}, w0::Vector{Float64}, niter::Int)
  w, H = copy(w0), Array(Float64, length(w0), max(niter - 1, 0))
  niter > 1 && check(ccall((:gync_ebmodel_em, LIBGYNC), Cint, (Ptr{Void}, Ptr{Cdouble}, Cint, Ptr{Cdouble}), h, w, niter - 1, H))
  vcat(Vector[w0], [H[:, i] for i in 1:size(H, 2)])
end

# mple(m, reg, h, niter; autoh, w0) (likelihoodmodel.jl:33-41): the whole gradientascent loop on the device
function mple(hm::Ptr{Void}, ndata::Int, reg, h, niter; autoh=true, w0=Float64[])
  autoh && (h = h / ((1 - reg) * (ndata / (1 / 100)) + reg / (1 / 1000)))
  w, H = copy(w0), Array(Float64, length(w0), max(niter - 1, 0))
  niter > 1 && check(ccall((:gync_ebmodel_mple, LIBGYNC), Cint,
                           (Ptr{Void}, Ptr{Cdouble}, Cdouble, Cdouble, Cint, Cint, Ptr{Cdouble}),
                           hm, w, reg, h, niter - 1, DHZCONT ? 1 : 0, H))
  vcat(Vector[w0], [H[:, i] for i in 1:size(H, 2)])
end

# projectsimplex!(y) (projectsimplex.jl:8)
function projectsimplex!(y::Vector{Float64})
  check(ccall((:gync_projectsimplex, LIBGYNC), Cint, (Ptr{Cdouble}, Int64), y, length(y))); y
end

# --- WeightedChain(samplings, burnin) (src/gync/weightedchain.jl:29-71): the whole constructor on the device ---
function collapseduplicates(S::Matrix{Float64})          # weightedchain.jl:31-46 on the concatenated chains
  K = size(S, 1)
  rows, counts, n = Array(Float64, K, 116), Array(Int64, K), Int64[0]
  check(ccall((:gync_collapse_duplicates, LIBGYNC), Cint, (Ptr{Cdouble}, Int64, Ptr{Cdouble}, Ptr{Int64}, Ptr{Int64}),
              S, K, rows, counts, n))
  rows[1:n[1], :], counts[1:n[1]]
end

function WeightedChain(samplings::Vector{Sampling}, burnin=0)
  S = vcat([s.samples[(1+burnin):end, :] for s in samplings]...)
  samples, counts = collapseduplicates(S)                                            # :31-46
  K  = size(samples, 1)
  lp = Array(Float64, K)
  check(ccall((:gync_logprior_batch, LIBGYNC), Cint, (Ptr{Cdouble}, Int64, Ptr{Cdouble}, Ptr{Cdouble}),
              samples, K, priorvector(samplings[1].config), lp))                     # :49-51
  LL = loglikmatrix(samples, [data(s) for s in samplings], [s.config.sigma_rho for s in samplings])   # :53-56, one solve per sample
  L, w, d = wcnormalize(LL, lp, counts)                                              # :59-67
  WeightedChain(samples, L, w, d)                                                    # :69-70
end

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
function gync(y0::Vector, p::Vector, t::Vector; reltol=1e-8, abstol=1e-10)
  Y  = Array(Float64, length(t), length(y0))           # N = 1: exactly the reference's nT x 33 layout
  st = Cint[0]
  o  = Ref(SolverOpts(reltol, abstol))
  check(ccall((:gync_solve, LIBGYNC), Cint,
              (Ptr{Cdouble}, Ptr{Cdouble}, Int64, Ptr{Cdouble}, Cint, Ptr{SolverOpts}, Ptr{Cdouble}, Ptr{Cint}),
              y0, p, 1, convert(Vector{Float64}, t), length(t), o, Y, st))
  Y                                                    # all-NaN when st[1] != 0 (gyncycle.jl:18-19)
end

# --- llh(c,x) (model.jl:128-155), llh(s::Sampling) (sampling.jl:32) --------------------------
function loglikmatrix(X::Matrix{Float64}, datas::Vector{Matrix{Float64}}, sigmas::Vector{Float64}; reltol=1e-8, abstol=1e-10)
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
end

function sample!(s::Sampling, iters::Int)
  c, v  = s.config, s.variate::B200Variate
  thin  = c.thin
  n     = round(Int, iters / thin, RoundDown)
  out   = Array(Float64, n, 116)
  nacc  = Int64[0]
  state = reshape(copy(v.state), 1, 116)
  lf    = [v.logf]
  o     = Ref(SolverOpts(1e-8, 1e-10))
  check(ccall((:gync_mcmc_run, LIBGYNC), Cint,
              (Ptr{Cdouble}, Ptr{Cdouble}, Int64, Ptr{Cdouble}, Cdouble, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Ptr{Cint},
               Ptr{Cdouble}, Ptr{SolverOpts}, Int64, Cint, UInt64, UInt64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Int64}),
              state, lf, 1, v.SigmaL, 0.0, data(c), 1, [c.sigma_rho], Cint[0], priorvector(c), o,
              iters, thin, v.seed, v.iteration, C_NULL, C_NULL, out, nacc))
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

# B200Solver.jl — Julia shim binding libfpsb200.so behind FletcherPenaltySolver.jl's QDSolver surface.
#
# UNEXECUTED in this repository's environment (no Julia in the image); it documents exactly what a
# maintainer adds to the reference to use the B200 path.  Include it after
# `src/solve_linear_system.jl` (e.g. at the end of src/model-Fletcherpenaltynlp.jl) and register the
# two solver types in `qdsolver_correspondence` (src/parameters.jl:197):
#
#     const qdsolver_correspondence = Dict(:iterative => IterativeSolver, :ldlt => LDLtSolver,
#                                          :b200_iterative => B200IterativeSolver,
#                                          :b200_ldlt => B200LDLtSolver)
#
# then `fps_solve(nlp; qds_solver = :b200_ldlt)` (src/parameters.jl:290, :299) selects it.
# Only `ccall` is used: no CUDA.jl, no kernel generation on the Julia side.

const libfpsb = get(ENV, "FPSB200_LIB", "libfpsb200.so")

const FPSB_HOST = Cint(0)
const FPSB_DEVICE = Cint(1)

struct FpsbKrylovStats            # fpsb_krylov_stats (include/fpsb.h)
  niter::Int64
  solved::Int32
  inconsistent::Int32
  status::Int32
  pad::Int32
  rnorm::Float64
  arnorm::Float64
  anorm::Float64
  acond::Float64
  xnorm::Float64
end

struct FpsbIterOpts               # fpsb_iter_opts
  ls_atol::Float64; ls_rtol::Float64; ls_itmax::Int64
  ln_atol::Float64; ln_rtol::Float64; ln_btol::Float64; ln_conlim::Float64; ln_itmax::Int64
  ne_atol::Float64; ne_rtol::Float64; ne_etol::Float64; ne_conlim::Float64; ne_itmax::Int64
end

struct FpsbLdltOpts               # fpsb_ldlt_opts
  ldlt_tol::Float64; ldlt_r1::Float64; ldlt_r2::Float64
end

function _fpsb_check(rc::Cint, what)
  rc == 0 && return nothing
  msg = unsafe_string(ccall((:fpsb_last_error, libfpsb), Cstring, ()))
  error("$what failed (code $rc): $msg")      # ABI misuse / CUDA failure only, never numerical failure
end

mutable struct FpsbHandle
  ptr::Ptr{Cvoid}
  function FpsbHandle(nvar, ncon, jrows::Vector{Int}, jcols::Vector{Int}; device = 0)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    # index_base = 1: jac_structure! returns 1-based COO (src/solve_two_systems_struct.jl:333-337)
    rc = ccall((:fpsb_create, libfpsb), Cint,
      (Int64, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Cint, Cint, Ref{Ptr{Cvoid}}),
      nvar, ncon, length(jrows), jrows, jcols, 1, device, out)
    _fpsb_check(rc, "fpsb_create")
    h = new(out[])
    finalizer(x -> ccall((:fpsb_destroy, libfpsb), Cint, (Ptr{Cvoid},), x.ptr), h)
    return h
  end
end

function _structure(nlp, explicit_linear_constraints)
  nnzj = explicit_linear_constraints ? nlp.meta.nln_nnzj : nlp.meta.nnzj
  rows, cols = zeros(Int, nnzj), zeros(Int, nnzj)
  explicit_linear_constraints ? jac_nln_structure!(nlp, rows, cols) : jac_structure!(nlp, rows, cols)
  ncon = explicit_linear_constraints ? nlp.meta.nnln : nlp.meta.ncon
  return rows, cols, nnzj, ncon
end

# Every caller-owned vector may be a SubArray (test/nlpmodelstest.jl:46-48): copy if not contiguous.
_dense(v) = v isa Vector{Float64} ? v : Vector{Float64}(v)

"""
    B200IterativeSolver(nlp, ::T; kwargs...) <: QDSolver

Same keywords as `IterativeSolver` (src/solve_two_systems_struct.jl:94-131); LSQR / CRAIG / MINRES run
as fused CUDA kernels on the device-resident Jacobian.
"""
struct B200IterativeSolver{T} <: QDSolver
  h::FpsbHandle
  jvals::Vector{T}
  p1::Vector{T}; q1::Vector{T}; p2::Vector{T}; q2::Vector{T}   # solver-owned outputs
  u1::Vector{T}; u2::Vector{T}                                 # extras must not clobber p2 (SURVEY §8b)
  stats::Vector{FpsbKrylovStats}
end

function B200IterativeSolver(nlp::AbstractNLPModel{T, S}, ::T; explicit_linear_constraints = false,
    ls_atol::T = √eps(T), ls_rtol::T = √eps(T), ls_itmax::Integer = -1,
    ln_atol::T = √eps(T), ln_rtol::T = √eps(T), ln_btol::T = √eps(T), ln_conlim::T = 1 / √eps(T),
    ln_itmax::Integer = -1, ne_atol::T = √eps(T), ne_rtol::T = √eps(T), ne_etol::T = √eps(T),
    ne_itmax::Int = 0, ne_conlim::T = 1 / √eps(T), kwargs...) where {T, S}
  T == Float64 || error("libfpsb200 computes in Float64")
  rows, cols, nnzj, ncon = _structure(nlp, explicit_linear_constraints)
  nvar = nlp.meta.nvar
  h = FpsbHandle(nvar, ncon, rows, cols)
  itd = 5 * (ncon + nvar)
  opts = Ref(FpsbIterOpts(ls_atol, ls_rtol, ls_itmax < 0 ? itd : ls_itmax, ln_atol, ln_rtol, ln_btol,
    ln_conlim, ln_itmax < 0 ? itd : ln_itmax, ne_atol, ne_rtol, ne_etol, ne_conlim, ne_itmax))
  _fpsb_check(ccall((:fpsb_iter_setup, libfpsb), Cint, (Ptr{Cvoid}, Ref{FpsbIterOpts}), h.ptr, opts),
    "fpsb_iter_setup")
  return B200IterativeSolver{T}(h, zeros(T, nnzj), zeros(T, nvar), zeros(T, ncon), zeros(T, nvar),
    zeros(T, ncon), zeros(T, ncon), zeros(T, ncon), Vector{FpsbKrylovStats}(undef, 2))
end

"""
    B200LDLtSolver(nlp, ::T; ldlt_tol, ldlt_r1, ldlt_r2, P = nothing, ordering = :auto, kwargs...) <: QDSolver

`ldl_analyze` happens here (host ordering + symbolic analysis, uploaded to the GPU); `P` (1-based)
plays the role of `ldl_analyze(A, P)`.  Without `P`, `ordering` picks the built-in one: `:amd` (minimum
degree, what the reference does), `:dissection` (shallow elimination tree for the GPU) or `:auto`
(minimum degree below 20 000 unknowns, dissection from there on — same rule as the Python mirror).
"""
struct B200LDLtSolver{T} <: QDSolver
  h::FpsbHandle
  jvals::Vector{T}
  p1::Vector{T}; q1::Vector{T}; p2::Vector{T}; q2::Vector{T}
  u1::Vector{T}; u2::Vector{T}
  stats::Vector{FpsbKrylovStats}
end

function B200LDLtSolver(nlp::AbstractNLPModel{T, S}, ::T; explicit_linear_constraints = false,
    ldlt_tol = √eps(T), ldlt_r1 = √eps(T), ldlt_r2 = -√eps(T), P = nothing, ordering = :auto,
    kwargs...) where {T, S}
  T == Float64 || error("libfpsb200 computes in Float64")
  rows, cols, nnzj, ncon = _structure(nlp, explicit_linear_constraints)
  nvar = nlp.meta.nvar
  h = FpsbHandle(nvar, ncon, rows, cols)
  if P === nothing && (ordering == :dissection || (ordering == :auto && nvar + ncon >= 20_000))
    P = Vector{Int64}(undef, nvar + ncon)
    _fpsb_check(ccall((:fpsb_order_dissection, libfpsb), Cint,
      (Int64, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Cint, Cint, Ptr{Int64}),
      nvar, ncon, nnzj, rows, cols, 1, 0, P), "fpsb_order_dissection")
  end
  opts = Ref(FpsbLdltOpts(ldlt_tol, ldlt_r1, ldlt_r2))
  Pv = P === nothing ? Int64[] : Vector{Int64}(P)     # the converted copy is what the ccall reads: root IT
  GC.@preserve Pv begin
    Pp = P === nothing ? Ptr{Int64}(C_NULL) : pointer(Pv)
    _fpsb_check(ccall((:fpsb_ldlt_analyze, libfpsb), Cint,
      (Ptr{Cvoid}, Ptr{Int64}, Cint, Ref{FpsbLdltOpts}), h.ptr, Pp, 1, opts), "fpsb_ldlt_analyze")
  end
  return B200LDLtSolver{T}(h, zeros(T, nnzj), zeros(T, nvar), zeros(T, ncon), zeros(T, nvar),
    zeros(T, ncon), zeros(T, ncon), zeros(T, ncon), Vector{FpsbKrylovStats}(undef, 2))
end

function _refresh!(nlp, qds, x)
  # jac_coord! straight into the staging vector (src/solve_linear_system.jl:224-228)
  if nlp.explicit_linear_constraints
    jac_nln_coord!(nlp.nlp, x, qds.jvals)
  else
    jac_coord!(nlp.nlp, x, qds.jvals)
  end
  _fpsb_check(ccall((:fpsb_set_jac_values, libfpsb), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint),
    qds.h.ptr, qds.jvals, FPSB_HOST), "fpsb_set_jac_values")
end

# ---- solve_two_mixed (src/solve_linear_system.jl:107-140 and :206-252) ---------------------------
function solve_two_mixed(nlp::FletcherPenaltyNLP{T, S, A, P, B200IterativeSolver{T}}, x::AbstractVector,
    rhs1, rhs2) where {T, S, A, P}
  q = nlp.qdsolver
  _refresh!(nlp, q, x)
  r1, r2 = _dense(rhs1), _dense(rhs2)
  _fpsb_check(ccall((:fpsb_iter_solve_two_mixed, libfpsb), Cint,
    (Ptr{Cvoid}, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
     Ptr{Float64}, Cint, Ptr{FpsbKrylovStats}),
    q.h.ptr, Float64(nlp.δ), r1, r2, q.p1, q.q1, q.p2, q.q2, FPSB_HOST, q.stats), "fpsb_iter_solve_two_mixed")
  q.stats[1].solved != 0 || @warn "Failed solving 1st linear system lsqr in mixed."
  q.stats[2].solved != 0 || @warn "Failed solving 2nd linear system craig in mixed."
  return q.p1, q.q1, q.p2, q.q2
end

function solve_two_mixed(nlp::FletcherPenaltyNLP{T, S, A, P, B200LDLtSolver{T}}, x::AbstractVector,
    rhs1, rhs2) where {T, S, A, P}
  q = nlp.qdsolver
  _refresh!(nlp, q, x)
  ok = Ref{Cint}(0)
  r1, r2 = _dense(rhs1), _dense(rhs2)
  _fpsb_check(ccall((:fpsb_ldlt_solve_two_mixed, libfpsb), Cint,
    (Ptr{Cvoid}, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
     Ptr{Float64}, Cint, Ref{Cint}),
    q.h.ptr, Float64(nlp.δ), r1, r2, q.p1, q.q1, q.p2, q.q2, FPSB_HOST, ok), "fpsb_ldlt_solve_two_mixed")
  ok[] != 0 || @warn "_solve_ldlt_factorization: failed _factorization"
  return q.p1, q.q1, q.p2, q.q2
end

# ---- solve_two_least_squares (:79-105 and :161-204) ------------------------------------------------
function solve_two_least_squares(nlp::FletcherPenaltyNLP{T, S, A, P, B200IterativeSolver{T}},
    x::AbstractVector, rhs1, rhs2) where {T, S, A, P}
  q = nlp.qdsolver
  r1, r2 = _dense(rhs1), _dense(rhs2)
  _fpsb_check(ccall((:fpsb_iter_solve_two_least_squares, libfpsb), Cint,
    (Ptr{Cvoid}, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
     Ptr{Float64}, Cint, Ptr{FpsbKrylovStats}),
    q.h.ptr, Float64(nlp.δ), r1, r2, q.p1, q.q1, q.p2, q.q2, FPSB_HOST, q.stats),
    "fpsb_iter_solve_two_least_squares")
  q.stats[1].solved != 0 || @warn "Failed solving 1st linear system lsqr."
  q.stats[2].solved != 0 || @warn "Failed solving 2nd linear system lsqr."
  return q.p1, q.q1, q.p2, q.q2
end

function solve_two_least_squares(nlp::FletcherPenaltyNLP{T, S, A, P, B200LDLtSolver{T}},
    x::AbstractVector, rhs1, rhs2) where {T, S, A, P}
  q = nlp.qdsolver
  ok = Ref{Cint}(0)
  r1, r2 = _dense(rhs1), _dense(rhs2)
  _fpsb_check(ccall((:fpsb_ldlt_solve_two_least_squares, libfpsb), Cint,
    (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
     Cint, Ref{Cint}),
    q.h.ptr, r1, r2, q.p1, q.q1, q.p2, q.q2, FPSB_HOST, ok), "fpsb_ldlt_solve_two_least_squares")
  ok[] != 0 || @warn "_solve_ldlt_factorization: failed _factorization"
  return q.p1, q.q1, q.p2, q.q2
end

# ---- solve_two_extras (:45-77 and :142-159) ---------------------------------------------------------
function solve_two_extras(nlp::FletcherPenaltyNLP{T, S, A, P, B200IterativeSolver{T}}, x::AbstractVector,
    rhs1, rhs2) where {T, S, A, P}
  q = nlp.qdsolver
  r1, r2 = _dense(rhs1), _dense(rhs2)
  _fpsb_check(ccall((:fpsb_iter_solve_two_extras, libfpsb), Cint,
    (Ptr{Cvoid}, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint,
     Ptr{FpsbKrylovStats}),
    q.h.ptr, Float64(nlp.δ), r1, r2, q.u1, q.u2, FPSB_HOST, q.stats), "fpsb_iter_solve_two_extras")
  q.stats[1].solved != 0 || @warn "Failed solving 1st linear system lsqr in extra."
  q.stats[2].solved != 0 || @warn "Failed solving 2nd linear system minres in extra."
  return q.u1, q.u2
end

function solve_two_extras(nlp::FletcherPenaltyNLP{T, S, A, P, B200LDLtSolver{T}}, x::AbstractVector,
    rhs1, rhs2) where {T, S, A, P}
  q = nlp.qdsolver
  _refresh!(nlp, q, x)        # the reference re-evaluates jac_op here (:149-153)
  r1, r2 = _dense(rhs1), _dense(rhs2)
  _fpsb_check(ccall((:fpsb_ldlt_solve_two_extras, libfpsb), Cint,
    (Ptr{Cvoid}, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint,
     Ptr{FpsbKrylovStats}),
    q.h.ptr, Float64(nlp.δ), r1, r2, q.u1, q.u2, FPSB_HOST, q.stats), "fpsb_ldlt_solve_two_extras")
  return q.u1, q.u2
end

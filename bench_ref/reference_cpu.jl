# reference_cpu.jl — the TRUE reference arm of bench.py: FletcherPenaltySolver.jl's own solve_two_mixed
# (IterativeSolver = Krylov.jl LSQR + CRAIG, and LDLtSolver = LDLFactorizations.jl) on the workload dumped
# by bench_ref/dump_workload.py.
#
# UNEXECUTED in this repository's environment: there is no Julia in the image or on the GPU box, which is why
# `bench.py --impl reference` times the C restatement under oracle/ instead (DESIGN.md §2).  For whoever has
# Julia:   julia --project=/path/to/FletcherPenaltySolver.jl bench_ref/reference_cpu.jl dump_dir [nsolves]
# prints one JSON line per solver in bench.py's units (2-RHS KKT solves/s).
using LinearAlgebra, SparseArrays, Printf
using NLPModels, FletcherPenaltySolver

struct DumpedModel{T, S} <: AbstractNLPModel{T, S}
  meta::NLPModelMeta{T, S}
  counters::Counters
  jrow::Vector{Int}
  jcol::Vector{Int}
  jval::Vector{T}
  A::SparseMatrixCSC{T, Int}
end

function DumpedModel(dir)
  n, m, nnzj = parse.(Int, split(read(joinpath(dir, "meta.txt"), String)))
  rd(T, f, k) = (v = Vector{T}(undef, k); read!(joinpath(dir, f), v); v)
  jrow, jcol = rd(Int64, "jrow.i64", nnzj), rd(Int64, "jcol.i64", nnzj)
  jval = rd(Float64, "jval.f64", nnzj)
  meta = NLPModelMeta(n, ncon = m, nnzj = nnzj, x0 = zeros(n), lcon = zeros(m), ucon = zeros(m), name = "dumped-qp")
  return DumpedModel(meta, Counters(), jrow, jcol, jval, sparse(jrow, jcol, jval, m, n)), rd(Float64, "rhs1.f64", n), rd(Float64, "rhs2.f64", m)
end

# a linear-constraint model: exactly the calls the hot path makes (SURVEY §2.2)
NLPModels.obj(nlp::DumpedModel, x::AbstractVector) = dot(x, x) / 2
NLPModels.grad!(nlp::DumpedModel, x::AbstractVector, g::AbstractVector) = (g .= x)
NLPModels.cons!(nlp::DumpedModel, x::AbstractVector, c::AbstractVector) = mul!(c, nlp.A, x)
function NLPModels.jac_structure!(nlp::DumpedModel, rows::AbstractVector{<:Integer}, cols::AbstractVector{<:Integer})
  rows .= nlp.jrow; cols .= nlp.jcol
  return rows, cols
end
NLPModels.jac_coord!(nlp::DumpedModel, x::AbstractVector, vals::AbstractVector) = (vals .= nlp.jval)
NLPModels.jprod!(nlp::DumpedModel, x::AbstractVector, v::AbstractVector, Jv::AbstractVector) = mul!(Jv, nlp.A, v)
NLPModels.jtprod!(nlp::DumpedModel, x::AbstractVector, v::AbstractVector, Jtv::AbstractVector) = mul!(Jtv, nlp.A', v)
NLPModels.hprod!(nlp::DumpedModel, x::AbstractVector, y::AbstractVector, v::AbstractVector, Hv::AbstractVector; obj_weight = 1.0) = (Hv .= obj_weight .* v)

function main()
  dir = ARGS[1]
  nsolves = length(ARGS) > 1 ? parse(Int, ARGS[2]) : 3
  nlp, rhs1, rhs2 = DumpedModel(dir)
  x = nlp.meta.x0
  for (name, qds) in (("iterative", FletcherPenaltySolver.IterativeSolver(nlp, 0.0)), ("ldlt", FletcherPenaltySolver.LDLtSolver(nlp, 0.0)))
    fp = FletcherPenaltyNLP(nlp, 1.0, 0.0, name == "ldlt" ? sqrt(eps()) : 0.0, Val(2); qds = qds)
    FletcherPenaltySolver.solve_two_mixed(fp, x, rhs1, rhs2)            # warm-up (compilation)
    t = @elapsed for _ = 1:nsolves
      FletcherPenaltySolver.solve_two_mixed(fp, x, rhs1, rhs2)          # src/solve_linear_system.jl:107-140 / 206-252
    end
    @printf("{\"impl\": \"reference-julia\", \"qds_solver\": \"%s\", \"metric\": \"2-RHS KKT solves/s\", \"value\": %.6f, \"ms_per_step\": %.3f, \"threads\": %d}\n",
            name, nsolves / t, 1e3 * t / nsolves, Threads.nthreads())
  end
end

main()

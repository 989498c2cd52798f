# IIFB200.jl — reference-side binding of libiifb200.so (cannot be executed in the build image: no Julia).
#
# Drop-in boundary B3 (SURVEY.md §8b): overrides IncrementalInference.propagateBelief for graphs whose
# SolverParams.devParams[:backend] == "b200" and forwards to the C-ABI declared in include/iifb200.h.
# Everything above (factor graph, Bayes tree, CliqueStateMachine, solveTree!) is unchanged Julia.
module IIFB200

using IncrementalInference
using DistributedFactorGraphs
import IncrementalInference: propagateBelief
const IIF = IncrementalInference
const AMP = IIF.ApproxManifoldProducts

const LIB = get(ENV, "IIFB200_LIB", joinpath(@__DIR__, "..", "incrementalinference.jl_b200", "csrc", "libiifb200.so"))

# ---- mirrors of the C structs (include/iifb200.h) -------------------------------------------------
const MAX_DIM, MAX_ARITY, MAX_FACTORS = 4, 6, 8
struct SlotDesc;   dim::Int32; circ_mask::Int32; cap::Int32; pts_off::Int32; end
struct DistDesc;   kind::Int32; dim::Int32; ncomp::Int32; comp_kind::Int32; slot::Int32; poff::Int32; end
struct FactorDesc
  kind::Int32; arity::Int32; zdim::Int32; dist::Int32
  slot::NTuple{MAX_ARITY,Int32}; nmh::Int32; partial_mask::Int32
  mh::NTuple{MAX_ARITY,Float64}; nullhypo::Float64; inflation::Float64
end
struct SolverParamsC; spreadNH::Float64; nullSurplusAdd::Float64; inflateCycles::Int32; gibbsNiter::Int32; seed::UInt64; end
struct PropOp
  target_slot::Int32; out_slot::Int32; nfactors::Int32; N::Int32
  factor::NTuple{MAX_FACTORS,Int32}; sfidx::NTuple{MAX_FACTORS,Int32}; call_id::Int32; any_multihypo::Int32
end
struct SchedOp;  kind::Int32; a::Int32; b::Int32; lane::Int32; end      # kind: 1 PROPAGATE, 2 COPY, 3 DECONV
struct DeconvOp; factor::Int32; out_slot::Int32; N::Int32; call_id::Int32; end

check(ctx, st, what) = st == 0 || error("iifb200 $what failed ($st): " *
        unsafe_string(ccall((:iifb200_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx)))

function init(device::Integer = 0)
  ctx = Ref{Ptr{Cvoid}}(C_NULL)
  st = ccall((:iifb200_init, LIB), Int32, (Int32, Ref{Ptr{Cvoid}}), device, ctx)
  st == 0 || error("iifb200_init failed: " * unsafe_string(ccall((:iifb200_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
  return ctx[]
end

# factor kind / distribution lowering: only the built-in residual library runs on the device
factorkind(::Prior) = Int32(1); factorkind(::LinearRelative) = Int32(2)
factorkind(::PriorCircular) = Int32(3); factorkind(::CircularCircular) = Int32(4)
factorkind(::EuclidDistance) = Int32(5); factorkind(::IIF.MsgPrior) = Int32(6)
factorkind(::IIF.PartialPrior) = Int32(7)
factorkind(::ManifoldPrior) = Int32(8); factorkind(::IIF.ManifoldPriorPartial) = Int32(8)   # p is folded into Z's mean by _lower_factors
factorkind(f::ManifoldFactor) = _manifoldfactorkind(f.M)
_manifoldfactorkind(::Manifolds.SpecialEuclidean{2}) = Int32(9)     # hybrid tangent representation (testSpecialEuclidean2Mani.jl:14)
_manifoldfactorkind(::Manifolds.TranslationGroup) = Int32(2)
_manifoldfactorkind(::Manifolds.RealCircleGroup) = Int32(4)
_manifoldfactorkind(M) = error("IIFB200: ManifoldFactor on $(M) has no device residual")
factorkind(f) = error("IIFB200: factor $(typeof(f)) has no device residual (no CPU fallback on the b200 backend)")

# circular-coordinate mask of a variable type; SpecialEuclidean(2) points ArrayPartition(t, R) travel as (t1, t2, theta)
circmask(::Type{<:IIF.Circular}) = Int32(1)
circmask(::Type{T}) where {T <: InferenceVariable} = getManifold(T) isa Manifolds.SpecialEuclidean{2} ? Int32(0b100) : Int32(0)
circmask(::Any) = Int32(0)
se2coords(p) = (p.x[1][1], p.x[1][2], atan(p.x[2][2, 1], p.x[2][1, 1]))           # AMP.makeCoordsFromPoint
se2point(c)  = ArrayPartition(SA[c[1], c[2]], SA[cos(c[3]) -sin(c[3]); sin(c[3]) cos(c[3])])   # AMP.makePointFromCoords

"""
    propagateBelief(dfg, destvar, factors; N, ...)   — GraphProductOperations.jl:16-64

b200 backend: lowers the destination variable, its factors and their variables to the descriptor
tables, uploads the particle blocks (zero-copy for `Vector{SVector{d,Float64}}`; `Circular`'s
`Vector{Vector{Float64}}` is packed), runs iifb200_propagate_batch (F convolutions + KDE product) and
rebuilds the ManifoldKernelDensity from the returned points and bandwidths.
"""
function propagateBelief(dfg::AbstractDFG, destvar::DFGVariable, factors::AbstractVector;
                         solveKey::Symbol = :default, N::Integer = getSolverParams(dfg).N, kw...)
  get(getSolverParams(dfg).devParams, :backend, "") == "b200" ||
    return invoke(propagateBelief, Tuple{AbstractDFG, DFGVariable, AbstractVector}, dfg, destvar, factors; solveKey, N, kw...)
  ctx = _ctx()
  vars, slotof = _collect_variables(dfg, destvar, factors)           # labels -> slot index
  slots   = [SlotDesc(getDimension(v), circmask(getVariableType(v)), max(N, length(getVal(v; solveKey))), 0) for v in vars]
  dists, dparams, fdescs = _lower_factors(dfg, factors, slotof)       # Normal / MvNormal / Mixture / MsgPrior(MKD)
  sp = getSolverParams(dfg)
  spc = Ref(SolverParamsC(sp.spreadNH, sp.nullSurplusAdd, sp.inflateCycles, 1, rand(UInt64)))
  GC.@preserve slots dists dparams fdescs begin
    check(ctx, ccall((:iifb200_set_graph, LIB), Int32,
          (Ptr{Cvoid}, Int32, Ptr{SlotDesc}, Int32, Ptr{FactorDesc}, Int32, Ptr{DistDesc}, Int32, Ptr{Float64}, Ref{SolverParamsC}, Ptr{Cvoid}),
          ctx, length(slots), slots, length(fdescs), fdescs, length(dists), dists, length(dparams), dparams, spc, C_NULL), "set_graph")
    for (i, v) in enumerate(vars)
      pts = _packpoints(getVal(v; solveKey))                         # reinterpret(Float64, val) when contiguous
      bw  = getBW(v; solveKey)[:, 1]
      check(ctx, ccall((:iifb200_upload_belief, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Int32),
            ctx, i - 1, size(pts, 2), pts, bw, isInitialized(v, solveKey)), "upload_belief")
    end
    dest = slotof[getLabel(destvar)]
    op = Ref(PropOp(dest, dest, length(fdescs), N,
                    ntuple(i -> i <= length(fdescs) ? Int32(i - 1) : Int32(0), MAX_FACTORS),
                    ntuple(i -> i <= length(fdescs) ? Int32(findfirst(==(getLabel(destvar)), getVariableOrder(factors[i]))) : Int32(0), MAX_FACTORS),
                    0, any(IIF.isMultihypo.(factors))))
    # blocking ccall on a dedicated thread so the other cliques' Tasks keep running (SolverAPI.jl:59-96)
    st = @threadcall((:iifb200_propagate_batch, LIB), Int32, (Ptr{Cvoid}, Int32, Ref{PropOp}), ctx, 1, op)
    check(ctx, st, "propagate_batch")
    d = getDimension(destvar)
    pts = Matrix{Float64}(undef, d, N); bw = zeros(d); ipc = zeros(d); npts = Ref{Int32}(0)
    check(ctx, ccall((:iifb200_download_belief, LIB), Int32, (Ptr{Cvoid}, Int32, Ref{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
          ctx, dest, npts, pts, bw, ipc), "download_belief")
  end
  M = getManifold(getVariableType(destvar))
  mkd = AMP.manikde!(M, _unpackpoints(getVariableType(destvar), pts); bw)   # bw given => no re-selection (FGOSUtils.jl:118-128)
  return mkd, ipc
end

"""
    calcPPE(var, varType; solveKey)   — FGOSUtils.jl:237-278 (setPPE! at CSM step 5 funnels through it)

b200 backend: mean and KDE-max of a belief that is resident on the device (`slot` = its slot index in the
tables uploaded by the enclosing clique solve) from one `iifb200_ppe_batch` launch.
"""
function calcPPE_b200(slots::Vector{Int32})
  ctx = _ctx(); V = length(slots)
  mean = zeros(4, V); mx = zeros(4, V)
  check(ctx, ccall((:iifb200_ppe_batch, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}),
        ctx, V, slots, mean, mx), "ppe_batch")
  return mean, mx      # column v: coordinates of belief v; MeanMaxPPE(suggested = mean, max = mx, mean = mean)
end

"""
    approxDeconv(dfg, fctsym)   — DeconvUtils.jl:176-202,  mmd(p1, p2, varType) — SolverUtilities.jl:25-47
"""
function approxDeconv_b200(factor::Integer, N::Integer, zdim::Integer)
  ctx = _ctx()
  pred = zeros(zdim, N); meas = zeros(zdim, N)
  check(ctx, ccall((:iifb200_deconv_batch, LIB), Int32,
        (Ptr{Cvoid}, Int32, Ref{Int32}, Ref{Int32}, Ref{Int32}, Ptr{Float64}, Ptr{Float64}),
        ctx, 1, Int32(factor), Int32(N), Int32(rand(0:2^30)), pred, meas), "deconv_batch")
  return pred, meas
end

function mmd_b200(a::Matrix{Float64}, b::Matrix{Float64}; circmask::Integer = 0, bw::Float64 = 0.001)
  ctx = _ctx(); out = Ref(0.0)
  check(ctx, ccall((:iifb200_mmd, LIB), Int32,
        (Ptr{Cvoid}, Int32, Ref{Int32}, Ref{Int32}, Ref{Int32}, Ref{Int32}, Ptr{Float64}, Ptr{Float64}, Float64, Ref{Float64}),
        ctx, 1, Int32(size(a, 2)), Int32(size(b, 2)), Int32(size(a, 1)), Int32(circmask), a, b, bw, out), "mmd")
  return out[]
end

"""
    schedule(ctx, wave_off, ops, props, deconvs) -> id     — boundary B4 (whole clique / whole tree per replay)

`ops` in wave order (`wave_off[w]:wave_off[w+1]` are mutually independent), `props` the propagateBelief descriptors,
`deconvs` the differential-likelihood constructions of `useMsgLikelihoods=true` up messages
(addLikelihoodsDifferentialCHILD!, TreeMessageUtils.jl:279-335).  `SchedOp.lane` marks independent sub-trees that
the captured CUDA graph runs as parallel branches (0 = none).
"""
function schedule(ctx, wave_off::Vector{Int32}, ops::Vector{SchedOp}, props::Vector{PropOp}, deconvs::Vector{DeconvOp} = DeconvOp[])
  id = Ref{Int32}(-1)
  check(ctx, ccall((:iifb200_schedule_build_ex, LIB), Int32,
        (Ptr{Cvoid}, Int32, Ptr{Int32}, Int32, Ptr{SchedOp}, Int32, Ptr{PropOp}, Int32, Ptr{DeconvOp}, Ref{Int32}),
        ctx, length(wave_off) - 1, wave_off, length(ops), ops, length(props), props, length(deconvs), deconvs, id), "schedule_build_ex")
  return id[]
end
run!(ctx, id; first = 0, last = -1) = check(ctx, ccall((:iifb200_schedule_run, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Int32), ctx, id, first, last), "schedule_run")

# throughput mode (boundary B4): IIF.upGibbsCliqueDensity / localProductAndUpdate! are lowered per tree by
# iifb200_schedule_build and replayed by iifb200_schedule_run; see incrementalinference.jl_b200/tree.py for
# the lowering that a Julia implementation mirrors 1:1 (same descriptor structs).

end # module

# IIFB200.jl — reference-side binding of libiifb200.so.
#
# NOT EXECUTED in the build image (no Julia toolchain there); tests/test_julia_shim.py checks it as text: every
# `ccall` symbol is declared in include/iifb200.h, every struct mirrors its C layout byte for byte, every helper the
# module calls is defined in the module, every package it names is imported.
#
# Two drop-in boundaries (SURVEY.md §8b):
#   B3  `propagateBelief(dfg, destvar, factors)` — GraphProductOperations.jl:16-64.  Graphs whose
#       SolverParams.devParams[:backend] == "b200" forward ONE belief update per call to the C-ABI; everything above
#       (factor graph, Bayes tree, CliqueStateMachine, solveTree!) is unchanged Julia.
#   B4  `solveTree_b200!(dfg, tree)` — the whole up + down pass of a built Bayes tree in one schedule: the clique
#       table (`tree.bt`: parent, frontals, separators, potentials, Gibbs variable classes of setCliqMCIDs!) goes to
#       iifb200_plan_tree, the library lowers it to waves of independent belief updates (what the CSM would do clique
#       by clique: upGibbsCliqueDensity SolveTree.jl:164-239, solveCliqDownFrontalProducts!
#       CliqStateMachineUtils.jl:479-571), beliefs go up once, posteriors come back once.
module IIFB200

using IncrementalInference
using DistributedFactorGraphs
using Distributions
using LinearAlgebra
using Manifolds
using StaticArrays
using RecursiveArrayTools: ArrayPartition
import IncrementalInference: propagateBelief
const IIF = IncrementalInference
const AMP = IIF.ApproxManifoldProducts

const LIB = get(ENV, "IIFB200_LIB", joinpath(@__DIR__, "..", "incrementalinference.jl_b200", "csrc", "libiifb200.so"))

# ---- mirrors of the C structs (include/iifb200.h) -------------------------------------------------
const MAX_DIM, MAX_ARITY, MAX_FACTORS, MAX_POINTS = 4, 6, 16, 256
struct SlotDesc;   dim::Int32; circ_mask::Int32; cap::Int32; pts_off::Int32; end
struct DistDesc;   kind::Int32; dim::Int32; ncomp::Int32; comp_kind::Int32; slot::Int32; poff::Int32; end
struct FactorDesc
  kind::Int32; arity::Int32; zdim::Int32; dist::Int32
  slot::NTuple{6,Int32}; nmh::Int32; partial_mask::Int32; solver::Int32; _pad::Int32
  mh::NTuple{6,Float64}; nullhypo::Float64; inflation::Float64; aux::NTuple{4,Float64}
end
struct SolverParamsC; spreadNH::Float64; nullSurplusAdd::Float64; inflateCycles::Int32; gibbsNiter::Int32; seed::UInt64; end
struct PropOp
  target_slot::Int32; out_slot::Int32; nfactors::Int32; N::Int32
  factor::NTuple{16,Int32}; sfidx::NTuple{16,Int32}; call_id::Int32; any_multihypo::Int32
end
struct SchedOp;  kind::Int32; a::Int32; b::Int32; lane::Int32; end      # kind: 1 PROPAGATE, 2 COPY, 3 DECONV
struct DeconvOp; factor::Int32; out_slot::Int32; N::Int32; call_id::Int32; end
struct GraphDesc
  nvars::Int32; vars::Ptr{SlotDesc}; nfactors::Int32; factors::Ptr{FactorDesc}
  ndists::Int32; dists::Ptr{DistDesc}; nparams::Int32; dparams::Ptr{Float64}
  factor_type::Ptr{Int32}; var_type::Ptr{Int32}; var_relative_kind::Ptr{Int32}; var_relative_type::Ptr{Int32}
  msgprior_type::Int32; _pad::Int32
end
struct TreeDesc
  ncliques::Int32; parent::Ptr{Int32}
  frontal_off::Ptr{Int32}; frontals::Ptr{Int32}
  separator_off::Ptr{Int32}; separators::Ptr{Int32}
  potential_off::Ptr{Int32}; potentials::Ptr{Int32}
  directFrtlMsg_off::Ptr{Int32}; directFrtlMsg::Ptr{Int32}
  msgskip_off::Ptr{Int32}; msgskip::Ptr{Int32}
  itervar_off::Ptr{Int32}; itervar::Ptr{Int32}
  directPriorMsg_off::Ptr{Int32}; directPriorMsg::Ptr{Int32}
end
struct PlanOpts
  N::Int32; gibbsIters::Int32; downIters::Int32; downsolve::Int32
  lanes::Int32; forward_copies::Int32; useMsgLikelihoods::Int32; call_base::Int32
  inflation::Float64
end

check(ctx, st, what) = st == 0 || error("iifb200 $what failed ($st): " *
        unsafe_string(ccall((:iifb200_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx)))

function init(device::Integer = 0)
  ctx = Ref{Ptr{Cvoid}}(C_NULL)
  st = ccall((:iifb200_init, LIB), Int32, (Int32, Ref{Ptr{Cvoid}}), device, ctx)
  st == 0 || error("iifb200_init failed: " * unsafe_string(ccall((:iifb200_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
  return ctx[]
end

# one context per process (one process per GPU; the ordinal comes from the launcher's LOCAL_RANK)
const _CTX = Ref{Ptr{Cvoid}}(C_NULL)
const _CTX_LOCK = ReentrantLock()
const _CALLS = Threads.Atomic{Int32}(0)              # Philox call ids: 16 per belief update, never reused
function _ctx()
  lock(_CTX_LOCK) do
    _CTX[] == C_NULL && (_CTX[] = init(parse(Int, get(ENV, "LOCAL_RANK", "0"))))
    return _CTX[]
  end
end
_nextcall(n::Integer = 16) = Threads.atomic_add!(_CALLS, Int32(n))

# B3 context pool: solveTree! runs sibling cliques as concurrent Tasks (SolverAPI.jl:59-96); every propagateBelief call
# checks a context out (own stream, arena, tables and scratch), so independent calls overlap on the device instead of
# queueing behind one lock.  IIFB200_CONTEXTS sets the pool size (default 8; a lone propagateBelief occupies 3-6 SMs).
const _POOL = Ref{Union{Nothing, Channel{Ptr{Cvoid}}}}(nothing)
function _pool()
  lock(_CTX_LOCK) do
    if _POOL[] === nothing
      k = parse(Int, get(ENV, "IIFB200_CONTEXTS", "8")); dev = parse(Int, get(ENV, "LOCAL_RANK", "0"))
      ch = Channel{Ptr{Cvoid}}(k)
      for _ in 1:k
        put!(ch, init(dev))
      end
      _POOL[] = ch
    end
    return _POOL[]
  end
end
function _with_context(f)
  pool = _pool(); ctx = take!(pool)
  try
    return f(ctx)
  finally
    put!(pool, ctx)
  end
end

# ---- lowering of variables, distributions and factors ---------------------------------------------
# factor kind: only the built-in residual library runs on the device
factorkind(::Prior) = Int32(1); factorkind(::LinearRelative) = Int32(2)
factorkind(::PriorCircular) = Int32(3); factorkind(::CircularCircular) = Int32(4)
factorkind(::EuclidDistance) = Int32(5); factorkind(::IIF.MsgPrior) = Int32(6)
factorkind(::IIF.PartialPrior) = Int32(7)
factorkind(f::ManifoldPrior) = _so3prior(f) ? Int32(10) : Int32(8)                             # 8: p is folded into Z's mean by _lower_dist
factorkind(::IIF.ManifoldPriorPartial) = Int32(8)
factorkind(f::ManifoldFactor) = _manifoldfactorkind(f.M)
factorkind(f::Mixture) = factorkind(f.mechanics)
_manifoldfactorkind(::Manifolds.SpecialEuclidean{2}) = Int32(9)     # hybrid tangent representation (testSpecialEuclidean2Mani.jl:14)
_manifoldfactorkind(::Manifolds.TranslationGroup) = Int32(2)
_manifoldfactorkind(::Manifolds.RealCircleGroup) = Int32(4)
_manifoldfactorkind(::Manifolds.SpecialOrthogonal{3}) = Int32(11)   # rotation-vector coordinates, IIF_MANI_SO3 slots
_manifoldfactorkind(M) = error("IIFB200: ManifoldFactor on $(M) has no device residual")
_so3prior(f) = f isa ManifoldPrior && f.M isa Manifolds.SpecialOrthogonal{3}
_so3coords(R) = collect(Float64, vee(SpecialOrthogonal(3), Matrix(1.0I, 3, 3), log(SpecialOrthogonal(3), Matrix(1.0I, 3, 3), Matrix(R))))
factorkind(f) = error("IIFB200: factor $(typeof(f)) has no device residual (no CPU fallback on the b200 backend)")

# circular-coordinate mask of a variable type; SpecialEuclidean(2) points ArrayPartition(t, R) travel as (t1, t2, theta)
_isse2(T) = getManifold(T) isa Manifolds.SpecialEuclidean{2}
_isso3(T) = getManifold(T) isa Manifolds.SpecialOrthogonal{3}
circmask(T::Type{<:InferenceVariable}) = T <: IIF.Circular ? Int32(1) : (_isse2(T) ? Int32(0b100) : (_isso3(T) ? Int32(0x100) : Int32(0)))
circmask(v::DFGVariable) = circmask(typeof(getVariableType(v)))
se2coords(p) = (p.x[1][1], p.x[1][2], atan(p.x[2][2, 1], p.x[2][1, 1]))           # AMP.makeCoordsFromPoint
se2point(c)  = ArrayPartition(SA[c[1], c[2]], SA[cos(c[3]) -sin(c[3]); sin(c[3]) cos(c[3])])   # AMP.makePointFromCoords

# points <-> the d x N coordinate block the device stores (== Vector{SVector{d,Float64}} in memory)
function _packpoints(vartype, val::AbstractVector)
  d = getDimension(vartype)
  out = Matrix{Float64}(undef, d, length(val))
  se2 = _isse2(typeof(vartype)); so3 = _isso3(typeof(vartype))
  for (n, p) in enumerate(val)
    c = se2 ? se2coords(p) : (so3 ? _so3coords(p) : p)
    for k in 1:d
      out[k, n] = c[k]
    end
  end
  return out
end
function _unpackpoints(vartype, pts::AbstractMatrix{Float64})
  d = size(pts, 1)
  _isse2(typeof(vartype)) && return [se2point(view(pts, :, n)) for n in 1:size(pts, 2)]
  _isso3(typeof(vartype)) && return [exp(SpecialOrthogonal(3), Matrix(1.0I, 3, 3), hat(SpecialOrthogonal(3), Matrix(1.0I, 3, 3), pts[:, n])) for n in 1:size(pts, 2)]
  vartype isa IIF.Circular && return [[pts[1, n]] for n in 1:size(pts, 2)]         # Vector{Vector{Float64}}
  return [SVector{d, Float64}(view(pts, :, n)) for n in 1:size(pts, 2)]
end
_bw(v::DFGVariable, solveKey) = Vector{Float64}(getSolverData(v, solveKey).bw[:, 1])

# an extra device slot that only holds the kernels of a ManifoldKernelDensity (the measurement of a MsgPrior)
struct _BeliefSlot
  vartype::Any
  pts::Matrix{Float64}
  bw::Vector{Float64}
end

# destination first, then the variables of every factor in factor order, each once
function _collect_variables(dfg::AbstractDFG, destvar::DFGVariable, factors::AbstractVector)
  vars = DFGVariable[destvar]
  slotof = Dict{Symbol, Int32}(getLabel(destvar) => Int32(0))
  for f in factors, lbl in getVariableOrder(f)
    haskey(slotof, lbl) && continue
    push!(vars, getVariable(dfg, lbl))
    slotof[lbl] = Int32(length(vars) - 1)
  end
  return vars, slotof
end

# SamplableBelief -> (kind, dim, ncomp, comp_kind, slot) + parameter block; `shift` is added to the mean (ManifoldPrior's p)
function _simpleblock(Z::Normal, shift)
  return Int32(1), 1, Float64[mean(Z) + (shift === nothing ? 0.0 : shift[1]), std(Z)]
end
_simpleblock(Z::Uniform, shift) = (Int32(5), 1, Float64[minimum(Z), maximum(Z)])
function _simpleblock(Z::AbstractMvNormal, shift)
  d = length(Z)
  mu = Vector{Float64}(mean(Z)) .+ (shift === nothing ? zeros(d) : collect(Float64, shift))
  Lc = Matrix(cholesky(Symmetric(Matrix(cov(Z)))).L)
  return Int32(2), d, vcat(mu, vec(permutedims(Lc)))          # row-major lower factor
end
_simpleblock(Z, shift) = error("IIFB200: distribution $(typeof(Z)) has no device sampler")

function _lower_dist!(dists::Vector{DistDesc}, dparams::Vector{Float64}, extra::Vector{_BeliefSlot}, nvars::Int, Z, shift = nothing)
  poff = Int32(length(dparams))
  if Z isa AMP.ManifoldKernelDensity                      # MsgPrior(belief): kernels live in an extra slot
    M = Z.manifold
    vt = M isa Manifolds.RealCircleGroup ? IIF.Circular() : ContinuousEuclid(manifold_dimension(M))
    push!(extra, _BeliefSlot(vt, _packpoints(vt, getPoints(Z, false)), Vector{Float64}(getBW(Z)[:, 1])))
    push!(dists, DistDesc(Int32(4), Int32(manifold_dimension(M)), Int32(0), Int32(0), Int32(nvars + length(extra) - 1), poff))
  else
    kind, d, prm = _simpleblock(Z, shift)
    append!(dparams, prm)
    push!(dists, DistDesc(kind, Int32(d), Int32(0), Int32(0), Int32(-1), poff))
  end
  return Int32(length(dists) - 1)
end

function _lower_mixture!(dists, dparams, mix::Mixture)
  poff = Int32(length(dparams))
  comps = collect(values(mix.components))
  blocks = [_simpleblock(c, nothing) for c in comps]
  length(unique(b[1] for b in blocks)) == 1 && length(unique(b[2] for b in blocks)) == 1 ||
    error("IIFB200: Mixture components must share one distribution kind and dimension")
  append!(dparams, Vector{Float64}(probs(mix.diversity)))
  foreach(b -> append!(dparams, b[3]), blocks)
  push!(dists, DistDesc(Int32(3), Int32(blocks[1][2]), Int32(length(blocks)), blocks[1][1], Int32(-1), poff))
  return Int32(length(dists) - 1)
end

_padtuple(v, n, T) = ntuple(i -> i <= length(v) ? T(v[i]) : zero(T), n)

# factor objects -> descriptor tables; variables are referenced through `slotof`
function _lower_factors(dfg::AbstractDFG, factors::AbstractVector, slotof::Dict{Symbol, Int32}, nvars::Int)
  dists, dparams, fdescs, extra = DistDesc[], Float64[], FactorDesc[], _BeliefSlot[]
  for f in factors
    fnc = getFactorType(f)
    ccw = IIF._getCCW(f)
    vo = getVariableOrder(f)
    length(vo) <= MAX_ARITY || error("IIFB200: factor $(getLabel(f)) has more than $MAX_ARITY variables")
    shift = nothing
    if (fnc isa ManifoldPrior && !_so3prior(fnc)) || fnc isa IIF.ManifoldPriorPartial
      p = fnc isa ManifoldPrior ? fnc.p : nothing        # sample = retract(M, p, hat(Z)) = p (+) Z on these groups
      shift = p === nothing ? nothing : (p isa ArrayPartition ? collect(se2coords(p)) : collect(Float64, p))
    end
    di = fnc isa Mixture ? _lower_mixture!(dists, dparams, fnc) : _lower_dist!(dists, dparams, extra, nvars, fnc.Z, shift)
    pmask = Int32(0)
    if hasfield(typeof(fnc), :partial)
      for c in fnc.partial
        pmask |= Int32(1) << (Int(c) - 1)
      end
    end
    mh = ccw.hyporecipe.hypotheses === nothing ? Float64[] : Vector{Float64}(probs(ccw.hyporecipe.hypotheses))
    push!(fdescs, FactorDesc(factorkind(fnc), Int32(length(vo)), dists[di + 1].dim, di,
                             _padtuple([slotof[l] for l in vo], MAX_ARITY, Int32), Int32(length(mh)), pmask,
                             Int32(get(ENV, "IIFB200_NUMERIC_SOLVE", "0") == "1"), Int32(0),
                             _padtuple(mh, MAX_ARITY, Float64), ccw.nullhypo, ccw.inflation,
                             _padtuple(_so3prior(fnc) ? _so3coords(fnc.p) : Float64[], MAX_DIM, Float64)))
  end
  return dists, dparams, fdescs, extra
end

_solverparams(sp, seed = rand(UInt64)) = SolverParamsC(sp.spreadNH, sp.nullSurplusAdd, sp.inflateCycles, 1, seed)

function _set_graph(ctx, slots, fdescs, dists, dparams, spc)
  GC.@preserve slots dists dparams fdescs begin
    check(ctx, ccall((:iifb200_set_graph, LIB), Int32,
          (Ptr{Cvoid}, Int32, Ptr{SlotDesc}, Int32, Ptr{FactorDesc}, Int32, Ptr{DistDesc}, Int32, Ptr{Float64}, Ref{SolverParamsC}, Ptr{Cvoid}),
          ctx, length(slots), slots, length(fdescs), fdescs, length(dists), dists, length(dparams), dparams, Ref(spc), C_NULL), "set_graph")
  end
end
# all beliefs of a freshly set graph in ONE asynchronous transfer: slot i occupies cap_i * dim_i doubles of `pts`
function _upload_all(ctx, slots::Vector{SlotDesc}, blocks::Vector{Matrix{Float64}}, bws::Vector{Vector{Float64}}, inits::Vector{Bool})
  ns = length(slots)
  pts = zeros(sum(Int(s.cap) * Int(s.dim) for s in slots)); bw = zeros(MAX_DIM, ns)
  npts = zeros(Int32, ns); flags = zeros(Int32, ns)
  off = 0
  for i in 1:ns
    p = blocks[i]
    pts[(off + 1):(off + length(p))] .= vec(p)
    off += Int(slots[i].cap) * Int(slots[i].dim)
    bw[1:length(bws[i]), i] .= bws[i]; npts[i] = size(p, 2); flags[i] = inits[i]
  end
  check(ctx, ccall((:iifb200_upload_slots, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}),
        ctx, 0, ns, pts, bw, npts, flags), "upload_slots")
  return pts   # keep alive until the next synchronising call
end

"""
    propagateBelief(dfg, destvar, factors; N, ...)   — GraphProductOperations.jl:16-64   (boundary B3)

b200 backend: lowers the destination variable, its factors and their variables to the descriptor tables and hands
them, with the particle blocks, to iifb200_propagate_once (F convolutions + KDE product in two launches, one stream
synchronisation), then rebuilds the ManifoldKernelDensity from the returned points and bandwidths.  The library keeps its
arena, tables and scratch between calls (grow-only), so a call costs the transfers and two kernels.
"""
function propagateBelief(dfg::AbstractDFG, destvar::DFGVariable, factors::AbstractVector;
                         solveKey::Symbol = :default, N::Integer = getSolverParams(dfg).N, kw...)
  get(getSolverParams(dfg).devParams, :backend, "") == "b200" ||
    return invoke(propagateBelief, Tuple{AbstractDFG, DFGVariable, AbstractVector}, dfg, destvar, factors; solveKey, N, kw...)
  length(factors) <= MAX_FACTORS || error("IIFB200: $(length(factors)) factors exceed IIF_MAX_FACTORS = $MAX_FACTORS")
  vars, slotof = _collect_variables(dfg, destvar, factors)           # labels -> slot index (destination = slot 0)
  dists, dparams, fdescs, extra = _lower_factors(dfg, factors, slotof, length(vars))
  slots = SlotDesc[SlotDesc(getDimension(v), circmask(v), max(N, length(getVal(v; solveKey)), 1), 0) for v in vars]
  for e in extra
    push!(slots, SlotDesc(size(e.pts, 1), circmask(typeof(e.vartype)), max(size(e.pts, 2), 1), 0))
  end
  blocks = vcat([_packpoints(getVariableType(v), getVal(v; solveKey)) for v in vars], [e.pts for e in extra])
  bws = vcat([_bw(v, solveKey) for v in vars], [e.bw for e in extra])
  inits = vcat(Bool[isInitialized(v, solveKey) for v in vars], fill(true, length(extra)))
  ns = length(slots)
  hpts = zeros(sum(Int(s.cap) * Int(s.dim) for s in slots)); hbw = zeros(MAX_DIM, ns)
  hn = zeros(Int32, ns); hfl = zeros(Int32, ns)
  off = 0
  for i in 1:ns                                                      # slot i occupies cap_i * dim_i doubles of `hpts`
    p = blocks[i]
    hpts[(off + 1):(off + length(p))] .= vec(p)
    off += Int(slots[i].cap) * Int(slots[i].dim)
    hbw[1:length(bws[i]), i] .= bws[i]; hn[i] = size(p, 2); hfl[i] = inits[i]
  end
  dlbl = getLabel(destvar)
  op = Ref(PropOp(0, 0, length(fdescs), N,
                  _padtuple(collect(0:(length(fdescs) - 1)), MAX_FACTORS, Int32),
                  _padtuple([findfirst(==(dlbl), getVariableOrder(f)) for f in factors], MAX_FACTORS, Int32),
                  _nextcall(), any(IIF.isMultihypo.(factors))))
  d = getDimension(destvar)
  pts = Matrix{Float64}(undef, d, N); bw = zeros(MAX_DIM); ipc = zeros(MAX_DIM); npts = Ref{Int32}(0)
  spc = Ref(_solverparams(getSolverParams(dfg)))
  _with_context() do ctx                                             # one context per call in flight (pool above)
    # ONE C-ABI call: descriptor tables + beliefs in, F convolutions + product, posterior out.  Blocking, so it runs on a
    # dedicated thread and the other cliques' Tasks keep going (SolverAPI.jl:59-96)
    GC.@preserve slots fdescs dists dparams hpts hbw hn hfl pts bw ipc begin
      st = @threadcall((:iifb200_propagate_once, LIB), Int32,
                       (Ptr{Cvoid}, Int32, Ptr{SlotDesc}, Int32, Ptr{FactorDesc}, Int32, Ptr{DistDesc}, Int32, Ptr{Float64},
                        Ref{SolverParamsC}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ref{PropOp}, Ref{Int32},
                        Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                       ctx, ns, slots, length(fdescs), fdescs, length(dists), dists, length(dparams), dparams, spc,
                       hpts, hbw, hn, hfl, op, npts, pts, bw, ipc)
      check(ctx, st, "propagate_once")
    end
  end
  pts = pts[:, 1:npts[]]; bw = bw[1:d]; ipc = ipc[1:d]
  M = getManifold(getVariableType(destvar))
  mkd = AMP.manikde!(M, _unpackpoints(getVariableType(destvar), pts); bw)   # bw given => no re-selection (FGOSUtils.jl:118-128)
  return mkd, ipc
end

"""
    parallelEliminationOrder(dfg; method = :is, slack = 1) -> Vector{Symbol}

Variable elimination order from the library, for `solveTree!(dfg; eliminationOrder = IIFB200.parallelEliminationOrder(dfg))`
(SolverAPI.jl:338) or `buildTreeReset!`: the Bayes tree of a pose chain gets depth O(log n) instead of n, so the cliques
of a level solve side by side.  `:is` = rounds of independent low-degree variables (generalised odd-even reduction,
iifb200_elimination_order_is), `:nd` = level-set bisection (iifb200_elimination_order_nd).
"""
function parallelEliminationOrder(dfg::AbstractDFG; method::Symbol = :is, slack::Integer = 1)
  vlabels = listVariables(dfg); sort!(vlabels; by = l -> getVariable(dfg, l).nstime)   # graph (insertion) order
  vidx = Dict(l => Int32(i - 1) for (i, l) in enumerate(vlabels))
  flists = [Int32[vidx[v] for v in getVariableOrder(getFactor(dfg, f))] for f in listFactors(dfg)]
  off, flat = _csr(flists)
  isempty(flat) && push!(flat, Int32(0))
  order = zeros(Int32, max(length(vlabels), 1))
  st = method == :nd ?
    ccall((:iifb200_elimination_order_nd, LIB), Int32, (Int32, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}),
          length(vlabels), length(flists), off, flat, order) :
    ccall((:iifb200_elimination_order_is, LIB), Int32, (Int32, Int32, Ptr{Int32}, Ptr{Int32}, Int32, Ptr{Int32}),
          length(vlabels), length(flists), off, flat, slack, order)
  st == 0 || error("iifb200_elimination_order failed ($st): " * unsafe_string(ccall((:iifb200_plan_error, LIB), Cstring, ())))
  return Symbol[vlabels[i + 1] for i in order[1:length(vlabels)]]
end

# ---- boundary B4: the whole tree pass ---------------------------------------------------------------
_csr(lists) = (Int32[0; cumsum(length.(lists))], Int32[x for l in lists for x in l])

"""
    solveTree_b200!(dfg, tree; solveKey, lanes, downsolve)

Up + down pass of an already built Bayes tree (`buildTreeReset!`) on the device: `tree.bt`'s clique table goes through
iifb200_plan_tree / iifb200_plan_upload, the graph's beliefs through iifb200_upload_slots, one iifb200_schedule_run
replays the pass as a CUDA graph and iifb200_download_slots returns every posterior, which is written back with
setValKDE! (CSM step 5, updateFromSubgraph).  SolverParams.useMsgLikelihoods is honoured: the library builds the joint
up messages (differentials + one MsgPrior per class) itself.
"""
solveTree_b200!(dfg::AbstractDFG; kw...) =                      # order from the library, tree from the reference
  solveTree_b200!(dfg, IIF.buildTreeReset!(dfg, parallelEliminationOrder(dfg)); kw...)
function solveTree_b200!(dfg::AbstractDFG, tree; solveKey::Symbol = :default, lanes::Integer = 4,
                         downsolve::Bool = getSolverParams(dfg).downsolve)
  sp = getSolverParams(dfg)
  ctx = _ctx()
  vlbls = listVariables(dfg); sort!(vlbls; by = l -> getVariable(dfg, l).nstime)       # graph (insertion) order
  flbls = listFactors(dfg);   sort!(flbls; by = l -> getFactor(dfg, l).nstime)
  vidx = Dict(l => Int32(i - 1) for (i, l) in enumerate(vlbls))
  fidx = Dict(l => Int32(i - 1) for (i, l) in enumerate(flbls))
  vars = [getVariable(dfg, l) for l in vlbls]
  fcts = [getFactor(dfg, l) for l in flbls]
  N = sp.N
  gslots = SlotDesc[SlotDesc(getDimension(v), circmask(v), max(N, length(getVal(v; solveKey)), 1), 0) for v in vars]
  dists, dparams, fdescs, extra = _lower_factors(dfg, fcts, vidx, length(vars))
  isempty(extra) || error("IIFB200: graph factors with belief-valued measurements are not supported in a tree plan")
  # cliques numbered parents first: breadth-first from the roots over `tree.bt`
  cliqs = collect(values(IIF.getCliques(tree)))
  order = IIF.TreeClique[]; queue = [c for c in cliqs if isempty(IIF.getParent(tree, c))]
  while !isempty(queue)
    c = popfirst!(queue); push!(order, c); append!(queue, IIF.getChildren(tree, c))
  end
  cid = Dict(c.id => Int32(i - 1) for (i, c) in enumerate(order))
  parent = Int32[isempty(IIF.getParent(tree, c)) ? Int32(-1) : cid[IIF.getParent(tree, c)[1].id] for c in order]
  lists(get, idx) = _csr([[idx[x] for x in get(IIF.getCliqueData(c))] for c in order])
  fo, fr = _csr([[vidx[x] for x in IIF.getCliqFrontalVarIds(c)] for c in order])
  so, se = _csr([[vidx[x] for x in IIF.getCliqSeparatorVarIds(c)] for c in order])
  po, pt = lists(d -> d.potentials, fidx)
  d1o, d1 = lists(d -> d.directFrtlMsgIDs, vidx)
  d2o, d2 = lists(d -> d.msgskipIDs, vidx)
  d3o, d3 = lists(d -> d.itervarIDs, vidx)
  d4o, d4 = lists(d -> d.directPriorMsgIDs, vidx)
  # type tables for the joint up messages of useMsgLikelihoods = true (ids are positions in `tnames`)
  tnames = String[]
  function _tid(n::String)
    i = findfirst(==(n), tnames)
    i === nothing || return Int32(i - 1)
    push!(tnames, n)
    return Int32(length(tnames) - 1)
  end
  ftype = Int32[_tid(string(nameof(typeof(getFactorType(f))))) for f in fcts]
  vtype = Int32[_tid("var:" * string(typeof(getVariableType(v)))) for v in vars]
  msgprior_t = _tid("MsgPrior")
  vrelkind = zeros(Int32, length(vars)); vreltype = fill(Int32(-1), length(vars))
  for (i, v) in enumerate(vars)        # selectFactorType(T, T), DefaultNodeTypes.jl:12-31
    T = typeof(getVariableType(v))
    sft = try IIF.selectFactorType(T, T) catch; nothing end
    sft === nothing && continue
    vrelkind[i] = sft <: LinearRelative ? Int32(2) : (sft <: CircularCircular ? Int32(4) : Int32(0))
    vrelkind[i] == 0 && continue       # no device kind for this default relative: the pair gets no differential
    vreltype[i] = _tid(string(nameof(sft)))
  end
  plan = Ref{Ptr{Cvoid}}(C_NULL); sid = Ref{Int32}(-1)
  nd = sum(s.cap * s.dim for s in gslots)
  hpts = zeros(nd); hbw = zeros(MAX_DIM, length(vars)); hipc = zeros(MAX_DIM, length(vars))
  hn = zeros(Int32, length(vars)); hfl = ones(Int32, length(vars))
  off = 0
  for (i, v) in enumerate(vars)
    p = _packpoints(getVariableType(v), getVal(v; solveKey))
    hpts[(off + 1):(off + length(p))] .= vec(p); off += gslots[i].cap * gslots[i].dim
    hbw[1:size(p, 1), i] .= _bw(v, solveKey); hn[i] = size(p, 2); hfl[i] = isInitialized(v, solveKey)
  end
  GC.@preserve gslots fdescs dists dparams parent fo fr so se po pt d1o d1 d2o d2 d3o d3 d4o d4 ftype vtype vrelkind vreltype begin
    gd = Ref(GraphDesc(length(vars), pointer(gslots), length(fdescs), pointer(fdescs), length(dists), pointer(dists),
                       length(dparams), pointer(dparams), pointer(ftype), pointer(vtype), pointer(vrelkind), pointer(vreltype),
                       msgprior_t, 0))
    td = Ref(TreeDesc(length(order), pointer(parent), pointer(fo), pointer(fr), pointer(so), pointer(se), pointer(po), pointer(pt),
                      pointer(d1o), pointer(d1), pointer(d2o), pointer(d2), pointer(d3o), pointer(d3), pointer(d4o), pointer(d4)))
    opts = Ref(PlanOpts(N, sp.gibbsIters, 3, downsolve, lanes, 1, sp.useMsgLikelihoods, _nextcall(16 * 64 * length(vars)), sp.inflation))
    st = ccall((:iifb200_plan_tree, LIB), Int32, (Ref{GraphDesc}, Ref{TreeDesc}, Ref{PlanOpts}, Ref{Ptr{Cvoid}}), gd, td, opts, plan)
    st == 0 || error("iifb200_plan_tree failed ($st): " * unsafe_string(ccall((:iifb200_plan_error, LIB), Cstring, ())))
  end
  lock(_CTX_LOCK) do
    check(ctx, ccall((:iifb200_plan_upload, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{SolverParamsC}, Ptr{Cvoid}, Ref{Int32}),
          ctx, plan[], Ref(_solverparams(sp)), C_NULL, sid), "plan_upload")
    check(ctx, ccall((:iifb200_upload_slots, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}),
          ctx, 0, length(vars), hpts, hbw, hn, hfl), "upload_slots")
    check(ctx, @threadcall((:iifb200_schedule_run, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Int32), ctx, sid[], 0, -1), "schedule_run")
    check(ctx, ccall((:iifb200_download_slots, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
          ctx, 0, length(vars), hpts, hbw, hipc, hn), "download_slots")
    check(ctx, @threadcall((:iifb200_sync, LIB), Int32, (Ptr{Cvoid},), ctx), "sync")
    check(ctx, ccall((:iifb200_schedule_free, LIB), Int32, (Ptr{Cvoid}, Int32), ctx, sid[]), "schedule_free")
  end
  ccall((:iifb200_plan_free, LIB), Cvoid, (Ptr{Cvoid},), plan[])
  off = 0
  for (i, v) in enumerate(vars)
    d = Int(gslots[i].dim); n = Int(hn[i])
    pts = reshape(hpts[(off + 1):(off + d * n)], d, n); off += gslots[i].cap * d
    M = getManifold(getVariableType(v))
    mkd = AMP.manikde!(M, _unpackpoints(getVariableType(v), pts); bw = hbw[1:d, i])
    setValKDE!(v, mkd, true, hipc[1:d, i]; solveKey)                      # FactorGraph.jl:237-286
    setPPE!(v, getVariableType(v), solveKey)                               # FGOSUtils.jl:546-570
  end
  return dfg
end

"""
    calcPPE_b200(dfg, labels; solveKey)   — calcPPE, FGOSUtils.jl:237-278 (setPPE! at CSM step 5 funnels through it)

Mean and KDE-max of several variables' beliefs from one `iifb200_ppe_batch` launch (beliefs are uploaded first).
"""
function calcPPE_b200(dfg::AbstractDFG, labels::Vector{Symbol}; solveKey::Symbol = :default)
  ctx = _ctx(); vars = [getVariable(dfg, l) for l in labels]; V = length(vars)
  slots = SlotDesc[SlotDesc(getDimension(v), circmask(v), max(length(getVal(v; solveKey)), 1), 0) for v in vars]
  mean = zeros(MAX_DIM, V); mx = zeros(MAX_DIM, V)
  lock(_CTX_LOCK) do
    _set_graph(ctx, slots, FactorDesc[], DistDesc[], Float64[], _solverparams(getSolverParams(dfg)))
    staged = _upload_all(ctx, slots, [_packpoints(getVariableType(v), getVal(v; solveKey)) for v in vars],
                         [_bw(v, solveKey) for v in vars], fill(true, V))
    check(ctx, ccall((:iifb200_ppe_batch, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}),
          ctx, V, Int32.(0:(V - 1)), mean, mx), "ppe_batch")
  end
  return [MeanMaxPPE(solveKey, mean[1:getDimension(v), i], mx[1:getDimension(v), i], mean[1:getDimension(v), i])
          for (i, v) in enumerate(vars)]
end

"""
    approxDeconv_b200(dfg, fctsym; solveKey)   — approxDeconv, DeconvUtils.jl:176-202
"""
function approxDeconv_b200(dfg::AbstractDFG, fctsym::Symbol; solveKey::Symbol = :default)
  ctx = _ctx(); f = getFactor(dfg, fctsym)
  vo = getVariableOrder(f); vars = [getVariable(dfg, l) for l in vo]
  slotof = Dict(l => Int32(i - 1) for (i, l) in enumerate(vo))
  dists, dparams, fdescs, extra = _lower_factors(dfg, [f], slotof, length(vars))
  isempty(extra) || error("IIFB200: approxDeconv of a belief-valued prior is not supported")
  N = length(getVal(vars[1]; solveKey)); zdim = Int(fdescs[1].zdim)
  slots = SlotDesc[SlotDesc(getDimension(v), circmask(v), max(N, length(getVal(v; solveKey)), 1), 0) for v in vars]
  pred = zeros(zdim, N); meas = zeros(zdim, N)
  lock(_CTX_LOCK) do
    _set_graph(ctx, slots, fdescs, dists, dparams, _solverparams(getSolverParams(dfg)))
    staged = _upload_all(ctx, slots, [_packpoints(getVariableType(v), getVal(v; solveKey)) for v in vars],
                         [_bw(v, solveKey) for v in vars], fill(true, length(vars)))
    check(ctx, ccall((:iifb200_deconv_batch, LIB), Int32,
          (Ptr{Cvoid}, Int32, Ref{Int32}, Ref{Int32}, Ref{Int32}, Ptr{Float64}, Ptr{Float64}),
          ctx, 1, Int32(0), Int32(N), _nextcall(), pred, meas), "deconv_batch")
  end
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
    schedule(ctx, wave_off, ops, props, deconvs) -> id     — explicit wave lists (plans lowered by the caller, e.g.
useMsgLikelihoods = true with IIF_S_DECONV ops, addLikelihoodsDifferentialCHILD! TreeMessageUtils.jl:279-335)
"""
function schedule(ctx, wave_off::Vector{Int32}, ops::Vector{SchedOp}, props::Vector{PropOp}, deconvs::Vector{DeconvOp} = DeconvOp[])
  id = Ref{Int32}(-1)
  check(ctx, ccall((:iifb200_schedule_build_ex, LIB), Int32,
        (Ptr{Cvoid}, Int32, Ptr{Int32}, Int32, Ptr{SchedOp}, Int32, Ptr{PropOp}, Int32, Ptr{DeconvOp}, Ref{Int32}),
        ctx, length(wave_off) - 1, wave_off, length(ops), ops, length(props), props, length(deconvs), deconvs, id), "schedule_build_ex")
  return id[]
end
run!(ctx, id; first = 0, last = -1) = check(ctx, ccall((:iifb200_schedule_run, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Int32), ctx, id, first, last), "schedule_run")

end # module

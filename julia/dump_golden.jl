# dump_golden.jl — golden vectors from the reference itself, to PIN sample-level parity (SURVEY.md §8c item 5).
#
# NOT EXECUTED IN THE BUILD IMAGE (no Julia).  Run it wherever IncrementalInference.jl is installed, commit the output
# under tests/golden/julia/ and compare through the host-stream inputs of the C-ABI (`meas`, `mhidx`, `uinf` of
# iifb200_conv_batch; `randU`, `randN` of iifb200_product_batch): with the reference's own random draws passed in,
# hypothesis labels must match bit for bit and proposals within 1e-6 (1-D BFGS) / 1e-3 (Nelder-Mead).
#
# usage: julia julia/dump_golden.jl out_dir
# writes, per case, CSV files: <case>_src.csv (source particles), <case>_meas.csv, <case>_mhidx.csv,
# <case>_proposal.csv, <case>_bw.csv  (one row per particle / coordinate)
using IncrementalInference, Random, DelimitedFiles
const IIF = IncrementalInference
out = length(ARGS) > 0 ? ARGS[1] : "golden_julia"
mkpath(out)
Random.seed!(42)
N = 100

function dump_case(name, fg, fct::Symbol, target::Symbol)
  fc = getFactor(fg, fct)
  ccw = IIF._getCCW(fc)
  # the reference's own draws: fresh measurements and the hypothesis recipe for this convolution
  IIF.sampleFactor!(ccw, N)
  m0 = deepcopy(ccw.measurement)
  meas = [collect(Float64, m) for m in m0]
  # pass the drawn measurements explicitly so that the convolution uses exactly the dumped ones
  # (approxConvBelief(dfg, fc, target, measurement; N), ApproxConv.jl:4-12)
  bel = approxConvBelief(fg, fct, target, m0; N = N)
  pts = getPoints(bel, false)
  writedlm(joinpath(out, "$(name)_meas.csv"), reduce(hcat, meas)', ',')
  writedlm(joinpath(out, "$(name)_proposal.csv"), reduce(hcat, [collect(Float64, p) for p in pts])', ',')
  writedlm(joinpath(out, "$(name)_bw.csv"), getBW(bel)[:, 1], ',')
  for v in getVariableOrder(fc)
    vals = getVal(fg, v)
    writedlm(joinpath(out, "$(name)_src_$(v).csv"), reduce(hcat, [collect(Float64, p) for p in vals])', ',')
  end
end

# case 1: scalar prior + relative (tests/parity_cases.py "scalar_prior_and_relative")
fg = initfg(); getSolverParams(fg).N = N
addVariable!(fg, :x0, ContinuousScalar); addVariable!(fg, :x1, ContinuousScalar)
addFactor!(fg, [:x0], Prior(Normal(0.0, 1.0)))
addFactor!(fg, [:x0, :x1], LinearRelative(Normal(1.0, 0.1)))
initAll!(fg)
dump_case("scalar_prior", fg, :x0f1, :x0)
dump_case("scalar_relative_fwd", fg, :x0x1f1, :x1)
dump_case("scalar_relative_bwd", fg, :x0x1f1, :x0)

# case 2: product of two proposals with the Gibbs streams AMP consumed (manifoldProduct is called by propagateBelief)
mkd, ipc = propagateBelief(fg, :x0, :)
writedlm(joinpath(out, "product_x0_posterior.csv"), reduce(hcat, [collect(Float64, p) for p in getPoints(mkd, false)])', ',')
writedlm(joinpath(out, "product_x0_bw.csv"), getBW(mkd)[:, 1], ',')
# case 3: separated modes (test/testMultiHypo3Door.jl:40-124) - does the one-sweep multiscale Gibbs product lock onto
# node pairs that only overlap at a coarse level?  Dump both proposals of x1 and the posterior of their product.
fg3 = initfg(); getSolverParams(fg3).N = 200; getSolverParams(fg3).graphinit = false
for (k, l) in enumerate((0.0, 10.0, 20.0, 40.0))
  addVariable!(fg3, Symbol("l$(k-1)"), ContinuousScalar)
  addFactor!(fg3, [Symbol("l$(k-1)")], Prior(Normal(l, 0.01)))
end
addVariable!(fg3, :x0, ContinuousScalar); addVariable!(fg3, :x1, ContinuousScalar)
addFactor!(fg3, [:x0; :l0; :l1; :l2; :l3], LinearRelative(Normal(0, 0.25)), multihypo = [1.0; 0.25; 0.25; 0.25; 0.25])
addFactor!(fg3, [:x0; :x1], LinearRelative(Normal(10.0, 0.1)))
addFactor!(fg3, [:x1; :l0; :l1; :l2; :l3], LinearRelative(Normal(0, 0.25)), multihypo = [1.0; 0.25; 0.25; 0.25; 0.25])
initAll!(fg3)
dump_case("door_odometry_to_x1", fg3, :x0x1f1, :x1)
dump_case("door_multihypo_to_x1", fg3, :x1l0l1l2l3f1, :x1)
mkd3, = propagateBelief(fg3, :x1, :)
writedlm(joinpath(out, "door_x1_posterior.csv"), reduce(hcat, [collect(Float64, p) for p in getPoints(mkd3, false)])', ',')

# case 4: a whole tree solve (VERDICT r1, next-round 1d): 10-pose scalar chain, nested-dissection elimination order
# (odd-even reduction, as workloads.chain_nd_order), default SolverParams (useMsgLikelihoods = false), 20 seeds.
# Per seed and pose: posterior mean and standard deviation after solveTree!.  tests/test_golden.py compares the
# oracle's / the device's statistics over 20 seeds with these (band test: the reference's RNG is not reproduced).
function chain_nd_order(n)
  remaining = collect(0:(n - 1)); order = Int[]
  while length(remaining) > 2
    elim = isodd(length(remaining)) ? remaining[2:2:end] : remaining[2:2:(end - 1)]
    isempty(elim) && break
    append!(order, elim)
    remaining = [k for k in remaining if !(k in elim)]
  end
  append!(order, remaining)
  return [Symbol("x$k") for k in order]
end
let n = 10, rows = Vector{Vector{Float64}}()
  for seed in 1:20
    Random.seed!(seed)
    fgc = initfg(); getSolverParams(fgc).N = N; getSolverParams(fgc).graphinit = false
    for k in 0:(n - 1)
      addVariable!(fgc, Symbol("x$k"), ContinuousScalar)
    end
    addFactor!(fgc, [:x0], Prior(Normal(0.0, 0.1)))
    for k in 0:(n - 2)
      addFactor!(fgc, [Symbol("x$k"), Symbol("x$(k+1)")], LinearRelative(Normal(1.0, 0.1)))
    end
    for k in 0:(n - 1)    # the initial beliefs workloads.scalar_chain generates: x_k ~ N(k, 0.1 sqrt(k + 1))
      initVariable!(fgc, Symbol("x$k"), [[k + 0.1 * sqrt(k + 1.0) * randn()] for _ in 1:N])
    end
    solveTree!(fgc; eliminationOrder = chain_nd_order(n))
    for k in 0:(n - 1)
      p = [x[1] for x in getPoints(getBelief(fgc, Symbol("x$k")), false)]
      push!(rows, [seed, k, sum(p) / length(p), sqrt(sum(abs2, p .- sum(p) / length(p)) / (length(p) - 1))])
    end
  end
  writedlm(joinpath(out, "tree_chain10_nd.csv"), reduce(hcat, rows)', ',')     # seed, pose, mean, std
end
println("wrote golden vectors to ", out)

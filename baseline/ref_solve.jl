# ref_solve.jl — times the UNMODIFIED reference (IncrementalInference.jl, stock Julia CPU path) on the benchmark
# workload, for the `--impl reference` arm of bench.py (SURVEY.md §8d "CPU reference timing beside it").
#
# NOT EXECUTED IN THE BUILD IMAGE: no Julia toolchain exists there (nor on the GPU boxes, same image), so bench.py
# falls back to the CPU oracle port (kind "port").  bench.py runs this script only when `julia` is on PATH and
# `using IncrementalInference` succeeds; it then reports kind "reference".
#
# usage: JULIA_NUM_THREADS=$(nproc) julia baseline/ref_solve.jl <poses> <N> <solves> <multithread=true|false> <convolutions per solve>
# prints one JSON line: {"conv_per_s": ..., "convolutions": ..., "seconds": ..., "threads": ...}
using IncrementalInference, Random
const IIF = IncrementalInference

poses  = length(ARGS) > 0 ? parse(Int, ARGS[1]) : 100
N      = length(ARGS) > 1 ? parse(Int, ARGS[2]) : 100
solves = length(ARGS) > 2 ? parse(Int, ARGS[3]) : 3
mt     = length(ARGS) > 3 ? parse(Bool, ARGS[4]) : true

# convolutions per solve: supplied by bench.py from the mirrored plan (same tree and Gibbs-set rules, tree.compile_solve),
# so that both arms divide the same unit count by their own wall time
nconv  = length(ARGS) > 4 ? parse(Int, ARGS[5]) : 0

function build(poses, N)
  # generateGraph_LineStep(poses-1; poseEvery=1, landmarkEvery=0, posePriorsAt=[0]) with the benchmark's noise
  # (CanonicalGraphExamples.jl:154-240): x0..x{poses-1}, Prior(Normal(0,0.1)), LinearRelative(Normal(1,0.1))
  fg = initfg()
  getSolverParams(fg).N = N
  getSolverParams(fg).multiproc = false
  addVariable!(fg, :x0, ContinuousScalar)
  addFactor!(fg, [:x0], Prior(Normal(0.0, 0.1)))
  for k in 1:(poses - 1)
    addVariable!(fg, Symbol("x$k"), ContinuousScalar)
    addFactor!(fg, [Symbol("x$(k-1)"), Symbol("x$k")], LinearRelative(Normal(1.0, 0.1)))
  end
  initAll!(fg)
  return fg
end

# odd-even (nested dissection) elimination order of a chain, same as workloads.chain_nd_order
function chain_nd_order(n)
  remaining = collect(0:(n - 1)); order = Int[]
  while length(remaining) > 2
    elim = isodd(length(remaining)) ? remaining[2:2:end] : remaining[2:2:(end - 1)]
    isempty(elim) && break
    append!(order, elim)
    remaining = setdiff(remaining, elim)
  end
  append!(order, remaining)
  return [Symbol("x$k") for k in order]
end

Random.seed!(42)
fg = build(poses, N)
order = chain_nd_order(poses)
solveTree!(deepcopy(fg); eliminationOrder = order, multithread = mt)         # warm-up (JIT)
secs = 0.0
for s in 1:solves
  g = deepcopy(fg)
  Random.seed!(42 + s)
  global secs += @elapsed solveTree!(g; eliminationOrder = order, multithread = mt)
end
println("{\"conv_per_s\": $(nconv * solves / secs), \"convolutions\": $nconv, \"seconds\": $secs, " *
        "\"threads\": $(Threads.nthreads()), \"solves\": $solves, \"poses\": $poses, \"N\": $N}")

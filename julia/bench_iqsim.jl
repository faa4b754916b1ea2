# bench_iqsim.jl -- times the UNMODIFIED reference (ImageQuilting.jl) on a BASELINE config exported by
# scripts/export_config.py, for a machine that has Julia (this repository's build image does not: UNTESTED here).
#
#   python scripts/export_config.py --config 5 --out /tmp/iq_cfg5
#   julia -t auto julia/bench_iqsim.jl /tmp/iq_cfg5 [nreal]
#
# Prints voxels/s of iqsim (src/iqsim.jl:50-63) the way bench.py defines the metric: nreal * prod(simsize) / wall time.
# The reference's own parallelism is FFTW threads = physical cores (src/iqsim.jl:66).
using ImageQuilting, JSON, Random

dir = ARGS[1]
meta = JSON.parsefile(joinpath(dir, "meta.json"))
sz = Tuple(Int.(meta["size"]))
TI = reshape(reinterpret(Float32, read(joinpath(dir, "ti.f32"))), sz) |> collect
tilesize = Tuple(Int.(meta["tilesize"]))
nreal = length(ARGS) >= 2 ? parse(Int, ARGS[2]) : 1
overlap = Tuple(Float64.(meta["overlap"]))

kwargs = Dict{Symbol,Any}(:overlap => overlap, :nreal => nreal, :showprogress => false, :rng => MersenneTwister(2024))
if meta["soft"]
  asz = Tuple(Int.(meta["aux_size"]))
  AUX = reshape(reinterpret(Float32, read(joinpath(dir, "aux.f32"))), asz) |> collect
  AUXTI = reshape(reinterpret(Float32, read(joinpath(dir, "auxti.f32"))), sz) |> collect
  kwargs[:soft] = [(AUX, AUXTI)]
end
if !isempty(meta["hard"])
  kwargs[:hard] = Dict(CartesianIndex(Int.(h[1:end-1])...) => Float64(h[end]) for h in meta["hard"])
end

iqsim(TI, tilesize; kwargs..., nreal=1)            # compile + warm up
t = @elapsed reals = iqsim(TI, tilesize; kwargs...)
vox = nreal * prod(sz)
println("config ", meta["config"], ": ", nreal, " realization(s), ", round(t, digits=2), " s, ",
        round(vox / t, digits=1), " voxels/s on ", Threads.nthreads(), " Julia threads / ",
        Sys.CPU_THREADS, " CPU threads")

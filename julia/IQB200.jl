# IQB200.jl -- thin Julia binding of libiqb200.so (C ABI: include/iqb200.h).
#
# UNTESTED: Julia is not available in the environment this repository was built in.  The file is the
# binding a maintainer of ImageQuilting.jl would add; INTEGRATION.md shows where it plugs into
# src/iqsim.jl.  Struct layouts mirror include/iqb200.h field by field.
module IQB200

const lib = get(ENV, "IQB200_LIB", joinpath(@__DIR__, "..", "imagequilting.jl_b200", "libiqb200.so"))

struct CtxDesc
  ndim::Int32
  ti_size::NTuple{3,Int64}
  tile_size::NTuple{3,Int64}
  ti::Ptr{Float32}
  disabled::Ptr{UInt8}
  nsoft::Int32
  auxti::Ptr{Ptr{Float32}}
  device::Int32
  max_batch::Int32
end

struct Tile
  simdev::Ptr{Float32}
  hard_nnz::Int32
  hard_offset::Ptr{Int32}
  hard_value::Ptr{Float32}
  softdev::Ptr{Ptr{Float32}}
end

struct Result
  count::Int64
  idx::Ptr{Int64}
  prob::Ptr{Float64}
  picked::Int64
  relax_iters::Int32
  dmin::Float32
end

lasterror() = unsafe_string(ccall((:iq_last_error, lib), Cstring, ()))
check(rc) = rc == 0 || error("libiqb200: ", lasterror())

pad3(t::Dims{N}) where {N} = ntuple(i -> i <= N ? Int64(t[i]) : Int64(1), 3)

mutable struct Context
  handle::Ptr{Cvoid}
  keep::Vector{Any}   # arrays the library reads during create
end

"""
    Context(TI, tilesize; disabled=nothing, auxTIs=[], device=0, max_batch=1)

Uploads the (NaN-free) training image and auxiliary images once; replaces `array_kernel`
(src/utils.jl:57-61) and the `geoconfig` tuple (src/iqsim.jl:118-127).
"""
function Context(TI::AbstractArray{<:Real,N}, tilesize::Dims{N}; disabled=nothing, auxTIs=[], device=0, max_batch=1) where {N}
  ti = Array{Float32}(TI)
  aux = [Array{Float32}(a) for a in auxTIs]
  auxptr = [pointer(a) for a in aux]
  dis = disabled === nothing ? UInt8[] : Array{UInt8}(vec(disabled))
  out = Ref{Ptr{Cvoid}}(C_NULL)
  GC.@preserve ti aux auxptr dis begin
    desc = CtxDesc(N, pad3(size(TI)), pad3(tilesize), pointer(ti),
                   isempty(dis) ? Ptr{UInt8}(C_NULL) : pointer(dis), length(aux),
                   isempty(aux) ? Ptr{Ptr{Float32}}(C_NULL) : pointer(auxptr), device, max_batch)
    check(ccall((:iq_ctx_create, lib), Int32, (Ref{Ptr{Cvoid}}, Ref{CtxDesc}), out, desc))
  end
  ctx = Context(out[], Any[])
  finalizer(c -> ccall((:iq_ctx_destroy, lib), Int32, (Ptr{Cvoid},), c.handle), ctx)
  ctx
end

"""
    search(ctx, ovlmask, simdevs; hard=nothing, softdevs=nothing, tol=0.1)

One call per path step for a batch of tiles sharing `ovlmask`.  Returns, per tile, `(patterndb, probs)`
with 1-based linear indices into the distance map -- exactly what src/iqsim.jl:237-240 produces, ready
for `sample(rng, patterndb, weights(probs))` (src/iqsim.jl:243), which stays in Julia.
`hard` is `(offsets::Vector{Int32} (0-based, column-major inside the tile), values::Vector{Float32})` or
`nothing`; `softdevs[i]` is the vector of tile-sized soft events of tile i.
"""
function search(ctx::Context, ovlmask::AbstractArray{Bool}, simdevs::Vector; hard=nothing, softdevs=nothing, tol=0.1)
  n = length(simdevs)
  mask = Array{UInt8}(vec(ovlmask))
  devs = [Array{Float32}(vec(s)) for s in simdevs]
  softs = softdevs === nothing ? nothing : [[Array{Float32}(vec(a)) for a in sd] for sd in softdevs]
  softptrs = softs === nothing ? nothing : [[pointer(a) for a in sd] for sd in softs]
  hoff, hval = hard === nothing ? (Int32[], Float32[]) : (Vector{Int32}(hard[1]), Vector{Float32}(hard[2]))
  tiles = Vector{Tile}(undef, n)
  results = Vector{Result}(undef, n)
  out = Vector{Tuple{Vector{Int},Vector{Float64}}}(undef, n)
  GC.@preserve mask devs softs softptrs hoff hval tiles results begin
    for i in 1:n
      tiles[i] = Tile(pointer(devs[i]), length(hoff), isempty(hoff) ? Ptr{Int32}(C_NULL) : pointer(hoff),
                      isempty(hval) ? Ptr{Float32}(C_NULL) : pointer(hval),
                      softptrs === nothing ? Ptr{Ptr{Float32}}(C_NULL) : pointer(softptrs[i]))
    end
    check(ccall((:iq_search, lib), Int32, (Ptr{Cvoid}, Ptr{UInt8}, Ptr{Tile}, Int32, Float64, Ptr{Result}),
                ctx.handle, mask, tiles, n, tol, results))
    for i in 1:n
      r = results[i]
      idx = unsafe_wrap(Array, r.idx, r.count) .+ 1          # copy: buffers belong to the context
      out[i] = (Vector{Int}(idx), copy(unsafe_wrap(Array, r.prob, r.count)))
    end
  end
  out
end

# ---- device-resident simulation (iq_sim_*): grids, sampling, boundary cuts and paste on the device -----------------
struct SimDesc          # == iq_sim_desc
  pad_size::NTuple{3,Int64}
  ovl_size::NTuple{3,Int64}
  nreal::Int32
  ti64::Ptr{Float64}
  u::Ptr{Float64}
  npath::Int64
  tol::Float64
  debug::Int32
  aux::Ptr{Ptr{Float32}}
  hard_has::Ptr{UInt8}
  hard_val::Ptr{Float32}
  exact_cut::Int32      # 1: boundary cuts in exact 128-bit integer arithmetic (integer-valued / categorical images)
end

struct SimSlab          # == iq_sim_slab (tile coordinates, 0-based)
  dim::Int32
  prev::Int32
  lo::NTuple{3,Int32}
  sz::NTuple{3,Int32}
end

"""
    simulate!(outs, ctx, TI, padsize, ovlsize, steps, U; tol=0.1, aux=[])

Runs a whole simulation on the device.  `steps` is the path as a vector of
`(start0::NTuple{3,Int}, ovlmask::Array{Bool}, slabs::Vector{SimSlab})` (what src/iqsim.jl:177-205 derives from
`simpath` and `pasted`), `U` the `nreal x length(steps)` uniforms drawn in the reference's order, `outs` the
preallocated `Float32` realizations (cropped to `size(outs[1])`).  Returns the status word (0 = ok; otherwise fall
back to `search` + host cut/paste).  Steps with an empty mask on a context with soft data must be done with
`search` + `sample` and handed over through `iq_sim_step_picked` (not wrapped here).
"""
function simulate!(outs::Vector{<:Array{Float32}}, ctx::Context, TI::AbstractArray{<:Real,N}, padsize::Dims{N},
                   ovlsize::Dims{N}, steps, U::Matrix{Float64}; tol=0.1, aux=[]) where {N}
  ti64 = Array{Float64}(TI)
  u = permutedims(U)                       # library layout: [nreal][npath], row-major
  auxs = [Array{Float32}(a) for a in aux]
  auxptr = [pointer(a) for a in auxs]
  status = Ref{Int32}(0)
  GC.@preserve ti64 u auxs auxptr outs begin
    desc = SimDesc(pad3(padsize), pad3(ovlsize), length(outs), pointer(ti64), pointer(u), length(steps), tol, 0,
                   isempty(auxs) ? Ptr{Ptr{Float32}}(C_NULL) : pointer(auxptr), Ptr{UInt8}(C_NULL), Ptr{Float32}(C_NULL),
                   all(isinteger, TI) ? 1 : 0)
    check(ccall((:iq_sim_begin, lib), Int32, (Ptr{Cvoid}, Ref{SimDesc}), ctx.handle, desc))
    for (k, (start0, ovlmask, slabs)) in enumerate(steps)
      st = Int64[start0...]; mask = Array{UInt8}(vec(ovlmask))
      check(ccall((:iq_sim_step, lib), Int32, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{UInt8}, Ptr{SimSlab}, Int32, Int32),
                  ctx.handle, k - 1, st, mask, slabs, length(slabs), 0))
    end
    check(ccall((:iq_sim_sync, lib), Int32, (Ptr{Cvoid}, Ptr{Int64}, Ref{Int32}), ctx.handle, C_NULL, status))
    if status[] == 0
      crop = Int64[pad3(size(outs[1]))...]
      ptrs = [Ptr{Cvoid}(pointer(o)) for o in outs]
      check(ccall((:iq_sim_fetch_all, lib), Int32, (Ptr{Cvoid}, Int32, Ptr{Int64}, Ptr{Ptr{Cvoid}}, Int32),
                  ctx.handle, 1, crop, ptrs, 0))
    end
    check(ccall((:iq_sim_end, lib), Int32, (Ptr{Cvoid},), ctx.handle))
  end
  status[]
end

"""
    dependency_levels(tilesize, ovlsize, ntiles, path0) -> (levels, nlevels)

Dependency level of every step of a simulation path (`path0`: 0-based column-major tile indices).  Steps of one level
touch disjoint windows of the simulation grid and may be handed to `iq_sim_step_multi` together (after registering their
overlap shapes with `iq_sim_define_shape`); running the levels in order reproduces the sequential loop of
src/iqsim.jl:172-285 bit for bit.  Raster paths: level(i, j, k) = i + 2j + 4k.
"""
function dependency_levels(tilesize::Dims{N}, ovlsize::Dims{N}, ntiles::Dims{N}, path0::Vector{Int64}) where {N}
  levels = Vector{Int32}(undef, length(path0)); nl = Ref{Int32}(0)
  check(ccall((:iqh_dependency_levels, lib), Int32,
              (Int32, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Int64, Ptr{Int32}, Ref{Int32}),
              N, Int64[tilesize...], Int64[ovlsize...], Int64[ntiles...], path0, length(path0), levels, nl))
  levels, nl[]
end

end # module

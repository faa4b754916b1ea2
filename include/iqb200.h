/*
 * iqb200.h -- C ABI of libiqb200.so, the B200-native (sm_100a) replacement for the per-tile
 * overlap-distance search of ImageQuilting.jl v1.3.1.
 *
 * What it replaces (citations relative to the reference checkout):
 *   - the backend shim  imfilter_kernel / array_kernel / view_kernel
 *         src/imfilter.jl:5-28, src/utils.jl:57-67
 *   - fastdistance       src/utils.jl:5-13
 *   - the distance assembly + candidate selection inside the tile loop of iqsim
 *         src/iqsim.jl:187-240  (overlap / hard / soft distances, disabled knock-out,
 *         threshold selection), relaxation src/relaxation.jl:5-48,
 *         taumodel src/taumodel.jl:5-45
 * What stays with the caller: RNG and sampling (src/iqsim.jl:243) unless the caller asks
 * for a fused pick and supplies the uniform itself; the boundary cut (src/graphcut.jl) and
 * the paste (src/iqsim.jl:251-278) -- unless the caller opts into the device-resident
 * simulation (iq_sim_*, below), which keeps the grids of all realizations on the device and
 * runs sampling (pre-drawn uniforms), cut and paste there as well.
 *
 * Conventions
 *   - All arrays are column-major (Julia layout, first index fastest), densely packed.
 *   - ndim is 2 or 3; sizes are passed as int64_t[ndim]; a 2-D problem is a 3-D one with nz = 1.
 *   - Linear indices returned are 0-BASED column-major indices into the distance map of size
 *     distsize = ti_size - tile_size + 1 (Julia callers add 1; src/iqsim.jl:244).
 *   - Every entry point returns an int32_t status: IQ_OK (0) or a negative IQ_ERR_* code; the
 *     message of the last failure on the calling thread is returned by iq_last_error().
 *   - No exceptions cross the ABI, no callbacks, no global state other than the thread-local
 *     error string.  No CPU fallback: iq_ctx_create fails with IQ_ERR_NO_DEVICE when there is
 *     no sm_100 device.
 *   - A context is bound to one device and one stream; calls on ONE context must be serialised
 *     by the caller; different contexts may be driven from different host threads (one per GPU).
 *   - Host buffers passed in belong to the caller and are only read during the call.  Buffers
 *     handed out in iq_result belong to the context and stay valid until the next iq_search* /
 *     iq_ctx_destroy on that context.
 */
#ifndef IQB200_H
#define IQB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IQ_ABI_VERSION 1

enum {
  IQ_OK = 0,
  IQ_ERR_INVALID = -1,    /* bad argument (NULL, size mismatch, out-of-range) */
  IQ_ERR_NO_DEVICE = -2,  /* no CUDA device of compute capability 10.x / bad device index */
  IQ_ERR_CUDA = -3,       /* a CUDA runtime call failed; see iq_last_error() */
  IQ_ERR_NOMEM = -4,      /* host or device allocation failed */
  IQ_ERR_STATE = -5       /* call sequence error (e.g. result read before a search) */
};

typedef struct iq_ctx iq_ctx;

/* Geometry + resident images.  Replaces `geoconfig` (src/iqsim.jl:118-127) and the uploads done
 * by imagepreproc/array_kernel (src/utils.jl:57-61,86-89). */
typedef struct iq_ctx_desc {
  int32_t ndim;               /* 2 or 3 */
  int64_t ti_size[3];         /* training image size (unused dims = 1) */
  int64_t tile_size[3];       /* tile size (unused dims = 1) */
  const float* ti;            /* training image, NaN/missing already replaced by 0 (src/utils.jl:96-102) */
  const uint8_t* disabled;    /* distsize bytes, nonzero = patch contains an inactive voxel
                                 (finddisabled, src/utils.jl:115-129); may be NULL */
  int32_t nsoft;              /* number of auxiliary variables (soft data), >= 0 */
  const float* const* auxti;  /* nsoft auxiliary training images, same size as ti (src/iqsim.jl:222-227) */
  int32_t device;             /* CUDA device ordinal */
  int32_t max_batch;          /* max tiles per iq_search call that share device buffers (>=1); larger
                                 batches are processed in chunks */
} iq_ctx_desc;

/* One tile of a batch.  All tiles of one iq_search call share the overlap mask (on any path the
 * mask depends only on the path step, not on the realization: src/iqsim.jl:188-205). */
typedef struct iq_tile {
  const float* simdev;          /* tile-sized current content of the simulation grid (src/iqsim.jl:185) */
  int32_t hard_nnz;             /* number of hard data inside this tile; 0 = none (src/iqsim.jl:210-219) */
  const int32_t* hard_offset;   /* hard_nnz 0-based column-major offsets inside the tile */
  const float* hard_value;      /* hard_nnz values */
  const float* const* softdev;  /* ctx.nsoft tile-sized views of the padded auxiliary grids (src/iqsim.jl:224) */
} iq_tile;

typedef struct iq_result {
  int64_t count;        /* number of candidates (length of patterndb, src/iqsim.jl:237) */
  const int64_t* idx;   /* ascending 0-based linear indices into the distance map */
  const double* prob;   /* taumodel output, unnormalised (src/taumodel.jl:44) */
  int64_t picked;       /* iq_search_pick only: the sampled linear index (src/iqsim.jl:243), else -1 */
  int32_t relax_iters;  /* number of relaxation rounds used (0 on the threshold path) */
  float dmin;           /* minimum of the primary distance over enabled positions */
} iq_result;

/* library / device info */
int32_t iq_abi_version(void);
int32_t iq_device_count(void);
const char* iq_last_error(void);

/* lifetime */
int32_t iq_ctx_create(iq_ctx** out, const iq_ctx_desc* desc);
int32_t iq_ctx_destroy(iq_ctx* ctx);
int32_t iq_ctx_npos(const iq_ctx* ctx, int64_t* npos, int64_t* nenabled);
/* Would iq_ctx_create(desc) build a context identical to `ctx`?  Same device, geometry, nsoft and max_batch, the same
 * disabled patches, and bitwise the same images: desc's images are uploaded and compared with the resident copies on
 * the device.  *same = 1 lets a caller that simulates again on the same training image keep the context -- its
 * summed-volume tables, cached A2 maps, image spectra and work buffers -- instead of rebuilding it.  0 while a
 * simulation is open on the context. */
int32_t iq_ctx_matches(iq_ctx* ctx, const iq_ctx_desc* desc, int32_t* same);
/* The same test against the images of `ref`, a context on the same device that the caller has just created from, or
 * matched against, these host arrays (device-to-device comparison: a host that keeps several contexts per call -- the
 * lockstep groups of iqh_run -- uploads the images once). */
int32_t iq_ctx_matches_ctx(iq_ctx* ctx, const iq_ctx_desc* desc, const iq_ctx* ref, int32_t* same);

/* The hot path: for each of `ntile` tiles sharing `ovlmask` (tile-sized, nonzero = voxel belongs
 * to the overlap with an already pasted neighbour) compute the overlap / hard / soft distance maps,
 * knock out disabled patches, select the candidate set (threshold rule when there is no auxiliary
 * distance, relaxation otherwise) and evaluate the tau model.  results[i] describes tile i.
 * Replaces src/iqsim.jl:187-240. */
int32_t iq_search(iq_ctx* ctx, const uint8_t* ovlmask, const iq_tile* tiles, int32_t ntile,
                  double tol, iq_result* results);

/* Same, plus the inverse-CDF walk of StatsBase.sample (src/iqsim.jl:243) with caller-supplied
 * uniforms u[i] in [0,1): results[i].picked is the chosen pattern.  idx/prob are still exposed. */
int32_t iq_search_pick(iq_ctx* ctx, const uint8_t* ovlmask, const iq_tile* tiles, int32_t ntile,
                       double tol, const double* u, iq_result* results);

/* Parity / debugging entry: one full distance map (distsize floats, +Inf on disabled patches).
 *   which = -1 : overlap distance   fastdistance(TI, simdev, weights=ovlmask)
 *   which = -2 : hard distance      fastdistance(TI, harddev, weights=hardmask)
 *   which >= 0 : soft distance for auxiliary variable `which`
 * Replaces one fastdistance call (src/utils.jl:5-13) + the knock-out (src/iqsim.jl:207,216,226). */
int32_t iq_distance(iq_ctx* ctx, int32_t which, const uint8_t* ovlmask, const iq_tile* tile,
                    float* out_map);

/* Fetch a tile-sized patch of the resident training image starting at linear index `pos` of the
 * distance map (view_kernel, src/utils.jl:63-67). */
int32_t iq_fetch_tile(iq_ctx* ctx, int64_t pos, float* out_tile);

/* Position-slice mode (second axis): for ONE large realization every rank holds a slab of the
 * training image (its context is created on the cropped image) and owns the patch positions of that slab.
 * A tile search becomes: iq_slice_distance on every rank (overlap distance + local minimum) -> the host
 * all-reduces the minimum -> iq_slice_select with the global minimum (threshold rule of src/iqsim.jl:237 on the
 * local positions) -> the host all-gathers the short candidate lists (local index + offset of the slab) and
 * evaluates the tau model (iq_taumodel) and the sampling walk (iq_sample) on the merged list.  No hard data.
 * Buffers handed out by iq_slice_candidates stay valid until the next iq_slice_* call. */
int32_t iq_slice_distance(iq_ctx* ctx, const uint8_t* ovlmask, const iq_tile* tiles, int32_t ntile, float* dmin_local);
int32_t iq_slice_select(iq_ctx* ctx, double tol, const float* dmin_global, int64_t* counts);
int32_t iq_slice_candidates(const iq_ctx* ctx, int32_t tile, const int64_t** idx, const float** val);
/* Relaxation path of position-slice mode (contexts with auxiliary images; src/relaxation.jl:5-48 with every k-th key
 * selected over ALL slabs, SURVEY 8(e) "all-reduce of radix histograms").  Sources: 0 = overlap distance, 1 + i = soft
 * distance i (iq_slice_distance computes them from tile->softdev).
 *   iq_slice_minmax  float bits of the local [min, max] of every source (all-zero test of relaxation.jl:11)
 *   iq_slice_hist    for up to 8 requests: local 256-bin histogram of digit `level` (0 = top byte) of the VALUE bits of
 *                    a source, restricted to the values whose `level` higher digits equal `prefix`; the host sums the
 *                    histograms of all ranks, picks the bin that holds the k-th value and descends (4 levels)
 *   iq_slice_kth     ties on the k-th value are broken by position (partialsortperm, relaxation.jl:12,27): the rank whose
 *                    slab holds the k-th entry selects its local k_local-th smallest (value, position) key exactly
 *   iq_slice_pick    local candidates = positions whose key (value bits << 32 | local position) is <= kth[s] in every
 *                    source s < nsrc; iq_slice_candidates then returns idx[count] and val[source][count] */
int32_t iq_slice_minmax(iq_ctx* ctx, int32_t tile, uint32_t* minbits, uint32_t* maxbits);
int32_t iq_slice_hist(iq_ctx* ctx, int32_t tile, int32_t nreq, const int32_t* src, const int32_t* level, const uint32_t* prefix,
                      int64_t* hist);
int32_t iq_slice_kth(iq_ctx* ctx, int32_t tile, int32_t src, int64_t k_local, uint64_t* key);
int32_t iq_slice_pick(iq_ctx* ctx, int32_t tile, int32_t nsrc, const uint64_t* kth, int64_t* count);
/* Host FP64 pieces exposed for hosts that merge candidate lists themselves: taumodel (src/taumodel.jl:5-45) on
 * vals[source][candidate] and the StatsBase.sample walk (src/iqsim.jl:243). */
int32_t iq_taumodel(int64_t n, int32_t nsrc, const float* vals, double* prob);
int32_t iq_sample(const double* prob, int64_t n, double u, int64_t* pos);

/* Optional device boundary cut (the reference keeps graphcut on the host, src/graphcut.jl:5-84; this entry
 * exists because a multi-GPU node has few host cores per GPU).  Each task is one overlap slab: A = content
 * already pasted, B = new patch, both column-major with size sz (unused dims = 1), cut along `dim`;
 * keep[i] = 1 where the pasted voxel is kept (exactly graphcut(A, B, dim), same answer as the host routine
 * iqh_graphcut on well-conditioned slabs).  Slabs too large for the shared-memory kernel are cut on the host.
 * iters (may be NULL) receives the number of push-relabel sweeps per task (0 = host path). */
typedef struct iq_cut_task {
  const double* A;
  const double* B;
  int32_t sz[3];
  int32_t dim;
  uint8_t* keep;
} iq_cut_task;
int32_t iq_cut_batch(iq_ctx* ctx, const iq_cut_task* tasks, int32_t ntask, int32_t* iters);

/* Device-resident simulation (simulation grids, tile paste and boundary cuts on the device).
 * The grids of `nreal` realizations live in device memory for the whole simulation; every path step is enqueued
 * on the context's stream WITHOUT any host synchronisation:
 *     template gather from the grids (iqsim.jl:185) -> overlap distance (utils.jl:5-13) -> threshold selection
 *     (iqsim.jl:237) -> tau model (taumodel.jl:5-45) -> StatsBase.sample walk with the pre-drawn uniform
 *     (iqsim.jl:243) -> boundary cuts (graphcut.jl:5-84, device kernel of iq_cut_batch) -> paste (iqsim.jl:278).
 * Scope: overlap slabs that fit the shared-memory cut kernel (iq_sim_begin returns IQ_ERR_STATE otherwise and the
 * caller uses iq_search_pick + its own paste instead).  With soft data, and on tiles with hard data, every step runs
 * up to three relaxation rounds (src/relaxation.jl:5-39: radix select of the dbsize / softk smallest keys per source,
 * intersection) on the device.  Steps whose overlap mask is empty on a context with soft data (their candidate set is a tenth of
 * all patterns, far above the device tau model's 32 768 entries) are done by the caller through iq_search +
 * iq_sample and handed over with iq_sim_step_picked.
 * A data-dependent condition the device path does not cover (a candidate set of more than 32 768 entries on a
 * non-empty mask, an empty first relaxation round, a cut hitting its iteration cap) is reported by iq_sim_sync
 * through `status` != 0; the grids are then invalid and the caller reruns the simulation through iq_search_pick.
 * While a simulation is open the context may be used for iq_search / iq_distance between steps (they wait for the
 * enqueued steps), nothing else. */
typedef struct iq_sim_desc {
  int64_t pad_size[3];   /* padded simulation grid (src/iqsim.jl:106), unused dims = 1 */
  int64_t ovl_size[3];   /* overlap size per dimension (src/iqsim.jl:92): fixes the largest cut slab */
  int32_t nreal;         /* realizations held by this context; max_batch of the context must be a multiple of it
                            (max_batch / nreal = tiles one iq_sim_step_multi launch may carry) */
  const double* ti64;    /* training image in FP64 (the values that are pasted and cut), ti_size doubles; NULL = the
                            context's FP32 image is exact (FP32 input) and is widened on the device */
  const double* u;       /* [nreal][npath] uniforms in the reference's draw order (src/iqsim.jl:243) */
  int64_t npath;         /* number of path steps */
  double tol;
  int32_t debug;         /* nonzero: also keep the boundary-cut grids (src/iqsim.jl:281) */
  const float* const* aux; /* contexts with soft data: ctx.nsoft padded auxiliary grids (pad_size floats each,
                              symmetric-padded, NaN -> 0; src/utils.jl:74-89), else NULL */
  const uint8_t* hard_has; /* hard data (src/utils.jl:18-36): pad_size bytes, nonzero = voxel carries a non-NaN datum;
                              NULL = no hard data */
  const float* hard_val;   /* pad_size floats: the datum where hard_has */
  int32_t exact_cut;       /* nonzero: boundary cuts in exact integer arithmetic (the FP64 capacities of graphcut.jl:52 as
                              128-bit integers: no rounding in the flow, one well-defined cut).  For integer-valued
                              (categorical) images, whose degenerate capacities -- (Du+Dv)/eps next to O(1) terms -- make an
                              FP64 max-flow return whichever equal-cost cut its rounding favours.  Needs twice the shared
                              memory per slab voxel; a slab whose capacities span more than ~2^56 raises status bit 2. */
} iq_sim_desc;
typedef struct iq_sim_slab {   /* one overlap slab of the current tile, in tile coordinates */
  int32_t dim;           /* dimension of the overlap */
  int32_t prev;          /* 1 = overlap with the previous tile along dim, 0 = with the next one */
  int32_t lo[3], sz[3];
} iq_sim_slab;
int32_t iq_sim_begin(iq_ctx* ctx, const iq_sim_desc* desc);
/* Enqueues one path step for all realizations: tile origin `start` (0-based voxel coordinates in the padded grid),
 * the overlap mask of the step and the slabs whose union it is (nslab may be 0: nothing pasted around the tile).
 * hard_tile != 0: the tile contains hard data (the caller knows: it owns the grids it passed to iq_sim_begin); the
 * hard distance then is the primary source and the overlap distance the first auxiliary one (src/iqsim.jl:230-231). */
int32_t iq_sim_step(iq_ctx* ctx, int64_t step, const int64_t* start, const uint8_t* ovlmask, const iq_sim_slab* slabs,
                    int32_t nslab, int32_t hard_tile);
/* Several MUTUALLY INDEPENDENT tiles in one launch (dependency-level batching).  A tile only reads and writes its own
 * window of the simulation grid, so it depends exactly on the earlier path tiles whose windows intersect it; tiles that
 * share a dependency level (iqh_dependency_levels in iqb200_host.h: i + 2j + 4k on a raster path) may run together and
 * the result is bit for bit that of the sequential path -- every (realization, step) keeps its own uniform
 * u[r][steps[k]].  iq_sim_define_shape registers an overlap shape (mask + the slabs whose union it is) once and returns
 * its id; iq_sim_step_multi launches ntile tiles whose windows do not intersect (checked): tile k has path step
 * steps[k], origin starts[3k..3k+2] and overlap shape shapes[k].  Shapes may be mixed freely, except that tiles without
 * any pasted neighbour (empty mask) cannot share a launch with others.  ntile * nreal must not exceed the max_batch of the
 * context.  hard_tiles != 0: EVERY tile of the launch contains hard data (hard distance primary, src/iqsim.jl:230-231);
 * tiles with and without data go in separate launches.  Soft data: one auxiliary distance map per tile and source is
 * computed and shared by the realizations of that tile. */
int32_t iq_sim_define_shape(iq_ctx* ctx, const uint8_t* ovlmask, const iq_sim_slab* slabs, int32_t nslab, int32_t* shape);
int32_t iq_sim_step_multi(iq_ctx* ctx, int32_t ntile, const int64_t* steps, const int64_t* starts, const int32_t* shapes,
                          int32_t hard_tiles);
/* A step whose patterns the caller chose itself (picks[r] = 0-based linear index of the pattern of realization r):
 * the whole tile is pasted (no pasted neighbour, hence no cut).  Used for empty-mask steps of soft-data simulations. */
int32_t iq_sim_step_picked(iq_ctx* ctx, int64_t step, const int64_t* start, const int64_t* picks);
/* Waits for the enqueued steps.  picks (may be NULL): [nreal][npath] chosen patterns; status: 0 = ok. */
int32_t iq_sim_sync(iq_ctx* ctx, int64_t* picks, int32_t* status);
/* Copies realization r, cropped to crop[3] (unused dims = 1), to the host as FP64 (dtype 0) or FP32 (dtype 1). */
int32_t iq_sim_fetch(iq_ctx* ctx, int32_t r, int32_t dtype, const int64_t* crop, void* out);
/* Same for all realizations of the context: out[r] receives realization r.  Export kernel, device->host copy into a
 * pinned double buffer and the host copy into out[r] (nthreads host threads, 0 = 4) are pipelined. */
int32_t iq_sim_fetch_all(iq_ctx* ctx, int32_t dtype, const int64_t* crop, void* const* out, int32_t nthreads);
/* Copies the boundary-cut grid of realization r (pad_size bytes; debug simulations only). */
int32_t iq_sim_fetch_cut(iq_ctx* ctx, int32_t r, uint8_t* out);
/* Device time (ms, CUDA events on the context's stream) of the simulation since iq_sim_begin: whole stream, the
 * distance computations, the selection + tau + sampling kernels, the boundary cuts.  Valid after iq_sim_sync. */
int32_t iq_sim_times(const iq_ctx* ctx, double* total_ms, double* dist_ms, double* select_ms, double* cut_ms);
int32_t iq_sim_end(iq_ctx* ctx);

/* Timing hooks for benchmarks: device time (ms, CUDA events on the context's stream) and number
 * of kernels launched by the most recent iq_search* call. */
int32_t iq_last_search_stats(const iq_ctx* ctx, double* device_ms, int64_t* kernel_launches);

/* Device time (ms) spent inside the distance kernels (direct kernel or FFT passes) alone during the most recent
 * iq_search* call, measured with CUDA events on the context's stream, and the number of its launches. */
int32_t iq_last_search_kernel_ms(const iq_ctx* ctx, double* dist_ms, int64_t* dist_launches);

/* Which distance kernels the most recent iq_search* call used: tile searches served by the direct
 * correlation kernel, by the FFT path, the algorithmic bytes the FFT passes moved and their device time. */
int32_t iq_last_search_path(const iq_ctx* ctx, int64_t* direct_searches, int64_t* fft_searches, double* fft_bytes,
                            double* fft_ms);

/* Direct-kernel launches since the context was created: through the TMA-staged kernel (k_dist_flat: tensor-map boxes of
 * the image, double-buffered) and through the register-staged fallback (k_dist_flat_ldg: option "variant" = 1, mask
 * boxes beyond TMA's 256-element box sides, more than 8 boxes, or a driver without cuTensorMapEncodeTiled). */
int32_t iq_ctx_direct_kernel_launches(const iq_ctx* ctx, int64_t* tma_staged, int64_t* register_staged);

/* FP32 FMA issue-rate microbenchmark on `device` (register-operand FFMA chains on every SM): writes the
 * measured rate in TFMA/s (1 FMA = 2 flop).  This is the denominator of the kernel's FMA roofline. */
int32_t iq_bench_fma_peak(int32_t device, double* tfma_per_s);
/* Same with the packed fma.rn.f32x2 instruction (two FMAs per issue slot on sm_100). */
int32_t iq_bench_fma2_peak(int32_t device, double* tfma_per_s);

/* Device buffers are taken from the device's stream-ordered memory pool with the release threshold lifted, so the
 * memory of a destroyed context / finished simulation is reused by the next one instead of going back to the driver.
 * iq_release_device_memory returns everything that is currently unused on `device` to the driver. */
int32_t iq_release_device_memory(int32_t device);
/* Free / total memory of `device` in bytes (cudaMemGetInfo), counting the library's cached pool memory as free. */
int32_t iq_device_free_memory(int32_t device, size_t* free_bytes, size_t* total_bytes);

/* Page-locked host memory for result arrays: iq_sim_fetch_all copies device -> host straight into destinations that
 * are page-locked (cudaHostAlloc / cudaHostRegister memory: no bounce buffer, no host-side copy, full PCIe / C2C rate);
 * pageable destinations go through the library's pinned double buffer.  A host language that wants the fast path
 * allocates its result arrays here (the Python mirror pools them). */
int32_t iq_host_alloc(size_t bytes, void** out);
int32_t iq_host_free(void* p);

/* Tuning knobs (benchmarks/tests): key is one of "rb" (tiles per CTA pass: 0 = auto,
 * 1, 2, 4), "variant" (direct kernel: 0 TMA-staged, 1 register-staged), "fft" (-1 never, 0 auto crossover, 1 always),
 * "cut_exact" (iq_cut_batch in exact integer arithmetic, for integer-valued slabs; see iq_sim_desc.exact_cut). */
int32_t iq_ctx_set_option(iq_ctx* ctx, const char* key, int64_t value);

#ifdef __cplusplus
}
#endif
#endif /* IQB200_H */

/*
 * iqb200_host.h -- optional native host driver shipped inside libiqb200.so.
 *
 * The drop-in boundary of this project is include/iqb200.h (the search).  Julia is not available in
 * the build image, so the host side of the reference -- the tile loop of iqsim
 * (/root/reference/src/iqsim.jl:163-312), the boundary cut (/root/reference/src/graphcut.jl:5-84) and
 * the paste -- is restated here in C++ ABOVE that boundary: it only talks to the device through the
 * public entry points of iqb200.h.  The Python mirror of the reference API
 * (imagequilting.jl_b200/api.py) calls iqh_run; a Julia host would keep its own loop and ccall
 * iq_search instead.
 *
 * All realizations advance in lockstep along the (shared) simulation path.  Two pipelines:
 *   host-staged     one iq_search_pick call per path step carries the templates of every realization of a group,
 *                   then the cuts/pastes of that step run on host threads (or iq_cut_batch);
 *   device-resident the grids stay on the device and every step is enqueued through iq_sim_step without any host
 *                   synchronisation (threshold path: no soft / hard data).
 */
#ifndef IQB200_HOST_H
#define IQB200_HOST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct iqh_desc {
  int32_t ndim;                /* 2 or 3 */
  int64_t ti_size[3];
  int64_t tile_size[3];
  int64_t ovl_size[3];         /* ceil(overlap * tilesize), src/iqsim.jl:92 */
  int64_t ntiles[3];           /* src/iqsim.jl:103 */
  int64_t pad_size[3];         /* src/iqsim.jl:106 */
  const double* ti;            /* prepared training image (NaN -> 0) in FP64: the values that are pasted and cut.  NULL = the
                                  image IS FP32 (ti_f32 holds it exactly); the FP64 copy is then derived from ti_f32 */
  const float* ti_f32;         /* same image in FP32: what the device searches */
  const uint8_t* disabled;     /* distsize bytes or NULL (finddisabled, src/utils.jl:115-129) */
  int32_t nsoft;
  const float* const* aux;     /* nsoft padded auxiliary grids (pad_size each), symmetric-padded, NaN -> 0 */
  const float* const* auxti;   /* nsoft auxiliary training images */
  const uint8_t* hard_has;     /* pad_size bytes: voxel carries a non-NaN datum; NULL = no hard data */
  const float* hard_val;       /* pad_size floats: datum value where hard_has */
  const int64_t* path;         /* visited tiles in simulation order: 0-based column-major tile indices,
                                  skipped tiles (findskipped, src/utils.jl:131-156) already removed */
  int64_t npath;
  double tol;
  int32_t nreal;
  const double* u;             /* [nreal][npath] uniforms in the reference's draw order (src/iqsim.jl:243) */
  int32_t debug;               /* nonzero: also return the boundary-cut grids */
  int32_t device;
  int32_t batch;               /* realizations per iq_search_pick call (<= nreal); 0 = all */
  int32_t nthreads;            /* host threads for cut + paste; 0 = hardware concurrency */
  int32_t ngroups;             /* lockstep groups pipelined against the host cuts; 0 = auto */
  int32_t cut_mode;            /* host-staged boundary cuts: 0 = auto (iq_cut_batch on the device when fewer than 6 host
                                  threads and the image is not integer-valued), 1 = host, 2 = device (FP64) */
  int32_t fft_mode;            /* -1 never, 0 auto crossover, 1 always: distance path selection */
  int32_t pipeline;            /* 0 = auto (device-resident whenever the simulation qualifies), 1 = host-staged (grids,
                                  cuts and paste on the host, one iq_search_pick per step), 2 = device-resident
                                  (iq_sim_*: grids, cuts and paste on the device; error if it does not qualify) */
  void* const* out_real;       /* optional: nreal pointers to sim_size-shaped column-major arrays that receive the
                                  realizations cropped to sim_size (src/iqsim.jl:303); out_grids may then be NULL */
  int32_t out_real_f32;        /* element type of out_real: 0 = FP64, 1 = FP32 */
  int64_t sim_size[3];         /* crop of out_real (unused dims = 1) */
} iqh_desc;

typedef struct iqh_stats {
  double search_ms;            /* wall time spent inside iq_search_pick */
  double search_device_ms;     /* device time reported by iq_last_search_stats */
  double cut_ms;               /* wall time of the cut + paste phases */
  double total_ms;
  int64_t searches;            /* tile searches performed (nreal * npath) */
  int64_t kernel_launches;
  int64_t candidates;          /* sum of candidate-set sizes */
  double setup_ms;             /* context creation: uploads + summed-volume tables */
  double dist_kernel_ms;       /* device time inside the dense correlation kernel alone */
  int64_t dist_launches;
  int64_t fft_searches;        /* tile searches served by the FFT path / by the direct kernel */
  int64_t direct_searches;
  double fft_bytes;            /* algorithmic bytes moved by the FFT passes */
  double fft_ms;               /* device time inside the FFT passes */
  int32_t resident;            /* 1 = the device-resident pipeline produced the result */
  int32_t resident_status;     /* status word of a resident attempt that had to be redone host-staged (0 = none) */
  double device_ms;            /* resident: device time of the whole simulation (max over the lockstep groups) */
  double select_ms;            /* resident: device time of selection + tau model + sampling (sum over groups) */
  double cut_device_ms;        /* resident: device time of the boundary cuts (sum over groups) */
  double fetch_ms;             /* resident: wall time of exporting the realizations to the host */
  int64_t max_candidates;      /* host-staged: largest candidate set of any tile search */
  int32_t dep_levels;          /* resident: dependency levels of the path (iqh_dependency_levels) */
  int32_t tiles_per_launch;    /* resident: most tiles one step launch may carry (job slots / realizations per group) */
  int64_t step_launches;       /* resident: step launches issued per lockstep group (npath without batching) */
  double run_ms;               /* resident: wall time from the first enqueue to the end of iq_sim_sync */
  double teardown_ms;          /* resident: wall time of destroying the contexts */
} iqh_stats;

/* Runs the whole simulation.  `out_grids` receives nreal padded grids (pad_size doubles each,
 * column-major; may be NULL when desc->out_real is given); `out_cuts` (may be NULL unless debug) nreal pad_size byte grids with the boundary-cut
 * masks (src/iqsim.jl:281); `out_picks` (may be NULL) the chosen pattern per realization and step,
 * [nreal][npath], for parity tests.  Returns IQ_OK or an IQ_ERR_* code (message: iq_last_error()). */
int32_t iqh_run(const iqh_desc* desc, double* out_grids, uint8_t* out_cuts, int64_t* out_picks, iqh_stats* stats);

/* iqh_run parks the contexts of a finished device-resident simulation (up to 4) and reuses one when the next call
 * would build an identical context (iq_ctx_matches: same geometry, same job slots, bitwise the same images -- they are
 * uploaded and compared on every call).  The parked contexts keep their device buffers; iqh_cache_clear destroys them
 * (IQB200_CTX_CACHE=0 disables the cache). */
int32_t iqh_cache_clear(void);

/* Dependency levels of a simulation path (the schedule of the device-resident pipeline).  A tile only reads and writes
 * its own window of the simulation grid (template src/iqsim.jl:185, cut slabs :251-275, paste :278), so step s depends
 * exactly on the earlier steps whose tile windows intersect its own: levels[s] = 1 + max level of those, 0 without any.
 * Steps of one level are mutually independent; executing the levels in order, each level as one batch, reproduces the
 * sequential path bit for bit (the overlap masks are still those of the path order: which neighbours were pasted
 * EARLIER IN THE PATH).  Raster paths: levels[tile (i,j,k)] = i + 2j + 4k.  path: npath distinct 0-based column-major
 * tile indices; tile_size / ovl_size / ntiles: ndim entries. */
int32_t iqh_dependency_levels(int32_t ndim, const int64_t* tile_size, const int64_t* ovl_size, const int64_t* ntiles,
                              const int64_t* path, int64_t npath, int32_t* levels, int32_t* nlevels);

/* The boundary cut alone: keep-mask (1 = keep the already pasted voxel) for overlap slabs A (old) and
 * B (new) of size sz (ndim entries, column-major) cut along `dim`.  Restates graphcut(A, B, dim). */
int32_t iqh_graphcut(const double* A, const double* B, int32_t ndim, const int64_t* sz, int32_t dim, uint8_t* keep);
/* iqh_graphcut picks the arithmetic by the data: integer-valued slabs are cut in exact integer arithmetic (the FP64
 * capacities of graphcut.jl:52 scaled to 128-bit integers; no rounding in the flow, hence the one "cannot reach the
 * sink" set of those capacities whatever the max-flow algorithm), everything else in FP64.  On integer-valued
 * (categorical) slabs the capacities are degenerate -- (Du+Dv)/eps next to O(1) terms, many equal-cost cuts -- and an
 * FP64 max-flow, GraphsFlows' Boykov-Kolmogorov included, returns whichever cut its own rounding favours.
 * iqh_graphcut_mode forces one or the other (exact != 0: IQ_ERR_INVALID when the capacities span more than 128 bits). */
int32_t iqh_graphcut_mode(const double* A, const double* B, int32_t ndim, const int64_t* sz, int32_t dim, int32_t exact,
                          uint8_t* keep);

#ifdef __cplusplus
}
#endif
#endif /* IQB200_HOST_H */

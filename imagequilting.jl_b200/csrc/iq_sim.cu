// iq_sim.cu -- device-resident simulation: the grids of all realizations of a context stay in device memory and a
// whole path step is enqueued on the context's stream without any host synchronisation (include/iqb200.h,
// iq_sim_*).  SURVEY.md 8(f) rank 2 ("tile paste / simgrid on device + realization-lockstep driver").
//
// Restated reference code (relative to /root/reference/):
//   tile view of the simulation grid        src/iqsim.jl:185          -> k_sim_templates
//   overlap distance + threshold selection  src/iqsim.jl:187-237      -> existing distance / pick kernels
//   tau model                               src/taumodel.jl:5-45      -> k_tau_rank / k_tau_prob
//   sample(rng, patterndb, weights(p))      src/iqsim.jl:243          -> k_sim_sample (pre-drawn uniforms)
//   boundary cut per overlap slab           src/iqsim.jl:251-275, src/graphcut.jl:5-84 -> k_sim_slabs + k_graphcut
//   simdev[.!cutmask] = TIdev[.!cutmask]    src/iqsim.jl:278          -> k_sim_paste
// The host only decides what depends on the path alone (tile origin, overlap mask, slabs) and enqueues kernels.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "iq_ctx.h"

namespace iqimpl {

struct SlabDev {
  int dim, prev;
  int lo[3], sz[3];
  int n0, n1, L;        // cut layout: [L][n1][n0], cut dimension slowest
  int st_a, st_b, st_k; // strides (in tile voxels) of the layout axes inside the tile
};
struct SlabSet {
  int n;
  SlabDev s[6];
};

// One tile of a (multi-tile) step: origin in the padded grid and the path step it belongs to (index of its uniform /
// pick).  A step launch carries ntile x R jobs: job j = tile j / R of realization j % R.
struct TileInfo { int sx, sy, sz, step, shape; };

// One overlap shape of the simulation: the mask, its A2 map (sum of img^2 over the mask per position) and the slabs
// whose union the mask is.  Tiles of different shapes may share a launch (they only differ in these three).
struct ShapeDev {
  const uint8_t* mask;   // [tilevol]
  const float* a2;       // [npos], nullptr for the empty mask
  SlabSet S;
};
constexpr int kMaxShapes = 256;

struct SimState {
  int R = 0;
  int J = 0;                      // job slots of one launch (= max_batch of the context, a multiple of R)
  TileInfo* d_tiles = nullptr;    // [npath] tile records in launch order (every path step is launched exactly once)
  TileInfo* h_tiles = nullptr;    // pinned mirror
  size_t tile_cursor = 0;         // tiles launched so far
  ShapeDev* d_shapes = nullptr;   // [kMaxShapes] shape table (entries are written once, stream ordered)
  ShapeDev* h_shapes = nullptr;   // pinned mirror
  struct ShapeHost { MaskEntry* e; std::vector<int> sig; size_t smem; int nslab; };
  std::vector<ShapeHost> shapes;
  const float** d_a2list = nullptr;  // [J] A2 map of every job of the current launch (written by k_sim_templates)
  int4* d_cutdims = nullptr;         // [J * maxslabs] {n0, n1, L, -} of every cut task of the current launch (L = 0: none)
  int pad[3] = {1, 1, 1};
  long long padvol = 0;
  int64_t npath = 0;
  double tol = 0.1;
  int debug = 0;
  bool exact = false;             // boundary cuts in exact integer arithmetic (integer-valued images)
  double* d_grid = nullptr;       // [R][padvol]
  uint8_t* d_cutgrid = nullptr;   // [R][padvol] (debug)
  double* d_ti64 = nullptr;       // [nimg]
  double* d_u = nullptr;          // [R][npath]
  std::vector<double> h_u;
  float* d_tmpl = nullptr;        // [R][tilevol] dense masked templates
  float* d_pack = nullptr;        // packed templates of the direct kernel
  size_t pack_cap = 0;
  double* d_b2 = nullptr;         // [R]
  double* d_plane = nullptr;      // [R][tz] plane sums of B2
  unsigned* d_ticket = nullptr;   // [R]
  long long* d_picked = nullptr;  // [R] pattern chosen in the current step
  long long* d_picks = nullptr;   // [R][npath]
  long long* h_pickstage = nullptr;  // pinned [npath][R]: picks of empty-mask steps (computed on the host)
  int* d_status = nullptr;
  // soft data (relaxation path)
  int S = 0;
  std::vector<float*> d_aux_pad;     // [S] padded auxiliary grids
  float* d_soft_tmpl = nullptr;      // [tilevol] soft template of the step (the same for every realization)
  double* d_soft_b2 = nullptr;
  double* d_soft_plane = nullptr;    // [tz]
  unsigned* d_soft_ticket = nullptr;
  iq::PickJob* d_pickjobs = nullptr; // [R] selection jobs of the simulation (the context's own array serves iq_search)
  // hard data (tiles that contain data: hard distance primary, overlap distance first auxiliary; iqsim.jl:230-231)
  uint8_t* d_hard_has = nullptr;     // [padvol]
  float* d_hard_val = nullptr;       // [padvol]
  int* d_hard_ptr = nullptr;         // [2] CSR row of the step's tile
  long long* d_hard_off = nullptr;   // [tilevol] image offsets of the data voxels relative to the patch origin
  float* d_hard_list = nullptr;      // [tilevol] their values
  iq::PickJob* d_pickjobs_hard = nullptr;  // [R] selection jobs of hard tiles
  int* d_pending = nullptr;          // [R] relaxation: realization still without candidates (next round needed)
  double* d_cutA = nullptr;
  double* d_cutB = nullptr;
  uint8_t* d_keep = nullptr;
  int* d_cut_iters = nullptr;
  size_t maxslab = 0;             // voxels of the largest slab
  int maxslabs = 0;               // slabs per tile at most
  void* d_export = nullptr;       // cropped realization in the output type
  size_t export_cap = 0;
  char* h_export[2] = {nullptr, nullptr};  // pinned double buffer of iq_sim_fetch_all
  void* d_export2[2] = {nullptr, nullptr};
  size_t export2_cap = 0;
  cudaEvent_t ev_export[2] = {nullptr, nullptr};
  cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
  std::vector<cudaEvent_t> ev;    // per step: select start, select stop (= cut start), cut stop
  size_t ev_used = 0;
  double total_ms = 0, dist_ms = 0, select_ms = 0, cut_ms = 0;
  bool synced = false;
};

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------

// Dense masked templates (zeros outside the overlap mask) of every realization + B2 = sum of their squares in the
// library's fixed order (b2_ordered in iq_ctx.cu): one CTA per (z plane, realization).
template <typename GT>
__global__ void __launch_bounds__(256) k_sim_templates(const GT* __restrict__ grid, long long padvol, int p0, int p1,
                                                       const TileInfo* __restrict__ tiles, int R,
                                                       const ShapeDev* __restrict__ shapes, const uint8_t* __restrict__ mask,
                                                       const float** __restrict__ a2list, int tx,
                                                       int ty, int tz, float* __restrict__ tmpl, double* __restrict__ plane,
                                                       double* __restrict__ b2, unsigned* __restrict__ ticket) {
  const int z = blockIdx.x, r = blockIdx.y, tid = threadIdx.x;  // r = job: tile r / R of realization r % R
  const int pl = tx * ty;
  const TileInfo T = tiles[r / R];
  const int sx = T.sx, sy = T.sy, sz = T.sz;
  if (shapes) {  // the job's own overlap shape (else: the one mask passed in, e.g. the all-ones soft-data mask)
    mask = shapes[T.shape].mask;
    if (z == 0 && tid == 0 && a2list) a2list[r] = shapes[T.shape].a2;
  }
  const GT* g = grid + (long long)(r % R) * padvol + ((long long)(sz + z) * p1 + sy) * p0 + sx;
  const uint8_t* m = mask + (long long)z * pl;
  float* out = tmpl + ((long long)r * tz + z) * pl;
  double acc = 0.0;
  for (int i = tid; i < pl; i += 256) {
    const int y = i / tx, x = i - y * tx;
    const float v = m[i] ? (float)g[(long long)y * p0 + x] : 0.f;
    out[i] = v;
    acc = __dadd_rn(acc, __dmul_rn((double)v, (double)v));
  }
  __shared__ double s[256];
  __shared__ int s_last;
  s[tid] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) s[tid] = __dadd_rn(s[tid], s[tid + o]);
    __syncthreads();
  }
  if (tid == 0) {
    plane[(long long)r * tz + z] = s[0];
    __threadfence();
    s_last = (atomicAdd(&ticket[r], 1u) == (unsigned)(tz - 1));
  }
  __syncthreads();
  if (!s_last || tid != 0) return;
  __threadfence();
  const volatile double* pv = plane + (long long)r * tz;
  double tot = pv[0];
  for (int k = 1; k < tz; ++k) tot = __dadd_rn(tot, pv[k]);
  b2[r] = tot;
  ticket[r] = 0u;
}

// Dense masked templates -> layout of the direct kernel [grp][box][qz][qy][chunk][r(rb)][8] (zero padded).
__global__ void __launch_bounds__(256) k_sim_pack(const float* __restrict__ tmpl, long long tilevol, int tx, int ty,
                                                  const iq::BoxDesc* __restrict__ boxes, int nbox, long long tmpl_floats,
                                                  int rb, int R, long long total, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const long long gstride = tmpl_floats * rb;
  const int g = (int)(i / gstride);
  long long w = i - (long long)g * gstride;
  int b = nbox - 1;
  while (b > 0 && (long long)boxes[b].tmpl_off * rb > w) --b;
  const iq::BoxDesc B = boxes[b];
  w -= (long long)B.tmpl_off * rb;
  const int tap = (int)(w & 7);
  long long q = w >> 3;
  const int ri = (int)(q % rb); q /= rb;
  const int ch = (int)(q % B.nch); q /= B.nch;
  const int qy = (int)(q % B.h);
  const int qz = (int)(q / B.h);
  const int x = ch * 8 + tap - B.pad, r = g * rb + ri;  // B.pad leading zero taps per packed row
  float v = 0.f;
  if (r < R && x >= 0 && x < B.w) v = tmpl[(long long)r * tilevol + ((long long)(B.z0 + qz) * ty + (B.y0 + qy)) * tx + B.x0 + x];
  out[i] = v;
}

// Base.sum's pairwise reduction (Base.mapreduce_impl, block 1024), as julia_sum in iq_ctx.cu: sequential inside a
// leaf (hi - lo < 1024), left + right above.
__device__ double dev_julia_sum(const double* w, int lo, int hi) {
  if (hi - lo < 1024) {
    double v = w[lo];
    for (int i = lo + 1; i <= hi; ++i) v = __dadd_rn(v, w[i]);
    return v;
  }
  const int mid = lo + ((hi - lo) >> 1);
  return __dadd_rn(dev_julia_sum(w, lo, mid), dev_julia_sum(w, mid + 1, hi));
}

// The same sum by one warp, bit for bit: lane l walks down the recursion tree following the bits of l (most
// significant first) for at most 5 levels; a node that is a leaf before level 5 belongs to the lane whose remaining
// bits are zero, the other lanes below it own nothing.  Every owner sums its sub-range with the sequential/recursive
// routine, then the tree is folded upwards: a parent is left + right when both halves exist, else the half that does.
__device__ double warp_julia_sum(const double* w, int n, int lane) {
  int lo = 0, hi = n - 1;
  bool owner = true;
#pragma unroll
  for (int level = 4; level >= 0; --level) {
    if (hi - lo < 1024) {                       // leaf reached early
      if (lane & ((2 << level) - 1)) owner = false;  // only the lane with all remaining bits clear keeps it
      break;
    }
    const int mid = lo + ((hi - lo) >> 1);
    if ((lane >> level) & 1) lo = mid + 1; else hi = mid;
  }
  double v = owner ? dev_julia_sum(w, lo, hi) : 0.0;
  int have = owner ? 1 : 0;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double ov = __shfl_down_sync(0xffffffffu, v, o);
    const int oh = __shfl_down_sync(0xffffffffu, have, o);
    if ((lane & (2 * o - 1)) == 0) {            // left child of a level-o pair
      if (have && oh) v = __dadd_rn(v, ov);
      else if (oh) { v = ov; have = 1; }
    }
  }
  return __shfl_sync(0xffffffffu, v, 0);
}

// StatsBase.sample walk (iqsim.jl:243) on the candidate list of every realization; one WARP each: the total is the
// pairwise sum above, the cumulative walk is inherently sequential in FP64 (cw += p[i]) -- the lanes fetch 32
// consecutive weights with one coalesced load (next batch prefetched) and lane 0's running sum is fed by shuffles,
// so the chain costs one DADD latency per candidate instead of a dependent global load.
// status bits: 1 = candidate set outside 1..kTauMax, 2 = a boundary cut hit its iteration cap, 4 = relaxation rounds.
__global__ void __launch_bounds__(128) k_sim_sample(const iq::PickJob* __restrict__ jobs, const double* __restrict__ prob,
                                                    const double* __restrict__ u, long long npath,
                                                    const TileInfo* __restrict__ tiles, int R, int njobs,
                                                    long long* __restrict__ picked, long long* __restrict__ picks,
                                                    int* __restrict__ status) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;  // r = job
  if (r >= njobs) return;
  const long long ustep = (long long)(r % R) * npath + tiles[r / R].step;  // uniform / pick slot of (realization, step)
  const iq::PickJob& J = jobs[r];
  const unsigned n = *J.total;
  unsigned pos = 0;
  if (n == 0u || n > (unsigned)iq::kTauMax) {
    if (lane == 0) atomicOr(status, 1);
  } else if (n > 1u) {
    const double* p = prob + (long long)r * iq::kTauMax;
    const double t = __dmul_rn(u[ustep], warp_julia_sum(p, (int)n, lane));
    // first i with cw_i >= t among i < n-1, else n-1, where cw_0 = p[0], cw_i = fl(cw_{i-1} + p[i])
    // Every lane carries the same running sum (the only dependent chain: one DADD per candidate, cw_0 = 0 + p[0]
    // exactly); lane j keeps the value after candidate base + j and the crossing test runs once per batch of 32.
    double cw = 0.0;
    bool found = false;
    double cur = lane < n ? p[lane] : 0.0;
    for (unsigned base = 0; base < n && !found; base += 32) {
      const unsigned nb = base + 32;
      const double nxt = (nb + lane < n) ? p[nb + lane] : 0.0;  // prefetch the next batch
      double mine = 0.0;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        cw = __dadd_rn(cw, __shfl_sync(0xffffffffu, cur, j));
        if (lane == j) mine = cw;
      }
      const unsigned i = base + lane;
      const bool hit = i < n && (!(mine < t) || i == n - 1u);
      const unsigned bal = __ballot_sync(0xffffffffu, hit);
      if (bal) { pos = base + (unsigned)(__ffs(bal) - 1); found = true; }
      cur = nxt;
    }
  }
  if (lane == 0) {
    const long long pk = (n == 0u) ? 0 : (long long)J.cand_idx[pos];
    picked[r] = pk;
    picks[ustep] = pk;
  }
}

// Relaxation rounds (src/relaxation.jl:7-36) of every realization on the device.  Round 0: radix-select jobs for the
// dbsize smallest keys of the primary map and the softk = ceil(frac * npatterns) smallest of every auxiliary map, with
// dbsize = all(D .== 0) ? npatterns : ceil(tol * npatterns) and frac = 0.1 * dbsize / npatterns; round j > 0 (only for
// realizations whose intersection was empty): frac = min(frac + 0.1, 1), auxiliary maps re-selected -- the same FP64
// operations as the host driver of iq_search.  The auxiliary maps are shared by all realizations, so a realization
// whose k equals realization 0's only copies its thresholds (k_sim_copykth).  One block per realization, one thread
// per source.
__global__ void k_sim_seljobs(iq::SelJob* __restrict__ jobs, const iq::PickJob* __restrict__ pick, int maxS,
                              const unsigned* __restrict__ maxbits, int maxbits_stride, int maxbits_tile_stride, int R,
                              const unsigned* __restrict__ minmax, int B, int hard, unsigned shared_mask, double tol,
                              long long npatterns, long long npos, int round, int* __restrict__ pending,
                              unsigned long long* __restrict__ selbuf, unsigned selcap) {
  const int r = blockIdx.x, s = threadIdx.x;  // r = job; the jobs of one tile (r / R) share that tile's auxiliary maps
  const int lead = (r / R) * R;               // first job of the tile: it runs the selection on the shared maps
  if (s >= maxS) return;
  if (round > 0 && pending[r] == 0) return;  // candidates found: its jobs stay finished
  if (round > 0 && s == 0) return;           // primary threshold already known
  iq::SelJob& J = jobs[(long long)r * maxS + s];
  const iq::PickJob& P = pick[r];
  J.k = 0; J.prefix = 0; J.mask = 0; J.kth = 0; J.pass = 0; J.active = 0; J.ticket = 0;
  J.sub = 0; J.nv = 0;
  J.cbuf = selbuf + ((long long)r * maxS + s) * selcap; J.ccap = selcap; J.ccount = 0; J.compact = 0;
  for (int i = 0; i < 256; ++i) J.hist[i] = 0;
  if (s == 0 && round == 0) pending[r] = 1;
  if (s >= P.nsrc) return;
  const long long tileoff = (long long)(r / R) * maxbits_tile_stride;
  const bool allzero = maxbits[(long long)r * maxbits_stride + tileoff] == 0u;
  const bool allzero0 = maxbits[(long long)lead * maxbits_stride + tileoff] == 0u;
  const long long dbsize = allzero ? npatterns : (long long)ceil(__dmul_rn(tol, (double)npatterns));
  double frac = __dmul_rn(0.1, __ddiv_rn((double)dbsize, (double)npatterns));
  for (int j = 0; j < round; ++j) frac = fmin(__dadd_rn(frac, 0.1), 1.0);  // relaxation.jl:35, once per empty round
  const long long softk = (long long)ceil(__dmul_rn(frac, (double)npatterns));
  long long k = s == 0 ? dbsize : softk;
  k = max(1ll, min(k, npos));
  J.map = P.src[s];
  J.k = (unsigned long long)k;
  {
    // range-adaptive digits: [min, max] of this source map over the enabled positions, written by its distance
    // epilogue -- kind 0 = overlap map of job r, 1 = hard map of the tile, 2 + i = soft map i of the tile
    const int tile = r / R;
    int kind, slot;
    if (hard) { kind = s == 0 ? 1 : (s == 1 ? 0 : s); slot = s == 1 ? r : tile; }
    else { kind = s == 0 ? 0 : 1 + s; slot = s == 0 ? r : tile; }
    const unsigned lo = minmax[(size_t)(kind * 2 + 0) * B + slot], hi = minmax[(size_t)(kind * 2 + 1) * B + slot];
    if (lo <= hi && hi < 0x7f800000u) {
      const unsigned range = hi - lo;
      const int t = range ? 32 - __clz(range) : 1;  // the values span t bits
      int sh = 32 + max(t - 8, 0), nv = 0;
      for (;;) {
        J.vshift[nv++] = sh;
        if (sh == 32 || nv == 4) break;
        sh = max(sh - 8, 32);
      }
      J.vshift[nv - 1] = 32;  // the last value digit always ends at the lowest value bit
      J.nv = nv;
      J.sub = lo;
      J.mask = ~((1ull << (J.vshift[0] + 8)) - 1ull);  // in-range keys have nothing above the first digit; +Inf has
    }
  }
  // map shared by all realizations (bit s of shared_mask) and the same k as realization 0, which runs its selection
  // in this round: copy afterwards.  (k depends on the all-zero flag and the round only; pending[] is not written
  // during rounds > 0.)
  const bool same_as_first = ((shared_mask >> s) & 1u) && r != lead && allzero == allzero0 && (round == 0 || pending[lead] != 0);
  J.active = same_as_first ? 0 : 1;
  J.ticket = same_as_first ? 0xffffffffu : 0u;  // marker read by k_sim_copykth
}

// Thresholds of the shared auxiliary maps for the realizations that did not run their own selection.
__global__ void k_sim_copykth(iq::SelJob* __restrict__ jobs, int maxS, int R) {
  const int r = blockIdx.x, s = threadIdx.x;
  const int lead = (r / R) * R;
  if (r == lead || s >= maxS) return;
  iq::SelJob& J = jobs[(long long)r * maxS + s];
  if (J.ticket != 0xffffffffu) return;
  J.kth = jobs[(long long)lead * maxS + s].kth;
  J.ticket = 0u;
}

// After a round: a realization stays pending iff its intersection is empty; `left` counts them at the last round.
__global__ void k_sim_relax_check(const iq::PickJob* __restrict__ pick, int R, int* __restrict__ pending, int last,
                                  int* __restrict__ status) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R || pending[r] == 0) return;
  if (*pick[r].total > 0u) pending[r] = 0;
  else if (last) atomicOr(status, 4);  // still empty after the rounds run on the device
}

// Hard data of the step's tile (indicator! / event!, utils.jl:18-36) as the sparse list k_dist_sparse consumes, in
// ascending tile offset (the order the host driver builds it in: the FP32 accumulation order of the distance).
__global__ void __launch_bounds__(256) k_sim_hardlist(const uint8_t* __restrict__ has, const float* __restrict__ val, int p0,
                                                      int p1, const TileInfo* __restrict__ tiles, int tx, int ty, int tz,
                                                      int nx, int ny, int* __restrict__ ptr, long long* __restrict__ off,
                                                      float* __restrict__ list) {
  __shared__ unsigned s_w[8];
  __shared__ unsigned s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tv = tx * ty * tz;
  // one CTA per tile of the launch: its list occupies the segment [tile * tv, tile * tv + count), ptr holds the pair
  const TileInfo T = tiles[blockIdx.x];
  const int sx = T.sx, sy = T.sy, sz = T.sz;
  off += (long long)blockIdx.x * tv;
  list += (long long)blockIdx.x * tv;
  ptr += 2 * blockIdx.x;
  const int seg0 = blockIdx.x * tv;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int q0 = 0; q0 < tv; q0 += 256) {
    const int q = q0 + tid;
    bool hit = false;
    long long gi = 0;
    int qx = 0, qy = 0, qz = 0;
    if (q < tv) {
      qx = q % tx; const int qt = q / tx; qy = qt % ty; qz = qt / ty;
      gi = ((long long)(sz + qz) * p1 + (sy + qy)) * p0 + sx + qx;
      hit = has[gi] != 0;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) s_w[warp] = __popc(bal);
    __syncthreads();
    unsigned before = 0, all = 0;
    for (int w = 0; w < 8; ++w) { const unsigned c = s_w[w]; if (w < warp) before += c; all += c; }
    const unsigned base = s_base;
    if (hit) {
      const unsigned slot = base + before + __popc(bal & ((1u << lane) - 1u));
      off[slot] = ((long long)qz * ny + qy) * nx + qx;
      list[slot] = val[gi];
    }
    __syncthreads();
    if (tid == 0) s_base = base + all;
    __syncthreads();
  }
  if (tid == 0) { ptr[0] = seg0; ptr[1] = seg0 + (int)s_base; }
}

__global__ void k_sim_store_picks(const long long* __restrict__ picked, long long* __restrict__ picks, long long npath,
                                  const TileInfo* __restrict__ tiles, int R, int njobs) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;  // job
  if (r < njobs) picks[(long long)(r % R) * npath + tiles[r / R].step] = picked[r];
}

// Overlap slabs of every job in the cut kernel's layout: A = pasted content, B = chosen patch.  One CTA per
// (job, slab slot k < maxn); a job whose shape has fewer slabs marks the slot empty (dims.z = 0).
__global__ void __launch_bounds__(256) k_sim_slabs(const double* __restrict__ grid, long long padvol, int p0, int p1,
                                                   const TileInfo* __restrict__ tiles, int R,
                                                   const ShapeDev* __restrict__ shapes, int maxn, int maxslabs,
                                                   const double* __restrict__ ti, int nx, int ny, int nxo,
                                                   int nyo, const long long* __restrict__ picked, int tx, int ty,
                                                   double* __restrict__ A, double* __restrict__ B, int4* __restrict__ dims,
                                                   long long maxslab) {
  const int r = blockIdx.x / maxn, k = blockIdx.x - r * maxn;  // r = job
  const TileInfo T = tiles[r / R];
  const SlabSet& S = shapes[T.shape].S;
  const int task = r * maxslabs + k;
  if (k >= S.n) {
    if (threadIdx.x == 0) dims[task] = make_int4(0, 0, 0, 0);
    return;
  }
  const SlabDev& s = S.s[k];
  if (threadIdx.x == 0) dims[task] = make_int4(s.n0, s.n1, s.L, 0);
  const int sx = T.sx, sy = T.sy, sz = T.sz;
  const long long pk = picked[r];
  const int rx = (int)(pk % nxo), ry = (int)((pk / nxo) % nyo), rz = (int)(pk / ((long long)nxo * nyo));
  const double* g = grid + (long long)(r % R) * padvol;
  double* a = A + (long long)task * maxslab;
  double* b = B + (long long)task * maxslab;
  const int nv = s.n0 * s.n1 * s.L;
  const int q0 = (s.lo[2] * ty + s.lo[1]) * tx + s.lo[0];
  for (int o = threadIdx.x; o < nv; o += 256) {
    const int aa = o % s.n0, t = o / s.n0, bb = t % s.n1, kk = t / s.n1;
    const int q = q0 + aa * s.st_a + bb * s.st_b + kk * s.st_k;  // voxel inside the tile
    const int qx = q % tx, qt = q / tx, qy = qt % ty, qz = qt / ty;
    a[o] = g[((long long)(sz + qz) * p1 + (sy + qy)) * p0 + sx + qx];
    b[o] = ti[((long long)(rz + qz) * ny + (ry + qy)) * nx + rx + qx];
  }
}

// cutmask = OR over the slabs of (prev ? keep : !keep) (iqsim.jl:264,273); simdev[.!cutmask] = TIdev[.!cutmask].
__global__ void __launch_bounds__(256) k_sim_paste(double* __restrict__ grid, uint8_t* __restrict__ cutgrid, long long padvol,
                                                   int p0, int p1, const TileInfo* __restrict__ tiles, int R,
                                                   const ShapeDev* __restrict__ shapes, int maxslabs,
                                                   const double* __restrict__ ti,
                                                   int nx, int ny, int nxo, int nyo, const long long* __restrict__ picked,
                                                   int tx, int ty, int tz, const uint8_t* __restrict__ keep,
                                                   long long maxslab, const int* __restrict__ cut_iters,
                                                   int* __restrict__ status) {
  const int r = blockIdx.y;  // job
  const TileInfo T = tiles[r / R];
  const int sx = T.sx, sy = T.sy, sz = T.sz;
  const int nslab = shapes ? shapes[T.shape].S.n : 0;  // shapes == nullptr: the whole tile is pasted (no pasted neighbour)
  const int q = blockIdx.x * 256 + threadIdx.x;
  if (q == 0 && nslab > 0) {
    bool bad = false;
    for (int k = 0; k < nslab; ++k) bad |= cut_iters[r * maxslabs + k] < 0;
    if (bad) atomicOr(status, 2);
  }
  if (q >= tx * ty * tz) return;
  const int qx = q % tx, qt = q / tx, qy = qt % ty, qz = qt / ty;
  unsigned cm = 0;
  for (int k = 0; k < nslab; ++k) {
    const SlabDev& s = shapes[T.shape].S.s[k];
    const int lx = qx - s.lo[0], ly = qy - s.lo[1], lz = qz - s.lo[2];
    if (lx < 0 || ly < 0 || lz < 0 || lx >= s.sz[0] || ly >= s.sz[1] || lz >= s.sz[2]) continue;
    const int l[3] = {lx, ly, lz};
    const int da = s.dim == 0 ? 1 : 0, db = s.dim == 2 ? 1 : 2;
    const int o = (l[s.dim] * s.n1 + l[db]) * s.n0 + l[da];
    const unsigned kv = keep[(long long)(r * maxslabs + k) * maxslab + o];
    cm |= s.prev ? kv : (kv ^ 1u);
  }
  const long long pk = picked[r];
  const int rx = (int)(pk % nxo), ry = (int)((pk / nxo) % nyo), rz = (int)(pk / ((long long)nxo * nyo));
  const long long gi = (long long)(r % R) * padvol + ((long long)(sz + qz) * p1 + (sy + qy)) * p0 + sx + qx;
  if (!cm) grid[gi] = ti[((long long)(rz + qz) * ny + (ry + qy)) * nx + rx + qx];
  if (cutgrid) cutgrid[gi] = (uint8_t)cm;
}

__global__ void __launch_bounds__(256) k_sim_widen(const float* __restrict__ in, double* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i < n) out[i] = (double)in[i];
}

template <typename OT>
__global__ void __launch_bounds__(256) k_sim_export(const double* __restrict__ grid, int p0, int p1, int c0, int c1, int c2,
                                                    OT* __restrict__ out) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long n = (long long)c0 * c1 * c2;
  if (i >= n) return;
  const int x = (int)(i % c0);
  const long long t = i / c0;
  const int y = (int)(t % c1), z = (int)(t / c1);
  out[i] = (OT)grid[((long long)z * p1 + y) * p0 + x];
}

void sim_destroy(iq_ctx* c) {
  SimState* s = c->sim;
  if (!s) return;
  cudaFree(s->d_grid); cudaFree(s->d_cutgrid); cudaFree(s->d_ti64); cudaFree(s->d_u); cudaFree(s->d_tmpl);
  cudaFree(s->d_pack); cudaFree(s->d_b2); cudaFree(s->d_plane); cudaFree(s->d_ticket); cudaFree(s->d_picked);
  cudaFree(s->d_picks); cudaFree(s->d_status); cudaFree(s->d_cutA); cudaFree(s->d_cutB); cudaFree(s->d_keep);
  cudaFree(s->d_cut_iters); cudaFree(s->d_export);
  for (auto p : s->d_aux_pad) cudaFree(p);
  cudaFree(s->d_soft_tmpl); cudaFree(s->d_soft_b2); cudaFree(s->d_soft_plane); cudaFree(s->d_soft_ticket);
  cudaFree(s->d_pickjobs); cudaFree(s->d_pending);
  cudaFree(s->d_hard_has); cudaFree(s->d_hard_val); cudaFree(s->d_hard_ptr); cudaFree(s->d_hard_off);
  cudaFree(s->d_hard_list); cudaFree(s->d_pickjobs_hard);
  if (s->h_pickstage) cudaFreeHost(s->h_pickstage);
  cudaFree(s->d_tiles);
  if (s->h_tiles) cudaFreeHost(s->h_tiles);
  cudaFree(s->d_shapes);
  if (s->h_shapes) cudaFreeHost(s->h_shapes);
  cudaFree(s->d_a2list);
  cudaFree(s->d_cutdims);
  for (int i = 0; i < 2; ++i) {
    if (s->h_export[i]) cudaFreeHost(s->h_export[i]);
    cudaFree(s->d_export2[i]);
    if (s->ev_export[i]) cudaEventDestroy(s->ev_export[i]);
  }
  for (auto e : s->ev) cudaEventDestroy(e);
  if (s->ev_begin) cudaEventDestroy(s->ev_begin);
  if (s->ev_end) cudaEventDestroy(s->ev_end);
  delete s;
  c->sim = nullptr;
}

namespace {

constexpr size_t kCutSmemLimit = 220 * 1024;

bool slab_fits(int n0, int n1, int L, bool exact) {
  if (L < 2) return false;
  if (L == 2) return true;  // no inner voxel: the kernel only writes the source / sink slices
  return iq::graphcut_smem(n0, n1, L, exact) <= kCutSmemLimit && (long long)(L - 2) * n0 * n1 <= 4096;
}

int sim_events(SimState* s, cudaEvent_t* out, int n) {
  while (s->ev_used + n > s->ev.size()) {
    cudaEvent_t e;
    CK(cudaEventCreate(&e));
    s->ev.push_back(e);
  }
  for (int i = 0; i < n; ++i) out[i] = s->ev[s->ev_used + i];
  s->ev_used += n;
  return IQ_OK;
}

}  // namespace
}  // namespace iqimpl

using namespace iqimpl;

extern "C" {

int32_t iq_sim_begin(iq_ctx* c, const iq_sim_desc* d) {
  if (!c || !d || !d->u) return fail(IQ_ERR_INVALID, "iq_sim_begin: NULL argument");
  if (d->nreal < 1 || d->nreal > c->max_batch) return fail(IQ_ERR_INVALID, "iq_sim_begin: nreal must be in 1..max_batch");
  if (c->max_batch % d->nreal != 0)
    return fail(IQ_ERR_INVALID, "iq_sim_begin: max_batch of the context must be a multiple of nreal (job slots = tiles per step x nreal)");
  if (d->npath < 0) return fail(IQ_ERR_INVALID, "iq_sim_begin: npath < 0");
  if (!(d->tol > 0.0 && d->tol <= 1.0)) return fail(IQ_ERR_INVALID, "tolerance must be in range (0,1]");
  if (c->nsoft > 0 && !d->aux) return fail(IQ_ERR_INVALID, "iq_sim_begin: the context has soft data but desc.aux is NULL");
  for (int i = 0; i < c->nsoft; ++i)
    if (!d->aux[i]) return fail(IQ_ERR_INVALID, "iq_sim_begin: aux[%d] is NULL", i);
  CK(cudaSetDevice(c->device));
  sim_destroy(c);
  const int t[3] = {c->tx, c->ty, c->tz};
  int pad[3] = {1, 1, 1}, ov[3] = {1, 1, 1};
  for (int i = 0; i < c->ndim; ++i) {
    pad[i] = (int)d->pad_size[i];
    ov[i] = (int)d->ovl_size[i];
    if (pad[i] < t[i] || ov[i] < 0 || ov[i] >= t[i] + 1) return fail(IQ_ERR_INVALID, "iq_sim_begin: bad pad_size / ovl_size");
  }
  size_t maxslab = 1;
  int maxslabs = 0;
  for (int dd = 0; dd < c->ndim; ++dd) {
    if (ov[dd] <= 1) continue;
    const int da = dd == 0 ? 1 : 0, db = dd == 2 ? 1 : 2;
    if (!slab_fits(t[da], t[db], ov[dd], d->exact_cut != 0))
      return fail(IQ_ERR_STATE, "overlap slab %d x %d x %d does not fit the shared-memory cut kernel", t[da], t[db], ov[dd]);
    maxslab = std::max(maxslab, (size_t)t[da] * t[db] * ov[dd]);
    maxslabs += 2;
  }
  SimState* s = new SimState();
  c->sim = s;
  s->R = d->nreal;
  s->J = c->max_batch;
  for (int i = 0; i < 3; ++i) s->pad[i] = pad[i];
  s->padvol = (long long)pad[0] * pad[1] * pad[2];
  s->npath = d->npath;
  s->tol = d->tol;
  s->debug = d->debug;
  s->exact = d->exact_cut != 0;
  s->maxslab = maxslab;
  s->maxslabs = std::max(maxslabs, 1);
  const size_t R = (size_t)s->R, J = (size_t)s->J, nimg = (size_t)c->nx * c->ny * c->nz, np = (size_t)std::max<int64_t>(d->npath, 1);
  CK(iq::dmalloc((void**)&s->d_tiles, np * sizeof(TileInfo)));
  CK(cudaMallocHost((void**)&s->h_tiles, np * sizeof(TileInfo)));
  CK(iq::dmalloc((void**)&s->d_shapes, kMaxShapes * sizeof(ShapeDev)));
  CK(cudaMallocHost((void**)&s->h_shapes, kMaxShapes * sizeof(ShapeDev)));
  CK(iq::dmalloc((void**)&s->d_a2list, J * sizeof(float*)));
  CK(iq::dmalloc((void**)&s->d_cutdims, J * s->maxslabs * sizeof(int4)));
  CK(iq::dmalloc((void**)&s->d_grid, R * s->padvol * sizeof(double)));
  CK(cudaMemsetAsync(s->d_grid, 0, R * s->padvol * sizeof(double), c->stream));  // simgrid = zeros (iqsim.jl:165)
  if (s->debug) {
    CK(iq::dmalloc((void**)&s->d_cutgrid, R * s->padvol));
    CK(cudaMemsetAsync(s->d_cutgrid, 0, R * s->padvol, c->stream));
  }
  CK(iq::dmalloc((void**)&s->d_ti64, nimg * sizeof(double)));
  if (d->ti64) {
    CK(cudaMemcpyAsync(s->d_ti64, d->ti64, nimg * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  } else {
    k_sim_widen<<<(unsigned)((nimg + 255) / 256), 256, 0, c->stream>>>(c->d_ti, s->d_ti64, (long long)nimg);
    CK(cudaGetLastError());
  }
  s->h_u.assign(d->u, d->u + R * (size_t)d->npath);
  CK(iq::dmalloc((void**)&s->d_u, R * np * sizeof(double)));
  if (d->npath > 0)
    CK(cudaMemcpyAsync(s->d_u, d->u, R * (size_t)d->npath * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CK(iq::dmalloc((void**)&s->d_tmpl, J * c->tilevol * sizeof(float)));
  CK(iq::dmalloc((void**)&s->d_b2, J * sizeof(double)));
  CK(iq::dmalloc((void**)&s->d_plane, J * c->tz * sizeof(double)));
  CK(iq::dmalloc((void**)&s->d_ticket, J * sizeof(unsigned)));
  CK(cudaMemsetAsync(s->d_ticket, 0, J * sizeof(unsigned), c->stream));
  CK(iq::dmalloc((void**)&s->d_picked, J * sizeof(long long)));
  CK(iq::dmalloc((void**)&s->d_picks, R * np * sizeof(long long)));
  CK(cudaMemsetAsync(s->d_picks, 0xff, R * np * sizeof(long long), c->stream));
  CK(cudaMallocHost((void**)&s->h_pickstage, R * np * sizeof(long long)));
  CK(iq::dmalloc((void**)&s->d_status, sizeof(int)));
  CK(cudaMemsetAsync(s->d_status, 0, sizeof(int), c->stream));
  const size_t ntask = J * s->maxslabs;
  CK(iq::dmalloc((void**)&s->d_cutA, ntask * maxslab * sizeof(double)));
  CK(iq::dmalloc((void**)&s->d_cutB, ntask * maxslab * sizeof(double)));
  CK(iq::dmalloc((void**)&s->d_keep, ntask * maxslab));
  CK(iq::dmalloc((void**)&s->d_cut_iters, ntask * sizeof(int)));
  CK(cudaMemsetAsync(s->d_cut_iters, 0, ntask * sizeof(int), c->stream));
  s->S = c->nsoft;
  s->d_aux_pad.assign(s->S, nullptr);
  for (int i = 0; i < s->S; ++i) {
    CK(iq::dmalloc((void**)&s->d_aux_pad[i], (size_t)s->padvol * sizeof(float)));
    CK(cudaMemcpyAsync(s->d_aux_pad[i], d->aux[i], (size_t)s->padvol * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  }
  const size_t T = J / R;  // tiles one launch may carry: the auxiliary (soft / hard) maps belong to a tile, one set per tile
  if (s->S > 0) {
    CK(iq::dmalloc((void**)&s->d_soft_tmpl, T * c->tilevol * sizeof(float)));
    CK(iq::dmalloc((void**)&s->d_soft_b2, T * sizeof(double)));
    CK(iq::dmalloc((void**)&s->d_soft_plane, T * c->tz * sizeof(double)));
    CK(iq::dmalloc((void**)&s->d_soft_ticket, T * sizeof(unsigned)));
    CK(cudaMemsetAsync(s->d_soft_ticket, 0, T * sizeof(unsigned), c->stream));
  }
  CK(iq::dmalloc((void**)&s->d_pickjobs, J * sizeof(iq::PickJob)));
  CK(iq::dmalloc((void**)&s->d_pending, J * sizeof(int)));
  // selection jobs never change during the simulation: threshold rule on the overlap distance of job slot r, or
  // (soft data: one tile per step, slots 0..R-1) the intersection of the radix-select thresholds of the overlap map and
  // the shared soft maps
  for (int r = 0; r < s->J; ++r) {
    iq::PickJob& J = c->h_pick[r];
    std::memset(&J, 0, sizeof J);
    J.mode = s->S > 0 ? 1 : 0;
    J.nsrc = 1 + s->S;
    for (int i = 0; i < s->S; ++i) J.src[1 + i] = c->d_Dsoft[i] + (size_t)(r / s->R) * c->npos;  // one soft map per tile serves every realization
    J.pending = s->S > 0 ? s->d_pending + r : nullptr;
    J.src[0] = c->d_Dovl + (size_t)r * c->npos;
    J.sel = c->d_sel + (size_t)r * c->max_src;
    J.tol = s->tol;
    J.minbits = c->d_minmax + r;
    J.blockcount = c->d_blockcount + (size_t)r * iq::pick_nblk(c->npos);
    J.total = c->d_total + r;
    J.cand_idx = c->d_cand_idx + (size_t)r * c->npos;
    J.cand_val = c->d_cand_val + (size_t)r * c->max_src * c->npos;
    J.cap = c->npos;
    J.chunkmin = c->d_chunkmin + (size_t)r * c->chunk_stride;
  }
  CK(cudaMemcpyAsync(s->d_pickjobs, c->h_pick, J * sizeof(iq::PickJob), cudaMemcpyHostToDevice, c->stream));
  if (d->hard_has) {
    if (!d->hard_val) return fail(IQ_ERR_INVALID, "iq_sim_begin: hard_has without hard_val");
    CK(cudaStreamSynchronize(c->stream));  // h_pick is rewritten below
    CK(iq::dmalloc((void**)&s->d_hard_has, (size_t)s->padvol));
    CK(iq::dmalloc((void**)&s->d_hard_val, (size_t)s->padvol * sizeof(float)));
    CK(cudaMemcpyAsync(s->d_hard_has, d->hard_has, (size_t)s->padvol, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(s->d_hard_val, d->hard_val, (size_t)s->padvol * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CK(iq::dmalloc((void**)&s->d_hard_ptr, 2 * T * sizeof(int)));
    CK(iq::dmalloc((void**)&s->d_hard_off, T * c->tilevol * sizeof(long long)));
    CK(iq::dmalloc((void**)&s->d_hard_list, T * c->tilevol * sizeof(float)));
    CK(iq::dmalloc((void**)&s->d_pickjobs_hard, J * sizeof(iq::PickJob)));
    // hard tiles: sources = [hard distance (one map per tile, for all realizations), overlap distance, soft maps...]
    for (int r = 0; r < s->J; ++r) {
      const size_t tl = (size_t)(r / s->R);
      iq::PickJob& J = c->h_pick[r];
      std::memset(&J, 0, sizeof J);
      J.mode = 1;
      J.nsrc = 2 + s->S;
      J.src[0] = c->d_Dhard + tl * c->npos;
      J.src[1] = c->d_Dovl + (size_t)r * c->npos;
      for (int i = 0; i < s->S; ++i) J.src[2 + i] = c->d_Dsoft[i] + tl * c->npos;
      J.pending = s->d_pending + r;
      J.sel = c->d_sel + (size_t)r * c->max_src;
      J.tol = s->tol;
      J.minbits = c->d_minmax + (size_t)(1 * 2 + 0) * c->max_batch + tl;
      J.blockcount = c->d_blockcount + (size_t)r * iq::pick_nblk(c->npos);
      J.total = c->d_total + r;
      J.cand_idx = c->d_cand_idx + (size_t)r * c->npos;
      J.cand_val = c->d_cand_val + (size_t)r * c->max_src * c->npos;
      J.cap = c->npos;
    }
    CK(cudaMemcpyAsync(s->d_pickjobs_hard, c->h_pick, J * sizeof(iq::PickJob), cudaMemcpyHostToDevice, c->stream));
  }
  CK(cudaEventCreate(&s->ev_begin));
  CK(cudaEventCreate(&s->ev_end));
  CK(cudaStreamSynchronize(c->stream));
  c->dist_ev_used = 0;
  c->last_fft_bytes = 0.0;
  c->last_fft_searches = c->last_direct_searches = 0;
  CK(cudaEventRecord(s->ev_begin, c->stream));
  return IQ_OK;
}

// Registers (or finds) the overlap shape (mask + slab set) and returns its slot in the device shape table.
static int sim_shape(iq_ctx* c, const uint8_t* ovlmask, const iq_sim_slab* slabs, int32_t nslab, int* slot) {
  SimState* s = c->sim;
  const int t[3] = {c->tx, c->ty, c->tz};
  if (nslab < 0 || nslab > 6 || nslab > s->maxslabs || (nslab > 0 && !slabs)) return fail(IQ_ERR_INVALID, "iq_sim: bad slab list");
  // ---- slab geometry (host side of iqsim.jl:251-275) ----
  SlabSet S{};
  S.n = nslab;
  size_t smem = 0;
  std::vector<int> sig;
  for (int k = 0; k < nslab; ++k) {
    const iq_sim_slab& in = slabs[k];
    SlabDev& o = S.s[k];
    if (in.dim < 0 || in.dim >= c->ndim) return fail(IQ_ERR_INVALID, "iq_sim: slab %d: bad dim", k);
    o.dim = in.dim;
    o.prev = in.prev ? 1 : 0;
    for (int i = 0; i < 3; ++i) {
      o.lo[i] = i < c->ndim ? in.lo[i] : 0;
      o.sz[i] = i < c->ndim ? in.sz[i] : 1;
      if (o.lo[i] < 0 || o.sz[i] < 1 || o.lo[i] + o.sz[i] > t[i]) return fail(IQ_ERR_INVALID, "iq_sim: slab %d outside the tile", k);
    }
    const int da = o.dim == 0 ? 1 : 0, db = o.dim == 2 ? 1 : 2;
    const int tst[3] = {1, t[0], t[0] * t[1]};
    o.n0 = o.sz[da]; o.n1 = o.sz[db]; o.L = o.sz[o.dim];
    o.st_a = tst[da]; o.st_b = tst[db]; o.st_k = tst[o.dim];
    if ((size_t)o.n0 * o.n1 * o.L > s->maxslab || !slab_fits(o.n0, o.n1, o.L, s->exact))
      return fail(IQ_ERR_INVALID, "iq_sim: slab %d is larger than the overlap declared at iq_sim_begin", k);
    if (o.L > 2) smem = std::max(smem, iq::graphcut_smem(o.n0, o.n1, o.L, s->exact));
    for (int v : {o.dim, o.prev, o.lo[0], o.lo[1], o.lo[2], o.sz[0], o.sz[1], o.sz[2]}) sig.push_back(v);
  }
  MaskEntry* e = nullptr;
  int rc = get_mask(c, ovlmask, &e);
  if (rc) return rc;
  // the mask alone does not determine the slabs (overlap >= 0.5: {px,nx,py} and {px,py,ny} cover the same voxels with
  // different slabs), so a shape is (mask, slab list)
  for (size_t i = 0; i < s->shapes.size(); ++i)
    if (s->shapes[i].e == e && s->shapes[i].sig == sig) { *slot = (int)i; return IQ_OK; }
  if ((int)s->shapes.size() >= kMaxShapes) return fail(IQ_ERR_STATE, "iq_sim: more than %d distinct overlap shapes", kMaxShapes);
  const float* a2 = nullptr;
  if (e->nnz > 0) {
    rc = get_a2(c, e, -1, &a2);
    if (rc) return rc;
  }
  const int id = (int)s->shapes.size();
  s->shapes.push_back({e, sig, smem, nslab});
  ShapeDev& h = s->h_shapes[id];  // written once, never modified: the pinned entry stays valid for the queued copy
  h.mask = e->d_mask;
  h.a2 = a2;
  h.S = S;
  CK(cudaMemcpyAsync(s->d_shapes + id, &h, sizeof(ShapeDev), cudaMemcpyHostToDevice, c->stream));
  *slot = id;
  return IQ_OK;
}

// One launch of `ntile` mutually independent tiles (no two windows intersect): ntile x R jobs, job j = tile j / R of
// realization j % R; tile k has overlap shape shapes[k].  Tiles with hard data and contexts with soft data take one
// tile per launch (their auxiliary maps belong to the tile, and one map per source is kept).
static int sim_step_impl(iq_ctx* c, int ntile, const int64_t* steps, const int64_t* starts, const int32_t* shapes,
                         int32_t hard_tile) {
  SimState* s = c->sim;
  const bool hardt = hard_tile != 0;
  const int R = s->R;
  if (ntile < 1 || (long long)ntile * R > s->J)
    return fail(IQ_ERR_INVALID, "iq_sim_step: %d tiles x %d realizations exceed the %d job slots (max_batch) of the context", ntile, R, s->J);
  if (s->tile_cursor + (size_t)ntile > (size_t)std::max<int64_t>(s->npath, 1))
    return fail(IQ_ERR_STATE, "iq_sim_step: more tiles launched than path steps declared at iq_sim_begin");
  CK(cudaSetDevice(c->device));
  const int t[3] = {c->tx, c->ty, c->tz};
  TileInfo* ht = s->h_tiles + s->tile_cursor;
  int nempty = 0, maxn = 0;
  size_t smem = 0;
  const MaskEntry* ebig = nullptr;   // the shape with the most mask voxels (FFT / direct decision)
  bool one_shape = true;
  for (int k = 0; k < ntile; ++k) {
    if (steps[k] < 0 || steps[k] >= s->npath) return fail(IQ_ERR_INVALID, "iq_sim_step: step out of range");
    if (shapes[k] < 0 || shapes[k] >= (int)s->shapes.size()) return fail(IQ_ERR_INVALID, "iq_sim_step: unknown shape id");
    int st3[3] = {0, 0, 0};
    for (int i = 0; i < c->ndim; ++i) {
      st3[i] = (int)starts[3 * k + i];
      if (st3[i] < 0 || st3[i] + t[i] > s->pad[i]) return fail(IQ_ERR_INVALID, "iq_sim_step: tile outside the padded grid");
    }
    ht[k].sx = st3[0]; ht[k].sy = st3[1]; ht[k].sz = st3[2]; ht[k].step = (int)steps[k]; ht[k].shape = shapes[k];
    for (int m = 0; m < k; ++m) {  // the tiles of one launch read and write disjoint windows
      const bool apart = std::abs(ht[m].sx - st3[0]) >= t[0] || std::abs(ht[m].sy - st3[1]) >= t[1] || std::abs(ht[m].sz - st3[2]) >= t[2];
      if (!apart) return fail(IQ_ERR_INVALID, "iq_sim_step_multi: tiles %d and %d overlap; only independent tiles can share a step", m, k);
    }
    const SimState::ShapeHost& sh = s->shapes[(size_t)shapes[k]];
    if (sh.e->nnz == 0) ++nempty;
    maxn = std::max(maxn, sh.nslab);
    smem = std::max(smem, sh.smem);
    if (!ebig || sh.e->nnz > ebig->nnz) ebig = sh.e;
    if (shapes[k] != shapes[0]) one_shape = false;
  }
  if (nempty != 0 && nempty != ntile)
    return fail(IQ_ERR_INVALID, "iq_sim_step_multi: tiles without any pasted neighbour cannot share a step with others");
  const TileInfo* dt = s->d_tiles + s->tile_cursor;
  CK(cudaMemcpyAsync(s->d_tiles + s->tile_cursor, ht, (size_t)ntile * sizeof(TileInfo), cudaMemcpyHostToDevice, c->stream));
  const size_t cursor0 = s->tile_cursor;
  s->tile_cursor += (size_t)ntile;
  const int NJ = ntile * R;  // jobs of this launch
  s->synced = false;
  MaskEntry* e = s->shapes[(size_t)shapes[0]].e;
  int rc = IQ_OK;
  const int64_t l0 = c->launches;
  cudaEvent_t ev[3];
  rc = sim_events(s, ev, 3);
  if (rc) return rc;

  const unsigned gJ = (unsigned)((NJ + 127) / 128);
  if (nempty == ntile && s->S == 0 && !hardt) {
    // nothing pasted around the tile: every enabled patch with equal probability (iqsim.jl:237 on an all-zero map);
    // the walk is evaluated on the host from the cached cumulative weights
    rc = build_uniform(c);
    if (rc) return rc;
    const int64_t n = (int64_t)c->enabled_idx.size();
    if (n == 0) return fail(IQ_ERR_INVALID, "all patches of the training image are disabled");
    long long* hp = s->h_pickstage + cursor0 * R;
    for (int j = 0; j < NJ; ++j) {
      const double tt = s->h_u[(size_t)(j % R) * s->npath + ht[j / R].step] * c->uniform_sum;
      const auto it = std::lower_bound(c->uniform_cum.begin(), c->uniform_cum.end() - 1, tt);
      hp[j] = c->enabled_idx[(size_t)(it - c->uniform_cum.begin())];
    }
    CK(cudaEventRecord(ev[0], c->stream));
    CK(cudaMemcpyAsync(s->d_picked, hp, (size_t)NJ * sizeof(long long), cudaMemcpyHostToDevice, c->stream));
    k_sim_store_picks<<<gJ, 128, 0, c->stream>>>(s->d_picked, s->d_picks, s->npath, dt, R, NJ);
    CK(cudaGetLastError());
    c->launches++;
  } else {
    for (int kind = 0; kind < 2 + s->S; ++kind) {
      if (kind == 1 && !hardt) continue;
      CK(iq::launch_fill_u32(c->d_minmax + (size_t)(kind * 2 + 0) * c->max_batch, 0x7f800000u, c->max_batch, c->stream));
      CK(iq::launch_fill_u32(c->d_minmax + (size_t)(kind * 2 + 1) * c->max_batch, 0u, c->max_batch, c->stream));
      c->launches += 2;
    }
    // direct kernel on `nt` dense templates of ONE mask starting at job slot job0: pack into its layout (scratch of the
    // simulation, stream ordered) and launch
    auto direct = [&](MaskEntry* me, int image, const float* d_dense, const double* d_b2, int nt, float* d_out, int kind, int job0) -> int {
      const int rb = pick_rb(c, nt);
      const int ngrp = (nt + rb - 1) / rb;
      const long long total = (long long)ngrp * me->tmpl_floats * rb;
      if ((size_t)total > s->pack_cap) {
        CK(cudaStreamSynchronize(c->stream));
        cudaFree(s->d_pack);
        s->d_pack = nullptr;
        s->pack_cap = 0;
        CK(iq::dmalloc((void**)&s->d_pack, (size_t)total * 2 * sizeof(float)));
        s->pack_cap = (size_t)total * 2;
      }
      if (total > 0) {
        k_sim_pack<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(d_dense, c->tilevol, c->tx, c->ty, me->d_boxes,
                                                                            (int)me->boxes.size(), me->tmpl_floats, rb, nt, total,
                                                                            s->d_pack);
        CK(cudaGetLastError());
        c->launches++;
      }
      return launch_direct(c, me, image, s->d_pack, d_b2, nt, rb, d_out, kind, job0);
    };
    k_sim_templates<double><<<dim3((unsigned)c->tz, (unsigned)NJ), 256, 0, c->stream>>>(
        s->d_grid, s->padvol, s->pad[0], s->pad[1], dt, R, s->d_shapes, nullptr, s->d_a2list, c->tx, c->ty, c->tz, s->d_tmpl,
        s->d_plane, s->d_b2, s->d_ticket);
    CK(cudaGetLastError());
    c->launches += 1;
    bool done = false;
    if (want_fft(c, ebig, NJ)) {
      rc = ensure_fft(c, -1);
      if (rc == IQ_OK) {
        // templates are copies of training-image voxels (or zeros): integer-valued whenever the image is.  The FFT
        // passes do not depend on the mask: one call serves every shape of the launch (A2 map per template).
        rc = launch_fft(c, e, -1, s->d_tmpl, s->d_b2, NJ, c->image_is_int[-1], c->d_Dovl, 0, 0, one_shape ? nullptr : s->d_a2list);
        if (rc) return rc;
        done = true;
      } else if (!(rc == IQ_ERR_STATE && c->fft_failed)) {
        return rc;
      }
    }
    if (!done) {
      // the direct kernel is specialised on the mask (its box decomposition): one launch per run of equal shapes
      for (int k0 = 0; k0 < ntile;) {
        int k1 = k0 + 1;
        while (k1 < ntile && shapes[k1] == shapes[k0]) ++k1;
        const int job0 = k0 * R, nt = (k1 - k0) * R;
        rc = direct(s->shapes[(size_t)shapes[k0]].e, -1, s->d_tmpl + (size_t)job0 * c->tilevol, s->d_b2 + job0, nt,
                    c->d_Dovl + (size_t)job0 * c->npos, 0, job0);
        if (rc) return rc;
        k0 = k1;
      }
    }
    // hard-data distance (iqsim.jl:210-219): the data are the same for every realization -> one map per tile
    if (hardt) {
      k_sim_hardlist<<<ntile, 256, 0, c->stream>>>(s->d_hard_has, s->d_hard_val, s->pad[0], s->pad[1], dt, c->tx,
                                                   c->ty, c->tz, c->nx, c->ny, s->d_hard_ptr, s->d_hard_off, s->d_hard_list);
      CK(cudaGetLastError());
      iq::SparseParams sp{};
      sp.img = c->d_ti;
      sp.nx = c->nx; sp.ny = c->ny; sp.nz = c->nz; sp.nxo = c->nxo; sp.nyo = c->nyo; sp.nzo = c->nzo;
      sp.npos = c->npos;
      sp.ptr = s->d_hard_ptr;
      sp.ptr_stride = 2;
      sp.off = s->d_hard_off;
      sp.val = s->d_hard_list;
      sp.disabled = c->d_disabled;
      sp.out = c->d_Dhard;
      sp.minbits = c->d_minmax + (size_t)(1 * 2 + 0) * c->max_batch;
      sp.maxbits = c->d_minmax + (size_t)(1 * 2 + 1) * c->max_batch;
      sp.R = ntile;
      CK(iq::launch_dist_sparse(sp, c->stream));
      c->launches += 2;
    }
    // soft-data distances (iqsim.jl:222-227): the auxiliary tile is the same for every realization -> one map per tile
    for (int si = 0; si < s->S; ++si) {
      k_sim_templates<float><<<dim3((unsigned)c->tz, (unsigned)ntile), 256, 0, c->stream>>>(
          s->d_aux_pad[si], 0, s->pad[0], s->pad[1], dt, 1, nullptr, c->full_mask->d_mask, nullptr, c->tx, c->ty, c->tz,
          s->d_soft_tmpl, s->d_soft_plane, s->d_soft_b2, s->d_soft_ticket);
      CK(cudaGetLastError());
      c->launches++;
      bool sdone = false;
      if (want_fft(c, c->full_mask, ntile)) {
        rc = ensure_fft(c, si);
        if (rc == IQ_OK) {
          rc = launch_fft(c, c->full_mask, si, s->d_soft_tmpl, s->d_soft_b2, ntile, false, c->d_Dsoft[si], 2 + si);
          if (rc) return rc;
          sdone = true;
        } else if (!(rc == IQ_ERR_STATE && c->fft_failed)) {
          return rc;
        }
      }
      if (!sdone) {
        rc = direct(c->full_mask, si, s->d_soft_tmpl, s->d_soft_b2, ntile, c->d_Dsoft[si], 2 + si, 0);
        if (rc) return rc;
      }
    }
    CK(cudaEventRecord(ev[0], c->stream));
    iq::PickJob* pj = hardt ? s->d_pickjobs_hard : s->d_pickjobs;
    if (s->S == 0 && !hardt) {
      rc = ensure_chunkmin(c, NJ);
      if (rc) return rc;
      CK(iq::launch_pick_chunks(pj, NJ, c->npos, c->chunk_len, c->chunk_n, c->stream));
      c->launches += 1;
    } else {
      // sources shared by all realizations: the hard map (source 0 of a hard tile) and every soft map
      const int nsrc = 1 + (hardt ? 1 : 0) + s->S;
      unsigned shared_mask = 0;
      for (int i = 0; i < nsrc; ++i)
        if (hardt ? (i != 1) : (i != 0)) shared_mask |= 1u << i;
      // all-zero test of the primary map (relaxation.jl:11): per job (overlap map) or per tile (hard map)
      const unsigned* pmax = hardt ? c->d_minmax + (size_t)(1 * 2 + 1) * c->max_batch : c->d_minmax + (size_t)c->max_batch;
      const int pmax_stride = hardt ? 0 : 1, pmax_tile_stride = hardt ? 1 : 0;
      // relaxation rounds on the device: kRelaxRounds rounds are enqueued unconditionally, the kernels of a round
      // return at once for realizations that already have candidates; a realization still empty after the last
      // round raises the status word (the caller then reruns host-staged, where the round count is unbounded)
      // (5 rounds normally, frac up to 0.4 + 0.1 tol; tiles with hard data or an empty overlap mask often need many -- an
      // all-zero overlap map selects an index prefix -- and get all 11, after which frac = 1 guarantees a non-empty
      // intersection)
      const int kRelaxRounds = (hardt || nempty > 0) ? 11 : 5;
      for (int round = 0; round < kRelaxRounds; ++round) {
        k_sim_seljobs<<<NJ, 32, 0, c->stream>>>(c->d_sel, pj, c->max_src, pmax, pmax_stride, pmax_tile_stride, R, c->d_minmax,
                                               c->max_batch, hardt ? 1 : 0, shared_mask, s->tol, c->nenabled, c->npos, round,
                                               s->d_pending, c->d_selbuf, c->sel_cap);
        CK(cudaGetLastError());
        {
          int nl = 0;
          CK(iq::launch_select_all(c->d_sel, NJ * c->max_src, c->npos, c->d_shifts, c->nshift, c->stream, &nl, c->d_sel_list));
          c->launches += nl;
        }
        k_sim_copykth<<<NJ, 32, 0, c->stream>>>(c->d_sel, c->max_src, R);
        CK(cudaGetLastError());
        CK(iq::launch_pick_count(pj, NJ, c->npos, c->stream));
        k_sim_relax_check<<<gJ, 128, 0, c->stream>>>(pj, NJ, s->d_pending, round == kRelaxRounds - 1 ? 1 : 0, s->d_status);
        CK(cudaGetLastError());
        c->launches += 4;
      }
      CK(iq::launch_pick_write(pj, NJ, c->npos, c->stream));
      c->launches += 1;
    }
    CK(iq::launch_tau(pj, NJ, c->max_src, c->d_rank, c->d_colsum, c->d_prob, c->stream));
    k_sim_sample<<<(unsigned)((NJ + 3) / 4), 128, 0, c->stream>>>(pj, c->d_prob, s->d_u, s->npath, dt, R, NJ,
                                                                  s->d_picked, s->d_picks, s->d_status);
    CK(cudaGetLastError());
    c->launches += 3;
  }
  CK(cudaEventRecord(ev[1], c->stream));

  // ---- boundary cuts: one CTA per (job, slab) of an implicit task grid (iq_cutgpu.cu); jobs with fewer slabs than the
  //      launch's maximum leave their extra slots empty ----
  if (maxn > 0) {
    k_sim_slabs<<<NJ * maxn, 256, 0, c->stream>>>(s->d_grid, s->padvol, s->pad[0], s->pad[1], dt, R, s->d_shapes, maxn,
                                                  s->maxslabs, s->d_ti64, c->nx, c->ny, c->nxo, c->nyo, s->d_picked, c->tx, c->ty,
                                                  s->d_cutA, s->d_cutB, s->d_cutdims, (long long)s->maxslab);
    CK(cudaGetLastError());
    CK(iq::launch_graphcut_grid(s->d_cutA, s->d_cutB, s->d_keep, s->d_cut_iters, s->d_cutdims, (long long)s->maxslab,
                                s->maxslabs, maxn, NJ, std::max<size_t>(smem, 64), s->exact, c->stream));
    c->launches += 2;
  }
  CK(cudaEventRecord(ev[2], c->stream));
  k_sim_paste<<<dim3((unsigned)((c->tilevol + 255) / 256), (unsigned)NJ), 256, 0, c->stream>>>(
      s->d_grid, s->d_cutgrid, s->padvol, s->pad[0], s->pad[1], dt, R, s->d_shapes, s->maxslabs, s->d_ti64, c->nx, c->ny,
      c->nxo, c->nyo, s->d_picked, c->tx, c->ty, c->tz, s->d_keep, (long long)s->maxslab, s->d_cut_iters, s->d_status);
  CK(cudaGetLastError());
  c->launches++;
  c->last_launches = c->launches - l0;
  return IQ_OK;
}

int32_t iq_sim_define_shape(iq_ctx* c, const uint8_t* ovlmask, const iq_sim_slab* slabs, int32_t nslab, int32_t* shape) {
  if (!c || !c->sim) return fail(IQ_ERR_STATE, "iq_sim_define_shape: no simulation open on this context");
  if (!ovlmask || !shape) return fail(IQ_ERR_INVALID, "iq_sim_define_shape: NULL argument");
  CK(cudaSetDevice(c->device));
  int slot = -1;
  const int rc = sim_shape(c, ovlmask, slabs, nslab, &slot);
  if (rc) return rc;
  *shape = slot;
  return IQ_OK;
}

int32_t iq_sim_step(iq_ctx* c, int64_t step, const int64_t* start, const uint8_t* ovlmask, const iq_sim_slab* slabs,
                    int32_t nslab, int32_t hard_tile) {
  if (!c || !c->sim) return fail(IQ_ERR_STATE, "iq_sim_step: no simulation open on this context");
  if (!start || !ovlmask) return fail(IQ_ERR_INVALID, "iq_sim_step: bad argument");
  if (hard_tile && !c->sim->d_hard_has) return fail(IQ_ERR_INVALID, "iq_sim_step: hard_tile on a simulation opened without hard data");
  CK(cudaSetDevice(c->device));
  int32_t slot = -1;
  const int rc = sim_shape(c, ovlmask, slabs, nslab, &slot);
  if (rc) return rc;
  const int64_t st[3] = {start[0], c->ndim > 1 ? start[1] : 0, c->ndim > 2 ? start[2] : 0};
  return sim_step_impl(c, 1, &step, st, &slot, hard_tile);
}

int32_t iq_sim_step_multi(iq_ctx* c, int32_t ntile, const int64_t* steps, const int64_t* starts, const int32_t* shapes,
                          int32_t hard_tiles) {
  if (!c || !c->sim) return fail(IQ_ERR_STATE, "iq_sim_step_multi: no simulation open on this context");
  if (!steps || !starts || !shapes) return fail(IQ_ERR_INVALID, "iq_sim_step_multi: NULL argument");
  if (hard_tiles && !c->sim->d_hard_has) return fail(IQ_ERR_INVALID, "iq_sim_step_multi: hard_tiles on a simulation opened without hard data");
  return sim_step_impl(c, ntile, steps, starts, shapes, hard_tiles);
}

int32_t iq_sim_step_picked(iq_ctx* c, int64_t step, const int64_t* start, const int64_t* picks) {
  if (!c || !c->sim) return fail(IQ_ERR_STATE, "iq_sim_step_picked: no simulation open on this context");
  SimState* s = c->sim;
  if (!start || !picks) return fail(IQ_ERR_INVALID, "iq_sim_step_picked: NULL argument");
  if (step < 0 || step >= s->npath) return fail(IQ_ERR_INVALID, "iq_sim_step_picked: step out of range");
  if (s->tile_cursor + 1 > (size_t)std::max<int64_t>(s->npath, 1))
    return fail(IQ_ERR_STATE, "iq_sim_step_picked: more tiles launched than path steps declared at iq_sim_begin");
  CK(cudaSetDevice(c->device));
  const int t[3] = {c->tx, c->ty, c->tz};
  int st3[3] = {0, 0, 0};
  for (int i = 0; i < c->ndim; ++i) {
    st3[i] = (int)start[i];
    if (st3[i] < 0 || st3[i] + t[i] > s->pad[i]) return fail(IQ_ERR_INVALID, "iq_sim_step_picked: tile outside the padded grid");
  }
  const int R = s->R;
  s->synced = false;
  long long* hp = s->h_pickstage + s->tile_cursor * R;
  for (int r = 0; r < R; ++r) {
    if (picks[r] < 0 || picks[r] >= c->npos) return fail(IQ_ERR_INVALID, "iq_sim_step_picked: pick %d out of range", r);
    hp[r] = picks[r];
  }
  TileInfo* ht = s->h_tiles + s->tile_cursor;
  ht->sx = st3[0]; ht->sy = st3[1]; ht->sz = st3[2]; ht->step = (int)step; ht->shape = 0;
  const TileInfo* dt = s->d_tiles + s->tile_cursor;
  CK(cudaMemcpyAsync(s->d_tiles + s->tile_cursor, ht, sizeof(TileInfo), cudaMemcpyHostToDevice, c->stream));
  s->tile_cursor += 1;
  const unsigned gR = (unsigned)((R + 127) / 128);
  CK(cudaMemcpyAsync(s->d_picked, hp, (size_t)R * sizeof(long long), cudaMemcpyHostToDevice, c->stream));
  k_sim_store_picks<<<gR, 128, 0, c->stream>>>(s->d_picked, s->d_picks, s->npath, dt, R, R);
  CK(cudaGetLastError());
  k_sim_paste<<<dim3((unsigned)((c->tilevol + 255) / 256), (unsigned)R), 256, 0, c->stream>>>(
      s->d_grid, s->d_cutgrid, s->padvol, s->pad[0], s->pad[1], dt, R, nullptr, s->maxslabs, s->d_ti64, c->nx, c->ny,
      c->nxo, c->nyo, s->d_picked, c->tx, c->ty, c->tz, s->d_keep, (long long)s->maxslab, s->d_cut_iters, s->d_status);
  CK(cudaGetLastError());
  c->launches += 2;
  c->last_launches = 2;
  return IQ_OK;
}

int32_t iq_sim_sync(iq_ctx* c, int64_t* picks, int32_t* status) {
  if (!c || !c->sim) return fail(IQ_ERR_STATE, "iq_sim_sync: no simulation open on this context");
  SimState* s = c->sim;
  CK(cudaSetDevice(c->device));
  CK(cudaEventRecord(s->ev_end, c->stream));
  int st = 0;
  CK(cudaMemcpyAsync(&st, s->d_status, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  if (picks && s->npath > 0)
    CK(cudaMemcpyAsync(picks, s->d_picks, (size_t)s->R * s->npath * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (status) *status = st;
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, s->ev_begin, s->ev_end));
  s->total_ms = ms;
  int rc = collect_dist_times(c);
  if (rc) return rc;
  s->dist_ms = c->last_dist_ms;
  s->select_ms = s->cut_ms = 0.0;
  for (size_t i = 0; i + 3 <= s->ev_used; i += 3) {
    float a = 0.f, b = 0.f;
    CK(cudaEventElapsedTime(&a, s->ev[i], s->ev[i + 1]));
    CK(cudaEventElapsedTime(&b, s->ev[i + 1], s->ev[i + 2]));
    s->select_ms += a;
    s->cut_ms += b;
  }
  s->synced = true;
  return IQ_OK;
}

int32_t iq_sim_fetch(iq_ctx* c, int32_t r, int32_t dtype, const int64_t* crop, void* out) {
  if (!c || !c->sim) return fail(IQ_ERR_STATE, "iq_sim_fetch: no simulation open on this context");
  SimState* s = c->sim;
  if (r < 0 || r >= s->R || !crop || !out || (dtype != 0 && dtype != 1)) return fail(IQ_ERR_INVALID, "iq_sim_fetch: bad argument");
  CK(cudaSetDevice(c->device));
  int cr[3] = {1, 1, 1};
  for (int i = 0; i < c->ndim; ++i) {
    cr[i] = (int)crop[i];
    if (cr[i] < 1 || cr[i] > s->pad[i]) return fail(IQ_ERR_INVALID, "iq_sim_fetch: crop outside the padded grid");
  }
  const size_t n = (size_t)cr[0] * cr[1] * cr[2], bytes = n * (dtype == 0 ? sizeof(double) : sizeof(float));
  if (bytes > s->export_cap) {
    CK(cudaStreamSynchronize(c->stream));
    cudaFree(s->d_export);
    s->d_export = nullptr;
    s->export_cap = 0;
    CK(iq::dmalloc(&s->d_export, bytes));
    s->export_cap = bytes;
  }
  const double* g = s->d_grid + (size_t)r * s->padvol;
  const unsigned nb = (unsigned)((n + 255) / 256);
  if (dtype == 0) k_sim_export<double><<<nb, 256, 0, c->stream>>>(g, s->pad[0], s->pad[1], cr[0], cr[1], cr[2], (double*)s->d_export);
  else k_sim_export<float><<<nb, 256, 0, c->stream>>>(g, s->pad[0], s->pad[1], cr[0], cr[1], cr[2], (float*)s->d_export);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, s->d_export, bytes, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return IQ_OK;
}

int32_t iq_sim_fetch_all(iq_ctx* c, int32_t dtype, const int64_t* crop, void* const* out, int32_t nthreads) {
  if (!c || !c->sim) return fail(IQ_ERR_STATE, "iq_sim_fetch_all: no simulation open on this context");
  SimState* s = c->sim;
  if (!crop || !out || (dtype != 0 && dtype != 1)) return fail(IQ_ERR_INVALID, "iq_sim_fetch_all: bad argument");
  for (int r = 0; r < s->R; ++r)
    if (!out[r]) return fail(IQ_ERR_INVALID, "iq_sim_fetch_all: out[%d] is NULL", r);
  CK(cudaSetDevice(c->device));
  int cr[3] = {1, 1, 1};
  for (int i = 0; i < c->ndim; ++i) {
    cr[i] = (int)crop[i];
    if (cr[i] < 1 || cr[i] > s->pad[i]) return fail(IQ_ERR_INVALID, "iq_sim_fetch_all: crop outside the padded grid");
  }
  const size_t n = (size_t)cr[0] * cr[1] * cr[2], bytes = n * (dtype == 0 ? sizeof(double) : sizeof(float));
  // page-locked destinations (iq_host_alloc, cudaHostRegister): device -> host copies go straight into them
  bool pinned = true;
  for (int r = 0; r < s->R && pinned; ++r) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, out[r]) != cudaSuccess) { cudaGetLastError(); pinned = false; break; }
    pinned = at.type == cudaMemoryTypeHost;
  }
  if (bytes > s->export2_cap || (!pinned && !s->h_export[0])) {
    CK(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < 2; ++i) {
      if (s->h_export[i]) cudaFreeHost(s->h_export[i]);
      cudaFree(s->d_export2[i]);
      s->h_export[i] = nullptr;
      s->d_export2[i] = nullptr;
      if (!pinned) CK(cudaMallocHost((void**)&s->h_export[i], bytes));
      CK(iq::dmalloc(&s->d_export2[i], bytes));
      if (!s->ev_export[i]) CK(cudaEventCreateWithFlags(&s->ev_export[i], cudaEventDisableTiming));
    }
    s->export2_cap = bytes;
  }
  const unsigned nb = (unsigned)((n + 255) / 256);
  if (pinned) {
    for (int r = 0; r < s->R; ++r) {
      const int b = r & 1;  // two device buffers: the export of realization r + 1 does not wait for the copy of r - 1 ... r
      const double* g = s->d_grid + (size_t)r * s->padvol;
      if (dtype == 0) k_sim_export<double><<<nb, 256, 0, c->stream>>>(g, s->pad[0], s->pad[1], cr[0], cr[1], cr[2], (double*)s->d_export2[b]);
      else k_sim_export<float><<<nb, 256, 0, c->stream>>>(g, s->pad[0], s->pad[1], cr[0], cr[1], cr[2], (float*)s->d_export2[b]);
      CK(cudaGetLastError());
      CK(cudaMemcpyAsync(out[r], s->d_export2[b], bytes, cudaMemcpyDeviceToHost, c->stream));
    }
    CK(cudaStreamSynchronize(c->stream));
    return IQ_OK;
  }
  auto enqueue = [&](int r) -> int {
    const int b = r & 1;
    const double* g = s->d_grid + (size_t)r * s->padvol;
    if (dtype == 0) k_sim_export<double><<<nb, 256, 0, c->stream>>>(g, s->pad[0], s->pad[1], cr[0], cr[1], cr[2], (double*)s->d_export2[b]);
    else k_sim_export<float><<<nb, 256, 0, c->stream>>>(g, s->pad[0], s->pad[1], cr[0], cr[1], cr[2], (float*)s->d_export2[b]);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(s->h_export[b], s->d_export2[b], bytes, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaEventRecord(s->ev_export[b], c->stream));
    return IQ_OK;
  };
  // realization r+1 is exported and copied into the other pinned buffer while the host moves realization r
  const int nt = std::max(1, std::min(nthreads > 0 ? nthreads : 4, 16));
  int rc = enqueue(0);
  if (rc) return rc;
  for (int r = 0; r < s->R; ++r) {
    if (r + 1 < s->R) {
      rc = enqueue(r + 1);
      if (rc) return rc;
    }
    CK(cudaEventSynchronize(s->ev_export[r & 1]));
    const char* src = s->h_export[r & 1];
    char* dst = (char*)out[r];
    const size_t chunk = (bytes + nt - 1) / nt;
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int k = 0; k < nt; ++k) {
      const size_t o = (size_t)k * chunk;
      if (o < bytes) std::memcpy(dst + o, src + o, std::min(chunk, bytes - o));
    }
  }
  return IQ_OK;
}

int32_t iq_sim_fetch_cut(iq_ctx* c, int32_t r, uint8_t* out) {
  if (!c || !c->sim) return fail(IQ_ERR_STATE, "iq_sim_fetch_cut: no simulation open on this context");
  SimState* s = c->sim;
  if (r < 0 || r >= s->R || !out) return fail(IQ_ERR_INVALID, "iq_sim_fetch_cut: bad argument");
  if (!s->d_cutgrid) return fail(IQ_ERR_STATE, "iq_sim_fetch_cut: simulation was not opened with debug");
  CK(cudaSetDevice(c->device));
  CK(cudaMemcpyAsync(out, s->d_cutgrid + (size_t)r * s->padvol, (size_t)s->padvol, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return IQ_OK;
}

int32_t iq_sim_times(const iq_ctx* c, double* total_ms, double* dist_ms, double* select_ms, double* cut_ms) {
  if (!c || !c->sim) return fail(IQ_ERR_STATE, "iq_sim_times: no simulation open on this context");
  if (!c->sim->synced) return fail(IQ_ERR_STATE, "iq_sim_times: call iq_sim_sync first");
  if (total_ms) *total_ms = c->sim->total_ms;
  if (dist_ms) *dist_ms = c->sim->dist_ms;
  if (select_ms) *select_ms = c->sim->select_ms;
  if (cut_ms) *cut_ms = c->sim->cut_ms;
  return IQ_OK;
}

int32_t iq_sim_end(iq_ctx* c) {
  if (!c) return fail(IQ_ERR_INVALID, "NULL context");
  if (!c->sim) return IQ_OK;
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  sim_destroy(c);
  return IQ_OK;
}

}  // extern "C"

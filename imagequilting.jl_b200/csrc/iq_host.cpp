// iq_host.cpp -- native host driver above the C ABI (include/iqb200_host.h).
//
// Restates, in C++, the host side of the reference that north_star keeps on the host:
//   the tile loop            /root/reference/src/iqsim.jl:163-312
//   the boundary cut         /root/reference/src/graphcut.jl:5-84  (lattice max-flow: iq_cut.cpp)
//   the paste                /root/reference/src/iqsim.jl:278
// It reaches the GPU only through the public entry points of include/iqb200.h.
//
// Scheduling.  The simulation path is shared by all realizations (iqsim.jl:139) and a realization only
// depends on its own previous tiles, so realizations advance in lockstep: one iq_search_pick call carries
// the templates of a whole group.  The realizations are split into two groups with a context (stream)
// each; while the GPU searches step k of one group, a pool of host threads cuts and pastes step k of the
// other group(s), so the host cut hides behind the device search (or vice versa).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/iqb200.h"
#include "../../include/iqb200_host.h"
#include "iq_cut.h"

extern "C" void iq_post_error(const char* msg);  // iq_ctx.cu: sets the calling thread's iq_last_error() string

namespace {

int fail_exact_range() {
  iq_post_error("iqh_graphcut_mode: the capacities of this slab span more than 128 bits; no exact cut");
  return IQ_ERR_INVALID;
}

using clk = std::chrono::steady_clock;
double ms_since(clk::time_point t0) { return std::chrono::duration<double, std::milli>(clk::now() - t0).count(); }

struct Slab {
  int d;       // dimension of the overlap
  bool prev;   // overlap with the previous (true) or next (false) tile
  int lo[3], sz[3];
};

// Persistent pool of host threads that executes cut / paste tasks of every group from one shared queue:
// tasks of different groups interleave, so the cores stay busy even when a group has fewer tasks than
// there are threads.
class Pool {
 public:
  explicit Pool(int n) {
    for (int i = 0; i < n; ++i) th_.emplace_back([this, i] { loop(i); });
  }
  ~Pool() {
    {
      std::lock_guard<std::mutex> l(m_);
      quit_ = true;
    }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  void push(std::function<void(int)> f) {
    {
      std::lock_guard<std::mutex> l(m_);
      q_.push_back(std::move(f));
    }
    cv_.notify_one();
  }
  int size() const { return (int)th_.size(); }

 private:
  void loop(int tid) {
    for (;;) {
      std::function<void(int)> f;
      {
        std::unique_lock<std::mutex> l(m_);
        cv_.wait(l, [this] { return quit_ || !q_.empty(); });
        if (q_.empty()) return;  // quit requested and nothing left
        f = std::move(q_.front());
        q_.pop_front();
      }
      f(tid);
    }
  }
  std::mutex m_;
  std::condition_variable cv_;
  std::deque<std::function<void(int)>> q_;
  bool quit_ = false;
  std::vector<std::thread> th_;
};

// Completion latch of one group's cut+paste job.
struct Latch {
  std::mutex m;
  std::condition_variable cv;
  bool busy = false;
  void begin() {
    std::lock_guard<std::mutex> l(m);
    busy = true;
  }
  void finish() {
    {
      std::lock_guard<std::mutex> l(m);
      busy = false;
    }
    cv.notify_all();
  }
  void wait() {
    std::unique_lock<std::mutex> l(m);
    cv.wait(l, [this] { return !busy; });
  }
};

struct Geo {
  int N;
  int n[3], t[3], ov[3], nt[3], pad[3], sp[3], dist[3];
  long long tilevol, padvol, ntile_total;
};

// A group of realizations that share one context and advance together.
struct Group {
  iq_ctx* ctx = nullptr;
  int r0 = 0, R = 0;
  std::vector<uint8_t> pasted, ovlmask;
  std::vector<Slab> slabs;
  int start[3] = {0, 0, 0};
  std::vector<float> simdev;
  std::vector<std::vector<float>> softdev;
  std::vector<const float*> softptr;
  std::vector<int32_t> hoff;
  std::vector<float> hval;
  std::vector<iq_tile> tiles;
  std::vector<iq_result> results;
  std::vector<int64_t> picked;
  std::vector<double> ustep;
  std::vector<std::vector<uint8_t>> keepbuf;
  std::vector<std::vector<double>> slabA, slabB;  // device cut: gathered slabs
  std::vector<iq_cut_task> cut_tasks;
  std::unique_ptr<Latch> latch{new Latch()};
  std::atomic<int> cuts_left{0}, pastes_left{0};
  clk::time_point cut_t0;
  double search_ms = 0, search_dev_ms = 0, cut_ms = 0, dist_ms = 0, fft_bytes = 0, fft_ms = 0;
  int64_t launches = 0, dist_launches = 0, ncand = 0, nfft = 0, ndirect = 0, maxcand = 0;
};

}  // namespace

extern "C" int32_t iqh_graphcut(const double* A, const double* B, int32_t ndim, const int64_t* sz64, int32_t dim, uint8_t* keep) {
  if (!A || !B || !sz64 || !keep || ndim < 1 || ndim > 3 || dim < 0 || dim >= ndim) return IQ_ERR_INVALID;
  int sz[3] = {1, 1, 1};
  for (int i = 0; i < ndim; ++i) sz[i] = (int)sz64[i];
  if (sz[dim] < 2) return IQ_ERR_INVALID;
  iqcut::Work w;
  // integer-valued slabs: exact integer arithmetic (one well-defined cut), FP64 otherwise
  if (!(iqcut::integer_valued(A, B, sz[0] * sz[1] * sz[2]) && iqcut::graphcut_exact(A, B, sz, dim, keep, w)))
    iqcut::graphcut(A, B, sz, dim, keep, w);
  return IQ_OK;
}

extern "C" int32_t iqh_graphcut_mode(const double* A, const double* B, int32_t ndim, const int64_t* sz64, int32_t dim, int32_t exact,
                                     uint8_t* keep) {
  if (!A || !B || !sz64 || !keep || ndim < 1 || ndim > 3 || dim < 0 || dim >= ndim) return IQ_ERR_INVALID;
  int sz[3] = {1, 1, 1};
  for (int i = 0; i < ndim; ++i) sz[i] = (int)sz64[i];
  if (sz[dim] < 2) return IQ_ERR_INVALID;
  iqcut::Work w;
  if (exact) {
    if (!iqcut::graphcut_exact(A, B, sz, dim, keep, w)) return fail_exact_range();
  } else {
    iqcut::graphcut(A, B, sz, dim, keep, w);
  }
  return IQ_OK;
}


namespace {

int fail_invalid(const char* msg) {
  iq_post_error(msg);
  return IQ_ERR_INVALID;
}

// Overlap slabs of tile `ind` with its already pasted neighbours (iqsim.jl:188-205) and the overlap mask.
void tile_slabs(const Geo& G, const std::vector<uint8_t>& pasted, int64_t ind, std::vector<Slab>& slabs, std::vector<uint8_t>& mask,
                int start[3]) {
  const int* t = G.t;
  const int ti3[3] = {(int)(ind % G.nt[0]), (int)((ind / G.nt[0]) % G.nt[1]), (int)(ind / ((long long)G.nt[0] * G.nt[1]))};
  for (int i = 0; i < 3; ++i) start[i] = ti3[i] * G.sp[i];
  const long long tstride[3] = {1, G.nt[0], (long long)G.nt[0] * G.nt[1]};
  slabs.clear();
  for (int d = 0; d < G.N; ++d) {
    if (G.ov[d] <= 1) continue;
    if (ti3[d] > 0 && pasted[(size_t)(ind - tstride[d])]) {
      Slab s{d, true, {0, 0, 0}, {t[0], t[1], t[2]}};
      s.sz[d] = G.ov[d];
      slabs.push_back(s);
    }
    if (ti3[d] + 1 < G.nt[d] && pasted[(size_t)(ind + tstride[d])]) {
      Slab s{d, false, {0, 0, 0}, {t[0], t[1], t[2]}};
      s.lo[d] = G.sp[d];
      s.sz[d] = t[d] - G.sp[d];
      slabs.push_back(s);
    }
  }
  std::fill(mask.begin(), mask.end(), 0);
  for (const Slab& s : slabs)
    for (int z = s.lo[2]; z < s.lo[2] + s.sz[2]; ++z)
      for (int y = s.lo[1]; y < s.lo[1] + s.sz[1]; ++y)
        std::memset(&mask[((size_t)z * t[1] + y) * t[0] + s.lo[0]], 1, (size_t)s.sz[0]);
}

// Integer-valued (categorical) training image?  Its cut capacities (graphcut.jl:52) are degenerate (division by eps next
// to O(1) terms): equal-cost cuts abound and which one comes out depends on the max-flow algorithm's rounding.
bool image_is_integer(const iqh_desc* D, const Geo& G) {
  const long long nvox = (long long)G.n[0] * G.n[1] * G.n[2];
  for (long long i = 0; i < nvox; ++i)
    if (D->ti_f32[i] != std::nearbyint(D->ti_f32[i])) return false;
  return true;
}

// Contexts kept between iqh_run calls.  Building a context costs tens of milliseconds (summed-volume tables, pinned
// staging, image spectra, work buffers) and destroying it more (pinned frees), which is a large share of a call on a
// small shard (8 realizations per GPU).  A finished simulation parks its contexts here; the next call takes one back
// if iq_ctx_matches says a fresh context would be identical (same geometry and job slots, bitwise the same images --
// they are uploaded and compared every time, so the caller may change its arrays freely).  iqh_cache_clear empties it.
struct CtxCache {
  std::mutex m;
  std::vector<iq_ctx*> parked;
  static constexpr size_t kMax = 4;
  // `ref`: a context of this call already known to hold cd's images (the other parked contexts are then compared with it
  // on the device instead of uploading the images again)
  iq_ctx* take(const iq_ctx_desc& cd, const iq_ctx* ref = nullptr) {
    std::vector<iq_ctx*> cand;
    {
      std::lock_guard<std::mutex> l(m);
      cand.swap(parked);
    }
    iq_ctx* hit = nullptr;
    std::vector<iq_ctx*> rest;
    for (iq_ctx* c : cand) {
      int32_t same = 0;
      if (!hit && (ref ? iq_ctx_matches_ctx(c, &cd, ref, &same) : iq_ctx_matches(c, &cd, &same)) == IQ_OK && same) hit = c;
      else rest.push_back(c);
    }
    std::lock_guard<std::mutex> l(m);
    for (iq_ctx* c : rest) parked.push_back(c);
    return hit;
  }
  void park(iq_ctx* c) {
    iq_ctx* evict = nullptr;
    {
      std::lock_guard<std::mutex> l(m);
      parked.push_back(c);
      if (parked.size() > kMax) { evict = parked.front(); parked.erase(parked.begin()); }
    }
    if (evict) iq_ctx_destroy(evict);
  }
  void clear() {
    std::vector<iq_ctx*> all;
    {
      std::lock_guard<std::mutex> l(m);
      all.swap(parked);
    }
    for (iq_ctx* c : all) iq_ctx_destroy(c);
  }
};
CtxCache g_cache;
bool cache_enabled() {
  const char* ev = std::getenv("IQB200_CTX_CACHE");
  return !(ev && ev[0] == '0');
}

// Dependency levels of a simulation path.  A tile reads (template, slabs) and writes (paste) only its own window, so it
// depends exactly on the tiles EARLIER IN THE PATH whose windows intersect its own: level = 1 + max level of those (0
// without any).  Tiles of one level are mutually independent (two tiles with intersecting windows are ordered by the
// path, hence on different levels), and running the levels in order -- each as one batch -- gives bit for bit the
// result of the sequential path.  On a raster path the level of tile (i, j, k) is i + 2j + 4k (3-D) / i + 2j (2-D): 50
// levels instead of 512 sequential steps on an 8 x 8 x 8 grid.  Returns the number of levels, -1 on a bad path.
int dependency_levels(const Geo& G, const int64_t* path, int64_t npath, std::vector<int>& level) {
  level.assign((size_t)npath, 0);
  std::vector<int64_t> step_of((size_t)G.ntile_total, -1);
  const long long tstride[3] = {1, G.nt[0], (long long)G.nt[0] * G.nt[1]};
  int reach[3];  // windows of tiles up to `reach` indices apart along a dimension intersect
  for (int d = 0; d < 3; ++d) reach[d] = G.sp[d] > 0 ? std::min(G.nt[d] - 1, (G.t[d] + G.sp[d] - 1) / G.sp[d] - 1) : G.nt[d] - 1;
  int nlevels = 0;
  for (int64_t step = 0; step < npath; ++step) {
    const int64_t ind = path[step];
    if (ind < 0 || ind >= G.ntile_total || step_of[(size_t)ind] >= 0) return -1;
    const int ti3[3] = {(int)(ind % G.nt[0]), (int)((ind / G.nt[0]) % G.nt[1]), (int)(ind / ((long long)G.nt[0] * G.nt[1]))};
    int lv = 0;
    for (int dz = -reach[2]; dz <= reach[2]; ++dz)
      for (int dy = -reach[1]; dy <= reach[1]; ++dy)
        for (int dx = -reach[0]; dx <= reach[0]; ++dx) {
          const int q[3] = {ti3[0] + dx, ti3[1] + dy, ti3[2] + dz};
          if (q[0] < 0 || q[1] < 0 || q[2] < 0 || q[0] >= G.nt[0] || q[1] >= G.nt[1] || q[2] >= G.nt[2]) continue;
          const int64_t so = step_of[(size_t)(q[0] * tstride[0] + q[1] * tstride[1] + q[2] * tstride[2])];
          if (so < 0) continue;  // not visited yet (or skipped): no dependency
          // windows [i * sp, i * sp + t) intersect along every dimension?
          const bool apart = std::abs(dx) * G.sp[0] >= G.t[0] || std::abs(dy) * G.sp[1] >= G.t[1] || std::abs(dz) * G.sp[2] >= G.t[2];
          if (!apart) lv = std::max(lv, level[(size_t)so] + 1);
        }
    level[(size_t)step] = lv;
    nlevels = std::max(nlevels, lv + 1);
    step_of[(size_t)ind] = step;
  }
  return nlevels;
}

// Device-resident pipeline: every lockstep group owns a context whose stream carries the whole simulation of its
// realizations (iq_sim_*); this thread only enqueues.  Returns IQ_ERR_STATE when the simulation does not qualify
// (the caller then runs host-staged), `*status` != 0 when a data-dependent condition invalidated the result.
int run_resident(const iqh_desc* D, const Geo& G, double* out_grids, uint8_t* out_cuts, int64_t* out_picks, iqh_stats* stats,
                 int* status) {
  *status = 0;
  const int S = D->nsoft;
  // Integer-valued (categorical) images make the cut capacities of graphcut.jl:52 degenerate (division by eps next
  // to O(1) terms): equal-cost cuts abound and which one an FP64 max-flow returns depends on its rounding.  Their cuts
  // therefore run in exact integer arithmetic -- on the device here, on the host in the staged pipeline -- which has
  // one well-defined answer; continuous images have a unique minimum cut with a margin and keep FP64.
  const bool exact_cut = image_is_integer(D, G);
  const auto t_start = clk::now();
  const int R = D->nreal;
  // one lockstep group by default: a launch already carries every realization (FFT pairs, one cut CTA per slab).
  // Several groups (streams) overlap one group's cut tail with another's FFT passes: -10 % device time with 4 groups
  // of 16 on config 5, at the price of per-kernel timings that no longer describe a kernel alone.
  // Few realizations (what one rank of a strong-scaling run over nreal = 64 holds): a level's cut launch then lasts as
  // long as its slowest cut on a mostly idle GPU, and four groups of two realizations fill that tail with each other's
  // FFT passes (measured on config 5 with 8 realizations: 219 -> 189 ms per call, with 4: 134 -> 119, with 32: 730 -> 666,
  // with 64: 1398 -> 1335; with 16 the device time drops too (353 -> 329 ms) but the call does not (370 -> 392); a loss
  // with 8 groups, and with exactly 2).  The default stays at one group from 9 realizations up: the 5-9 % at 32 / 64 are
  // left to `ngroups = 4`, because kernels of concurrent groups time-slice the SMs and their CUDA-event durations (the
  // roofline figures of bench.py) then no longer describe a kernel running alone.  2-D simulations are launch-latency bound and gain likewise (config 2 with
  // 16 realizations: 26.9 -> 21.6 ms).  Threshold path only: with soft or hard data four groups LOSE 6-11 % (configs
  // 3 and 4: the selection kernels of a group already fill the GPU).
  int ngroups = 1;
  if (D->ngroups > 0) ngroups = D->ngroups;
  else if (const char* ev = std::getenv("IQB200_GROUPS")) ngroups = std::max(1, std::atoi(ev));
  else if (S == 0 && !D->hard_has && ((G.N == 3 && R >= 4 && R <= 8) || (G.N == 2 && R >= 8 && R <= 16))) ngroups = 4;
  ngroups = std::max(1, std::min(ngroups, R));
  // Tiles per launch (dependency-level batching, see below): a launch carries tiles x realizations jobs.  Soft data keep
  // one tile per launch (one auxiliary map per source is kept).  IQB200_JOBS overrides the job slots per launch.
  const int Rg = (R + ngroups - 1) / ngroups;
  int tiles_per_launch = 1;
  {
    int jobs_cap = 256;
    if (const char* ev = std::getenv("IQB200_JOBS")) jobs_cap = std::max(1, std::atoi(ev));
    tiles_per_launch = (int)std::max<long long>(1, std::min<long long>(jobs_cap / std::max(Rg, 1), G.ntile_total));
    // device memory: every job slot holds an overlap map, a candidate list with values, radix-select scratch, its
    // share of the FFT work space and its cut slabs
    const double npos = (double)G.dist[0] * G.dist[1] * G.dist[2];
    auto p2 = [](int n) { double v = 1; while (v < n) v *= 2; return v; };
    const double fftws = 8.0 * p2(G.n[0]) * ((double)G.t[2] * G.t[1] + (G.n[2] > 1 ? (double)G.t[2] * p2(G.n[1]) + 2.0 * G.dist[2] * p2(G.n[1]) : 0.0) +
                                              (double)G.dist[2] * G.dist[1]) / 2.0;
    const double per_job = npos * 4.0 * (6.0 + 2.0 * S) + npos / 8.0 * 8.0 * (2.0 + S) + fftws + 6.0 * 17.0 * G.tilevol +
                           (2.0 + S) * 32768 * 12.0;
    size_t free_b = 0, total_b = 0;
    if (iq_device_free_memory(D->device, &free_b, &total_b) == IQ_OK) {
      const double budget = 0.6 * (double)free_b - (double)R * 8.0 * G.padvol;
      while (tiles_per_launch > 1 && (double)tiles_per_launch * Rg * ngroups * per_job > budget) --tiles_per_launch;
    }
  }
  struct RG { iq_ctx* ctx = nullptr; int r0 = 0, R = 0; };
  std::vector<RG> groups(ngroups);
  const bool use_cache = cache_enabled();
  auto destroy_all = [&] {
    for (auto& g : groups)
      if (g.ctx) { iq_ctx_destroy(g.ctx); g.ctx = nullptr; }
  };
  // after a successful simulation the contexts are parked for the next call instead of destroyed
  auto park_all = [&] {
    for (auto& g : groups)
      if (g.ctx) {
        if (use_cache && iq_sim_end(g.ctx) == IQ_OK) g_cache.park(g.ctx);
        else iq_ctx_destroy(g.ctx);
        g.ctx = nullptr;
      }
  };
  const auto t_setup = clk::now();
  int rc = IQ_OK;
  for (int gi = 0; gi < ngroups && rc == IQ_OK; ++gi) {
    RG& g = groups[gi];
    g.r0 = (int)((long long)R * gi / ngroups);
    g.R = (int)((long long)R * (gi + 1) / ngroups) - g.r0;
    iq_ctx_desc cd{};
    cd.ndim = G.N;
    for (int i = 0; i < 3; ++i) { cd.ti_size[i] = G.n[i]; cd.tile_size[i] = G.t[i]; }
    cd.ti = D->ti_f32;
    cd.disabled = D->disabled;
    cd.nsoft = S;
    cd.auxti = D->auxti;
    cd.device = D->device;
    cd.max_batch = g.R * tiles_per_launch;
    g.ctx = use_cache ? g_cache.take(cd, gi > 0 ? groups[0].ctx : nullptr) : nullptr;
    if (!g.ctx) rc = iq_ctx_create(&g.ctx, &cd);
    if (rc != IQ_OK) break;
    iq_ctx_set_option(g.ctx, "fft", D->fft_mode);
    iq_sim_desc sd{};
    for (int i = 0; i < 3; ++i) { sd.pad_size[i] = G.pad[i]; sd.ovl_size[i] = G.ov[i]; }
    sd.nreal = g.R;
    sd.ti64 = D->ti;
    sd.u = D->u + (size_t)g.r0 * D->npath;
    sd.npath = D->npath;
    sd.tol = D->tol;
    sd.debug = D->debug;
    sd.aux = S ? D->aux : nullptr;
    sd.hard_has = D->hard_has;
    sd.hard_val = D->hard_has ? D->hard_val : nullptr;
    sd.exact_cut = exact_cut ? 1 : 0;
    rc = iq_sim_begin(g.ctx, &sd);
  }
  if (rc != IQ_OK) { destroy_all(); return rc; }
  const double setup_ms = ms_since(t_setup);

  // ---- what depends on the path alone, for every step (iqsim.jl:172-205): tile origin, slabs / overlap mask from the
  //      neighbours pasted EARLIER IN THE PATH, hard-data flag ----
  struct StepInfo {
    int64_t ind;
    int start[3];
    unsigned key;      // bit 2d: overlap with the previous tile along d, bit 2d+1: with the next one (fixes slabs and mask)
    bool hard_tile, host_pick;
    int level;
  };
  std::vector<int> levels;
  const int nlevels = dependency_levels(G, D->path, D->npath, levels);
  if (nlevels < 0) { destroy_all(); return fail_invalid("iqh_run: path holds an out-of-range or repeated tile index"); }
  std::vector<StepInfo> steps((size_t)D->npath);
  std::vector<uint8_t> pasted((size_t)G.ntile_total, 0), mask((size_t)G.tilevol);
  std::vector<Slab> slabs;
  std::vector<iq_sim_slab> sl;
  for (int64_t step = 0; step < D->npath && rc == IQ_OK; ++step) {
    const int64_t ind = D->path[step];
    StepInfo& si = steps[(size_t)step];
    si.ind = ind;
    tile_slabs(G, pasted, ind, slabs, mask, si.start);
    si.key = 0;
    for (const Slab& sb : slabs) si.key |= 1u << (2 * sb.d + (sb.prev ? 0 : 1));
    // does the tile contain hard data?  (indicator!, utils.jl:31-36: the same for every realization)
    si.hard_tile = false;
    if (D->hard_has) {
      for (int z = 0; z < G.t[2] && !si.hard_tile; ++z)
        for (int y = 0; y < G.t[1] && !si.hard_tile; ++y) {
          const uint8_t* row = D->hard_has + ((long long)(si.start[2] + z) * G.pad[1] + (si.start[1] + y)) * G.pad[0] + si.start[0];
          for (int x = 0; x < G.t[0]; ++x)
            if (row[x]) { si.hard_tile = true; break; }
        }
    }
    si.host_pick = S > 0 && slabs.empty() && !si.hard_tile;
    si.level = levels[(size_t)step];
    pasted[(size_t)ind] = 1;
  }
  // launch order: level by level; inside a level the tiles that share slabs / mask are batched (path order otherwise)
  std::vector<int64_t> order((size_t)D->npath);
  for (int64_t i = 0; i < D->npath; ++i) order[(size_t)i] = i;
  const bool batching = tiles_per_launch > 1;
  if (batching)
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
      const StepInfo &x = steps[(size_t)a], &y = steps[(size_t)b];
      if (x.level != y.level) return x.level < y.level;
      if (x.hard_tile != y.hard_tile) return x.hard_tile < y.hard_tile;
      if (x.host_pick != y.host_pick) return x.host_pick < y.host_pick;
      return x.key < y.key;
    });
  std::vector<float> zero_tile((size_t)G.tilevol, 0.f);
  std::vector<std::vector<float>> soft_tile((size_t)S, std::vector<float>((size_t)G.tilevol));
  std::vector<const float*> soft_ptr((size_t)S, nullptr);
  std::vector<int64_t> bsteps, bstarts;
  std::vector<int32_t> bshapes;
  int64_t launches = 0, nlaunch_steps = 0;
  // slabs / mask of an overlap key (bit 2d: previous tile along d pasted, bit 2d+1: next one)
  auto key_shape = [&](unsigned key) {
    slabs.clear();
    for (int d = 0; d < G.N; ++d) {
      if ((key >> (2 * d)) & 1u) {
        Slab sb{d, true, {0, 0, 0}, {G.t[0], G.t[1], G.t[2]}};
        sb.sz[d] = G.ov[d];
        slabs.push_back(sb);
      }
      if ((key >> (2 * d + 1)) & 1u) {
        Slab sb{d, false, {0, 0, 0}, {G.t[0], G.t[1], G.t[2]}};
        sb.lo[d] = G.sp[d];
        sb.sz[d] = G.t[d] - G.sp[d];
        slabs.push_back(sb);
      }
    }
    std::fill(mask.begin(), mask.end(), 0);
    for (const Slab& sb : slabs)
      for (int z = sb.lo[2]; z < sb.lo[2] + sb.sz[2]; ++z)
        for (int y = sb.lo[1]; y < sb.lo[1] + sb.sz[1]; ++y)
          std::memset(&mask[((size_t)z * G.t[1] + y) * G.t[0] + sb.lo[0]], 1, (size_t)sb.sz[0]);
    sl.resize(slabs.size());
    for (size_t k = 0; k < slabs.size(); ++k) {
      sl[k].dim = slabs[k].d;
      sl[k].prev = slabs[k].prev ? 1 : 0;
      for (int i = 0; i < 3; ++i) { sl[k].lo[i] = slabs[k].lo[i]; sl[k].sz[i] = slabs[k].sz[i]; }
    }
  };
  // shape ids per lockstep group and key (registered on first use)
  std::vector<std::vector<int32_t>> shape_of(groups.size(), std::vector<int32_t>(64, -1));
  const auto t_enq = clk::now();
  for (size_t oi = 0; oi < order.size() && rc == IQ_OK;) {
    const StepInfo& si = steps[(size_t)order[oi]];
    const int64_t step = order[oi];
    const int* start = si.start;
    const int64_t st64[3] = {start[0], start[1], start[2]};
    // the batch: the following steps of the same level; tiles with hard data only go with their like (their primary
    // source is the hard distance), and so do tiles without a pasted neighbour (key 0: sampled from the uniform
    // distribution on the host, or -- with soft data -- searched on the host)
    size_t oe = oi + 1;
    if (batching && !si.host_pick)
      while (oe < order.size() && (int)(oe - oi) < tiles_per_launch) {
        const StepInfo& o = steps[(size_t)order[oe]];
        if (o.level != si.level || o.hard_tile != si.hard_tile || o.host_pick || (o.key == 0) != (si.key == 0)) break;
        ++oe;
      }
    const int ntile = (int)(oe - oi);
    if (si.host_pick) {
      // Soft data and nothing pasted around the tile: the candidate set is a tenth of all patterns (relaxation.jl:11,20),
      // far above the device tau model.  One search (the tile is the same for every realization), the sampling walk
      // per realization on the host, and the picks handed to the device (iq_sim_step_picked).
      std::fill(zero_tile.begin(), zero_tile.end(), 0.f);
      for (int sidx = 0; sidx < S; ++sidx) {
        for (int z = 0; z < G.t[2]; ++z)
          for (int y = 0; y < G.t[1]; ++y) {
            const long long gi = ((long long)(start[2] + z) * G.pad[1] + (start[1] + y)) * G.pad[0] + start[0];
            std::memcpy(&soft_tile[sidx][((size_t)z * G.t[1] + y) * G.t[0]], D->aux[sidx] + gi, sizeof(float) * G.t[0]);
          }
        soft_ptr[sidx] = soft_tile[sidx].data();
      }
    }
    for (size_t gi = 0; gi < groups.size() && rc == IQ_OK; ++gi) {
      RG& g = groups[gi];
      if (si.host_pick) {
        key_shape(si.key);
        iq_tile tile{};
        tile.simdev = zero_tile.data();
        tile.softdev = soft_ptr.data();
        iq_result res{};
        rc = iq_search(g.ctx, mask.data(), &tile, 1, D->tol, &res);
        if (rc != IQ_OK) break;
        if (res.count <= 0) { rc = IQ_ERR_STATE; break; }
        std::vector<int64_t> pk((size_t)g.R);
        for (int r = 0; r < g.R && rc == IQ_OK; ++r) {
          int64_t pos = 0;
          rc = iq_sample(res.prob, res.count, D->u[(size_t)(g.r0 + r) * D->npath + step], &pos);
          pk[(size_t)r] = res.idx[pos];
        }
        if (rc == IQ_OK) rc = iq_sim_step_picked(g.ctx, step, st64, pk.data());
      } else if (ntile == 1) {
        key_shape(si.key);
        rc = iq_sim_step(g.ctx, step, st64, mask.data(), sl.data(), (int32_t)sl.size(), si.hard_tile ? 1 : 0);
      } else {
        // (sort order inside a level: key, then hard flag -- see below)
        bsteps.resize((size_t)ntile);
        bstarts.resize((size_t)ntile * 3);
        bshapes.resize((size_t)ntile);
        for (int k = 0; k < ntile && rc == IQ_OK; ++k) {
          const StepInfo& o = steps[(size_t)order[oi + k]];
          bsteps[(size_t)k] = order[oi + k];
          for (int i = 0; i < 3; ++i) bstarts[(size_t)k * 3 + i] = o.start[i];
          int32_t& sid = shape_of[gi][o.key & 63u];
          if (sid < 0) {
            key_shape(o.key);
            rc = iq_sim_define_shape(g.ctx, mask.data(), sl.data(), (int32_t)sl.size(), &sid);
          }
          bshapes[(size_t)k] = sid;
        }
        if (rc == IQ_OK) rc = iq_sim_step_multi(g.ctx, ntile, bsteps.data(), bstarts.data(), bshapes.data(), si.hard_tile ? 1 : 0);
      }
      if (rc != IQ_OK) break;
      double dms = 0;
      int64_t nl = 0;
      iq_last_search_stats(g.ctx, &dms, &nl);
      launches += nl;
    }
    ++nlaunch_steps;
    oi = oe;
  }
  const double enqueue_ms = ms_since(t_enq);
  double device_ms = 0, dist_ms = 0, select_ms = 0, cut_ms = 0, fft_bytes = 0, fft_ms = 0;
  int64_t nfft = 0, ndirect = 0, dist_launches = 0;
  for (auto& g : groups) {
    if (rc != IQ_OK) break;
    int32_t st = 0;
    rc = iq_sim_sync(g.ctx, out_picks ? out_picks + (size_t)g.r0 * D->npath : nullptr, &st);
    if (rc != IQ_OK) break;
    *status |= st;
    double a = 0, b = 0, c2 = 0, d2 = 0;
    iq_sim_times(g.ctx, &a, &b, &c2, &d2);
    device_ms = std::max(device_ms, a);
    dist_ms += b; select_ms += c2; cut_ms += d2;
    int64_t nd = 0, nf = 0;
    double fb = 0, fm = 0;
    iq_last_search_path(g.ctx, &nd, &nf, &fb, &fm);
    ndirect += nd; nfft += nf; fft_bytes += fb; fft_ms += fm;
    double dm = 0;
    int64_t dl = 0;
    iq_last_search_kernel_ms(g.ctx, &dm, &dl);
    dist_launches += dl;
  }
  const double run_ms = ms_since(t_enq);
  const auto t_fetch = clk::now();
  if (rc == IQ_OK && *status == 0) {
    const int64_t padc[3] = {G.pad[0], G.pad[1], G.pad[2]};
    const int hw = std::max(1, (int)std::thread::hardware_concurrency());
    const int nth = std::max(1, std::min(D->nthreads > 0 ? D->nthreads : hw, 8));
    for (auto& g : groups) {
      if (D->out_real) rc = iq_sim_fetch_all(g.ctx, D->out_real_f32 ? 1 : 0, D->sim_size, D->out_real + g.r0, nth);
      for (int r = 0; r < g.R && rc == IQ_OK; ++r) {
        if (rc == IQ_OK && out_grids) rc = iq_sim_fetch(g.ctx, r, 0, padc, out_grids + (size_t)(g.r0 + r) * G.padvol);
        if (rc == IQ_OK && D->debug) rc = iq_sim_fetch_cut(g.ctx, r, out_cuts + (size_t)(g.r0 + r) * G.padvol);
      }
      if (rc != IQ_OK) break;
    }
  }
  const double fetch_ms = ms_since(t_fetch);
  const auto t_down = clk::now();
  if (rc == IQ_OK && *status == 0) park_all();
  else destroy_all();
  const double teardown_ms = ms_since(t_down);
  if (rc != IQ_OK) return rc;
  if (stats && *status == 0) {
    std::memset(stats, 0, sizeof *stats);
    stats->resident = 1;
    stats->search_ms = enqueue_ms;        // wall time this thread spent enqueueing
    stats->search_device_ms = dist_ms;
    stats->cut_ms = 0.0;
    stats->total_ms = ms_since(t_start);
    stats->searches = (int64_t)R * D->npath;
    stats->kernel_launches = launches;
    stats->setup_ms = setup_ms;
    stats->dist_kernel_ms = dist_ms;
    stats->dist_launches = dist_launches;
    stats->fft_searches = nfft;
    stats->direct_searches = ndirect;
    stats->fft_bytes = fft_bytes;
    stats->fft_ms = fft_ms;
    stats->device_ms = device_ms;
    stats->select_ms = select_ms;
    stats->cut_device_ms = cut_ms;
    stats->fetch_ms = fetch_ms;
    stats->teardown_ms = teardown_ms;
    stats->run_ms = run_ms;
    stats->dep_levels = nlevels;
    stats->tiles_per_launch = tiles_per_launch;
    stats->step_launches = nlaunch_steps;
  }
  return IQ_OK;
}

bool is_oom(int rc) {
  if (rc == IQ_ERR_NOMEM) return true;
  if (rc != IQ_ERR_CUDA) return false;
  const std::string m = iq_last_error();
  return m.find("out of memory") != std::string::npos;
}

// Realizations are independent, so a simulation whose resident state would not fit in device memory is run as
// consecutive waves of at most `rmax` realizations (each wave a complete run_resident on its rows of the uniform
// stream).  rmax comes from the free device memory and a per-realization estimate (FP64 grid, distance maps,
// candidate lists, FFT work space, cut slabs); IQB200_MAX_RESIDENT_REAL overrides it (tests).
int run_resident_waves(const iqh_desc* D, const Geo& G, double* out_grids, uint8_t* out_cuts, int64_t* out_picks, iqh_stats* stats,
                       int* status) {
  const int R = D->nreal;
  const double npos = (double)G.dist[0] * G.dist[1] * G.dist[2];
  auto p2 = [](int n) { double v = 1; while (v < n) v *= 2; return v; };
  const double fftws = 8.0 * p2(G.n[0]) * ((double)G.t[2] * G.t[1] + (G.n[2] > 1 ? (double)G.t[2] * p2(G.n[1]) + 2.0 * G.dist[2] * p2(G.n[1]) : 0.0) +
                                            (double)G.dist[2] * G.dist[1]) / 2.0;  // per template (two per transform)
  const double per_real = 8.0 * G.padvol * (D->debug ? 1.125 : 1.0) + npos * 4.0 * (2.0 + 1.0 + (2.0 + D->nsoft)) +
                          npos / 16.0 * 8.0 * (2.0 + D->nsoft) + fftws + 6.0 * 17.0 * G.tilevol;
  size_t free_b = 0, total_b = 0;
  int rmax = R;
  if (iq_device_free_memory(D->device, &free_b, &total_b) == IQ_OK && per_real > 0)
    rmax = (int)std::max(1.0, std::min((double)R, 0.7 * (double)free_b / per_real));
  if (const char* ev = std::getenv("IQB200_MAX_RESIDENT_REAL")) rmax = std::max(1, std::min(R, std::atoi(ev)));
  // The estimate is a heuristic (it ignores the per-context fixed costs and the pinned buffers): a wave that runs out
  // of device memory is retried with half as many realizations instead of failing the call.
  if (rmax >= R) {
    const int rc = run_resident(D, G, out_grids, out_cuts, out_picks, stats, status);
    if (!is_oom(rc) || R < 2) return rc;
    iq_release_device_memory(D->device);
    rmax = (R + 1) / 2;
  }
  iqh_stats acc{};
  *status = 0;
  for (int r0 = 0; r0 < R; r0 += rmax) {
    const int n = std::min(rmax, R - r0);
    iqh_desc sub = *D;
    sub.nreal = n;
    sub.u = D->u + (size_t)r0 * D->npath;
    sub.out_real = D->out_real ? D->out_real + r0 : nullptr;
    iqh_stats st{};
    int wst = 0;
    const int rc = run_resident(&sub, G, out_grids ? out_grids + (size_t)r0 * G.padvol : nullptr,
                                out_cuts ? out_cuts + (size_t)r0 * G.padvol : nullptr,
                                out_picks ? out_picks + (size_t)r0 * D->npath : nullptr, &st, &wst);
    if (is_oom(rc) && rmax > 1) {  // redo this wave (and the rest) with smaller waves
      iq_release_device_memory(D->device);
      rmax = (rmax + 1) / 2;
      r0 -= rmax;  // the loop increment brings r0 back to the start of the failed wave
      continue;
    }
    if (rc != IQ_OK) return rc;
    *status |= wst;
    if (wst) return IQ_OK;  // the caller redoes everything host-staged
    acc.resident = 1;
    acc.search_ms += st.search_ms; acc.search_device_ms += st.search_device_ms; acc.total_ms += st.total_ms;
    acc.searches += st.searches; acc.kernel_launches += st.kernel_launches; acc.setup_ms += st.setup_ms;
    acc.dist_kernel_ms += st.dist_kernel_ms; acc.dist_launches += st.dist_launches; acc.fft_searches += st.fft_searches;
    acc.direct_searches += st.direct_searches; acc.fft_bytes += st.fft_bytes; acc.fft_ms += st.fft_ms;
    acc.device_ms += st.device_ms; acc.select_ms += st.select_ms; acc.cut_device_ms += st.cut_device_ms;
    acc.fetch_ms += st.fetch_ms;
    acc.teardown_ms += st.teardown_ms; acc.run_ms += st.run_ms;
    acc.dep_levels = st.dep_levels; acc.tiles_per_launch = st.tiles_per_launch; acc.step_launches += st.step_launches;
  }
  if (stats) *stats = acc;
  return IQ_OK;
}

}  // namespace

extern "C" int32_t iqh_cache_clear(void) {
  g_cache.clear();
  return IQ_OK;
}

extern "C" int32_t iqh_dependency_levels(int32_t ndim, const int64_t* tile_size, const int64_t* ovl_size, const int64_t* ntiles,
                                         const int64_t* path, int64_t npath, int32_t* levels, int32_t* nlevels) {
  if (ndim < 1 || ndim > 3 || !tile_size || !ovl_size || !ntiles || (npath > 0 && (!path || !levels))) return IQ_ERR_INVALID;
  Geo G{};
  G.N = ndim;
  for (int i = 0; i < 3; ++i) { G.t[i] = G.nt[i] = 1; G.ov[i] = 1; }
  for (int i = 0; i < ndim; ++i) { G.t[i] = (int)tile_size[i]; G.ov[i] = (int)ovl_size[i]; G.nt[i] = (int)ntiles[i]; }
  for (int i = 0; i < 3; ++i) G.sp[i] = G.t[i] - G.ov[i];
  for (int i = ndim; i < 3; ++i) G.sp[i] = 1;
  G.ntile_total = (long long)G.nt[0] * G.nt[1] * G.nt[2];
  std::vector<int> lv;
  const int n = dependency_levels(G, path, npath, lv);
  if (n < 0) return fail_invalid("iqh_dependency_levels: path holds an out-of-range or repeated tile index");
  for (int64_t i = 0; i < npath; ++i) levels[i] = lv[(size_t)i];
  if (nlevels) *nlevels = n;
  return IQ_OK;
}

extern "C" int32_t iqh_run(const iqh_desc* D, double* out_grids, uint8_t* out_cuts, int64_t* out_picks, iqh_stats* stats) {
  if (!D || !D->ti_f32 || !D->path || !D->u) return IQ_ERR_INVALID;
  if (!out_grids && !D->out_real) return IQ_ERR_INVALID;
  if (D->debug && !out_cuts) return IQ_ERR_INVALID;
  const auto t_start = clk::now();
  Geo G{};
  G.N = D->ndim;
  if (G.N < 2 || G.N > 3) return IQ_ERR_INVALID;
  for (int i = 0; i < 3; ++i) { G.n[i] = G.t[i] = G.ov[i] = G.nt[i] = G.pad[i] = 1; }
  for (int i = 0; i < G.N; ++i) {
    G.n[i] = (int)D->ti_size[i]; G.t[i] = (int)D->tile_size[i]; G.ov[i] = (int)D->ovl_size[i];
    G.nt[i] = (int)D->ntiles[i]; G.pad[i] = (int)D->pad_size[i];
  }
  for (int i = 0; i < 3; ++i) { G.sp[i] = G.t[i] - G.ov[i]; G.dist[i] = G.n[i] - G.t[i] + 1; }
  G.tilevol = (long long)G.t[0] * G.t[1] * G.t[2];
  G.padvol = (long long)G.pad[0] * G.pad[1] * G.pad[2];
  G.ntile_total = (long long)G.nt[0] * G.nt[1] * G.nt[2];
  const int R = D->nreal;
  const int S = D->nsoft;
  const int* n = G.n;
  const int* t = G.t;
  const int* pad = G.pad;
  if (D->pipeline < 0 || D->pipeline > 2) return IQ_ERR_INVALID;

  int resident_status = 0;
  if (D->pipeline != 1) {
    const int rcr = run_resident_waves(D, G, out_grids, out_cuts, out_picks, stats, &resident_status);
    if (rcr == IQ_OK && resident_status == 0) return IQ_OK;
    // explicit request or a real error; in auto mode "does not qualify" includes "does not fit in device memory"
    if (rcr != IQ_OK && !((rcr == IQ_ERR_STATE || is_oom(rcr)) && D->pipeline == 0)) return rcr;
    if (rcr != IQ_OK && is_oom(rcr)) iq_release_device_memory(D->device);
    // otherwise: does not qualify (or a data-dependent bail-out): host-staged below, still on the GPU
  }
  // FP64 copy of the training image for the host-side cut and paste (exact: the image is FP32 when desc.ti is NULL)
  std::vector<double> own_ti64;
  const double* ti64 = D->ti;
  if (!ti64) {
    const size_t nimg = (size_t)G.n[0] * G.n[1] * G.n[2];
    own_ti64.resize(nimg);
    for (size_t i = 0; i < nimg; ++i) own_ti64[i] = (double)D->ti_f32[i];
    ti64 = own_ti64.data();
  }
  std::vector<double> own_grids;
  if (!out_grids) {
    own_grids.resize((size_t)G.padvol * R);
    out_grids = own_grids.data();
  }

  // one core stays with the thread that drives the GPU; the rest form the cut/paste team
  const int hw = std::max(1, (int)std::thread::hardware_concurrency());
  // (the driver thread spins in stream synchronisation and the worker thread is itself a team member, so the
  // team may own at most hw - 2 cores or its barriers wait on descheduled members)
  const int nthreads = std::max(1, std::min(D->nthreads > 0 ? D->nthreads : hw, hw - 2));
  // more groups = more searches in flight while cuts run (the per-step critical path is search + slowest
  // cut of ONE group); fewer groups = larger, more efficient search batches
  int ngroups = D->ngroups > 0 ? D->ngroups : (R >= 8 ? 4 : (R >= 2 ? 2 : 1));
  ngroups = std::max(1, std::min(ngroups, R));

  const auto t_setup = clk::now();
  std::vector<Group> groups(ngroups);
  int rc = IQ_OK;
  auto destroy_all = [&] {
    for (auto& g : groups)
      if (g.ctx) { iq_ctx_destroy(g.ctx); g.ctx = nullptr; }
  };
  for (int gi = 0; gi < ngroups; ++gi) {
    Group& g = groups[gi];
    g.r0 = (int)((long long)R * gi / ngroups);
    g.R = (int)((long long)R * (gi + 1) / ngroups) - g.r0;
    iq_ctx_desc cd{};
    cd.ndim = G.N;
    for (int i = 0; i < 3; ++i) { cd.ti_size[i] = n[i]; cd.tile_size[i] = t[i]; }
    cd.ti = D->ti_f32;
    cd.disabled = D->disabled;
    cd.nsoft = S;
    cd.auxti = D->auxti;
    cd.device = D->device;
    cd.max_batch = D->batch > 0 ? std::min(D->batch, g.R) : g.R;
    rc = iq_ctx_create(&g.ctx, &cd);
    if (rc != IQ_OK) { destroy_all(); return rc; }
    iq_ctx_set_option(g.ctx, "fft", D->fft_mode);
    g.pasted.assign((size_t)G.ntile_total, 0);
    g.ovlmask.resize((size_t)G.tilevol);
    g.simdev.resize((size_t)G.tilevol * g.R);
    g.softdev.assign(S, std::vector<float>((size_t)G.tilevol));
    g.softptr.resize(S);
    g.tiles.resize(g.R);
    g.results.resize(g.R);
    g.picked.resize(g.R);
    g.ustep.resize(g.R);
  }
  const double setup_ms = ms_since(t_setup);

  std::memset(out_grids, 0, sizeof(double) * G.padvol * R);  // simgrid = zeros (iqsim.jl:165)
  if (D->debug) std::memset(out_cuts, 0, (size_t)G.padvol * R);
  std::vector<iqcut::Work> work(nthreads);  // one scratch set per pool thread

  // ---- search of one step for one group (iqsim.jl:177-243) ----
  auto do_search = [&](Group& g, int64_t step) -> int {
    const int64_t ind = D->path[step];
    if (ind < 0 || ind >= G.ntile_total) return IQ_ERR_INVALID;
    const int ti3[3] = {(int)(ind % G.nt[0]), (int)((ind / G.nt[0]) % G.nt[1]), (int)(ind / ((long long)G.nt[0] * G.nt[1]))};
    for (int i = 0; i < 3; ++i) g.start[i] = ti3[i] * G.sp[i];
    const int* start = g.start;
    const long long tstride[3] = {1, G.nt[0], (long long)G.nt[0] * G.nt[1]};
    // overlap slabs with pasted neighbours (iqsim.jl:188-205)
    g.slabs.clear();
    for (int d = 0; d < G.N; ++d) {
      if (G.ov[d] <= 1) continue;
      if (ti3[d] > 0 && g.pasted[(size_t)(ind - tstride[d])]) {
        Slab s{d, true, {0, 0, 0}, {t[0], t[1], t[2]}};
        s.sz[d] = G.ov[d];
        g.slabs.push_back(s);
      }
      if (ti3[d] + 1 < G.nt[d] && g.pasted[(size_t)(ind + tstride[d])]) {
        Slab s{d, false, {0, 0, 0}, {t[0], t[1], t[2]}};
        s.lo[d] = G.sp[d];
        s.sz[d] = t[d] - G.sp[d];
        g.slabs.push_back(s);
      }
    }
    std::fill(g.ovlmask.begin(), g.ovlmask.end(), 0);
    for (const Slab& s : g.slabs)
      for (int z = s.lo[2]; z < s.lo[2] + s.sz[2]; ++z)
        for (int y = s.lo[1]; y < s.lo[1] + s.sz[1]; ++y)
          std::memset(&g.ovlmask[((size_t)z * t[1] + y) * t[0] + s.lo[0]], 1, (size_t)s.sz[0]);
    // hard data inside the tile (indicator!/event!, utils.jl:18-36)
    g.hoff.clear();
    g.hval.clear();
    if (D->hard_has) {
      for (int z = 0; z < t[2]; ++z)
        for (int y = 0; y < t[1]; ++y) {
          const long long gi = ((long long)(start[2] + z) * pad[1] + (start[1] + y)) * pad[0] + start[0];
          for (int x = 0; x < t[0]; ++x)
            if (D->hard_has[gi + x]) {
              g.hoff.push_back((int32_t)(((long long)z * t[1] + y) * t[0] + x));
              g.hval.push_back(D->hard_val[gi + x]);
            }
        }
    }
    // soft data events (iqsim.jl:224)
    for (int s = 0; s < S; ++s) {
      for (int z = 0; z < t[2]; ++z)
        for (int y = 0; y < t[1]; ++y) {
          const long long gi = ((long long)(start[2] + z) * pad[1] + (start[1] + y)) * pad[0] + start[0];
          std::memcpy(&g.softdev[s][((size_t)z * t[1] + y) * t[0]], D->aux[s] + gi, sizeof(float) * t[0]);
        }
      g.softptr[s] = g.softdev[s].data();
    }
    // current content of every realization's tile (iqsim.jl:185)
    for (int r = 0; r < g.R; ++r) {
      const double* grid = out_grids + (size_t)(g.r0 + r) * G.padvol;
      float* dst = &g.simdev[(size_t)r * G.tilevol];
      for (int z = 0; z < t[2]; ++z)
        for (int y = 0; y < t[1]; ++y) {
          const double* src = grid + ((long long)(start[2] + z) * pad[1] + (start[1] + y)) * pad[0] + start[0];
          float* drow = dst + ((size_t)z * t[1] + y) * t[0];
          for (int x = 0; x < t[0]; ++x) drow[x] = (float)src[x];
        }
      g.tiles[r].simdev = dst;
      g.tiles[r].hard_nnz = (int32_t)g.hoff.size();
      g.tiles[r].hard_offset = g.hoff.data();
      g.tiles[r].hard_value = g.hval.data();
      g.tiles[r].softdev = S ? g.softptr.data() : nullptr;
      g.ustep[r] = D->u[(size_t)(g.r0 + r) * D->npath + step];
    }
    const auto ts = clk::now();
    const int rcs = iq_search_pick(g.ctx, g.ovlmask.data(), g.tiles.data(), g.R, D->tol, g.ustep.data(), g.results.data());
    if (rcs != IQ_OK) return rcs;
    g.search_ms += ms_since(ts);
    double dms = 0;
    int64_t nl = 0;
    iq_last_search_stats(g.ctx, &dms, &nl);
    g.search_dev_ms += dms;
    g.launches += nl;
    iq_last_search_kernel_ms(g.ctx, &dms, &nl);
    g.dist_ms += dms;
    g.dist_launches += nl;
    {
      int64_t nd = 0, nf = 0;
      double fb = 0, fm = 0;
      iq_last_search_path(g.ctx, &nd, &nf, &fb, &fm);
      g.ndirect += nd; g.nfft += nf; g.fft_bytes += fb; g.fft_ms += fm;
    }
    for (int r = 0; r < g.R; ++r) {
      g.ncand += g.results[r].count;
      g.maxcand = std::max<int64_t>(g.maxcand, g.results[r].count);
      g.picked[r] = g.results[r].picked;
      if (out_picks) out_picks[(size_t)(g.r0 + r) * D->npath + step] = g.picked[r];
      if (g.picked[r] < 0) return IQ_ERR_STATE;
    }
    g.pasted[(size_t)ind] = 1;
    return IQ_OK;
  };

  // ---- boundary cut + paste of one step for one group (iqsim.jl:244-281): one task per (realization,
  //      slab) cut; the last cut to finish enqueues one paste task per realization; the last paste
  //      releases the group's latch ----
  Pool pool(nthreads);
  // integer-valued images: exact integer cuts (see run_resident); iq_cut_batch is FP64, so never chosen for them by default
  const bool exact_cuts = image_is_integer(D, G);
  auto cut_task = [&](Group& g, int task, int tid) {
    iqcut::Work& w = work[tid];
    const int nslab = (int)g.slabs.size();
    const int r = task / nslab;
    const Slab& s = g.slabs[task % nslab];
    const int* start = g.start;
    const double* grid = out_grids + (size_t)(g.r0 + r) * G.padvol;
    const int64_t rind = g.picked[r];
    const int rs[3] = {(int)(rind % G.dist[0]), (int)((rind / G.dist[0]) % G.dist[1]),
                       (int)(rind / ((long long)G.dist[0] * G.dist[1]))};
    const int nv = s.sz[0] * s.sz[1] * s.sz[2];
    w.A.resize(nv);
    w.B.resize(nv);
    g.keepbuf[task].resize(nv);
    int i = 0;
    for (int z = 0; z < s.sz[2]; ++z)
      for (int y = 0; y < s.sz[1]; ++y) {
        const int qy = s.lo[1] + y, qz = s.lo[2] + z;
        const double* ga = grid + ((long long)(start[2] + qz) * pad[1] + (start[1] + qy)) * pad[0] + start[0] + s.lo[0];
        const double* gb = ti64 + ((long long)(rs[2] + qz) * n[1] + (rs[1] + qy)) * n[0] + rs[0] + s.lo[0];
        for (int x = 0; x < s.sz[0]; ++x, ++i) { w.A[i] = ga[x]; w.B[i] = gb[x]; }
      }
    if (!(exact_cuts && iqcut::graphcut_exact(w.A.data(), w.B.data(), s.sz, s.d, g.keepbuf[task].data(), w)))
      iqcut::graphcut(w.A.data(), w.B.data(), s.sz, s.d, g.keepbuf[task].data(), w);
  };
  auto paste_task = [&](Group& g, int r, int tid) {
    iqcut::Work& w = work[tid];
    const int nslab = (int)g.slabs.size();
    const int* start = g.start;
    double* grid = out_grids + (size_t)(g.r0 + r) * G.padvol;
    const int64_t rind = g.picked[r];
    const int rs[3] = {(int)(rind % G.dist[0]), (int)((rind / G.dist[0]) % G.dist[1]),
                       (int)(rind / ((long long)G.dist[0] * G.dist[1]))};
    w.cutmask.assign((size_t)G.tilevol, 0);
    for (int si = 0; si < nslab; ++si) {
      const Slab& s = g.slabs[si];
      const uint8_t* keep = g.keepbuf[r * nslab + si].data();
      int i = 0;
      for (int z = 0; z < s.sz[2]; ++z)
        for (int y = 0; y < s.sz[1]; ++y)
          for (int x = 0; x < s.sz[0]; ++x, ++i) {
            const size_t q = ((size_t)(s.lo[2] + z) * t[1] + (s.lo[1] + y)) * t[0] + s.lo[0] + x;
            w.cutmask[q] |= s.prev ? keep[i] : (uint8_t)!keep[i];  // iqsim.jl:264 / :273
          }
    }
    // simdev[.!cutmask] = TIdev[.!cutmask]  (iqsim.jl:278)
    uint8_t* cg = D->debug ? out_cuts + (size_t)(g.r0 + r) * G.padvol : nullptr;
    for (int z = 0; z < t[2]; ++z)
      for (int y = 0; y < t[1]; ++y) {
        const long long gi = ((long long)(start[2] + z) * pad[1] + (start[1] + y)) * pad[0] + start[0];
        const double* src = ti64 + ((long long)(rs[2] + z) * n[1] + (rs[1] + y)) * n[0] + rs[0];
        const uint8_t* cm = &w.cutmask[((size_t)z * t[1] + y) * t[0]];
        for (int x = 0; x < t[0]; ++x) {
          if (!cm[x]) grid[gi + x] = src[x];
          if (cg) cg[gi + x] = cm[x];
        }
      }
  };
  std::function<void(Group*)> start_pastes = [&](Group* gp) {
    gp->pastes_left.store(gp->R);
    for (int r = 0; r < gp->R; ++r)
      pool.push([&, gp, r](int tid) {
        paste_task(*gp, r, tid);
        if (gp->pastes_left.fetch_sub(1) == 1) {
          gp->cut_ms += ms_since(gp->cut_t0);
          gp->latch->finish();
        }
      });
  };
  // few host threads per GPU (multi-GPU nodes): the cuts of a step run on the device instead -- but never by default
  // on integer-valued (categorical) images, whose degenerate capacities make the cut depend on the max-flow algorithm
  // (same rule as run_resident): the result must not depend on the host's core count or the nthreads argument
  const bool device_cut = D->cut_mode == 2 || (D->cut_mode == 0 && nthreads < 6 && !exact_cuts);
  std::atomic<int> cut_error{IQ_OK};
  std::mutex cut_error_m;
  std::string cut_error_msg;
  auto device_cut_job = [&](Group& g) {
    const int nslab = (int)g.slabs.size();
    const int ntask = g.R * nslab;
    if ((int)g.slabA.size() < ntask) { g.slabA.resize(ntask); g.slabB.resize(ntask); }
    g.cut_tasks.resize(ntask);
    const int* start = g.start;
    for (int task = 0; task < ntask; ++task) {
      const int r = task / nslab;
      const Slab& s = g.slabs[task % nslab];
      const double* grid = out_grids + (size_t)(g.r0 + r) * G.padvol;
      const int64_t rind = g.picked[r];
      const int rs[3] = {(int)(rind % G.dist[0]), (int)((rind / G.dist[0]) % G.dist[1]),
                         (int)(rind / ((long long)G.dist[0] * G.dist[1]))};
      const int nv = s.sz[0] * s.sz[1] * s.sz[2];
      g.slabA[task].resize(nv);
      g.slabB[task].resize(nv);
      g.keepbuf[task].resize(nv);
      int i = 0;
      for (int z = 0; z < s.sz[2]; ++z)
        for (int y = 0; y < s.sz[1]; ++y) {
          const int qy = s.lo[1] + y, qz = s.lo[2] + z;
          const double* ga = grid + ((long long)(start[2] + qz) * pad[1] + (start[1] + qy)) * pad[0] + start[0] + s.lo[0];
          const double* gb = ti64 + ((long long)(rs[2] + qz) * n[1] + (rs[1] + qy)) * n[0] + rs[0] + s.lo[0];
          for (int x = 0; x < s.sz[0]; ++x, ++i) { g.slabA[task][i] = ga[x]; g.slabB[task][i] = gb[x]; }
        }
      iq_cut_task& T = g.cut_tasks[task];
      T.A = g.slabA[task].data();
      T.B = g.slabB[task].data();
      T.sz[0] = s.sz[0]; T.sz[1] = s.sz[1]; T.sz[2] = s.sz[2];
      T.dim = s.d;
      T.keep = g.keepbuf[task].data();
    }
    const int rcc = iq_cut_batch(g.ctx, g.cut_tasks.data(), ntask, nullptr);
    if (rcc != IQ_OK) {
      std::lock_guard<std::mutex> l(cut_error_m);
      if (cut_error.load() == IQ_OK) cut_error_msg = iq_last_error();  // thread-local on this pool thread
      cut_error.store(rcc);
    }
  };
  auto submit_cut = [&](Group& g) {
    const int nslab = (int)g.slabs.size();
    const int ntask = g.R * nslab;
    if ((int)g.keepbuf.size() < ntask) g.keepbuf.resize(ntask);
    g.latch->begin();
    g.cut_t0 = clk::now();
    Group* gp = &g;
    if (ntask == 0) { start_pastes(gp); return; }
    if (device_cut) {
      pool.push([&, gp](int) {
        device_cut_job(*gp);
        start_pastes(gp);
      });
      return;
    }
    g.cuts_left.store(ntask);
    for (int task = 0; task < ntask; ++task)
      pool.push([&, gp, task](int tid) {
        cut_task(*gp, task, tid);
        if (gp->cuts_left.fetch_sub(1) == 1) start_pastes(gp);
      });
  };

  // ---- pipelined main loop: this thread runs the searches; the cut+paste job of a group is in flight
  //      on the pool until that group's next search needs its grid ----
  for (int64_t step = 0; step < D->npath && rc == IQ_OK; ++step) {
    for (int gi = 0; gi < ngroups; ++gi) {
      Group& g = groups[gi];
      g.latch->wait();          // cut+paste of this group's previous step
      rc = do_search(g, step);  // overlaps the cuts of the other groups
      if (rc != IQ_OK) break;
      submit_cut(g);
    }
  }
  for (auto& g : groups) g.latch->wait();
  if (rc == IQ_OK && cut_error.load() != IQ_OK) {
    rc = cut_error.load();
    iq_post_error(cut_error_msg.c_str());
  }
  double search_ms = 0, search_dev_ms = 0, cut_ms = 0, dist_ms = 0;
  int64_t launches = 0, dist_launches = 0, ncand = 0;
  for (auto& g : groups) {
    search_ms += g.search_ms; search_dev_ms += g.search_dev_ms; cut_ms += g.cut_ms; dist_ms += g.dist_ms;
    launches += g.launches; dist_launches += g.dist_launches; ncand += g.ncand;
  }
  destroy_all();
  if (rc != IQ_OK) return rc;
  if (D->out_real) {
    int cs[3] = {1, 1, 1};
    for (int i = 0; i < G.N; ++i) cs[i] = (int)std::min<int64_t>(D->sim_size[i], pad[i]);
    for (int r = 0; r < R; ++r) {
      const double* g = out_grids + (size_t)r * G.padvol;
      size_t o = 0;
      for (int z = 0; z < cs[2]; ++z)
        for (int y = 0; y < cs[1]; ++y) {
          const double* src = g + ((size_t)z * pad[1] + y) * pad[0];
          if (D->out_real_f32) {
            float* dst = (float*)D->out_real[r] + o;
            for (int x = 0; x < cs[0]; ++x) dst[x] = (float)src[x];
          } else {
            std::memcpy((double*)D->out_real[r] + o, src, sizeof(double) * cs[0]);
          }
          o += cs[0];
        }
    }
  }
  if (stats) {
    std::memset(stats, 0, sizeof *stats);
    stats->resident_status = resident_status;
    stats->search_ms = search_ms;
    stats->search_device_ms = search_dev_ms;
    stats->cut_ms = cut_ms;
    stats->total_ms = ms_since(t_start);
    stats->searches = (int64_t)R * D->npath;
    stats->kernel_launches = launches;
    stats->candidates = ncand;
    stats->setup_ms = setup_ms;
    stats->dist_kernel_ms = dist_ms;
    stats->dist_launches = dist_launches;
    stats->fft_searches = stats->direct_searches = 0;
    stats->fft_bytes = stats->fft_ms = 0;
    for (auto& g : groups) {
      stats->max_candidates = std::max(stats->max_candidates, g.maxcand);
      stats->fft_searches += g.nfft; stats->direct_searches += g.ndirect;
      stats->fft_bytes += g.fft_bytes; stats->fft_ms += g.fft_ms;
    }
  }
  return IQ_OK;
}

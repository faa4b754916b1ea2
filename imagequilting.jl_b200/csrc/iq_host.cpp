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
// each; while the GPU searches step k of one group, a host worker team cuts and pastes step k of the
// other group, so the host cut hides behind the device search (or vice versa).
#include <omp.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/iqb200.h"
#include "../../include/iqb200_host.h"
#include "iq_cut.h"

namespace {

using clk = std::chrono::steady_clock;
double ms_since(clk::time_point t0) { return std::chrono::duration<double, std::milli>(clk::now() - t0).count(); }

struct Slab {
  int d;       // dimension of the overlap
  bool prev;   // overlap with the previous (true) or next (false) tile
  int lo[3], sz[3];
};

// One persistent worker thread; it owns the OpenMP team used for cuts and pastes.
class Worker {
 public:
  Worker() : th_([this] { loop(); }) {}
  ~Worker() {
    {
      std::lock_guard<std::mutex> l(m_);
      quit_ = true;
    }
    cv_.notify_all();
    th_.join();
  }
  void submit(std::function<void()> f) {
    {
      std::lock_guard<std::mutex> l(m_);
      job_ = std::move(f);
      busy_ = true;
    }
    cv_.notify_all();
  }
  void wait() {
    std::unique_lock<std::mutex> l(m_);
    done_.wait(l, [this] { return !busy_; });
  }

 private:
  void loop() {
    for (;;) {
      std::function<void()> f;
      {
        std::unique_lock<std::mutex> l(m_);
        cv_.wait(l, [this] { return quit_ || busy_; });
        if (quit_) return;
        f = std::move(job_);
        job_ = nullptr;
      }
      f();
      {
        std::lock_guard<std::mutex> l(m_);
        busy_ = false;
      }
      done_.notify_all();
    }
  }
  std::mutex m_;
  std::condition_variable cv_, done_;
  std::function<void()> job_;
  bool busy_ = false, quit_ = false;
  std::thread th_;
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
  double search_ms = 0, search_dev_ms = 0, cut_ms = 0, dist_ms = 0;
  int64_t launches = 0, dist_launches = 0, ncand = 0;
};

}  // namespace

extern "C" int32_t iqh_graphcut(const double* A, const double* B, int32_t ndim, const int64_t* sz64, int32_t dim, uint8_t* keep) {
  if (!A || !B || !sz64 || !keep || ndim < 1 || ndim > 3 || dim < 0 || dim >= ndim) return IQ_ERR_INVALID;
  int sz[3] = {1, 1, 1};
  for (int i = 0; i < ndim; ++i) sz[i] = (int)sz64[i];
  if (sz[dim] < 2) return IQ_ERR_INVALID;
  iqcut::Work w;
  iqcut::graphcut(A, B, sz, dim, keep, w);
  return IQ_OK;
}

extern "C" int32_t iqh_run(const iqh_desc* D, double* out_grids, uint8_t* out_cuts, int64_t* out_picks, iqh_stats* stats) {
  if (!D || !out_grids || !D->ti || !D->ti_f32 || !D->path || !D->u) return IQ_ERR_INVALID;
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

  // one core stays with the thread that drives the GPU; the rest form the cut/paste team
  const int hw = std::max(1, (int)std::thread::hardware_concurrency());
  const int nthreads = D->nthreads > 0 ? D->nthreads : std::max(1, hw - 1);
  const int ngroups = (R >= 2) ? 2 : 1;

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
  std::vector<iqcut::Work> work(nthreads);

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
    for (int r = 0; r < g.R; ++r) {
      g.ncand += g.results[r].count;
      g.picked[r] = g.results[r].picked;
      if (out_picks) out_picks[(size_t)(g.r0 + r) * D->npath + step] = g.picked[r];
      if (g.picked[r] < 0) return IQ_ERR_STATE;
    }
    g.pasted[(size_t)ind] = 1;
    return IQ_OK;
  };

  // ---- boundary cut + paste of one step for one group (iqsim.jl:244-281): one (realization, slab)
  //      cut per task, then one paste per realization ----
  auto do_cut = [&](Group& g) {
    const auto tc = clk::now();
    const int nslab = (int)g.slabs.size();
    const int ntask = g.R * nslab;
    if ((int)g.keepbuf.size() < ntask) g.keepbuf.resize(ntask);
    const int* start = g.start;
    const int team = std::max(1, std::min(nthreads, std::max(ntask, g.R)));
#pragma omp parallel num_threads(team)
    {
      iqcut::Work& w = work[omp_get_thread_num()];
#pragma omp for schedule(dynamic, 1)
      for (int task = 0; task < ntask; ++task) {
        const int r = task / nslab;
        const Slab& s = g.slabs[task % nslab];
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
            const double* gb = D->ti + ((long long)(rs[2] + qz) * n[1] + (rs[1] + qy)) * n[0] + rs[0] + s.lo[0];
            for (int x = 0; x < s.sz[0]; ++x, ++i) { w.A[i] = ga[x]; w.B[i] = gb[x]; }
          }
        iqcut::graphcut(w.A.data(), w.B.data(), s.sz, s.d, g.keepbuf[task].data(), w);
      }
#pragma omp for schedule(static)
      for (int r = 0; r < g.R; ++r) {
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
            const double* src = D->ti + ((long long)(rs[2] + z) * n[1] + (rs[1] + y)) * n[0] + rs[0];
            const uint8_t* cm = &w.cutmask[((size_t)z * t[1] + y) * t[0]];
            for (int x = 0; x < t[0]; ++x) {
              if (!cm[x]) grid[gi + x] = src[x];
              if (cg) cg[gi + x] = cm[x];
            }
          }
      }
    }
    g.cut_ms += ms_since(tc);
  };

  // ---- pipelined main loop: search(group) on this thread, cut(group) on the worker thread ----
  {
    Worker worker;  // one worker = one OpenMP team: cuts of the two groups never oversubscribe the cores
    for (int64_t step = 0; step < D->npath && rc == IQ_OK; ++step) {
      for (int gi = 0; gi < ngroups; ++gi) {
        Group& g = groups[gi];
        // with two groups the only cut that can still be running is the one of the OTHER group's
        // current step -- and the one of this group's previous step has been waited for one
        // iteration ago; a single wait() here therefore covers both orders
        if (ngroups == 1) worker.wait();
        rc = do_search(g, step);  // overlaps the other group's cut
        if (rc != IQ_OK) break;
        worker.wait();
        Group* gp = &g;
        worker.submit([&do_cut, gp] { do_cut(*gp); });
      }
    }
    worker.wait();
  }
  double search_ms = 0, search_dev_ms = 0, cut_ms = 0, dist_ms = 0;
  int64_t launches = 0, dist_launches = 0, ncand = 0;
  for (auto& g : groups) {
    search_ms += g.search_ms; search_dev_ms += g.search_dev_ms; cut_ms += g.cut_ms; dist_ms += g.dist_ms;
    launches += g.launches; dist_launches += g.dist_launches; ncand += g.ncand;
  }
  destroy_all();
  if (rc != IQ_OK) return rc;
  if (stats) {
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
  }
  return IQ_OK;
}

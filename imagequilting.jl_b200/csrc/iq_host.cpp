// iq_host.cpp -- native host driver above the C ABI (include/iqb200_host.h).
//
// Restates, in C++, the host side of the reference that north_star keeps on the host:
//   the tile loop            /root/reference/src/iqsim.jl:163-312
//   the boundary cut         /root/reference/src/graphcut.jl:5-84  (lattice max-flow, here Dinic)
//   the paste                /root/reference/src/iqsim.jl:278
// It reaches the GPU only through iq_ctx_create / iq_search_pick / iq_last_search_stats / iq_ctx_destroy.
#include <omp.h>

#include <algorithm>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

#include "../../include/iqb200.h"
#include "../../include/iqb200_host.h"

namespace {

using clk = std::chrono::steady_clock;
double ms_since(clk::time_point t0) { return std::chrono::duration<double, std::milli>(clk::now() - t0).count(); }

// ------------------------------------------------------------------------------------------------
// Boundary cut.  The reference builds a lattice graph over one overlap slab, links the first slice along
// `dim` to a source and the last slice to a sink with infinite capacity, runs Boykov-Kolmogorov and keeps
// every voxel that is NOT in the sink tree at termination (labels 0/1, graphcut.jl:79-81).  At
// termination the sink tree is exactly the set of voxels that can still reach the sink in the residual
// graph, and that set is the same for every maximum flow -- so any exact max-flow algorithm gives the
// same mask.  Dinic on the implicit lattice is used here.
// ------------------------------------------------------------------------------------------------
struct CutWork {
  std::vector<double> cap;      // [nvox][2*3] residual capacities, dir = 2*d (+d) / 2*d+1 (-d)
  std::vector<int> level, it, queue, pnode, parc;
  std::vector<short> coord;     // [3][nvox]
  std::vector<uint8_t> reach;
  std::vector<double> A, B;     // slab extraction buffers
  std::vector<uint8_t> keep, cutmask;
};

void graphcut_impl(const double* A, const double* B, const int sz[3], int dim, uint8_t* keep, CutWork& w) {
  const int nvox = sz[0] * sz[1] * sz[2];
  const int stride[3] = {1, sz[0], sz[0] * sz[1]};
  constexpr int ND = 6;
  w.cap.assign((size_t)nvox * ND, 0.0);
  w.coord.resize((size_t)3 * nvox);
  for (int u = 0; u < nvox; ++u) {
    w.coord[u] = (short)(u % sz[0]);
    w.coord[nvox + u] = (short)((u / sz[0]) % sz[1]);
    w.coord[2 * nvox + u] = (short)(u / (sz[0] * sz[1]));
  }
  const double eps = std::numeric_limits<double>::epsilon();
  for (int d = 0; d < 3; ++d) {
    if (sz[d] < 2) continue;
    const short* cd = &w.coord[(size_t)d * nvox];
    for (int u = 0; u < nvox; ++u) {
      if (cd[u] >= sz[d] - 1) continue;
      const int v = u + stride[d];
      const double Du = std::fabs(A[u] - B[u]), Dv = std::fabs(A[v] - B[v]);
      const double gAu = std::fabs(A[v] - A[u]), gBu = std::fabs(B[v] - B[u]);
      double gAv = gAu, gBv = gBu;
      if (cd[u] + 2 < sz[d]) {
        const int x = v + stride[d];
        gAv = std::fabs(A[x] - A[v]);
        gBv = std::fabs(B[x] - B[v]);
      }
      const double c = (Du + Dv) / (gAu + gAv + gBu + gBv + eps);  // graphcut.jl:52
      w.cap[(size_t)u * ND + 2 * d] = c;
      w.cap[(size_t)v * ND + 2 * d + 1] = c;
    }
  }
  const short* cdim = &w.coord[(size_t)dim * nvox];
  const int last = sz[dim] - 1;
  auto neighbor = [&](int u, int dir) -> int {
    const int d = dir >> 1;
    const short c = w.coord[(size_t)d * nvox + u];
    if (dir & 1) return c > 0 ? u - stride[d] : -1;
    return c < sz[d] - 1 ? u + stride[d] : -1;
  };
  w.level.resize(nvox);
  w.it.resize(nvox);
  w.queue.resize(nvox);
  w.pnode.resize(nvox + 1);
  w.parc.resize(nvox + 1);

  for (;;) {
    // ---- BFS levels from the source slice ----
    std::fill(w.level.begin(), w.level.end(), -1);
    int qh = 0, qt = 0;
    for (int u = 0; u < nvox; ++u)
      if (cdim[u] == 0) { w.level[u] = 0; w.queue[qt++] = u; }
    int Lt = INT_MAX;
    while (qh < qt) {
      const int u = w.queue[qh++];
      if (cdim[u] == last) { Lt = std::min(Lt, w.level[u] + 1); continue; }  // u -> t (infinite)
      if (w.level[u] + 1 >= Lt) continue;
      for (int dir = 0; dir < ND; ++dir) {
        if (w.cap[(size_t)u * ND + dir] <= 0.0) continue;
        const int v = neighbor(u, dir);
        if (v >= 0 && w.level[v] < 0) { w.level[v] = w.level[u] + 1; w.queue[qt++] = v; }
      }
    }
    if (Lt == INT_MAX) break;
    // ---- blocking flow ----
    std::fill(w.it.begin(), w.it.end(), 0);
    for (int s0 = 0; s0 < nvox; ++s0) {
      if (cdim[s0] != 0) continue;
      int depth = 0;  // number of arcs on the current path
      w.pnode[0] = s0;
      for (;;) {
        const int u = w.pnode[depth];
        if (cdim[u] == last && w.level[u] + 1 == Lt) {
          if (depth == 0) break;  // source slice == sink slice cannot happen (ovlsize > 1)
          double f = std::numeric_limits<double>::infinity();
          for (int i = 0; i < depth; ++i) f = std::min(f, w.cap[(size_t)w.pnode[i] * ND + w.parc[i]]);
          int firstsat = depth;
          for (int i = 0; i < depth; ++i) {
            double& c = w.cap[(size_t)w.pnode[i] * ND + w.parc[i]];
            c -= f;
            w.cap[(size_t)w.pnode[i + 1] * ND + (w.parc[i] ^ 1)] += f;
            if (c <= 0.0 && firstsat == depth) firstsat = i;
          }
          depth = firstsat;  // resume from the tail of the first saturated arc
          continue;
        }
        bool advanced = false;
        if (!(cdim[u] == last)) {
          while (w.it[u] < ND) {
            const int dir = w.it[u];
            if (w.cap[(size_t)u * ND + dir] > 0.0) {
              const int v = neighbor(u, dir);
              if (v >= 0 && w.level[v] == w.level[u] + 1) {
                w.parc[depth] = dir;
                w.pnode[++depth] = v;
                advanced = true;
                break;
              }
            }
            ++w.it[u];
          }
        }
        if (advanced) continue;
        if (depth == 0) break;
        w.level[u] = -1;  // dead end in this phase
        --depth;
      }
    }
  }
  // ---- voxels that can still reach the sink in the residual graph ----
  w.reach.assign(nvox, 0);
  int qh = 0, qt = 0;
  for (int u = 0; u < nvox; ++u)
    if (cdim[u] == last) { w.reach[u] = 1; w.queue[qt++] = u; }
  while (qh < qt) {
    const int v = w.queue[qh++];
    for (int dir = 0; dir < ND; ++dir) {
      const int x = neighbor(v, dir);
      if (x < 0 || w.reach[x]) continue;
      if (w.cap[(size_t)x * ND + (dir ^ 1)] > 0.0) { w.reach[x] = 1; w.queue[qt++] = x; }
    }
  }
  for (int u = 0; u < nvox; ++u) keep[u] = w.reach[u] ? 0 : 1;
}

struct Slab {
  int d;       // dimension of the overlap
  bool prev;   // overlap with the previous (true) or next (false) tile
  int lo[3], sz[3];
};

}  // namespace

extern "C" int32_t iqh_graphcut(const double* A, const double* B, int32_t ndim, const int64_t* sz64, int32_t dim, uint8_t* keep) {
  if (!A || !B || !sz64 || !keep || ndim < 1 || ndim > 3 || dim < 0 || dim >= ndim) return IQ_ERR_INVALID;
  int sz[3] = {1, 1, 1};
  for (int i = 0; i < ndim; ++i) sz[i] = (int)sz64[i];
  if (sz[dim] < 2) return IQ_ERR_INVALID;
  CutWork w;
  graphcut_impl(A, B, sz, dim, keep, w);
  return IQ_OK;
}

extern "C" int32_t iqh_run(const iqh_desc* D, double* out_grids, uint8_t* out_cuts, int64_t* out_picks, iqh_stats* stats) {
  if (!D || !out_grids || !D->ti || !D->ti_f32 || !D->path || !D->u) return IQ_ERR_INVALID;
  if (D->debug && !out_cuts) return IQ_ERR_INVALID;
  const auto t_start = clk::now();
  const int N = D->ndim;
  int n[3] = {1, 1, 1}, t[3] = {1, 1, 1}, ov[3] = {1, 1, 1}, nt[3] = {1, 1, 1}, pad[3] = {1, 1, 1};
  for (int i = 0; i < N; ++i) {
    n[i] = (int)D->ti_size[i]; t[i] = (int)D->tile_size[i]; ov[i] = (int)D->ovl_size[i];
    nt[i] = (int)D->ntiles[i]; pad[i] = (int)D->pad_size[i];
  }
  int sp[3], dist[3];
  for (int i = 0; i < 3; ++i) { sp[i] = t[i] - ov[i]; dist[i] = n[i] - t[i] + 1; }
  if (N < 3) { ov[2] = 1; sp[2] = 0; }
  if (N < 2) return IQ_ERR_INVALID;
  const long long tilevol = (long long)t[0] * t[1] * t[2];
  const long long padvol = (long long)pad[0] * pad[1] * pad[2];
  const long long ntile_total = (long long)nt[0] * nt[1] * nt[2];
  const int R = D->nreal;
  const int S = D->nsoft;
  int nthreads = D->nthreads > 0 ? D->nthreads : (int)std::thread::hardware_concurrency();
  nthreads = std::max(1, std::min(nthreads, R));

  iq_ctx* ctx = nullptr;
  iq_ctx_desc cd{};
  cd.ndim = N;
  for (int i = 0; i < 3; ++i) { cd.ti_size[i] = n[i]; cd.tile_size[i] = t[i]; }
  cd.ti = D->ti_f32;
  cd.disabled = D->disabled;
  cd.nsoft = S;
  cd.auxti = D->auxti;
  cd.device = D->device;
  cd.max_batch = D->batch > 0 ? std::min(D->batch, R) : R;
  int rc = iq_ctx_create(&ctx, &cd);
  if (rc != IQ_OK) return rc;

  std::memset(out_grids, 0, sizeof(double) * padvol * R);  // simgrid = zeros (iqsim.jl:165)
  if (D->debug) std::memset(out_cuts, 0, (size_t)padvol * R);
  std::vector<uint8_t> pasted((size_t)ntile_total, 0);
  std::vector<uint8_t> ovlmask((size_t)tilevol);
  std::vector<float> simdev((size_t)tilevol * R);
  std::vector<std::vector<float>> softdev(S, std::vector<float>((size_t)tilevol));
  std::vector<const float*> softptr(S);
  std::vector<int32_t> hoff;
  std::vector<float> hval;
  std::vector<iq_tile> tiles(R);
  std::vector<iq_result> results(R);
  std::vector<double> ustep(R);
  std::vector<CutWork> work(nthreads);

  double search_ms = 0, search_dev_ms = 0, cut_ms = 0;
  int64_t launches = 0, ncand = 0;

  for (int64_t step = 0; step < D->npath; ++step) {
    const int64_t ind = D->path[step];
    if (ind < 0 || ind >= ntile_total) { iq_ctx_destroy(ctx); return IQ_ERR_INVALID; }
    const int ti3[3] = {(int)(ind % nt[0]), (int)((ind / nt[0]) % nt[1]), (int)(ind / ((long long)nt[0] * nt[1]))};
    const int start[3] = {ti3[0] * sp[0], ti3[1] * sp[1], ti3[2] * sp[2]};
    const long long tstride[3] = {1, nt[0], (long long)nt[0] * nt[1]};

    // ---- overlap slabs with pasted neighbours (iqsim.jl:188-205) ----
    std::vector<Slab> slabs;
    for (int d = 0; d < N; ++d) {
      if (ov[d] <= 1) continue;
      if (ti3[d] > 0 && pasted[(size_t)(ind - tstride[d])]) {
        Slab s{d, true, {0, 0, 0}, {t[0], t[1], t[2]}};
        s.sz[d] = ov[d];
        slabs.push_back(s);
      }
      if (ti3[d] + 1 < nt[d] && pasted[(size_t)(ind + tstride[d])]) {
        Slab s{d, false, {0, 0, 0}, {t[0], t[1], t[2]}};
        s.lo[d] = sp[d];
        s.sz[d] = t[d] - sp[d];
        slabs.push_back(s);
      }
    }
    std::fill(ovlmask.begin(), ovlmask.end(), 0);
    for (const Slab& s : slabs)
      for (int z = s.lo[2]; z < s.lo[2] + s.sz[2]; ++z)
        for (int y = s.lo[1]; y < s.lo[1] + s.sz[1]; ++y)
          std::memset(&ovlmask[((size_t)z * t[1] + y) * t[0] + s.lo[0]], 1, (size_t)s.sz[0]);

    // ---- hard data inside the tile (indicator!/event!, utils.jl:18-36) ----
    hoff.clear();
    hval.clear();
    if (D->hard_has) {
      for (int z = 0; z < t[2]; ++z)
        for (int y = 0; y < t[1]; ++y) {
          const long long g = ((long long)(start[2] + z) * pad[1] + (start[1] + y)) * pad[0] + start[0];
          for (int x = 0; x < t[0]; ++x)
            if (D->hard_has[g + x]) {
              hoff.push_back((int32_t)(((long long)z * t[1] + y) * t[0] + x));
              hval.push_back(D->hard_val[g + x]);
            }
        }
    }
    // ---- soft data events (iqsim.jl:224) ----
    for (int s = 0; s < S; ++s) {
      for (int z = 0; z < t[2]; ++z)
        for (int y = 0; y < t[1]; ++y) {
          const long long g = ((long long)(start[2] + z) * pad[1] + (start[1] + y)) * pad[0] + start[0];
          std::memcpy(&softdev[s][((size_t)z * t[1] + y) * t[0]], D->aux[s] + g, sizeof(float) * t[0]);
        }
      softptr[s] = softdev[s].data();
    }
    // ---- current content of every realization's tile (iqsim.jl:185) ----
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (int r = 0; r < R; ++r) {
      const double* grid = out_grids + (size_t)r * padvol;
      float* dst = &simdev[(size_t)r * tilevol];
      for (int z = 0; z < t[2]; ++z)
        for (int y = 0; y < t[1]; ++y) {
          const double* src = grid + ((long long)(start[2] + z) * pad[1] + (start[1] + y)) * pad[0] + start[0];
          float* drow = dst + ((size_t)z * t[1] + y) * t[0];
          for (int x = 0; x < t[0]; ++x) drow[x] = (float)src[x];
        }
    }
    for (int r = 0; r < R; ++r) {
      tiles[r].simdev = &simdev[(size_t)r * tilevol];
      tiles[r].hard_nnz = (int32_t)hoff.size();
      tiles[r].hard_offset = hoff.data();
      tiles[r].hard_value = hval.data();
      tiles[r].softdev = S ? softptr.data() : nullptr;
      ustep[r] = D->u[(size_t)r * D->npath + step];
    }

    // ---- the search (iqsim.jl:206-243) ----
    const auto ts = clk::now();
    rc = iq_search_pick(ctx, ovlmask.data(), tiles.data(), R, D->tol, ustep.data(), results.data());
    if (rc != IQ_OK) { iq_ctx_destroy(ctx); return rc; }
    search_ms += ms_since(ts);
    double dms = 0;
    int64_t nl = 0;
    iq_last_search_stats(ctx, &dms, &nl);
    search_dev_ms += dms;
    launches += nl;
    for (int r = 0; r < R; ++r) {
      ncand += results[r].count;
      if (out_picks) out_picks[(size_t)r * D->npath + step] = results[r].picked;
      if (results[r].picked < 0) { iq_ctx_destroy(ctx); return IQ_ERR_STATE; }
    }

    // ---- boundary cut + paste (iqsim.jl:244-281), one realization per host thread ----
    const auto tc = clk::now();
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 1)
    for (int r = 0; r < R; ++r) {
      CutWork& w = work[omp_get_thread_num()];
      double* grid = out_grids + (size_t)r * padvol;
      const int64_t rind = results[r].picked;
      const int rs[3] = {(int)(rind % dist[0]), (int)((rind / dist[0]) % dist[1]), (int)(rind / ((long long)dist[0] * dist[1]))};
      w.cutmask.assign((size_t)tilevol, 0);
      for (const Slab& s : slabs) {
        const int nv = s.sz[0] * s.sz[1] * s.sz[2];
        w.A.resize(nv);
        w.B.resize(nv);
        w.keep.resize(nv);
        int i = 0;
        for (int z = 0; z < s.sz[2]; ++z)
          for (int y = 0; y < s.sz[1]; ++y)
            for (int x = 0; x < s.sz[0]; ++x, ++i) {
              const int qx = s.lo[0] + x, qy = s.lo[1] + y, qz = s.lo[2] + z;
              w.A[i] = grid[((long long)(start[2] + qz) * pad[1] + (start[1] + qy)) * pad[0] + start[0] + qx];
              w.B[i] = D->ti[((long long)(rs[2] + qz) * n[1] + (rs[1] + qy)) * n[0] + rs[0] + qx];
            }
        graphcut_impl(w.A.data(), w.B.data(), s.sz, s.d, w.keep.data(), w);
        i = 0;
        for (int z = 0; z < s.sz[2]; ++z)
          for (int y = 0; y < s.sz[1]; ++y)
            for (int x = 0; x < s.sz[0]; ++x, ++i) {
              const size_t q = ((size_t)(s.lo[2] + z) * t[1] + (s.lo[1] + y)) * t[0] + s.lo[0] + x;
              const uint8_t k = s.prev ? w.keep[i] : (uint8_t)!w.keep[i];  // iqsim.jl:264 / :273
              w.cutmask[q] |= k;
            }
      }
      // simdev[.!cutmask] = TIdev[.!cutmask]  (iqsim.jl:278)
      uint8_t* cg = D->debug ? out_cuts + (size_t)r * padvol : nullptr;
      for (int z = 0; z < t[2]; ++z)
        for (int y = 0; y < t[1]; ++y) {
          const long long g = ((long long)(start[2] + z) * pad[1] + (start[1] + y)) * pad[0] + start[0];
          const double* src = D->ti + ((long long)(rs[2] + z) * n[1] + (rs[1] + y)) * n[0] + rs[0];
          const uint8_t* cm = &w.cutmask[((size_t)z * t[1] + y) * t[0]];
          for (int x = 0; x < t[0]; ++x) {
            if (!cm[x]) grid[g + x] = src[x];
            if (cg) cg[g + x] = cm[x];
          }
        }
    }
    cut_ms += ms_since(tc);
    pasted[(size_t)ind] = 1;
  }
  iq_ctx_destroy(ctx);
  if (stats) {
    stats->search_ms = search_ms;
    stats->search_device_ms = search_dev_ms;
    stats->cut_ms = cut_ms;
    stats->total_ms = ms_since(t_start);
    stats->searches = (int64_t)R * D->npath;
    stats->kernel_launches = launches;
    stats->candidates = ncand;
  }
  return IQ_OK;
}

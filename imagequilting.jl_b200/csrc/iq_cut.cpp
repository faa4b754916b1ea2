// iq_cut.cpp -- minimum boundary cut between two overlap slabs (host side).
//
// Restates /root/reference/src/graphcut.jl:5-84: lattice graph over the slab with capacities
// (|A-B|(u) + |A-B|(v)) / (grad A(u) + grad A(v) + grad B(u) + grad B(v) + eps) (graphcut.jl:32-52), the
// first slice along `dim` tied to the source and the last slice to the sink with infinite capacity
// (graphcut.jl:56-70), maximum flow by the Boykov-Kolmogorov augmenting-path algorithm (two search trees,
// grow / augment / adopt; Boykov & Kolmogorov, PAMI 2004 -- written from the paper, the reference calls
// GraphsFlows.jl for it, graphcut.jl:73) and the mask "not in the sink tree" (graphcut.jl:79-81).
//
// At termination the sink tree is the set of voxels that can still reach the sink in the residual graph;
// the routine recomputes that set with an explicit reverse BFS so the result does not depend on tree
// bookkeeping details.
#include "iq_cut.h"

#include <cmath>
#include <limits>

namespace iqcut {

namespace {
constexpr int8_t kTerminal = 6, kNone = 7, kOrphan = 8;

void setup_topology(Work& w, const int sz[3], int dim) {
  if (w.sz[0] == sz[0] && w.sz[1] == sz[1] && w.sz[2] == sz[2] && w.dim == dim) return;
  w.sz[0] = sz[0]; w.sz[1] = sz[1]; w.sz[2] = sz[2];
  w.dim = dim;
  const int nvox = sz[0] * sz[1] * sz[2];
  const int stride[3] = {1, sz[0], sz[0] * sz[1]};
  w.nbr.assign((size_t)nvox * 6, -1);
  w.term.assign(nvox, 0);
  int u = 0;
  for (int z = 0; z < sz[2]; ++z)
    for (int y = 0; y < sz[1]; ++y)
      for (int x = 0; x < sz[0]; ++x, ++u) {
        const int c[3] = {x, y, z};
        for (int d = 0; d < 3; ++d) {
          if (c[d] + 1 < sz[d]) w.nbr[(size_t)u * 6 + 2 * d] = u + stride[d];
          if (c[d] > 0) w.nbr[(size_t)u * 6 + 2 * d + 1] = u - stride[d];
        }
        if (c[dim] == 0) w.term[u] = 1;
        else if (c[dim] == sz[dim] - 1) w.term[u] = 2;
      }
}
}  // namespace

void graphcut(const double* A, const double* B, const int sz[3], int dim, uint8_t* keep, Work& w) {
  setup_topology(w, sz, dim);
  const int nvox = sz[0] * sz[1] * sz[2];
  const int* nbr = w.nbr.data();

  // ---- capacities (graphcut.jl:22-54) ----
  w.cap.assign((size_t)nvox * 6, 0.0);
  double* cap = w.cap.data();
  const double eps = std::numeric_limits<double>::epsilon();
  for (int d = 0; d < 3; ++d) {
    if (sz[d] < 2) continue;
    for (int u = 0; u < nvox; ++u) {
      const int v = nbr[(size_t)u * 6 + 2 * d];
      if (v < 0) continue;
      const double Du = std::fabs(A[u] - B[u]), Dv = std::fabs(A[v] - B[v]);
      const double gAu = std::fabs(A[v] - A[u]), gBu = std::fabs(B[v] - B[u]);
      double gAv = gAu, gBv = gBu;
      const int x = nbr[(size_t)v * 6 + 2 * d];
      if (x >= 0) {
        gAv = std::fabs(A[x] - A[v]);
        gBv = std::fabs(B[x] - B[v]);
      }
      const double c = (Du + Dv) / (gAu + gAv + gBu + gBv + eps);
      cap[(size_t)u * 6 + 2 * d] = c;
      cap[(size_t)v * 6 + 2 * d + 1] = c;
    }
  }

  // ---- Boykov-Kolmogorov ----
  w.tree.assign(w.term.begin(), w.term.end());
  w.par.assign(nvox, kNone);
  w.inactive_q.assign(nvox, 0);
  w.stamp.assign(nvox, 0);
  w.active.clear();
  w.orphans.clear();
  uint8_t* tree = w.tree.data();
  int8_t* par = w.par.data();
  uint8_t* inq = w.inactive_q.data();
  int* stamp = w.stamp.data();
  for (int u = 0; u < nvox; ++u)
    if (w.term[u]) { par[u] = kTerminal; w.active.push_back(u); inq[u] = 1; }
  size_t head = 0;
  int now = 0;

  auto activate = [&](int q) {
    if (!inq[q]) { inq[q] = 1; w.active.push_back(q); }
  };
  auto rooted = [&](int q) -> bool {
    int u = q;
    bool ok = false;
    for (;;) {
      if (stamp[u] == now) { ok = true; break; }
      const int8_t p = par[u];
      if (p == kTerminal) { ok = true; break; }
      if (p >= 6) break;  // kNone / kOrphan
      u = nbr[(size_t)u * 6 + p];
    }
    if (ok)
      for (u = q; stamp[u] != now; u = nbr[(size_t)u * 6 + par[u]]) {
        stamp[u] = now;
        if (par[u] == kTerminal) break;
      }
    return ok;
  };

  for (;;) {
    // ---- grow ----
    int s = -1, sdir = -1;
    while (head < w.active.size()) {
      const int p = w.active[head];
      const uint8_t tp = tree[p];
      if (!tp) { inq[p] = 0; ++head; continue; }
      for (int dir = 0; dir < 6 && s < 0; ++dir) {
        const int q = nbr[(size_t)p * 6 + dir];
        if (q < 0) continue;
        const double rc = (tp == 1) ? cap[(size_t)p * 6 + dir] : cap[(size_t)q * 6 + (dir ^ 1)];
        if (rc <= 0.0) continue;
        const uint8_t tq = tree[q];
        if (!tq) {
          tree[q] = tp;
          par[q] = (int8_t)(dir ^ 1);
          activate(q);
        } else if (tq != tp) {
          if (tp == 1) { s = p; sdir = dir; } else { s = q; sdir = dir ^ 1; }
        }
      }
      if (s >= 0) break;  // p stays active
      inq[p] = 0;
      ++head;
    }
    if (s < 0) break;
    if (head > (1u << 20)) {  // compact the queue now and then
      w.active.erase(w.active.begin(), w.active.begin() + head);
      head = 0;
    }

    // ---- augment ----
    const int t = nbr[(size_t)s * 6 + sdir];
    double f = cap[(size_t)s * 6 + sdir];
    for (int u = s; par[u] != kTerminal;) {
      const int pd = par[u], v = nbr[(size_t)u * 6 + pd];
      f = std::fmin(f, cap[(size_t)v * 6 + (pd ^ 1)]);
      u = v;
    }
    for (int u = t; par[u] != kTerminal;) {
      const int pd = par[u], v = nbr[(size_t)u * 6 + pd];
      f = std::fmin(f, cap[(size_t)u * 6 + pd]);
      u = v;
    }
    cap[(size_t)s * 6 + sdir] -= f;
    cap[(size_t)t * 6 + (sdir ^ 1)] += f;
    for (int u = s; par[u] != kTerminal;) {
      const int pd = par[u], v = nbr[(size_t)u * 6 + pd];
      double& c = cap[(size_t)v * 6 + (pd ^ 1)];
      c -= f;
      cap[(size_t)u * 6 + pd] += f;
      if (c <= 0.0) { par[u] = kOrphan; w.orphans.push_back(u); }
      u = v;
    }
    for (int u = t; par[u] != kTerminal;) {
      const int pd = par[u], v = nbr[(size_t)u * 6 + pd];
      double& c = cap[(size_t)u * 6 + pd];
      c -= f;
      cap[(size_t)v * 6 + (pd ^ 1)] += f;
      if (c <= 0.0) { par[u] = kOrphan; w.orphans.push_back(u); }
      u = v;
    }

    // ---- adopt ----
    ++now;
    while (!w.orphans.empty()) {
      const int u = w.orphans.back();
      w.orphans.pop_back();
      const uint8_t tr = tree[u];
      int best = -1;
      for (int dir = 0; dir < 6; ++dir) {
        const int q = nbr[(size_t)u * 6 + dir];
        if (q < 0 || tree[q] != tr) continue;
        const double rc = (tr == 1) ? cap[(size_t)q * 6 + (dir ^ 1)] : cap[(size_t)u * 6 + dir];
        if (rc <= 0.0) continue;
        if (rooted(q)) { best = dir; break; }
      }
      if (best >= 0) {
        par[u] = (int8_t)best;
        continue;
      }
      for (int dir = 0; dir < 6; ++dir) {
        const int q = nbr[(size_t)u * 6 + dir];
        if (q < 0 || tree[q] != tr) continue;
        const double rc = (tr == 1) ? cap[(size_t)q * 6 + (dir ^ 1)] : cap[(size_t)u * 6 + dir];
        if (rc > 0.0) activate(q);
        if (par[q] == (int8_t)(dir ^ 1)) { par[q] = kOrphan; w.orphans.push_back(q); }
      }
      tree[u] = 0;
      par[u] = kNone;
    }
  }

  // ---- voxels that can still reach the sink slice in the residual graph ----
  w.reach.assign(nvox, 0);
  w.queue.resize(nvox);
  int qh = 0, qt = 0;
  for (int u = 0; u < nvox; ++u)
    if (w.term[u] == 2) { w.reach[u] = 1; w.queue[qt++] = u; }
  while (qh < qt) {
    const int v = w.queue[qh++];
    for (int dir = 0; dir < 6; ++dir) {
      const int x = nbr[(size_t)v * 6 + dir];
      if (x < 0 || w.reach[x]) continue;
      if (cap[(size_t)x * 6 + (dir ^ 1)] > 0.0) { w.reach[x] = 1; w.queue[qt++] = x; }
    }
  }
  for (int u = 0; u < nvox; ++u) keep[u] = w.reach[u] ? 0 : 1;
}

}  // namespace iqcut

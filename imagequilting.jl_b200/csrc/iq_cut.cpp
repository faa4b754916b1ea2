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
// that set is the same for every maximum flow, so the routine (a) starts from a cheap valid flow -- every
// straight source-to-sink line along `dim` is saturated first, which removes about a third of the
// augmentations -- and (b) recomputes the set with an explicit reverse BFS, so the result does not depend
// on tree bookkeeping details.  Nodes are 64-byte records (capacities + tree state in one cache line) and
// neighbours are found by index arithmetic.
#include "iq_cut.h"

#include <algorithm>
#include <cmath>
#include <limits>

namespace iqcut {

namespace {
constexpr int8_t kTerminal = 6, kNone = 7, kOrphan = 8;
}

// Capacity arithmetic of the two instantiations: FP64 (what an FP64 max-flow code computes) and unsigned 128-bit
// integers (exact).
template <typename T> struct Cap;
template <> struct Cap<double> {
  static double inf() { return std::numeric_limits<double>::infinity(); }
  static bool pos(double c) { return c > 0.0; }
};
template <> struct Cap<u128> {
  static u128 inf() { return ~(u128)0; }
  static bool pos(u128 c) { return c != 0; }
};

// conv(c): the FP64 capacity of graphcut.jl:52 in the arithmetic of the cut
template <typename T, typename Conv>
static void graphcut_impl(const double* A, const double* B, const int sz[3], int dim, uint8_t* keep, Work& w,
                          std::vector<NodeT<T>>& nodes, Conv conv) {
  using Node = NodeT<T>;
  const int nvox = sz[0] * sz[1] * sz[2];
  const int off[6] = {1, -1, sz[0], -sz[0], sz[0] * sz[1], -sz[0] * sz[1]};
  nodes.resize(nvox);
  Node* nd = nodes.data();

  // ---- topology + initial tree state: the two terminal slices are the roots ----
  {
    int u = 0;
    for (int z = 0; z < sz[2]; ++z)
      for (int y = 0; y < sz[1]; ++y)
        for (int x = 0; x < sz[0]; ++x, ++u) {
          const int c[3] = {x, y, z};
          uint8_t v = 0;
          for (int d = 0; d < 3; ++d) {
            if (c[d] + 1 < sz[d]) v |= (uint8_t)(1 << (2 * d));
            if (c[d] > 0) v |= (uint8_t)(1 << (2 * d + 1));
          }
          Node& n = nd[u];
          n.valid = v;
          n.term = c[dim] == 0 ? 1 : (c[dim] == sz[dim] - 1 ? 2 : 0);
          for (int k = 0; k < 6; ++k) n.cap[k] = 0;
          n.stamp = 0;
          n.tree = n.term;
          n.par = n.term ? kTerminal : kNone;
          n.inq = 0;
        }
  }

  // ---- capacities (graphcut.jl:22-54) ----
  const double eps = std::numeric_limits<double>::epsilon();
  for (int d = 0; d < 3; ++d) {
    if (sz[d] < 2) continue;
    const uint8_t fwd = (uint8_t)(1 << (2 * d));
    for (int u = 0; u < nvox; ++u) {
      if (!(nd[u].valid & fwd)) continue;
      const int v = u + off[2 * d];
      const double Du = std::fabs(A[u] - B[u]), Dv = std::fabs(A[v] - B[v]);
      const double gAu = std::fabs(A[v] - A[u]), gBu = std::fabs(B[v] - B[u]);
      double gAv = gAu, gBv = gBu;  // gradient repeated on the border (graphcut.jl:43-47)
      if (nd[v].valid & fwd) {
        const int x = v + off[2 * d];
        gAv = std::fabs(A[x] - A[v]);
        gBv = std::fabs(B[x] - B[v]);
      }
      const T c = conv((Du + Dv) / (gAu + gAv + gBu + gBv + eps));
      nd[u].cap[2 * d] = c;
      nd[v].cap[2 * d + 1] = c;
    }
  }

  // ---- straight-line pre-augmentation along `dim` ----
  {
    const int dd = 2 * dim;
    for (int u0 = 0; u0 < nvox; ++u0) {
      if (nd[u0].term != 1) continue;
      T f = Cap<T>::inf();
      for (int u = u0; nd[u].term != 2; u += off[dd]) f = std::min(f, nd[u].cap[dd]);
      if (!Cap<T>::pos(f)) continue;
      for (int u = u0; nd[u].term != 2;) {
        const int v = u + off[dd];
        nd[u].cap[dd] -= f;
        nd[v].cap[dd + 1] += f;
        u = v;
      }
    }
  }

  // ---- Boykov-Kolmogorov ----
  w.active.clear();
  w.orphans.clear();
  for (int u = 0; u < nvox; ++u)
    if (nd[u].term) { w.active.push_back(u); nd[u].inq = 1; }
  size_t head = 0;
  int now = 0;
  auto activate = [&](int q) {
    if (!nd[q].inq) { nd[q].inq = 1; w.active.push_back(q); }
  };
  // does the parent chain of q end at a terminal?  Successful walks are cached for the current adoption.
  auto rooted = [&](int q) -> bool {
    int u = q;
    bool ok = false;
    for (;;) {
      if (nd[u].stamp == now) { ok = true; break; }
      const int8_t p = nd[u].par;
      if (p == kTerminal) { ok = true; break; }
      if (p >= 6) break;  // kNone / kOrphan
      u += off[p];
    }
    if (ok)
      for (u = q; nd[u].stamp != now; u += off[nd[u].par]) {
        nd[u].stamp = now;
        if (nd[u].par == kTerminal) break;
      }
    return ok;
  };

  for (;;) {
    // ---- grow: find an arc from the source tree to the sink tree ----
    int s = -1, sdir = -1;
    while (head < w.active.size()) {
      const int p = w.active[head];
      Node& np = nd[p];
      const uint8_t tp = np.tree;
      if (!tp) { np.inq = 0; ++head; continue; }
      for (int dir = 0; dir < 6 && s < 0; ++dir) {
        if (!(np.valid & (1 << dir))) continue;
        const int q = p + off[dir];
        Node& nq = nd[q];
        const T rc = (tp == 1) ? np.cap[dir] : nq.cap[dir ^ 1];
        if (!Cap<T>::pos(rc)) continue;
        const uint8_t tq = nq.tree;
        if (!tq) {
          nq.tree = tp;
          nq.par = (int8_t)(dir ^ 1);
          activate(q);
        } else if (tq != tp) {
          if (tp == 1) { s = p; sdir = dir; } else { s = q; sdir = dir ^ 1; }
        }
      }
      if (s >= 0) break;  // p stays active
      np.inq = 0;
      ++head;
    }
    if (s < 0) break;
    if (head > (1u << 20)) {  // compact the queue now and then
      w.active.erase(w.active.begin(), w.active.begin() + head);
      head = 0;
    }

    // ---- augment along source-root .. s -> t .. sink-root ----
    const int t = s + off[sdir];
    T f = nd[s].cap[sdir];
    for (int u = s; nd[u].par != kTerminal;) {
      const int pd = nd[u].par, v = u + off[pd];
      f = std::min(f, nd[v].cap[pd ^ 1]);
      u = v;
    }
    for (int u = t; nd[u].par != kTerminal;) {
      const int pd = nd[u].par, v = u + off[pd];
      f = std::min(f, nd[u].cap[pd]);
      u = v;
    }
    nd[s].cap[sdir] -= f;
    nd[t].cap[sdir ^ 1] += f;
    for (int u = s; nd[u].par != kTerminal;) {
      const int pd = nd[u].par, v = u + off[pd];
      T& c = nd[v].cap[pd ^ 1];
      c -= f;
      nd[u].cap[pd] += f;
      if (!Cap<T>::pos(c)) { nd[u].par = kOrphan; w.orphans.push_back(u); }
      u = v;
    }
    for (int u = t; nd[u].par != kTerminal;) {
      const int pd = nd[u].par, v = u + off[pd];
      T& c = nd[u].cap[pd];
      c -= f;
      nd[v].cap[pd ^ 1] += f;
      if (!Cap<T>::pos(c)) { nd[u].par = kOrphan; w.orphans.push_back(u); }
      u = v;
    }

    // ---- adopt orphans ----
    ++now;
    while (!w.orphans.empty()) {
      const int u = w.orphans.back();
      w.orphans.pop_back();
      Node& nu = nd[u];
      const uint8_t tr = nu.tree;
      int best = -1;
      for (int dir = 0; dir < 6; ++dir) {
        if (!(nu.valid & (1 << dir))) continue;
        const int q = u + off[dir];
        if (nd[q].tree != tr) continue;
        const T rc = (tr == 1) ? nd[q].cap[dir ^ 1] : nu.cap[dir];
        if (!Cap<T>::pos(rc)) continue;
        if (rooted(q)) { best = dir; break; }
      }
      if (best >= 0) {
        nu.par = (int8_t)best;
        continue;
      }
      for (int dir = 0; dir < 6; ++dir) {
        if (!(nu.valid & (1 << dir))) continue;
        const int q = u + off[dir];
        Node& nq = nd[q];
        if (nq.tree != tr) continue;
        const T rc = (tr == 1) ? nq.cap[dir ^ 1] : nu.cap[dir];
        if (Cap<T>::pos(rc)) activate(q);
        if (nq.par == (int8_t)(dir ^ 1)) { nq.par = kOrphan; w.orphans.push_back(q); }
      }
      nu.tree = 0;
      nu.par = kNone;
    }
  }

  // ---- voxels that can still reach the sink slice in the residual graph ----
  w.reach.assign(nvox, 0);
  w.queue.resize(nvox);
  int qh = 0, qt = 0;
  for (int u = 0; u < nvox; ++u)
    if (nd[u].term == 2) { w.reach[u] = 1; w.queue[qt++] = u; }
  while (qh < qt) {
    const int v = w.queue[qh++];
    for (int dir = 0; dir < 6; ++dir) {
      if (!(nd[v].valid & (1 << dir))) continue;
      const int x = v + off[dir];
      if (w.reach[x]) continue;
      if (Cap<T>::pos(nd[x].cap[dir ^ 1])) { w.reach[x] = 1; w.queue[qt++] = x; }
    }
  }
  for (int u = 0; u < nvox; ++u) keep[u] = w.reach[u] ? 0 : 1;
}

void graphcut(const double* A, const double* B, const int sz[3], int dim, uint8_t* keep, Work& w) {
  graphcut_impl<double>(A, B, sz, dim, keep, w, w.nodes, [](double c) { return c; });
}

bool integer_valued(const double* A, const double* B, int n) {
  for (int i = 0; i < n; ++i)
    if (A[i] != std::nearbyint(A[i]) || B[i] != std::nearbyint(B[i])) return false;
  return true;
}

bool graphcut_exact(const double* A, const double* B, const int sz[3], int dim, uint8_t* keep, Work& w) {
  // exponent range of the capacities (same formula and operation order as graphcut_impl)
  const int nvox = sz[0] * sz[1] * sz[2];
  const int off[3] = {1, sz[0], sz[0] * sz[1]};
  const double eps = std::numeric_limits<double>::epsilon();
  int emin = 1 << 30, emax = -(1 << 30);
  for (int d = 0; d < 3; ++d) {
    if (sz[d] < 2) continue;
    for (int u = 0; u < nvox; ++u) {
      const int cd = (u / off[d]) % sz[d];
      if (cd + 1 >= sz[d]) continue;
      const int v = u + off[d];
      const double Du = std::fabs(A[u] - B[u]), Dv = std::fabs(A[v] - B[v]);
      const double gAu = std::fabs(A[v] - A[u]), gBu = std::fabs(B[v] - B[u]);
      double gAv = gAu, gBv = gBu;
      if (cd + 2 < sz[d]) {
        const int x = v + off[d];
        gAv = std::fabs(A[x] - A[v]);
        gBv = std::fabs(B[x] - B[v]);
      }
      const double c = (Du + Dv) / (gAu + gAv + gBu + gBv + eps);
      if (!(c > 0.0)) continue;
      if (!std::isfinite(c)) return false;
      int ex;
      std::frexp(c, &ex);
      emin = std::min(emin, ex - 53);
      emax = std::max(emax, ex - 53);
    }
  }
  // every capacity < 2^(54 + emax - emin); sums over at most all 3 nvox arcs occur (pre-augmentation, path bottlenecks)
  int sumbits = 1;
  while ((1ll << sumbits) < 3ll * nvox + 1) ++sumbits;
  if (emax >= emin && 54 + (emax - emin) + sumbits > 128) return false;
  graphcut_impl<u128>(A, B, sz, dim, keep, w, w.nodes_x, [emin](double c) -> u128 {
    if (!(c > 0.0)) return 0;
    int ex;
    const double fr = std::frexp(c, &ex);
    return (u128)(uint64_t)std::ldexp(fr, 53) << (ex - 53 - emin);
  });
  return true;
}

}  // namespace iqcut

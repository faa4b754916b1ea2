/*
 * oracle/iq_oracle_c.c -- plain-C restatement of the two pieces of ImageQuilting.jl v1.3.1 that the NumPy oracle
 * (oracle/iq_oracle.py) is too slow for at BASELINE.json's full sizes.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as iq_oracle.py: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it).  PARITY UNPINNED against real Julia (see the
 * header of iq_oracle.py); this file is pinned against iq_oracle.py itself (tests/test_oracle_c.py) and is written
 * independently of the product's cut (csrc/iq_cut.cpp is Boykov-Kolmogorov, csrc/iq_cutgpu.cu push-relabel; this is
 * Dinic), so three unrelated max-flow codes have to agree on the keep-mask.
 *
 *   iqo_graphcut      /root/reference/src/graphcut.jl:5-84
 *   iqo_fastdistance  /root/reference/src/utils.jl:5-13 evaluated by its definition
 *                     D[p] = | sum_q w[q] (img[p+q] - kern[q])^2 |   (FP64, no FFT round-off)
 *
 * Arrays are column-major (Julia layout), sizes int64[ndim], ndim <= 3.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------------------------------------------------------
 * graphcut(A, B, dim)  (src/graphcut.jl:5-84).  Lattice arcs u<->v with capacity
 *     (|A[u]-B[u]| + |A[v]-B[v]|) / (|A[v]-A[u]| + gAv + |B[v]-B[u]| + gBv + eps())        (:37-52)
 * where gAv/gBv are the gradients one step further along the direction, repeated at the border (:41-47);
 * source = first slice along dim, sink = last slice, both with infinite arcs (:56-70).
 * Mask (:77-81) = labels in {free, source tree} at Boykov-Kolmogorov termination = complement of "can still reach
 * the sink in the residual graph of a maximum flow", which is the same set for every maximum flow.  Max-flow here:
 * Dinic (BFS levels + iterative blocking-flow DFS).
 * ------------------------------------------------------------------------------------------------------------- */
/* lattice capacities of graphcut.jl:22-54 in FP64, edge list (u, v, c) in the order d = 0, 1, 2 / column-major u */
static int lattice_caps(const double* A, const double* B, int ndim, const int sz[3], int* eu, int* ev, double* ec) {
  const int stride[3] = {1, sz[0], sz[0] * sz[1]};
  const double eps = 2.220446049250313e-16; /* eps(Float64) */
  int ne = 0;
  for (int d = 0; d < ndim; ++d) {
    if (sz[d] < 2) continue;
    for (int z = 0; z < sz[2]; ++z)
      for (int y = 0; y < sz[1]; ++y)
        for (int x = 0; x < sz[0]; ++x) {
          const int c[3] = {x, y, z};
          if (c[d] + 1 >= sz[d]) continue;
          const int u = x + y * stride[1] + z * stride[2], v = u + stride[d];
          const double Du = fabs(A[u] - B[u]), Dv = fabs(A[v] - B[v]);
          const double gAu = fabs(A[v] - A[u]), gBu = fabs(B[v] - B[u]);
          double gAv = gAu, gBv = gBu;
          if (c[d] + 2 < sz[d]) {
            const int w = v + stride[d];
            gAv = fabs(A[w] - A[v]);
            gBv = fabs(B[w] - B[v]);
          }
          eu[ne] = u; ev[ne] = v;
          ec[ne] = (Du + Dv) / (gAu + gAv + gBu + gBv + eps);
          ++ne;
        }
  }
  return ne;
}

/* Dinic (BFS levels + iterative blocking-flow DFS) + reverse reachability from the sink, generic in the capacity type:
 * double (what an FP64 max-flow code computes) or unsigned 128-bit integers (exact). */
#define DEFINE_MAXFLOW(NAME, CAP_T)                                                                                   \
  static int NAME(int nvox, int ne, const int* eu, const int* ev, const CAP_T* ec, CAP_T inf, int ndim,               \
                  const int sz[3], int dim, uint8_t* keep) {                                                          \
    const int stride[3] = {1, sz[0], sz[0] * sz[1]};                                                                  \
    const int s = nvox, t = nvox + 1, n = nvox + 2;                                                                   \
    const int maxarcs = 2 * (ne + 2 * nvox);                                                                          \
    int* head = (int*)malloc(sizeof(int) * n);                                                                        \
    int* next = (int*)malloc(sizeof(int) * maxarcs);                                                                  \
    int* to = (int*)malloc(sizeof(int) * maxarcs);                                                                    \
    CAP_T* cap = (CAP_T*)malloc(sizeof(CAP_T) * maxarcs);                                                             \
    int* level = (int*)malloc(sizeof(int) * n);                                                                       \
    int* it = (int*)malloc(sizeof(int) * n);                                                                          \
    int* queue = (int*)malloc(sizeof(int) * n);                                                                       \
    int* pathe = (int*)malloc(sizeof(int) * n);                                                                       \
    char* reach = (char*)calloc((size_t)n, 1);                                                                        \
    if (!head || !next || !to || !cap || !level || !it || !queue || !pathe || !reach) return -2;                      \
    int m = 0;                                                                                                        \
    for (int i = 0; i < n; ++i) head[i] = -1;                                                                         \
    for (int i = 0; i < ne; ++i) {                                                                                    \
      to[m] = ev[i]; cap[m] = ec[i]; next[m] = head[eu[i]]; head[eu[i]] = m++;                                        \
      to[m] = eu[i]; cap[m] = ec[i]; next[m] = head[ev[i]]; head[ev[i]] = m++;                                        \
    }                                                                                                                 \
    (void)ndim;                                                                                                       \
    for (int z = 0; z < sz[2]; ++z)                                                                                   \
      for (int y = 0; y < sz[1]; ++y)                                                                                 \
        for (int x = 0; x < sz[0]; ++x) {                                                                             \
          const int c[3] = {x, y, z};                                                                                 \
          const int u = x + y * stride[1] + z * stride[2];                                                            \
          if (c[dim] == 0) {                                                                                          \
            to[m] = u; cap[m] = inf; next[m] = head[s]; head[s] = m++;                                                \
            to[m] = s; cap[m] = 0; next[m] = head[u]; head[u] = m++;                                                  \
          }                                                                                                           \
          if (c[dim] == sz[dim] - 1) {                                                                                \
            to[m] = t; cap[m] = inf; next[m] = head[u]; head[u] = m++;                                                \
            to[m] = u; cap[m] = 0; next[m] = head[t]; head[t] = m++;                                                  \
          }                                                                                                           \
        }                                                                                                             \
    for (;;) {                                                                                                        \
      for (int i = 0; i < n; ++i) level[i] = -1;                                                                      \
      int qh = 0, qt = 0;                                                                                             \
      level[s] = 0;                                                                                                   \
      queue[qt++] = s;                                                                                                \
      while (qh < qt) {                                                                                               \
        const int u = queue[qh++];                                                                                    \
        for (int e = head[u]; e >= 0; e = next[e])                                                                    \
          if (cap[e] > 0 && level[to[e]] < 0) { level[to[e]] = level[u] + 1; queue[qt++] = to[e]; }                   \
      }                                                                                                               \
      if (level[t] < 0) break;                                                                                        \
      for (int i = 0; i < n; ++i) it[i] = head[i];                                                                    \
      for (;;) { /* one augmenting path per round of the blocking flow */                                            \
        int np = 0, u = s;                                                                                            \
        while (u != t) {                                                                                              \
          int adv = 0;                                                                                                \
          for (; it[u] >= 0; it[u] = next[it[u]]) {                                                                   \
            const int e = it[u];                                                                                      \
            if (cap[e] > 0 && level[to[e]] == level[u] + 1) { pathe[np++] = e; u = to[e]; adv = 1; break; }           \
          }                                                                                                           \
          if (!adv) {                                                                                                 \
            if (np == 0) break;                                                                                       \
            level[u] = -1; /* dead end */                                                                             \
            const int e = pathe[--np];                                                                                \
            u = to[e ^ 1];                                                                                            \
          }                                                                                                           \
        }                                                                                                             \
        if (u != t) break;                                                                                            \
        CAP_T f = inf;                                                                                                \
        for (int i = 0; i < np; ++i)                                                                                  \
          if (cap[pathe[i]] < f) f = cap[pathe[i]];                                                                   \
        for (int i = 0; i < np; ++i) { cap[pathe[i]] -= f; cap[pathe[i] ^ 1] += f; }                                  \
      }                                                                                                               \
    }                                                                                                                 \
    /* reverse reachability to t over arcs with residual capacity */                                                  \
    int qh = 0, qt = 0;                                                                                               \
    reach[t] = 1;                                                                                                     \
    queue[qt++] = t;                                                                                                  \
    while (qh < qt) {                                                                                                 \
      const int v = queue[qh++];                                                                                      \
      for (int e = head[v]; e >= 0; e = next[e]) {                                                                    \
        const int u = to[e];                                                                                          \
        if (!reach[u] && cap[e ^ 1] > 0) { reach[u] = 1; queue[qt++] = u; } /* arc u -> v is e^1 */                   \
      }                                                                                                               \
    }                                                                                                                 \
    for (int i = 0; i < nvox; ++i) keep[i] = reach[i] ? 0 : 1;                                                        \
    free(reach); free(head); free(next); free(to); free(cap); free(level); free(it); free(queue); free(pathe);        \
    return 0;                                                                                                         \
  }

typedef unsigned __int128 u128;
DEFINE_MAXFLOW(maxflow_f64, double)
DEFINE_MAXFLOW(maxflow_u128, u128)

static int graphcut_impl(const double* A, const double* B, int ndim, const int64_t* sz64, int dim, uint8_t* keep, int exact) {
  if (!A || !B || !sz64 || !keep || ndim < 1 || ndim > 3 || dim < 0 || dim >= ndim) return -1;
  int sz[3] = {1, 1, 1};
  for (int i = 0; i < ndim; ++i) sz[i] = (int)sz64[i];
  const int nvox = sz[0] * sz[1] * sz[2];
  int* eu = (int*)malloc(sizeof(int) * 3 * (size_t)nvox);
  int* ev = (int*)malloc(sizeof(int) * 3 * (size_t)nvox);
  double* ec = (double*)malloc(sizeof(double) * 3 * (size_t)nvox);
  if (!eu || !ev || !ec) return -2;
  const int ne = lattice_caps(A, B, ndim, sz, eu, ev, ec);
  int rc;
  if (!exact) {
    rc = maxflow_f64(nvox, ne, eu, ev, ec, INFINITY, ndim, sz, dim, keep);
  } else {
    /* every FP64 capacity is m * 2^e (m a 53-bit integer): scaled by 2^-emin they are all integers.  They must fit
       128 bits together with the sum over all arcs (the "infinite" terminal capacity): 54 + (emax - emin) +
       log2(arcs) <= 128, else -3 (the caller falls back to arbitrary precision). */
    int emin = 1 << 30, emax = -(1 << 30);
    for (int i = 0; i < ne; ++i) {
      if (!(ec[i] > 0.0)) continue;
      if (!isfinite(ec[i])) { free(eu); free(ev); free(ec); return -3; }
      int ex;
      frexp(ec[i], &ex);
      if (ex - 53 < emin) emin = ex - 53;
      if (ex - 53 > emax) emax = ex - 53;
    }
    int sumbits = 1;
    while ((1ll << sumbits) < (long long)ne + 2) ++sumbits;
    if (emax >= emin && 54 + (emax - emin) + sumbits > 128) { free(eu); free(ev); free(ec); return -3; }
    u128* ic = (u128*)malloc(sizeof(u128) * (size_t)(ne > 0 ? ne : 1));
    if (!ic) return -2;
    u128 total = 0;
    for (int i = 0; i < ne; ++i) {
      ic[i] = 0;
      if (!(ec[i] > 0.0)) continue;
      int ex;
      const double fr = frexp(ec[i], &ex);
      ic[i] = (u128)(uint64_t)ldexp(fr, 53) << (ex - 53 - emin);
      total += ic[i];
    }
    rc = maxflow_u128(nvox, ne, eu, ev, ic, total + 1, ndim, sz, dim, keep);
    free(ic);
  }
  free(eu); free(ev); free(ec);
  return rc;
}

int iqo_graphcut(const double* A, const double* B, int ndim, const int64_t* sz64, int dim, uint8_t* keep) {
  return graphcut_impl(A, B, ndim, sz64, dim, keep, 0);
}

/* The same cut with the FP64 capacities taken as exact integers (no rounding in the flow arithmetic): the one
 * well-defined answer on degenerate (integer-valued / categorical) slabs, see oracle/iq_oracle.py graphcut(exact=). */
int iqo_graphcut_exact(const double* A, const double* B, int ndim, const int64_t* sz64, int dim, uint8_t* keep) {
  return graphcut_impl(A, B, ndim, sz64, dim, keep, 1);
}

/* ---------------------------------------------------------------------------------------------------------------
 * fastdistance(img, kern; weights) by its definition (src/utils.jl:5-13 computes the same quantity as
 * A^2 - 2AB + B^2 through two FFT correlations): out[p] = | sum_q w[q] (img[p+q] - kern[q])^2 |, FP64,
 * accumulated term by term in column-major order of q.  out has size img - kern + 1 per dimension.
 * ------------------------------------------------------------------------------------------------------------- */
int iqo_fastdistance(const double* img, const int64_t* isz64, const double* kern, const double* weights,
                     const int64_t* ksz64, int ndim, double* out) {
  if (!img || !kern || !isz64 || !ksz64 || !out || ndim < 1 || ndim > 3) return -1;
  int64_t n[3] = {1, 1, 1}, k[3] = {1, 1, 1}, o[3] = {1, 1, 1};
  for (int i = 0; i < ndim; ++i) { n[i] = isz64[i]; k[i] = ksz64[i]; }
  for (int i = 0; i < 3; ++i) {
    o[i] = n[i] - k[i] + 1;
    if (o[i] < 1) return -1;
  }
  const int64_t kvol = k[0] * k[1] * k[2];
  int64_t nnz = 0;
  int64_t* qoff = (int64_t*)malloc(sizeof(int64_t) * (size_t)kvol);
  double* qk = (double*)malloc(sizeof(double) * (size_t)kvol);
  double* qw = (double*)malloc(sizeof(double) * (size_t)kvol);
  if (!qoff || !qk || !qw) return -2;
  for (int64_t qz = 0; qz < k[2]; ++qz)
    for (int64_t qy = 0; qy < k[1]; ++qy)
      for (int64_t qx = 0; qx < k[0]; ++qx) {
        const int64_t q = qx + k[0] * (qy + k[1] * qz);
        const double w = weights ? weights[q] : 1.0;
        if (w == 0.0) continue;
        qoff[nnz] = qx + n[0] * (qy + n[1] * qz);
        qk[nnz] = kern[q];
        qw[nnz] = w;
        ++nnz;
      }
#pragma omp parallel for collapse(2) schedule(static)
  for (int64_t pz = 0; pz < o[2]; ++pz)
    for (int64_t py = 0; py < o[1]; ++py) {
      const double* base = img + n[0] * (py + n[1] * pz);
      double* row = out + o[0] * (py + o[1] * pz);
      for (int64_t px = 0; px < o[0]; ++px) row[px] = 0.0;
      for (int64_t i = 0; i < nnz; ++i) {
        const double* src = base + qoff[i];
        const double kv = qk[i], w = qw[i];
        for (int64_t px = 0; px < o[0]; ++px) {
          const double d = src[px] - kv;
          row[px] += w * d * d;
        }
      }
      for (int64_t px = 0; px < o[0]; ++px) row[px] = fabs(row[px]);
    }
  free(qoff); free(qk); free(qw);
  return 0;
}

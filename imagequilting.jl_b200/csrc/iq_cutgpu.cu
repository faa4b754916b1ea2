// iq_cutgpu.cu -- boundary cut on the device: one CTA per overlap slab, whole problem in shared memory.
//
// Same graph and same answer as the host routine iq_cut.cpp (restating /root/reference/src/graphcut.jl:5-84):
// the keep-mask is the complement of "can still reach the sink slice in the residual graph of a maximum
// flow", a set that does not depend on which maximum (pre)flow is found.  The reference stays on the host
// (GraphsFlows' Boykov-Kolmogorov, graphcut.jl:73); this kernel exists because a multi-GPU box has only a
// few host cores per GPU and the host cut then bounds the whole simulation.
//
// Algorithm: push-relabel, phase 1 only (a maximum preflow already fixes the sink side), made DETERMINISTIC:
//   * slab re-laid out with the cut dimension slowest, so the source / sink slices are the first / last
//     layer and only the (L-2) inner layers are unknowns; all their state (6 residuals, excess, height) lives
//     in shared memory (61 B per voxel -> up to ~3600 inner voxels per CTA);
//   * voxels are 7-coloured by (a + 2b + 3k) mod 7, so that a voxel and its six neighbours all differ and two
//     voxels of one colour never share a neighbour.  A sweep discharges the active voxels colour by colour
//     (push along every admissible arc, relabel if excess is left): no two concurrent voxels touch the same
//     state, so there are no atomics and the floating-point result is independent of thread scheduling;
//   * the voxels of a colour that hold excess are first compacted into a list (dense warps), then discharged;
//   * every 8 sweeps an exact global relabel (frontier BFS from the sink slice in shared memory), which also
//     parks voxels that cannot reach the sink;
//   * the final global relabel IS the answer: height < HMAX <=> can reach the sink.
// FP64 throughout, capacities computed with the reference's formula and operation order
// ((Du+Dv)/(gAu+gAv+gBu+gBv+eps), graphcut.jl:52) with explicit round-to-nearest intrinsics (no contraction).
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>

#include "iq_internal.h"

namespace iq {

#ifndef IQ_CUT_THREADS
#define IQ_CUT_THREADS 512
#endif
#ifndef IQ_CUT_RELABEL
#define IQ_CUT_RELABEL 8
#endif
constexpr int kCutThreads = IQ_CUT_THREADS;

__device__ __forceinline__ double edge_cap(const double* __restrict__ A, const double* __restrict__ B, int lo, int off,
                                           bool has_next) {
  const int v = lo + off;
  const double Du = fabs(__dsub_rn(A[lo], B[lo])), Dv = fabs(__dsub_rn(A[v], B[v]));
  const double gAu = fabs(__dsub_rn(A[v], A[lo])), gBu = fabs(__dsub_rn(B[v], B[lo]));
  double gAv = gAu, gBv = gBu;
  if (has_next) {
    const int x = v + off;
    gAv = fabs(__dsub_rn(A[x], A[v]));
    gBv = fabs(__dsub_rn(B[x], B[v]));
  }
  const double den = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(gAu, gAv), gBu), gBv), 2.220446049250313e-16);
  return __ddiv_rn(__dadd_rn(Du, Dv), den);
}

// Flow arithmetic of the two instantiations: FP64 with explicit round-to-nearest operations (no contraction), and
// unsigned 128-bit integers -- the FP64 capacities scaled to exact integers, no rounding at all: the cut of
// integer-valued (categorical) slabs, whose degenerate capacities make an FP64 max-flow's answer depend on its
// rounding (iq_cut.h, graphcut_exact).
typedef unsigned __int128 u128;
template <typename T> struct DCap;
template <> struct DCap<double> {
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double mn(double a, double b) { return fmin(a, b); }
  static __device__ __forceinline__ bool pos(double a) { return a > 0.0; }
};
template <> struct DCap<u128> {
  static __device__ __forceinline__ u128 add(u128 a, u128 b) { return a + b; }
  static __device__ __forceinline__ u128 sub(u128 a, u128 b) { return a - b; }
  static __device__ __forceinline__ u128 mn(u128 a, u128 b) { return a < b ? a : b; }
  static __device__ __forceinline__ bool pos(u128 a) { return a != 0; }
};
// FP64 capacity -> arithmetic of the cut (emin: exponent of the unit of the exact representation)
template <typename T> __device__ __forceinline__ T cap_conv(double c, int emin);
template <> __device__ __forceinline__ double cap_conv<double>(double c, int) { return c; }
template <> __device__ __forceinline__ u128 cap_conv<u128>(double c, int emin) {
  if (!(c > 0.0)) return 0;
  int ex;
  const double fr = frexp(c, &ex);
  return (u128)(unsigned long long)ldexp(fr, 53) << (ex - 53 - emin);
}

// topology byte of an inner voxel: bits 0..5 = the neighbour in direction dir is an inner voxel,
// bit 6 = direction 4 (+cut axis) leads into the sink slice, bit 7 = checkerboard colour
constexpr unsigned kToSink = 64u, kColour = 128u;

template <typename CT>
__device__ __forceinline__ void graphcut_body(const CutTask& T) {
  using C = DCap<CT>;
  const int n0 = T.n0, n1 = T.n1, L = T.L;
  const int P = n0 * n1, nfree = (L - 2) * P;
  const double* __restrict__ A = T.A;
  const double* __restrict__ B = T.B;
  extern __shared__ __align__(16) unsigned char smraw[];
  CT* r = reinterpret_cast<CT*>(smraw);                  // [6][nfree]
  CT* e = r + 6 * (size_t)nfree;                         // [nfree]
  int* h = reinterpret_cast<int*>(e + nfree);            // [nfree]
  unsigned char* topo = reinterpret_cast<unsigned char*>(h + nfree);  // [nfree]
  const int rows = n1 * (L - 2), m7 = (n0 + 6) / 7, items = rows * m7;
  unsigned short* clist = reinterpret_cast<unsigned short*>(topo + ((nfree + 3) & ~3));  // [7][items] voxel of item or 0xffff
  unsigned short* fr0 = clist + 7 * items;   // [nfree] BFS frontier (ping) / active list of a colour phase
  unsigned short* fr1 = fr0 + nfree;         // [nfree] BFS frontier (pong)
  __shared__ int s_cnt[3];
  __shared__ int offs[6];  // neighbour offsets for the dynamically indexed BFS (a register array would live in local memory)
  const int tid = threadIdx.x;
  const int HMAX = nfree + 2;
  const int off[6] = {1, -1, n0, -n0, P, -P};
  if (tid < 6) offs[tid] = off[tid];

  // ---- exact arithmetic: exponent range of the capacities (unit 2^emin makes all of them integers) ----
  __shared__ int s_erange[2];
  int emin = 0;
  if (sizeof(CT) > sizeof(double)) {
    if (tid == 0) { s_erange[0] = 1 << 30; s_erange[1] = -(1 << 30); }
    __syncthreads();
    for (int i = tid; i < nfree; i += kCutThreads) {
      const int u = i + P;
      const int a = i % n0, b = (i / n0) % n1, k = i / P + 1;
      const int c3[3] = {a, b, k}, sz3[3] = {n0, n1, L};
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        double cv[2] = {0.0, 0.0};
        if (c3[d] + 1 < sz3[d]) cv[0] = edge_cap(A, B, u, off[2 * d], c3[d] + 2 < sz3[d]);
        if (c3[d] > 0) cv[1] = edge_cap(A, B, u - off[2 * d], off[2 * d], c3[d] + 1 < sz3[d]);
        for (int q = 0; q < 2; ++q)
          if (cv[q] > 0.0) {
            int ex2;
            frexp(cv[q], &ex2);
            if (isinf(cv[q])) ex2 = 1 << 20;
            atomicMin(&s_erange[0], ex2 - 53);
            atomicMax(&s_erange[1], ex2 - 53);
          }
      }
    }
    __syncthreads();
    emin = s_erange[0];
    // every capacity < 2^(54 + emax - emin); an excess is at most the sum of the arcs out of the source slice (P of them),
    // a residual at most twice a capacity: all of it must fit 128 bits
    int sumbits = 1;
    while ((1 << sumbits) < P + 1) ++sumbits;
    if (s_erange[1] >= emin && 54 + (s_erange[1] - emin) + sumbits + 1 > 128) {
      if (tid == 0 && T.iters) *T.iters = -2;
      return;
    }
  }

  // ---- capacities, topology and initial preflow (arcs out of the source slice are saturated) ----
  for (int i = tid; i < nfree; i += kCutThreads) {
    const int u = i + P;
    const int a = i % n0, b = (i / n0) % n1, k = i / P + 1;
    const int c3[3] = {a, b, k}, sz3[3] = {n0, n1, L};
    CT ex = 0;
    unsigned tp = ((a + b + k) & 1) ? kColour : 0u;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      CT cp = 0;  // arc towards +d
      if (c3[d] + 1 < sz3[d]) cp = cap_conv<CT>(edge_cap(A, B, u, off[2 * d], c3[d] + 2 < sz3[d]), emin);
      r[(2 * d) * nfree + i] = cp;
      CT cm = 0;  // arc towards -d (capacity defined from the lower voxel)
      if (c3[d] > 0) cm = cap_conv<CT>(edge_cap(A, B, u - off[2 * d], off[2 * d], c3[d] + 1 < sz3[d]), emin);
      if (d == 2 && k == 1) { ex = cm; cm = 0; }  // from the source slice: saturated, never pushed back
      r[(2 * d + 1) * nfree + i] = cm;
      if (d < 2) {
        if (c3[d] + 1 < sz3[d]) tp |= 1u << (2 * d);
        if (c3[d] > 0) tp |= 1u << (2 * d + 1);
      } else {
        if (k + 1 < L - 1) tp |= 1u << 4; else tp |= kToSink;
        if (k > 1) tp |= 1u << 5;
      }
    }
    e[i] = ex;
    h[i] = HMAX;
    topo[i] = (unsigned char)tp;
  }
  // work items of colour c: voxels with (a + 2b + 3k) mod 7 == c, enumerated row by row
  for (int w = tid; w < 7 * items; w += kCutThreads) {
    const int colour = w / items, ww = w - colour * items;
    const int row = ww / m7, j = ww - row * m7;
    const int b = row % n1, kk = row / n1;
    const int a = (((colour - 2 * b - 3 * kk) % 7) + 7) % 7 + 7 * j;
    clist[w] = (a < n0) ? (unsigned short)(row * n0 + a) : (unsigned short)0xffff;
  }
  __syncthreads();
#ifndef IQ_CUT_NO_PREAUG
  // Straight-line pre-augmentation (as the host routine does): every column along the cut axis carries
  // f = min(capacities along it) from the source slice to the sink slice before the push-relabel starts -- a valid
  // flow that takes the bulk of the excess out of the first layer.  One thread per column, no conflicts.
  for (int col = tid; col < (L > 2 ? P : 0); col += kCutThreads) {  // (L == 2: no inner layer, nothing to route)
    CT f = e[col];  // capacity of the (saturated) arc source slice -> layer 1
    for (int k = 0; k < L - 2; ++k) f = C::mn(f, r[4 * nfree + k * P + col]);
    if (C::pos(f)) {
      e[col] = C::sub(e[col], f);
      for (int k = 0; k < L - 2; ++k) {
        const int i = k * P + col;
        r[4 * nfree + i] = C::sub(r[4 * nfree + i], f);
        if (k + 1 < L - 2) r[5 * nfree + i + P] = C::add(r[5 * nfree + i + P], f);
      }
    }
  }
  __syncthreads();
#endif

  // Exact heights = BFS distance to the sink slice over residual arcs x -> v, level by level with explicit
  // frontiers in shared memory (every voxel is expanded once; levels are unique, so the result does not depend
  // on the order in which a frontier is filled).  Voxels that cannot reach the sink keep HMAX.
  auto global_relabel = [&]() {
    if (tid < 3) s_cnt[tid] = 0;
    __syncthreads();
    for (int i0 = 0; i0 < nfree; i0 += kCutThreads) {
      const int i = i0 + tid;
      const bool first = i < nfree && (topo[i] & kToSink) && C::pos(r[4 * nfree + i]);
      if (i < nfree) h[i] = first ? 1 : HMAX;
      const unsigned bal = __ballot_sync(0xffffffffu, first);
      if (bal) {
        int base = 0;
        if ((tid & 31) == 0) base = atomicAdd(&s_cnt[0], __popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (first) fr0[base + __popc(bal & ((1u << (tid & 31)) - 1u))] = (unsigned short)i;
      }
    }
    __syncthreads();
    // three rotating counters: level L reads s_cnt[L % 3], fills s_cnt[(L + 1) % 3] and clears s_cnt[(L + 2) % 3]
    // (the input of level L - 1, dead by now) -- one barrier per level
    unsigned short* cur = fr0;
    unsigned short* nxt = fr1;
    int ci = 0;
    for (int level = 1;; ++level) {
      const int n = s_cnt[ci];
      if (n == 0) break;
      const int cn = ci == 2 ? 0 : ci + 1, cz = cn == 2 ? 0 : cn + 1;
      if (tid == 0) s_cnt[cz] = 0;
      // 6 threads per frontier voxel: one per direction
      for (int t = tid; t < n * 6; t += kCutThreads) {
        const int v = cur[t / 6], dir = t % 6;
        if (!((topo[v] >> dir) & 1u)) continue;
        const int x = v + offs[dir];
        if (h[x] != HMAX) continue;
        if (!C::pos(r[(dir ^ 1) * nfree + x])) continue;  // arc x -> v must have residual capacity
        if (atomicCAS(&h[x], HMAX, level + 1) == HMAX) nxt[atomicAdd(&s_cnt[cn], 1)] = (unsigned short)x;
      }
      __syncthreads();
      ci = cn;
      unsigned short* tsw = cur; cur = nxt; nxt = tsw;
    }
    __syncthreads();
  };
#ifdef IQ_CUT_PROFILE
  long long t_push = 0, t_rel = 0, t_glob = 0, t0 = clock64();
  int nglob = 1;
#endif
  global_relabel();
#ifdef IQ_CUT_PROFILE
  t_glob += clock64() - t0;
#endif

  int iter = 0;
  const int max_iter = 200000;
  // 7-colouring (a + 2b + 3k) mod 7: a voxel and its six neighbours all have different colours, so two voxels
  // of one colour never share a neighbour.  All voxels of a colour can therefore be DISCHARGED at once (push
  // along every admissible arc, then relabel if excess is left) without atomics and with a result that does
  // not depend on thread scheduling; colours are processed one after the other (Gauss-Seidel across colours).
  for (;; ++iter) {
#ifdef IQ_CUT_PROFILE
    long long ta = clock64();
#endif
    int active = 0;
#pragma unroll 1
    for (int cstep = 0; cstep < 7; ++cstep) {
#ifndef IQ_CUT_ORDER_NATURAL
      // colours visited with stride 3: the neighbour of a voxel along the cut axis (+k changes the colour by 3) is
      // discharged in the very next phase, so flow injected next to the source slice can cross every inner layer
      // within one sweep (Gauss-Seidel wavefront along the flow direction)
      const int colour = (3 * cstep) % 7;
#else
      const int colour = cstep;
#endif
      const unsigned short* cl = clist + colour * items;
      // Every thread owns the voxels cl[tid], cl[tid + T], ... of this colour and discharges those that hold excess:
      // push along every admissible arc, relabel if excess is left.  No voxel of this colour shares a neighbour with
      // another one, so everything a discharge reads (own arcs, neighbour heights, reverse arcs, neighbour excesses)
      // is private to it during the phase: its 12 residuals / neighbour heights are loaded up front and the push logic
      // runs in registers (the first version interleaved loads and stores per direction and spent the phase in
      // shared-memory latency).
      for (int w = tid; w < items; w += kCutThreads) {
        const int i = cl[w];
        if (i == 0xffff) continue;
        CT ex = e[i];
        int hi = h[i];
        if (!C::pos(ex) || hi >= HMAX) continue;
        const unsigned tp = topo[i];
        CT rc[6];
        int hv[6];
#pragma unroll
        for (int dir = 0; dir < 6; ++dir) {
          const bool inner = (tp >> dir) & 1u;
          const bool sink = (dir == 4) && (tp & kToSink);
          rc[dir] = (inner || sink) ? r[dir * nfree + i] : (CT)0;
          hv[dir] = sink ? 0 : (inner ? h[i + off[dir]] : HMAX);
        }
        int mh = HMAX;  // lowest neighbour over the arcs that still have residual capacity after the pushes
#pragma unroll
        for (int dir = 0; dir < 6; ++dir) {
          CT rcd = rc[dir];
          if (!C::pos(rcd)) continue;
          if (C::pos(ex) && hi == hv[dir] + 1) {
            const CT d = C::mn(ex, rcd);
            rcd = C::sub(rcd, d);
            ex = C::sub(ex, d);
            r[dir * nfree + i] = rcd;
            if (!((dir == 4) && (tp & kToSink))) {
              const int v = i + off[dir];
              r[(dir ^ 1) * nfree + v] = C::add(r[(dir ^ 1) * nfree + v], d);  // only the few admissible arcs pay these
              e[v] = C::add(e[v], d);
              active = 1;  // v sits one level below: it can push on
            }
          }
          if (C::pos(rcd)) mh = min(mh, hv[dir]);
        }
        e[i] = ex;
        if (C::pos(ex)) {
          if (mh + 1 > hi) hi = min(mh + 1, HMAX);
          h[i] = hi;
          if (hi < HMAX) active = 1;
        }
      }
      __syncthreads();
    }
#ifdef IQ_CUT_PROFILE
    long long tb = clock64(); t_push += tb - ta;
#endif
    // (a voxel that received excess after its own colour was processed was flagged by the pusher)
    const int any = __syncthreads_or(active);
#ifdef IQ_CUT_PROFILE
    long long tc = clock64(); t_rel += tc - tb;
#endif
    if (!any || iter >= max_iter) break;
#ifdef IQ_CUT_LATE
    // experiment: exact heights more often once the bulk of the flow has settled (the long tail of sweeps is excess
    // crawling along stale labels)
    const bool relabel_now = iter >= 32 ? ((iter % IQ_CUT_LATE) == IQ_CUT_LATE - 1) : ((iter % IQ_CUT_RELABEL) == IQ_CUT_RELABEL - 1 || iter == 3);
#else
    const bool relabel_now = (iter % IQ_CUT_RELABEL) == IQ_CUT_RELABEL - 1 || iter == 3;
#endif
    if (relabel_now) {
      global_relabel();
#ifdef IQ_CUT_PROFILE
      t_glob += clock64() - tc; ++nglob;
#endif
    }
  }
  global_relabel();

  // ---- keep mask in the kernel's layout: source slice 1, sink slice 0, inner voxels = cannot reach the sink ----
  unsigned char* keep = T.keep;
  for (int i = tid; i < P; i += kCutThreads) {
    keep[i] = 1;
    keep[(L - 1) * P + i] = 0;
  }
  for (int i = tid; i < nfree; i += kCutThreads) keep[P + i] = (h[i] >= HMAX) ? 1 : 0;
  if (tid == 0 && T.iters) *T.iters = (iter >= max_iter) ? -1 : iter;
#ifdef IQ_CUT_PROFILE
  if (tid == 0 && blockIdx.x == 0) printf("cut nfree %d sweeps %d: cycles push %lld relabel %lld global %lld (n=%d) total %lld\n", nfree, iter, t_push, t_rel, t_glob, nglob, clock64() - t0);
#endif
}

// explicit task list (iq_cut_batch)
template <typename CT>
__global__ void __launch_bounds__(kCutThreads, 1) k_graphcut(const CutTask* tasks) { graphcut_body<CT>(tasks[blockIdx.x]); }

// Implicit task grid of the resident simulation: task (job, k) = slab k of job `job`, data at (job * maxslabs + k) * maxslab,
// dims[task] = {n0, n1, L, -} written by the slab gather (L = 0: the job's tile has no slab k -> nothing to do).  Launch
// position b -> slab index b / njobs, job b % njobs: the cuts of one slab kind (same size, similar cost) run side by
// side, the first kinds (x, y overlaps: the thick ones) lead and the thin z overlaps fill the tail.
template <typename CT>
__global__ void __launch_bounds__(kCutThreads, 1) k_graphcut_grid(const double* A, const double* B, unsigned char* keep, int* iters,
                                                                  const int4* __restrict__ dims, long long maxslab,
                                                                  int maxslabs, int njobs) {
  const int k = blockIdx.x / njobs, job = blockIdx.x - k * njobs;
  const int task = job * maxslabs + k;
  const int4 d = dims[task];
  if (d.z < 2) return;
  CutTask T;
  T.A = A + (long long)task * maxslab;
  T.B = B + (long long)task * maxslab;
  T.keep = keep + (long long)task * maxslab;
  T.n0 = d.x; T.n1 = d.y; T.L = d.z;
  T.iters = iters + task;
  graphcut_body<CT>(T);
}

size_t graphcut_smem(int n0, int n1, int L, bool exact) {
  const size_t nfree = (size_t)(L - 2) * n0 * n1;
  const size_t items = (size_t)n1 * (L - 2) * ((n0 + 6) / 7);
  const size_t flow = exact ? sizeof(u128) : sizeof(double);  // 6 residuals + the excess per voxel
  return nfree * (7 * flow + 4) + ((nfree + 3) & ~(size_t)3) + (7 * items + 2 * nfree) * sizeof(unsigned short) + 16;
}

cudaError_t launch_graphcut_grid(const double* A, const double* B, unsigned char* keep, int* iters, const int4* dims,
                                 long long maxslab, int maxslabs, int nslab, int njobs, size_t smem, bool exact, cudaStream_t s) {
  if (exact) {
    cudaError_t err = cudaFuncSetAttribute(k_graphcut_grid<u128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    k_graphcut_grid<u128><<<nslab * njobs, kCutThreads, smem, s>>>(A, B, keep, iters, dims, maxslab, maxslabs, njobs);
  } else {
    cudaError_t err = cudaFuncSetAttribute(k_graphcut_grid<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    k_graphcut_grid<double><<<nslab * njobs, kCutThreads, smem, s>>>(A, B, keep, iters, dims, maxslab, maxslabs, njobs);
  }
  return cudaGetLastError();
}

cudaError_t launch_graphcut(const CutTask* d_tasks, int ntask, size_t smem, cudaStream_t s, bool exact) {
  if (exact) {
    cudaError_t err = cudaFuncSetAttribute(k_graphcut<u128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    k_graphcut<u128><<<ntask, kCutThreads, smem, s>>>(d_tasks);
  } else {
    cudaError_t err = cudaFuncSetAttribute(k_graphcut<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    k_graphcut<double><<<ntask, kCutThreads, smem, s>>>(d_tasks);
  }
  return cudaGetLastError();
}

}  // namespace iq

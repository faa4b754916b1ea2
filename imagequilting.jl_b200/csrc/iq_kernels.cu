// iq_kernels.cu -- hand-written sm_100a kernels for the overlap-distance search.
//
//   k_dist_boxes   masked SSD cross-correlation of R templates against every patch position of the
//                  training image (replaces the two imfilter calls of fastdistance,
//                  /root/reference/src/utils.jl:5-13 + src/imfilter.jl:5-26), fused with the
//                  |A2 - 2AB + B2| combine (utils.jl:12), the disabled knock-out (iqsim.jl:207)
//                  and the min/max reduction needed by the selection (iqsim.jl:237, relaxation.jl:11)
//   k_dist_sparse  hard-data distance (iqsim.jl:210-219) as a direct sum over the few data voxels
//   k_sat_* / k_a2map   summed-volume table of img^2 and the per-mask A2 map (the template-
//                  independent half of fastdistance, utils.jl:8)
//   k_select_pass  radix select of the k-th smallest (value,index) key = Base.partialsortperm with
//                  the Perm ordering (relaxation.jl:12,27)
//   k_pick_count / k_pick_write   ordered compaction of the candidate set = findall (iqsim.jl:237)
//                  and fastintersect (relaxation.jl:41-48)
#include <algorithm>

#include "iq_internal.h"
#include "iq_tma.cuh"

#include <math_constants.h>

namespace iq {

// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ld8(float (&v)[8], const float* p) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// 8 template taps x 8 outputs x RB tiles of FMAs.  `lo`/`hi` hold 16 consecutive image values; output
// t and tap j use value t+j (sliding window kept in registers, indices are compile-time).
template <int RB>
__device__ __forceinline__ void fma_chunk(float (&acc)[RB][8], const float (&lo)[8], const float (&hi)[8],
                                          const float* __restrict__ kp) {
#pragma unroll
  for (int r = 0; r < RB; ++r) {
    float k[8];
    ld8(k, kp + r * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const float w = (t + j < 8) ? lo[t + j] : hi[t + j - 8];
        acc[r][t] = fmaf(w, k[j], acc[r][t]);
      }
    }
  }
}

__device__ __forceinline__ unsigned warp_min_u(unsigned v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ unsigned warp_max_u(unsigned v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ------------------------------------------------------------------------------------------------
// dense box correlation (direct path, below the FFT crossover)
// ------------------------------------------------------------------------------------------------
// Mask -> disjoint boxes (host).  The distance map of one z plane is cut in column panels of XT x-threads (8
// consecutive outputs each); a CTA of 256 threads takes 256 (x-thread, row) work items of a panel.  For every box of
// the mask and every z plane of the box the CTA needs the image patch under its items (plus the template plane of its
// RB tiles) in shared memory and slides the template rows over it with a 16-value register window: 64 RB FMAs per
// 2 + 2 RB LDS.128.
//
//   k_dist_flat      (default)  the patch is one cp.async.bulk.tensor.3d box of the image's tensor map, the template
//                    plane one cp.async.bulk copy, both landing on an mbarrier; two stage buffers, so the TMA engine
//                    fetches (box, plane) s+1 while the FMAs of stage s run -- no staging instructions, no registers, one
//                    __syncthreads per stage (one buffer when two would cost the second CTA of the SM).  TMA writes the
//                    box densely (row pitch = box width), so the conflict-free LDS.128 comes from the geometry instead
//                    of a swizzle: box width = 4 mod 8 floats and a quarter warp = an octet of work items, 4 x-threads
//                    of 2 consecutive rows or 1 x-thread of 8 rows.  Box starts are 16-byte aligned (BoxDesc::pad).
//   k_dist_flat_ldg  (variant 1, and the fallback for boxes beyond TMA's 256-element limit)  items numbered row-major,
//                    patch staged through registers with __ldg, single buffer, two barriers per stage; the two 16-byte
//                    halves of every 8-float chunk are swapped in odd 128-byte groups for conflict-free LDS.128.
constexpr int kFlatThreads = 256;

__device__ __forceinline__ void ld8sw(float (&v)[8], const float* rowp, int ci) {
  const float* a = rowp + ci * 8;
  const int sw = ((ci >> 2) & 1) << 2;
  const float4 lo = *reinterpret_cast<const float4*>(a + sw);
  const float4 hi = *reinterpret_cast<const float4*>(a + (sw ^ 4));
  v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w;
  v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
}

// Shared epilogue of the direct kernels: |A2 - 2AB + B2| in FP64 (utils.jl:12), disabled -> Inf (iqsim.jl:207),
// min / max of the enabled entries (iqsim.jl:237, relaxation.jl:11).
template <int RB>
__device__ __forceinline__ void flat_epilogue(const DistParams& P, const float (&tot)[RB][8], int grp, int pz, int row, int x0,
                                              bool valid, unsigned* s_min, unsigned* s_max) {
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid < 4) { s_min[tid] = 0x7f800000u; s_max[tid] = 0u; }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RB; ++r) {
    const int tile = grp * RB + r;
    unsigned vmin = 0x7f800000u, vmax = 0u;
    if (tile < P.R && valid) {
      const double b2 = P.b2[tile];
      float* orow = P.out + (long long)tile * P.npos;
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int x = x0 + t;
        if (x < P.nxo) {
          const long long p = ((long long)pz * P.nyo + row) * P.nxo + x;
          const double a2 = P.a2 ? (double)__ldg(P.a2 + p) : 0.0;
          float d = (float)fabs(a2 - 2.0 * (double)tot[r][t] + b2);
          const bool dis = P.disabled && P.disabled[p];
          if (dis) d = CUDART_INF_F;
          orow[p] = d;
          if (!dis) {
            const unsigned u = __float_as_uint(d);
            vmin = min(vmin, u);
            vmax = max(vmax, u);
          }
        }
      }
    }
    if (P.minbits) {
      vmin = warp_min_u(vmin);
      vmax = warp_max_u(vmax);
      if (lane == 0 && tile < P.R) {
        atomicMin(&s_min[r], vmin);
        atomicMax(&s_max[r], vmax);
      }
    }
  }
  if (P.minbits) {
    __syncthreads();
    if (tid < RB && grp * RB + tid < P.R) {
      atomicMin(P.minbits + grp * RB + tid, s_min[tid]);
      atomicMax(P.maxbits + grp * RB + tid, s_max[tid]);
    }
  }
}

// ---- TMA-staged (default) ---------------------------------------------------------------------------------------------
// Work items are numbered in octets of RS rows x (8 / RS) x-threads, a quarter warp = one octet.  With a box width of
// 4 mod 8 floats (an odd number of 16-byte units per row) both shapes give conflict-free LDS.128: RS = 2 (4 x-threads of 2
// rows; XT a multiple of 4) and RS = 8 (one x-thread of 8 rows; any XT) -- the host picks the one that pads the map
// least.  rows a CTA's 32 octets can span: they start anywhere in a row group and run over (XTB + 30) / XTB + 1 groups at most.
__host__ __device__ inline int flat_tma_rows(int XT, int RS) {
  const int xtb = XT / (8 / RS);
  return RS * ((xtb + 30) / xtb + 1);
}

template <int RB>
__global__ void __launch_bounds__(kFlatThreads, 2) k_dist_flat(const DistParams P, const __grid_constant__ FlatTmaMaps M) {
  using namespace iqtma;
  extern __shared__ __align__(128) unsigned char smraw_[];
  float* sm = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smraw_) + 127) & ~(uintptr_t)127);  // TMA boxes: 128-byte aligned
  const int NB = P.NB;  // stage buffers: 2 = the next stage is in flight during the FMAs, 1 = big boxes (two CTAs per SM overlap instead)
  auto patchS = [&](int buf) { return sm + buf * P.patch_floats; };
  auto tmplS = [&](int buf) { return sm + NB * P.patch_floats + buf * P.tmpl_floats; };
  __shared__ __align__(8) unsigned long long bar[2];
  __shared__ BoxDesc s_box[kFlatTmaMaxBox];
  __shared__ unsigned s_min[4], s_max[4];

  const int tid = threadIdx.x;
  const int XT = P.XT, RS = P.RS, XS = 8 / RS, XTB = XT / XS;
  const int X0 = blockIdx.x * XT * kT;
  const int noct = XTB * ((P.nyo + RS - 1) / RS);
  const int o0 = blockIdx.y * (kFlatThreads / 8);
  const int oct = min(o0 + (tid >> 3), noct - 1);  // clamped threads recompute the last octet; they never store
  const int rg = oct / XTB, xb = oct - rg * XTB;
  const int sub = tid & 7;
  const int row = RS * rg + sub / XS, xt = XS * xb + sub % XS;
  const bool valid = (o0 + (tid >> 3)) < noct && row < P.nyo;
  const int rlo = RS * (o0 / XTB);
  const int PHO = flat_tma_rows(XT, RS);
  const int pz = blockIdx.z % P.nzo, grp = blockIdx.z / P.nzo;
  const float* tgrp = P.tmpl + (long long)grp * P.tmpl_grp_stride;

  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_init_fence();
  }
  if (tid < P.nbox) s_box[tid] = P.boxes[tid];
  __syncthreads();

  // one elected thread feeds the TMA engine: patch box + template plane of stage (b, qz) into buffer `buf`
  auto issue = [&](int b, int qz, int buf) {
    const BoxDesc& bx = s_box[b];
    const int pitch = (XT + bx.nch) * 8 + 4;
    const unsigned pbytes = (unsigned)(pitch * (PHO + bx.h - 1)) * 4u;
    const int plane_floats = bx.h * bx.nch * 8 * RB;
    fence_async_smem();  // the buffer was last read by ordinary loads (released by the barrier of the previous stage)
    mbar_expect_tx(&bar[buf], pbytes + (unsigned)plane_floats * 4u);
    tma_load_3d(patchS(buf), &M.m[b], X0 + bx.x0 - bx.pad, rlo + bx.y0, pz + bx.z0 + qz, &bar[buf]);  // x start: multiple of 4 floats
    bulk_g2s(tmplS(buf), tgrp + (long long)bx.tmpl_off * RB + (long long)qz * plane_floats, (unsigned)plane_floats * 4u, &bar[buf]);
  };

  float tot[RB][8];
#pragma unroll
  for (int r = 0; r < RB; ++r)
#pragma unroll
    for (int t = 0; t < 8; ++t) tot[r][t] = 0.f;

  if (tid == 0 && P.nbox > 0) issue(0, 0, 0);
  int it = 0;
  for (int b = 0; b < P.nbox; ++b) {
    const int bh = s_box[b].h, bnch = s_box[b].nch, bd = s_box[b].d;
    const int pitch = (XT + bnch) * 8 + 4;
    for (int qz = 0; qz < bd; ++qz, ++it) {
      const int buf = NB == 2 ? (it & 1) : 0;
      int nb = b, nq = qz + 1;  // the stage after this one
      if (nq >= bd) { nq = 0; ++nb; }
      if (NB == 2 && tid == 0 && nb < P.nbox) issue(nb, nq, buf ^ 1);  // into the other buffer while this one is consumed
      mbar_wait_bounded(&bar[buf], (unsigned)((NB == 2 ? (it >> 1) : it) & 1));

      float acc[RB][8];
#pragma unroll
      for (int r = 0; r < RB; ++r)
#pragma unroll
        for (int t = 0; t < 8; ++t) acc[r][t] = 0.f;

      const float* rowp = patchS(buf) + (row - rlo) * pitch + xt * 8;
      const float* kp = tmplS(buf);
      for (int qy = 0; qy < bh; ++qy) {
        float A[8], B[8];
        const float* a = rowp;
        ld8(A, a);
        for (int c = 0; c < bnch; c += 2) {
          ld8(B, a + 8);
          fma_chunk<RB>(acc, A, B, kp);
          kp += RB * 8;
          if (c + 1 < bnch) {
            ld8(A, a + 16);
            fma_chunk<RB>(acc, B, A, kp);
            kp += RB * 8;
          }
          a += 16;
        }
        rowp += pitch;
      }
#pragma unroll
      for (int r = 0; r < RB; ++r)
#pragma unroll
        for (int t = 0; t < 8; ++t) tot[r][t] += acc[r][t];
      __syncthreads();  // everybody is done with this buffer before it is refilled
      if (NB == 1 && tid == 0 && nb < P.nbox) issue(nb, nq, 0);
    }
  }
  flat_epilogue<RB>(P, tot, grp, pz, row, X0 + xt * kT, valid, s_min, s_max);
}

// ---- register-staged (variant 1 / fallback) -----------------------------------------------------------------------
template <int RB>
__global__ void __launch_bounds__(kFlatThreads, 2) k_dist_flat_ldg(const DistParams P) {
  extern __shared__ __align__(16) float smem[];
  float* patch = smem;
  float* tmplS = smem + P.patch_floats;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int XT = P.XT;
  const int X0 = blockIdx.x * XT * kT;
  const int nitem = XT * P.nyo;
  const int i0 = blockIdx.y * kFlatThreads;
  const int item = min(i0 + tid, nitem - 1);  // clamped threads recompute the last item; they never store
  const bool valid = (i0 + tid) < nitem;
  const int row = item / XT, xt = item - row * XT;
  const int rlo = i0 / XT;
  const int rhi = min(i0 + kFlatThreads - 1, nitem - 1) / XT;
  const int PHO = rhi - rlo + 1;
  const int pz = blockIdx.z % P.nzo, grp = blockIdx.z / P.nzo;

  float tot[RB][8];
#pragma unroll
  for (int r = 0; r < RB; ++r)
#pragma unroll
    for (int t = 0; t < 8; ++t) tot[r][t] = 0.f;

  for (int b = 0; b < P.nbox; ++b) {
    const BoxDesc bx = P.boxes[b];
    const int PW = (XT + bx.nch) * 8;
    const int pitch = PW + 4;
    const int PH = PHO + bx.h - 1;
    const int plane_floats = bx.h * bx.nch * 8 * RB;
    const float* tsrc = P.tmpl + (long long)grp * P.tmpl_grp_stride + (long long)bx.tmpl_off * RB;

    for (int qz = 0; qz < bx.d; ++qz) {
      __syncthreads();
      const int gz = pz + bx.z0 + qz;
      const float* src = P.img + (long long)gz * P.nx * P.ny;
      for (int prow = warp; prow < PH; prow += kFlatThreads / 32) {
        const int gy = rlo + bx.y0 + prow;
        const bool rowok = gy < P.ny;
        const float* srow = src + (long long)gy * P.nx;
        float* drow = patch + prow * pitch;
        for (int col = lane; col < PW; col += 32) {
          const int gx = X0 + bx.x0 - bx.pad + col;
          const int pcol = col ^ ((((col >> 5) & 1)) << 2);  // swap the chunk halves in odd 32-float groups
          drow[pcol] = (rowok && gx < P.nx) ? __ldg(srow + gx) : 0.f;
        }
      }
      {
        const float4* t4 = reinterpret_cast<const float4*>(tsrc + (long long)qz * plane_floats);
        float4* d4 = reinterpret_cast<float4*>(tmplS);
        for (int i = tid; i < plane_floats / 4; i += kFlatThreads) d4[i] = __ldg(t4 + i);
      }
      __syncthreads();

      float acc[RB][8];
#pragma unroll
      for (int r = 0; r < RB; ++r)
#pragma unroll
        for (int t = 0; t < 8; ++t) acc[r][t] = 0.f;

      const float* rowp = patch + (row - rlo) * pitch;
      const float* kp = tmplS;
      for (int qy = 0; qy < bx.h; ++qy) {
        float A[8], B[8];
        int ci = xt;
        ld8sw(A, rowp, ci);
        for (int c = 0; c < bx.nch; c += 2) {
          ld8sw(B, rowp, ++ci);
          fma_chunk<RB>(acc, A, B, kp);
          kp += RB * 8;
          if (c + 1 < bx.nch) {
            ld8sw(A, rowp, ++ci);
            fma_chunk<RB>(acc, B, A, kp);
            kp += RB * 8;
          }
        }
        rowp += pitch;
      }
#pragma unroll
      for (int r = 0; r < RB; ++r)
#pragma unroll
        for (int t = 0; t < 8; ++t) tot[r][t] += acc[r][t];
    }
  }

  __shared__ unsigned s_min[4], s_max[4];
  flat_epilogue<RB>(P, tot, grp, pz, row, X0 + xt * kT, valid, s_min, s_max);
}

size_t dist_flat_ldg_smem(const BoxDesc* boxes, int nbox, int XT, int rb, int* patch_floats) {
  int pf = 0, tf = 0;
  const int pho = (kFlatThreads + XT - 1) / XT + 1;
  for (int b = 0; b < nbox; ++b) {
    const int pitch = (XT + boxes[b].nch) * 8 + 4;
    pf = max(pf, (pho + boxes[b].h - 1) * pitch);
    tf = max(tf, boxes[b].h * boxes[b].nch * 8 * rb);
  }
  pf = (pf + 3) & ~3;
  if (patch_floats) *patch_floats = pf;
  return (size_t)(pf + tf) * sizeof(float);
}

// TMA kernel: NB stage buffers of (largest patch box + largest template plane), each rounded up to 128 bytes, plus the
// alignment slack.  0 = this box list does not fit TMA at this panel shape (box sides <= 256 elements, at most
// kFlatTmaMaxBox boxes).  Boxes larger than the image or running over its edge are fine (zero-filled;
// scripts/probe/tma_probe.cu).
size_t dist_flat_smem(const BoxDesc* boxes, int nbox, int XT, int RS, int NB, int rb, int* patch_floats, int* tmpl_floats) {
  const int XS = 8 / RS;
  if (XT < XS || (XT % XS) || nbox > kFlatTmaMaxBox) return 0;
  int pf = 0, tf = 0;
  const int pho = flat_tma_rows(XT, RS);
  for (int b = 0; b < nbox; ++b) {
    const int pitch = (XT + boxes[b].nch) * 8 + 4, ph = pho + boxes[b].h - 1;
    if (pitch > 256 || ph > 256) return 0;
    pf = max(pf, ph * pitch);
    tf = max(tf, boxes[b].h * boxes[b].nch * 8 * rb);
  }
  pf = (pf + 31) & ~31;
  tf = (tf + 31) & ~31;
  if (patch_floats) *patch_floats = pf;
  if (tmpl_floats) *tmpl_floats = tf;
  return (size_t)NB * (pf + tf) * sizeof(float) + 128;
}

void dist_flat_box(const BoxDesc& b, int XT, int RS, int* width, int* rows) {
  *width = (XT + b.nch) * 8 + 4;
  *rows = flat_tma_rows(XT, RS) + b.h - 1;
}

template <int RB>
static cudaError_t launch_dist_flat_t(const DistParams& p, const FlatTmaMaps& maps, size_t smem, cudaStream_t s) {
  cudaError_t e = cudaFuncSetAttribute(k_dist_flat<RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int ngrp = (p.R + RB - 1) / RB;
  const int nxt = (p.nxo + kT - 1) / kT;
  const int noct = (p.XT / (8 / p.RS)) * ((p.nyo + p.RS - 1) / p.RS);
  dim3 grid((nxt + p.XT - 1) / p.XT, (noct + kFlatThreads / 8 - 1) / (kFlatThreads / 8), p.nzo * ngrp);
  k_dist_flat<RB><<<grid, kFlatThreads, smem, s>>>(p, maps);
  return cudaGetLastError();
}

cudaError_t launch_dist_flat(const DistParams& p, const FlatTmaMaps& maps, int rb, size_t smem, cudaStream_t s) {
  switch (rb) {
    case 1: return launch_dist_flat_t<1>(p, maps, smem, s);
    case 2: return launch_dist_flat_t<2>(p, maps, smem, s);
    case 4: return launch_dist_flat_t<4>(p, maps, smem, s);
    default: return cudaErrorInvalidValue;
  }
}

template <int RB>
static cudaError_t launch_dist_flat_ldg_t(const DistParams& p, size_t smem, cudaStream_t s) {
  cudaError_t e = cudaFuncSetAttribute(k_dist_flat_ldg<RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int ngrp = (p.R + RB - 1) / RB;
  const int nxt = (p.nxo + kT - 1) / kT;
  dim3 grid((nxt + p.XT - 1) / p.XT, (p.XT * p.nyo + kFlatThreads - 1) / kFlatThreads, p.nzo * ngrp);
  k_dist_flat_ldg<RB><<<grid, kFlatThreads, smem, s>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_dist_flat_ldg(const DistParams& p, int rb, size_t smem, cudaStream_t s) {
  switch (rb) {
    case 1: return launch_dist_flat_ldg_t<1>(p, smem, s);
    case 2: return launch_dist_flat_ldg_t<2>(p, smem, s);
    case 4: return launch_dist_flat_ldg_t<4>(p, smem, s);
    default: return cudaErrorInvalidValue;
  }
}

// ------------------------------------------------------------------------------------------------
// sparse (hard data) distance: D[p] = sum_i (img[p + off_i] - v_i)^2
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_dist_sparse(const SparseParams P) {
  const int r = blockIdx.y;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  __shared__ unsigned s_min, s_max;
  if (threadIdx.x == 0) { s_min = 0x7f800000u; s_max = 0u; }
  __syncthreads();
  unsigned vmin = 0x7f800000u, vmax = 0u;
  if (p < P.npos) {
    const int x = (int)(p % P.nxo);
    const long long q = p / P.nxo;
    const int y = (int)(q % P.nyo), z = (int)(q / P.nyo);
    const float* base = P.img + ((long long)z * P.ny + y) * P.nx + x;
    float sum = 0.f;
    const int i0 = P.ptr[r * P.ptr_stride], i1 = P.ptr[r * P.ptr_stride + 1];
    for (int i = i0; i < i1; ++i) {
      const float diff = __ldg(base + P.off[i]) - __ldg(P.val + i);
      sum = fmaf(diff, diff, sum);
    }
    const bool dis = P.disabled && P.disabled[p];
    if (dis) sum = CUDART_INF_F;
    P.out[(long long)r * P.npos + p] = sum;
    if (!dis) { vmin = __float_as_uint(sum); vmax = vmin; }
  }
  vmin = warp_min_u(vmin);
  vmax = warp_max_u(vmax);
  if ((threadIdx.x & 31) == 0) { atomicMin(&s_min, vmin); atomicMax(&s_max, vmax); }
  __syncthreads();
  if (threadIdx.x == 0 && P.minbits) { atomicMin(P.minbits + r, s_min); atomicMax(P.maxbits + r, s_max); }
}

cudaError_t launch_dist_sparse(const SparseParams& p, cudaStream_t s) {
  dim3 grid((unsigned)((p.npos + 255) / 256), p.R);
  k_dist_sparse<<<grid, 256, 0, s>>>(p);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// summed-volume table of img^2 in FP64: sat(x,y,z) = sum_{i<x,j<y,k<z} img(i,j,k)^2,
// dims (nx+1, ny+1, nz+1), zero on the x=0 / y=0 / z=0 faces.
// ------------------------------------------------------------------------------------------------
__global__ void k_sat_rows(const float* __restrict__ img, double* __restrict__ sat, int nx, int ny, int nz) {
  // one warp per (y,z) row: square + inclusive scan along x with shuffles
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= (long long)ny * nz) return;
  const int y = (int)(row % ny), z = (int)(row / ny);
  const float* src = img + row * nx;
  double* dst = sat + ((long long)(z + 1) * (ny + 1) + (y + 1)) * (nx + 1);
  double carry = 0.0;
  if (lane == 0) dst[0] = 0.0;
  for (int x0 = 0; x0 < nx; x0 += 32) {
    const int x = x0 + lane;
    double v = 0.0;
    if (x < nx) { const double f = (double)src[x]; v = f * f; }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double n = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += n;
    }
    v += carry;
    if (x < nx) dst[x + 1] = v;
    carry = __shfl_sync(0xffffffffu, v, 31);
  }
}
__global__ void k_sat_y(double* __restrict__ sat, int nx, int ny, int nz) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;  // 0..nx
  const int z = blockIdx.y;                             // 0..nz-1 -> plane z+1
  if (x > nx) return;
  double* pl = sat + (long long)(z + 1) * (ny + 1) * (nx + 1);
  double run = 0.0;
  pl[x] = 0.0;
  for (int y = 1; y <= ny; ++y) {
    run += pl[(long long)y * (nx + 1) + x];
    pl[(long long)y * (nx + 1) + x] = run;
  }
}
__global__ void k_sat_z(double* __restrict__ sat, int nx, int ny, int nz) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long plane = (long long)(ny + 1) * (nx + 1);
  if (i >= plane) return;
  double run = 0.0;
  sat[i] = 0.0;
  for (int z = 1; z <= nz; ++z) {
    run += sat[z * plane + i];
    sat[z * plane + i] = run;
  }
}

cudaError_t launch_sat_build(const float* img, double* sat, int nx, int ny, int nz, cudaStream_t s) {
  const long long rows = (long long)ny * nz;
  k_sat_rows<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(img, sat, nx, ny, nz);
  k_sat_y<<<dim3((nx + 1 + 127) / 128, nz), 128, 0, s>>>(sat, nx, ny, nz);
  const long long plane = (long long)(ny + 1) * (nx + 1);
  k_sat_z<<<(unsigned)((plane + 255) / 256), 256, 0, s>>>(sat, nx, ny, nz);
  return cudaGetLastError();
}

__global__ void k_a2map(const double* __restrict__ sat, int nx, int ny, int nz, const BoxDesc* __restrict__ boxes,
                        int nbox, float* __restrict__ a2, int nxo, int nyo, int nzo) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long npos = (long long)nxo * nyo * nzo;
  if (p >= npos) return;
  const int x = (int)(p % nxo);
  const long long q = p / nxo;
  const int y = (int)(q % nyo), z = (int)(q / nyo);
  const long long sx = 1, sy = nx + 1, sz = (long long)(nx + 1) * (ny + 1);
  double sum = 0.0;
  for (int b = 0; b < nbox; ++b) {
    const BoxDesc bx = boxes[b];
    const long long x0 = x + bx.x0, x1 = x0 + bx.w, y0 = y + bx.y0, y1 = y0 + bx.h, z0 = z + bx.z0, z1 = z0 + bx.d;
    sum += sat[z1 * sz + y1 * sy + x1 * sx] - sat[z1 * sz + y1 * sy + x0 * sx] - sat[z1 * sz + y0 * sy + x1 * sx] +
           sat[z1 * sz + y0 * sy + x0 * sx] - sat[z0 * sz + y1 * sy + x1 * sx] + sat[z0 * sz + y1 * sy + x0 * sx] +
           sat[z0 * sz + y0 * sy + x1 * sx] - sat[z0 * sz + y0 * sy + x0 * sx];
  }
  a2[p] = (float)sum;
}

cudaError_t launch_a2map(const double* sat, int nx, int ny, int nz, const BoxDesc* boxes, int nbox, float* a2, int nxo,
                         int nyo, int nzo, cudaStream_t s) {
  const long long npos = (long long)nxo * nyo * nzo;
  k_a2map<<<(unsigned)((npos + 255) / 256), 256, 0, s>>>(sat, nx, ny, nz, boxes, nbox, a2, nxo, nyo, nzo);
  return cudaGetLastError();
}

__global__ void k_fill_u32(unsigned* p, unsigned v, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
cudaError_t dmalloc(void** p, size_t bytes) {
  static thread_local int configured = -1;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (configured != dev) {
    cudaMemPool_t pool;
    if ((e = cudaDeviceGetDefaultMemPool(&pool, dev)) != cudaSuccess) return e;
    unsigned long long thr = ~0ull;
    if ((e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr)) != cudaSuccess) return e;
    configured = dev;
  }
  if ((e = cudaMallocAsync(p, bytes ? bytes : 1, cudaStreamPerThread)) != cudaSuccess) return e;
  return cudaStreamSynchronize(cudaStreamPerThread);
}

cudaError_t launch_fill_u32(unsigned* p, unsigned v, long long n, cudaStream_t s) {
  k_fill_u32<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p, v, n);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// radix select on the 64-bit key (float bits << 32 | position): values are >= 0 or +Inf so the
// unsigned order of the bits is the numeric order; the low word breaks ties by ascending position,
// exactly the Perm ordering of Base.partialsortperm.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long make_key(float v, long long p) {
  return ((unsigned long long)__float_as_uint(v) << 32) | (unsigned long long)p;
}
// key of the range-adaptive select (SelJob::sub); values below sub do not occur, values above the range (+Inf of
// disabled patches) keep high bits set and are filtered by the initial mask
__device__ __forceinline__ unsigned long long make_key_sub(float v, long long p, unsigned sub) {
  return ((unsigned long long)(__float_as_uint(v) - sub) << 32) | (unsigned long long)p;
}
// shift of pass `pass` of a job and its total number of passes (the context's schedule: 4 value bytes, then the
// position bytes; an adaptive job replaces the value bytes by its own nv digits)
__device__ __forceinline__ int job_shift(const SelJob* J, const int* __restrict__ shifts, int nshift, int pass, int* total) {
  if (J->nv == 0) { *total = nshift; return shifts[pass]; }
  *total = J->nv + (nshift - 4);
  return pass < J->nv ? J->vshift[pass] : shifts[4 + pass - J->nv];
}

// Grid = (blocks per job, jobs).  Inactive jobs (finished early, later relaxation rounds) return at once; the launcher
// keeps the blocks per job low when there are many jobs so that such rows stay cheap.
// Scratch of a select sequence: [0, njobs) active jobs, [njobs] their count, [njobs + 1, 2 njobs + 1) jobs marked for the
// gather (compact == 2), [2 njobs + 1] their count, [2 njobs + 2] CTAs of the running pass that have finished.
// Every CTA of a pass works from the same snapshot of the active jobs: it is rebuilt by the LAST CTA of the previous pass
// (and by k_select_list before the first one), so the later relaxation rounds -- whose passes find the list empty --
// cost a handful of near-empty launches instead of grids of tens of thousands of idle CTAs.
__device__ void select_build_lists(const SelJob* jobs, int njobs, int* scratch) {
  __shared__ int s_n[2];
  if (threadIdx.x < 2) s_n[threadIdx.x] = 0;
  __syncthreads();
  for (int j0 = 0; j0 < njobs; j0 += blockDim.x) {
    const int j = j0 + threadIdx.x;
    if (j < njobs && ((volatile const SelJob*)jobs)[j].active != 0) {
      scratch[atomicAdd(&s_n[0], 1)] = j;
      if (((volatile const SelJob*)jobs)[j].compact == 2) scratch[njobs + 1 + atomicAdd(&s_n[1], 1)] = j;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    scratch[njobs] = s_n[0];
    scratch[2 * njobs + 1] = s_n[1];
    scratch[2 * njobs + 2] = 0;
    __threadfence();
  }
}

__global__ void k_select_list(const SelJob* __restrict__ jobs, int njobs, int* __restrict__ scratch) {
  select_build_lists(jobs, njobs, scratch);
}

__global__ void __launch_bounds__(256) k_select_pass(SelJob* jobs, int njobs, int* scratch, int nb, long long npos,
                                                     const int* __restrict__ shifts, int nshift) {
  __shared__ unsigned h[8][256];  // one histogram per warp: the digits of an adaptive first pass spread over all bins
  __shared__ int s_active, s_pass, s_last;
  __shared__ unsigned long long s_prefix, s_mask;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int* list = scratch;
  const int nact = scratch[njobs];
  const int nwork = nact * nb;  // work item = (listed job, virtual block of nb), job fastest: the few blocks a short
                                // survivor list needs (vb < nbe below) are then spread over all CTAs
  for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
    const int vb = w / nact;
    SelJob* J = jobs + list[w % nact];
    __syncthreads();  // the previous item's shared state is dead
    if (tid == 0) {
      s_active = J->active;
      s_pass = J->pass;
      s_prefix = J->prefix;
      s_mask = J->mask;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) h[q][tid] = 0;
    __syncthreads();
    if (!s_active) continue;
    int npass_job;
    const int shift = job_shift(J, shifts, nshift, s_pass, &npass_job);
    const unsigned sub = J->sub;
    const unsigned long long prefix = s_prefix, mask = s_mask;
    const float* __restrict__ map = J->map;
    const bool comp = J->compact != 0;  // scan the compacted survivors instead of the map
    const unsigned long long* __restrict__ cbuf = J->cbuf;
    const long long n = comp ? (long long)J->ccount : npos;
    // a short survivor list does not need all nb virtual blocks: every block has a fixed cost (state loads, histogram
    // clear / merge, ticket) that dominated the list passes (650 us per pass for 128 jobs x 148 blocks on ~1 MB of keys)
    const int nbe = comp ? (int)max(1ll, min((long long)nb, (n + 8191) / 8192)) : nb;
    if (vb >= nbe) continue;
    // warp-aggregated histogram: distance values cluster in a few exponent bins, so per-lane shared-memory
    // atomics would serialise 32-way; lanes with the same bin elect one leader that adds their count
    const int lane = tid & 31;
    // four independent loads per thread and iteration (the pass streams whole maps: one 4-byte load in flight per
    // thread reached 0.7 TB/s); the trip count is warp-uniform (the vote below needs every lane)
    const long long stride = (long long)nbe * blockDim.x;
    const bool vote = s_pass == 0 && J->nv == 0;
    for (long long i0 = (long long)vb * blockDim.x + (tid & ~31); i0 < n; i0 += 4 * stride) {
      unsigned long long key[4];
      bool in[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long i = i0 + u * stride + lane;
        in[u] = i < n;
        key[u] = 0;
        if (in[u]) key[u] = comp ? cbuf[i] : make_key_sub(map[i], i, sub);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (i0 + u * stride >= n) break;  // warp-uniform
        unsigned bin = 0xffffffffu;
        if (in[u] && (key[u] & mask) == prefix) bin = (unsigned)(key[u] >> shift) & 255u;
        if (vote) {
          // first digit of the plain schedule (sign + high exponent bits): nearly every key falls into one or two bins --
          // aggregate per warp
          const unsigned peers = __match_any_sync(0xffffffffu, bin);
          if (bin != 0xffffffffu && lane == (__ffs(peers) - 1)) atomicAdd(&h[warp][bin], (unsigned)__popc(peers));
        } else if (bin != 0xffffffffu) {
          atomicAdd(&h[warp][bin], 1u);  // digits that spread over the bins: plain shared-memory atomics are cheaper than the vote
        }
      }
    }
    __syncthreads();
    {
      unsigned tot = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) tot += h[q][tid];
      if (tot) atomicAdd(&J->hist[tid], tot);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&J->ticket, 1u) == (unsigned)nbe - 1u);
    __syncthreads();
    if (!s_last) continue;
    if (tid == 0) {
      __threadfence();
      volatile unsigned* gh = J->hist;
      unsigned long long k = J->k, cum = 0, c = 0;
      int b = 0;
      for (; b < 256; ++b) {
        c = gh[b];
        if (cum + c >= k) break;
        cum += c;
      }
      if (b == 256) { b = 255; }  // k beyond the population: clamp (host never asks for this)
      k -= cum;
      const unsigned long long np = prefix | ((unsigned long long)b << shift);
      const bool exact = (c == k);  // every key sharing the new prefix is selected
      if (exact || s_pass + 1 == npass_job) {
        // back to plain keys (float bits << 32 | position): that is what the pick kernels compare with
        J->kth = (exact ? (np | ((shift == 0) ? 0ull : ((1ull << shift) - 1ull))) : np) + ((unsigned long long)sub << 32);
        J->active = 0;
      } else {
        J->mask = ~((1ull << shift) - 1ull);  // every bit from this digit upwards is decided
        J->pass = s_pass + 1;
        // c keys share the decided bits: once they fit (after the first pass of an adaptive job, after the second
        // otherwise), k_select_compact gathers them and the remaining passes scan that list
        if (J->compact == 0 && s_pass <= 1 && (J->nv > 0 || s_pass == 1) && npass_job > 3 && J->cbuf && c <= (unsigned long long)J->ccap) {
          J->compact = 2;
          J->ccount = 0;
        }
      }
      J->prefix = np;
      J->k = k;
      for (int i = 0; i < 256; ++i) gh[i] = 0;
      J->ticket = 0;
      __threadfence();
    }
  }
  // the last CTA of the pass snapshots the jobs for the next launches (next pass / gather)
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    s_last = (atomicAdd(&scratch[2 * njobs + 2], 1) == (int)gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) select_build_lists(jobs, njobs, scratch);
}

// Gathers the keys that match the decided prefix of every job marked by the second pass (compact == 2) into its
// scratch list (order is irrelevant: keys are unique) and switches the job to list mode (compact = 1).
__global__ void __launch_bounds__(256) k_select_compact(SelJob* jobs, int njobs, const int* __restrict__ scratch, int nb,
                                                        long long npos) {
 const int* list = scratch + njobs + 1;
 const int nact = scratch[2 * njobs + 1];
 const int nwork = nact * nb;
 for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
  const int vb = w / nact;
  SelJob* J = jobs + list[w % nact];
  if (!J->active || J->compact != 2) continue;
  const unsigned long long prefix = J->prefix, mask = J->mask;
  const unsigned sub = J->sub;
  const float* __restrict__ map = J->map;
  unsigned long long* __restrict__ cbuf = J->cbuf;
  const int tid = threadIdx.x, lane = tid & 31;
  const long long stride = (long long)nb * blockDim.x;
  for (long long i0 = (long long)vb * blockDim.x + (tid & ~31); i0 < npos; i0 += 4 * stride) {
    unsigned long long key[4];
    bool hit[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {  // four independent loads in flight
      const long long i = i0 + u * stride + lane;
      key[u] = 0;
      hit[u] = false;
      if (i < npos) {
        key[u] = make_key_sub(map[i], i, sub);
        hit[u] = (key[u] & mask) == prefix;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const unsigned bal = __ballot_sync(0xffffffffu, hit[u]);
      if (bal) {
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(&J->ccount, (unsigned)__popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (hit[u]) cbuf[base + __popc(bal & ((1u << lane) - 1u))] = key[u];
      }
    }
  }
 }
}
// one thread per job flips compact 2 -> 1 after the gather (a separate launch orders it after every block above)
__global__ void k_select_compact_done(SelJob* jobs, int njobs) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < njobs && jobs[j].compact == 2) jobs[j].compact = 1;
}

// scratch of the select launches: job list + count (one per stream would be needed for concurrent selects on one
// device from different contexts; the list lives next to the jobs: jobs[njobs] is followed by nothing we own, so the
// caller provides it)
static int select_nb(int njobs, long long npos) {
  long long nb = (npos + 256 * 8 - 1) / (256 * 8);
  const long long cap = njobs >= 16 ? 148 : 148 * 4;
  if (nb > cap) nb = cap;
  if (nb < 1) nb = 1;
  return (int)nb;
}

cudaError_t launch_select_all(SelJob* jobs, int njobs, long long npos, const int* shifts, int nshift, cudaStream_t s,
                              int* launches, int* scratch) {
  const int nb = select_nb(njobs, npos);
  const int grid = 148 * 6;  // persistent CTAs over (listed job, virtual block) work items (8 KB of histograms each)
  int nl = 0;
  k_select_list<<<1, 256, 0, s>>>(jobs, njobs, scratch);
  ++nl;
  for (int pass = 0; pass < nshift; ++pass) {
    k_select_pass<<<grid, 256, 0, s>>>(jobs, njobs, scratch, nb, npos, shifts, nshift);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    ++nl;
    if (pass <= 1 && nshift > 3) {  // gather point of adaptive jobs (after pass 0) and of plain ones (after pass 1)
      k_select_compact<<<grid, 256, 0, s>>>(jobs, njobs, scratch, nb, npos);
      k_select_compact_done<<<(njobs + 127) / 128, 128, 0, s>>>(jobs, njobs);
      if ((e = cudaGetLastError()) != cudaSuccess) return e;
      nl += 2;
    }
  }
  if (launches) *launches += nl;
  return cudaSuccess;
}

// ------------------------------------------------------------------------------------------------
// Position-slice mode (second axis), relaxation path: one 8-bit digit histogram of the VALUE bits of a
// local map slab, restricted to the values whose higher digits equal `prefix`.  The host sums the histograms of all
// ranks (all-gather), which turns the radix select of relaxation.jl:12,27 into a select over every slab at once; ties
// on the k-th value are resolved by position with a local select (iq_slice_kth), the slabs being ordered by rank.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_slice_hist(const SliceHistParams P) {
  const SliceHistReq rq = P.req[blockIdx.y];
  __shared__ unsigned h[8][256];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 8 * 256; i += 256) (&h[0][0])[i] = 0;
  __syncthreads();
  const int hi_shift = 32 - 8 * rq.level;  // bits above the current digit (level 0: none)
  const int shift = 24 - 8 * rq.level;
  for (long long q = (long long)blockIdx.x * 256; q < P.npos; q += (long long)gridDim.x * 256) {  // block-uniform trip count
    const long long p = q + tid;
    unsigned bin = 0xffffffffu;
    if (p < P.npos) {
      const unsigned bits = __float_as_uint(rq.map[p]);
      if (rq.level == 0 || (bits >> hi_shift) == rq.prefix) bin = (bits >> shift) & 255u;
    }
    // neighbouring positions carry similar distances: aggregate equal bins of a warp before the shared-memory atomic
    const unsigned peers = __match_any_sync(0xffffffffu, bin);
    if (bin != 0xffffffffu && lane == (__ffs(peers) - 1)) atomicAdd(&h[warp][bin], (unsigned)__popc(peers));
  }
  __syncthreads();
  unsigned t = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) t += h[w][tid];
  if (t) atomicAdd(&P.out[(size_t)blockIdx.y * 256 + tid], (unsigned long long)t);
}

cudaError_t launch_slice_hist(const SliceHistParams& P, cudaStream_t s) {
  long long nb = (P.npos + 256 * 16 - 1) / (256 * 16);
  nb = std::max<long long>(1, std::min<long long>(nb, 148 * 4));
  k_slice_hist<<<dim3((unsigned)nb, P.nreq), 256, 0, s>>>(P);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// candidate predicate + ordered compaction
// ------------------------------------------------------------------------------------------------
constexpr int kPickChunk = 4096;  // positions per CTA (256 threads x 16)

int pick_nblk(long long npos) { return (int)((npos + kPickChunk - 1) / kPickChunk); }

__device__ __forceinline__ bool pick_pred(const PickJob& J, long long p, double thr) {
  if (J.mode == 0) return (double)J.src[0][p] <= thr;
  // most selective source first: the auxiliary sets hold a tenth of the primary one (relaxation.jl:20-22), and the
  // auxiliary maps of a resident simulation are shared by all realizations (L2 hits) -- the per-realization primary
  // map is then read only where the auxiliary tests pass
  bool ok = true;
  for (int s = J.nsrc - 1; s >= 0 && ok; --s) ok = make_key(J.src[s][p], p) <= J.sel[s].kth;
  return ok;
}

__global__ void __launch_bounds__(256) k_pick_count(PickJob* jobs, long long npos) {
  PickJob& J = jobs[blockIdx.y];
  if (J.pending && *J.pending == 0) return;
  const int tid = threadIdx.x;
  const double thr = (J.mode == 0) ? (1.0 + J.tol) * (double)__uint_as_float(*J.minbits) : 0.0;
  const long long base = (long long)blockIdx.x * kPickChunk;
  unsigned cnt = 0;
#pragma unroll 4
  for (int it = 0; it < kPickChunk / 256; ++it) {
    const long long p = base + it * 256 + tid;
    if (p < npos && pick_pred(J, p, thr)) ++cnt;
  }
  __shared__ unsigned s_w[8];
  __shared__ int s_last;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((tid & 31) == 0) s_w[tid >> 5] = cnt;
  __syncthreads();
  if (tid == 0) {
    unsigned t = 0;
    for (int w = 0; w < 8; ++w) t += s_w[w];
    J.blockcount[blockIdx.x] = t;
    __threadfence();
    s_last = (atomicAdd(&J.ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  // last CTA of this job: exclusive scan of the block counts (in place) + total
  __threadfence();
  volatile unsigned* bc = J.blockcount;
  __shared__ unsigned s_scan[256];
  __shared__ unsigned s_carry;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  const int nblk = gridDim.x;
  for (int b0 = 0; b0 < nblk; b0 += 256) {
    const int i = b0 + tid;
    const unsigned v = (i < nblk) ? bc[i] : 0u;
    s_scan[tid] = v;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {
      const unsigned n = (tid >= o) ? s_scan[tid - o] : 0u;
      __syncthreads();
      s_scan[tid] += n;
      __syncthreads();
    }
    const unsigned incl = s_scan[tid];
    const unsigned carry = s_carry;
    if (i < nblk) bc[i] = carry + incl - v;
    __syncthreads();
    if (tid == 255) s_carry = carry + incl;
    __syncthreads();
  }
  if (tid == 0) {
    *J.total = s_carry;
    J.ticket = 0;
    __threadfence();
  }
}

__global__ void __launch_bounds__(256) k_pick_write(PickJob* jobs, long long npos) {
  const PickJob& J = jobs[blockIdx.y];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double thr = (J.mode == 0) ? (1.0 + J.tol) * (double)__uint_as_float(*J.minbits) : 0.0;
  // every warp owns kPickChunk / 8 consecutive positions of the CTA's chunk: the predicate is evaluated once (one bit
  // per iteration kept in a register), the warp totals are scanned with ONE barrier, then every warp writes its
  // candidates in ascending position (the first version took two barriers per 256 positions)
  constexpr int kPerWarp = kPickChunk / 8, kIters = kPerWarp / 32;
  const long long base = (long long)blockIdx.x * kPickChunk + (long long)warp * kPerWarp;
  __shared__ unsigned s_w[8];
  unsigned mybits = 0, wcount = 0;
#pragma unroll 4
  for (int it = 0; it < kIters; ++it) {
    const long long p = base + it * 32 + lane;
    const bool ok = (p < npos) && pick_pred(J, p, thr);
    mybits |= (ok ? 1u : 0u) << it;
    wcount += __popc(__ballot_sync(0xffffffffu, ok));
  }
  if (lane == 0) s_w[warp] = wcount;
  __syncthreads();
  unsigned running = J.blockcount[blockIdx.x];
  for (int w = 0; w < warp; ++w) running += s_w[w];
  if (wcount == 0) return;
  for (int it = 0; it < kIters; ++it) {
    const bool ok = (mybits >> it) & 1u;
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    if (ok) {
      const long long p = base + it * 32 + lane;
      const long long slot = (long long)running + __popc(bal & ((1u << lane) - 1u));
      if (slot < J.cap) {
        J.cand_idx[slot] = (unsigned)p;
        for (int s = 0; s < J.nsrc; ++s) J.cand_val[(long long)s * J.cap + slot] = J.src[s][p];
      }
    }
    running += __popc(bal);
  }
}

// Threshold rule (iqsim.jl:237) driven by chunk minima: one CTA per tile.  Phase 1 compacts, in ascending order, the
// chunks whose minimum passes the threshold; phase 2 counts the passing entries of those chunks (one warp per
// chunk); phase 3 writes the candidates in ascending linear index.  A map is read only where candidates can be.
__global__ void __launch_bounds__(1024) k_pick_chunks(PickJob* jobs, long long npos, int chunklen, int nchunk) {
  PickJob& J = jobs[blockIdx.x];
  if (J.mode != 0 || !J.chunkmin) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double thr = (1.0 + J.tol) * (double)__uint_as_float(*J.minbits);
  __shared__ unsigned s_rows[kPickMaxChunks];
  __shared__ unsigned s_cnt[kPickMaxChunks];
  __shared__ unsigned s_w[32];
  __shared__ unsigned s_nq;
  if (tid == 0) s_nq = 0;
  __syncthreads();
  for (int base = 0; base < nchunk; base += 1024) {
    const int ch = base + tid;
    const bool q = ch < nchunk && (double)__uint_as_float(J.chunkmin[ch]) <= thr;
    const unsigned bal = __ballot_sync(0xffffffffu, q);
    if (lane == 0) s_w[warp] = __popc(bal);
    __syncthreads();
    unsigned before = 0, all = 0;
    for (int w = 0; w < 32; ++w) {
      const unsigned c = s_w[w];
      if (w < warp) before += c;
      all += c;
    }
    const unsigned nq0 = s_nq;
    if (q) {
      const unsigned slot = nq0 + before + __popc(bal & ((1u << lane) - 1u));
      if (slot < (unsigned)kPickMaxChunks) s_rows[slot] = (unsigned)ch;
    }
    __syncthreads();
    if (tid == 0) s_nq = nq0 + all;
    __syncthreads();
  }
  const unsigned nq = s_nq;
  if (nq > (unsigned)kPickMaxChunks) {
    if (tid == 0) *J.total = kPickOverflow;
    return;
  }
  const float* map = J.src[0];
  for (unsigned k = warp; k < nq; k += 32) {
    const long long p0 = (long long)s_rows[k] * chunklen;
    const int len = (int)min((long long)chunklen, npos - p0);
    unsigned c = 0;
    for (int x = lane; x < len; x += 32) c += ((double)map[p0 + x] <= thr) ? 1u : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) s_cnt[k] = c;
  }
  __syncthreads();
  // exclusive scan of the chunk counts (nq <= 2048: two entries per thread)
  {
    const unsigned a0 = (2u * tid < nq) ? s_cnt[2 * tid] : 0u, a1 = (2u * tid + 1u < nq) ? s_cnt[2 * tid + 1] : 0u;
    unsigned incl = a0 + a1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      unsigned v = s_w[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
      }
      s_w[lane] = v;
    }
    __syncthreads();
    const unsigned excl = incl - (a0 + a1) + (warp > 0 ? s_w[warp - 1] : 0u);
    if (2u * tid < nq) s_cnt[2 * tid] = excl;
    if (2u * tid + 1u < nq) s_cnt[2 * tid + 1] = excl + a0;
    if (tid == 1023) *J.total = s_w[31];
  }
  __syncthreads();
  for (unsigned k = warp; k < nq; k += 32) {
    const long long p0 = (long long)s_rows[k] * chunklen;
    const int len = (int)min((long long)chunklen, npos - p0);
    unsigned off = s_cnt[k];
    for (int x0 = 0; x0 < len; x0 += 32) {
      const int x = x0 + lane;
      const float v = x < len ? map[p0 + x] : 0.f;
      const bool ok = x < len && (double)v <= thr;
      const unsigned bal = __ballot_sync(0xffffffffu, ok);
      if (ok) {
        const long long slot = (long long)off + __popc(bal & ((1u << lane) - 1u));
        if (slot < J.cap) {
          J.cand_idx[slot] = (unsigned)(p0 + x);
          J.cand_val[slot] = v;
        }
      }
      off += __popc(bal);
    }
  }
}

cudaError_t launch_pick_chunks(PickJob* jobs, int njobs, long long npos, int chunklen, int nchunk, cudaStream_t s) {
  k_pick_chunks<<<njobs, 1024, 0, s>>>(jobs, npos, chunklen, nchunk);
  return cudaGetLastError();
}

// Chunk minima (float bits) of njobs maps stored [job][npos]: one warp per chunk.  +Inf (disabled) entries never win.
__global__ void __launch_bounds__(256) k_chunkmin(const float* __restrict__ maps, long long npos, int chunklen, int nchunk,
                                                  unsigned* __restrict__ chunkmin, long long pitch) {
  const int ch = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (ch >= nchunk) return;
  const long long p0 = (long long)ch * chunklen;
  const int len = (int)min((long long)chunklen, npos - p0);
  const float* m = maps + (long long)blockIdx.y * npos + p0;
  unsigned mn = 0x7f800000u;
  for (int x = lane; x < len; x += 32) mn = min(mn, __float_as_uint(m[x]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  if (lane == 0) chunkmin[(long long)blockIdx.y * pitch + ch] = mn;
}

cudaError_t launch_chunkmin(const float* maps, int njobs, long long npos, int chunklen, int nchunk, unsigned* chunkmin,
                            long long pitch, cudaStream_t s) {
  k_chunkmin<<<dim3((unsigned)((nchunk + 7) / 8), njobs), 256, 0, s>>>(maps, npos, chunklen, nchunk, chunkmin, pitch);
  return cudaGetLastError();
}

cudaError_t launch_pick_count(PickJob* jobs, int njobs, long long npos, cudaStream_t s) {
  k_pick_count<<<dim3(pick_nblk(npos), njobs), 256, 0, s>>>(jobs, npos);
  return cudaGetLastError();
}
cudaError_t launch_pick_write(PickJob* jobs, int njobs, long long npos, cudaStream_t s) {
  k_pick_write<<<dim3(pick_nblk(npos), njobs), 256, 0, s>>>(jobs, npos);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// tau model on the device (src/taumodel.jl:5-45) for candidate sets of 2 .. kTauMax entries:
//   k_tau_rank : one CTA per (tile, source): bitonic sort of (value bits << 32 | slot) keys in shared memory,
//                dense ranks (ties share a rank, taumodel.jl:22-31) scattered back to candidate order, and the
//                integer column sum of (n - r + 1) (taumodel.jl:34-35)
//   k_tau_prob : per candidate, the FP64 combination of taumodel.jl:34-44 with every operation individually
//                rounded (no FMA contraction) so that the result is bit-identical to the host/oracle code
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_tau_rank(const PickJob* jobs, unsigned* rank_out, unsigned long long* colsum,
                                                   int maxS) {
  const PickJob& J = jobs[blockIdx.y];
  const int s = blockIdx.x;
  const unsigned n = *J.total;
  if (s >= J.nsrc || n < 2u || n > (unsigned)kTauMax) return;
  // value bits (keys) and candidate slots (payload) in separate arrays: 6 bytes per entry, 32 768 entries in 192 KB.
  // Ties need no tie-break: equal values share a dense rank whatever their order.
  extern __shared__ __align__(16) unsigned char tau_sm[];
  unsigned* keys = reinterpret_cast<unsigned*>(tau_sm);
  unsigned short* slot = reinterpret_cast<unsigned short*>(keys + kTauMax);
  __shared__ unsigned s_part[32];
  __shared__ unsigned long long s_sum;
  unsigned NP = 2;
  while (NP < n) NP <<= 1;
  const int tid = threadIdx.x;
  const float* vals = J.cand_val + (long long)s * J.cap;
  for (unsigned i = tid; i < NP; i += 1024) {
    keys[i] = i < n ? __float_as_uint(vals[i]) : 0xffffffffu;
    slot[i] = (unsigned short)i;
  }
  if (tid == 0) s_sum = 0ull;
  __syncthreads();
  for (unsigned k = 2; k <= NP; k <<= 1) {
    for (unsigned j = k >> 1; j > 0; j >>= 1) {
      for (unsigned t = tid; t < NP / 2; t += 1024) {
        const unsigned i = 2 * t - (t & (j - 1));
        const unsigned ixj = i + j;
        const bool up = (i & k) == 0;
        const unsigned a = keys[i], b = keys[ixj];
        if ((a > b) == up && a != b) {
          keys[i] = b; keys[ixj] = a;
          const unsigned short sa = slot[i];
          slot[i] = slot[ixj]; slot[ixj] = sa;
        }
      }
      __syncthreads();
    }
  }
  // dense ranks: contiguous segment per thread, block scan of the segment sums
  const unsigned seg = (NP + 1023) / 1024;
  const unsigned k0 = tid * seg, k1 = min(k0 + seg, n);
  unsigned local = 0;
  for (unsigned k = k0; k < k1; ++k) local += (k == 0 || keys[k] != keys[k - 1]) ? 1u : 0u;
  unsigned incl = local;
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_part[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    unsigned v = s_part[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned u = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += u;
    }
    s_part[lane] = v;
  }
  __syncthreads();
  unsigned r = incl - local + (warp > 0 ? s_part[warp - 1] : 0u);
  unsigned long long mysum = 0ull;
  unsigned* ro = rank_out + ((long long)blockIdx.y * maxS + s) * kTauMax;
  for (unsigned k = k0; k < k1; ++k) {
    if (k == 0 || keys[k] != keys[k - 1]) ++r;
    ro[slot[k]] = r;
    mysum += (unsigned long long)(n - r + 1u);
  }
  atomicAdd(&s_sum, mysum);
  __syncthreads();
  if (tid == 0) colsum[blockIdx.y * maxS + s] = s_sum;
}

__global__ void __launch_bounds__(256) k_tau_prob(const PickJob* jobs, const unsigned* rank, const unsigned long long* colsum,
                                                  double* prob, int maxS) {
  const PickJob& J = jobs[blockIdx.y];
  const unsigned n = *J.total;
  if (n < 2u || n > (unsigned)kTauMax) return;
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double dn = (double)n;
  const double inv = __ddiv_rn(1.0, dn);
  const double x0 = __ddiv_rn(__dsub_rn(1.0, inv), inv);  // (1 - 1/n) / (1/n), taumodel.jl:38
  double prod = 1.0;
  for (int s = 0; s < J.nsrc; ++s) {
    const double r = (double)rank[((long long)blockIdx.y * maxS + s) * kTauMax + i];
    const double P = __dadd_rn(__dsub_rn(dn, r), 1.0);                    // nevents - D + 1
    const double Pi = __ddiv_rn(P, (double)colsum[blockIdx.y * maxS + s]);  // / sum(P, dims=1)
    const double X = __ddiv_rn(__dsub_rn(1.0, Pi), Pi);                   // (1 - P) / P
    const double ratio = __ddiv_rn(X, x0);
    prod = (s == 0) ? ratio : __dmul_rn(prod, ratio);
  }
  prob[(long long)blockIdx.y * kTauMax + i] = __ddiv_rn(1.0, __dadd_rn(1.0, __dmul_rn(x0, prod)));
}

cudaError_t launch_tau(const PickJob* jobs, int njobs, int maxS, unsigned* rank, unsigned long long* colsum, double* prob,
                       cudaStream_t s) {
  const size_t smem = (size_t)kTauMax * (sizeof(unsigned) + sizeof(unsigned short));
  cudaError_t e = cudaFuncSetAttribute(k_tau_rank, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_tau_rank<<<dim3(maxS, njobs), 1024, smem, s>>>(jobs, rank, colsum, maxS);
  k_tau_prob<<<dim3(kTauMax / 256, njobs), 256, 0, s>>>(jobs, rank, colsum, prob, maxS);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// view_kernel (utils.jl:63-67): gather one tile-sized patch of the resident image
// ------------------------------------------------------------------------------------------------
__global__ void k_fetch_tile(const float* __restrict__ img, int nx, int ny, int nz, int tx, int ty, int tz,
                             long long x0, long long y0, long long z0, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)tx * ty * tz;
  if (i >= n) return;
  const int qx = (int)(i % tx);
  const long long q = i / tx;
  const int qy = (int)(q % ty), qz = (int)(q / ty);
  out[i] = img[((z0 + qz) * ny + (y0 + qy)) * nx + (x0 + qx)];
}
cudaError_t launch_fetch_tile(const float* img, int nx, int ny, int nz, int tx, int ty, int tz, long long x0,
                              long long y0, long long z0, float* out, cudaStream_t s) {
  const long long n = (long long)tx * ty * tz;
  k_fetch_tile<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(img, nx, ny, nz, tx, ty, tz, x0, y0, z0, out);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// FP32 FMA peak microbenchmark: 16 independent register-operand FFMA chains per thread
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fma_peak(int iters, float s, float t, float* out) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = (float)(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 8; ++rep)
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], s, t);
  }
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) sum += a[i];
  if (sum == 123.456f) out[0] = sum;  // never true in practice; keeps the chains alive
}
// packed variant: fma.rn.f32x2 (sm_100+), two FMAs per instruction on a 64-bit register pair
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__global__ void __launch_bounds__(256) k_fma2_peak(int iters, float s, float t, float* out) {
  unsigned long long a[16];
  const unsigned long long s2 = ((unsigned long long)__float_as_uint(s) << 32) | __float_as_uint(s);
  const unsigned long long t2 = ((unsigned long long)__float_as_uint(t) << 32) | __float_as_uint(t);
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = ((unsigned long long)__float_as_uint((float)(threadIdx.x + i)) << 32) | (unsigned)i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 4; ++rep)
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fma2(a[i], s2, t2);
  }
  unsigned long long x = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) x ^= a[i];
  if (x == 0x123456789abcull) out[0] = 1.f;
}
cudaError_t launch_fma_peak(int blocks, int iters, float* out, cudaStream_t s) {
  if (iters < 0) {  // negative iteration count selects the packed (f32x2) variant
    k_fma2_peak<<<blocks, 256, 0, s>>>(-iters, 0.999f, 0.001f, out);
  } else {
    k_fma_peak<<<blocks, 256, 0, s>>>(iters, 0.999f, 0.001f, out);
  }
  return cudaGetLastError();
}

}  // namespace iq

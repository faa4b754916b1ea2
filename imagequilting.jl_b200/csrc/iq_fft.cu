// iq_fft.cu -- hand-written shared-memory FFT cross-correlation for sm_100a (no cuFFT).
//
// Replaces the FFT route of the reference, imfilter(img, centered(krn), Inner(), Algorithm.FFT())
// (/root/reference/src/imfilter.jl:5-7) and its cuFFT twin real(ifft(fft(img) .* conj(fft(padkrn))))
// (src/imfilter.jl:14-25), for templates large enough that the direct kernel loses (measured crossover,
// DESIGN.md).  What is different from the reference's GPU path:
//   * the image spectrum is computed ONCE per context and cached (the reference re-transforms the image
//     for every tile, src/imfilter.jl:19);
//   * two real templates ride in one complex transform (re = template 2k, im = template 2k+1): correlation
//     with a real image is a real-linear operator, so the real/imaginary parts of the result are the two
//     correlations;
//   * transforms are pruned: the template only occupies tx*ty*tz of the padded volume and only the valid
//     region (distsize) of the result is needed, so most lines of the outer passes are never touched;
//   * the forward transform along the last axis, the spectrum product and the inverse transform along
//     that axis are one kernel; the last inverse pass writes |A2 - 2AB + B2| straight into the distance
//     maps (with the disabled knock-out and the min/max reduction) -- no padded kernel image, no crop
//     kernel, no device->host copy of the map (G3-G7 of SURVEY.md 2.1).
// Layout: complex float2, x fastest.  Sizes are padded to powers of two; Stockham autosort passes of radix
// 16 (then 8/4/2) with the butterflies in registers, lines in shared memory with a 1-in-16 padding that
// makes every pass and every global<->shared copy bank-conflict free.
#include "iq_fft.h"
#include "iq_internal.h"

#include "iq_tma.cuh"  // CUtensorMap, mbarrier / cp.async.bulk helpers; cuTensorMapEncodeTiled through cudaGetDriverEntryPoint
#include <math_constants.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <map>
#include <vector>

namespace iqfft {

constexpr int kThreads = 256;

__host__ __device__ constexpr int lines_per_block(int log2n) { return log2n >= 10 ? 4 : (log2n >= 9 ? 8 : 16); }
// shared-memory line layout: element i of a line lives at i + (i >> 4) (one float2 of padding per 16), lines
// are LS apart.  With this skew every access pattern of the radix-16/8/4/2 Stockham passes below and the
// line-fastest global<->shared copies is bank-conflict free.
__host__ __device__ constexpr int line_stride(int log2n) { return (1 << log2n) + ((1 << log2n) >> 4) + 1; }
__device__ __forceinline__ int PI(int i) { return i + (i >> 4); }
// In-place passes: when every pass of a transform is a single sweep of the CTA (lines_per_block * N / radix <= kThreads)
// each thread holds all its inputs in registers before anything is written, so source and destination can be the
// same buffer (one extra barrier per pass).  Halving the shared memory doubles the resident CTAs per SM, which is
// what these streaming kernels need to keep enough loads in flight.
__host__ __device__ constexpr bool inplace_ok(int log2n) {
  const int rem = log2n % 4;
  const int rmin = rem > 0 ? (1 << rem) : 16;
  return log2n >= 1 && lines_per_block(log2n) * ((1 << log2n) / rmin) <= kThreads;
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// multiply by exp(-+ 2 pi i m / 16) (minus sign = forward), compile-time m
template <int M, bool INV>
__device__ __forceinline__ float2 mul_w16(float2 v) {
  constexpr float c[8] = {1.f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f,
                          0.f, -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f};
  constexpr float sn[8] = {0.f, 0.38268343236508977f, 0.70710678118654752f, 0.92387953251128674f,
                           1.f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f};
  constexpr int m = M & 15;
  constexpr float wr = (m < 8) ? c[m] : -c[m - 8];
  constexpr float wi0 = (m < 8) ? -sn[m] : sn[m - 8];  // forward: exp(-i t) = cos t - i sin t
  constexpr float wi = INV ? -wi0 : wi0;
  if (m == 0) return v;
  return make_float2(v.x * wr - v.y * wi, v.x * wi + v.y * wr);
}

template <bool INV>
__device__ __forceinline__ void dft4(float2& a, float2& b, float2& c, float2& d) {
  const float2 apc = cadd(a, c), amc = csub(a, c), bpd = cadd(b, d), bmd = csub(b, d);
  const float2 jb = INV ? make_float2(bmd.y, -bmd.x) : make_float2(-bmd.y, bmd.x);  // forward: +i (b - d)
  a = cadd(apc, bpd);
  b = csub(amc, jb);
  c = csub(apc, bpd);
  d = cadd(amc, jb);
}

// R-point DFT in registers, natural order in and out (R = 2, 4, 8, 16)
template <int R, bool INV>
__device__ __forceinline__ void dft_reg(float2 (&v)[R]) {
  if constexpr (R == 2) {
    const float2 a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
  } else if constexpr (R == 4) {
    dft4<INV>(v[0], v[1], v[2], v[3]);
  } else if constexpr (R == 8) {
    // j = j0 + 2 j1, k = 4 k0 + k1
    float2 t0[4] = {v[0], v[2], v[4], v[6]}, t1[4] = {v[1], v[3], v[5], v[7]};
    dft4<INV>(t0[0], t0[1], t0[2], t0[3]);
    dft4<INV>(t1[0], t1[1], t1[2], t1[3]);
    t1[1] = mul_w16<2, INV>(t1[1]);
    t1[2] = mul_w16<4, INV>(t1[2]);
    t1[3] = mul_w16<6, INV>(t1[3]);
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
      v[k1] = cadd(t0[k1], t1[k1]);
      v[4 + k1] = csub(t0[k1], t1[k1]);
    }
  } else {
    // R == 16: j = j0 + 4 j1, k = 4 k0 + k1
    float2 t[4][4];
#pragma unroll
    for (int j0 = 0; j0 < 4; ++j0) {
      t[j0][0] = v[j0]; t[j0][1] = v[j0 + 4]; t[j0][2] = v[j0 + 8]; t[j0][3] = v[j0 + 12];
      dft4<INV>(t[j0][0], t[j0][1], t[j0][2], t[j0][3]);
    }
    t[1][1] = mul_w16<1, INV>(t[1][1]); t[1][2] = mul_w16<2, INV>(t[1][2]); t[1][3] = mul_w16<3, INV>(t[1][3]);
    t[2][1] = mul_w16<2, INV>(t[2][1]); t[2][2] = mul_w16<4, INV>(t[2][2]); t[2][3] = mul_w16<6, INV>(t[2][3]);
    t[3][1] = mul_w16<3, INV>(t[3][1]); t[3][2] = mul_w16<6, INV>(t[3][2]); t[3][3] = mul_w16<9, INV>(t[3][3]);
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
      dft4<INV>(t[0][k1], t[1][k1], t[2][k1], t[3][k1]);
      v[k1] = t[0][k1]; v[4 + k1] = t[1][k1]; v[8 + k1] = t[2][k1]; v[12 + k1] = t[3][k1];
    }
  }
}

// One Stockham autosort pass of radix R over LPB lines in shared memory (padded layout), butterflies in
// registers:  y[q + s (R p + k)] = w_n^(p k) * sum_j x[q + s (p + j n/R)] w_R^(j k).
// MUL: the inputs are multiplied by the spectrum tile `spec` (same layout) while they are read -- the
// spectrum product of the fused forward/inverse kernel costs no extra pass.
// LIN (in-place transforms only): the input of this pass is NOT in the padded line layout but where a bulk (TMA)
// copy dropped it: LIN = 1 lines of N consecutive elements, N apart; LIN = 2 element-major tiles
// [element][line] of LPB lines (threads then take the line index fastest so that the reads stay conflict free).
// The pass writes the padded layout, so every later pass is the ordinary one.
template <int LOG2N, int R, bool INV, bool MUL, int LIN = 0>
__device__ __forceinline__ void stockham_pass(const float2* src, float2* dst,
                                              const float2* __restrict__ tw, const float2* __restrict__ spec,
                                              int nlines, int n, int ls) {
  constexpr int N = 1 << LOG2N, LS = line_stride(LOG2N), NB = N / R;
  constexpr bool INPLACE = inplace_ok(LOG2N);  // src == dst: one sweep per thread, barrier between reads and writes
  const int s = 1 << ls, m = n / R, step = N / n;
  for (int idx = threadIdx.x; idx < (INPLACE ? kThreads : nlines * NB); idx += kThreads) {
    static_assert(LIN == 0 || (inplace_ok(LOG2N) && !MUL), "bulk-copy layouts need the in-place transform");
    constexpr int LPBC = lines_per_block(LOG2N);
    const bool live = idx < nlines * NB;
    int line, b;
    if (LIN == 2) { line = idx % LPBC; b = idx / LPBC; if (!live) { line = 0; b = 0; } }
    else { line = live ? idx / NB : 0; b = live ? idx - line * NB : 0; }
    const int p = b >> ls, q = b & (s - 1);
    const float2* xb = src + line * LS;
    float2* yb = dst + line * LS;
    float2 v[R];
    // Padded index of input j: e_j = (q + s p) + NBR j with q + s p < NBR = N / R.  When NBR is a multiple of 16 the
    // padding term grows by exactly NBR / 16 per j, so PI(e_j) = PI(q + s p) + (NBR + NBR / 16) j: one add, and the
    // j-dependent part folds into the load's immediate offset (the profile showed the index arithmetic of these passes
    // next to the butterflies in instruction count).
    constexpr int NBR = N / R;
    constexpr bool kInLinear = (NBR % 16) == 0;
    const int pin = PI(q + s * p);
    if (live) {
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const int e = q + s * (p + j * m);
        if (LIN == 1) v[j] = src[line * N + e];
        else if (LIN == 2) v[j] = src[e * LPBC + line];
        else {
          const int i = kInLinear ? pin + (NBR + NBR / 16) * j : PI(e);
          v[j] = xb[i];
          if (MUL) v[j] = cmul(v[j], spec[line * LS + i]);
        }
      }
    }
    if (INPLACE) __syncthreads();
    if (!live) continue;
    dft_reg<R, INV>(v);
    if (m > 1) {  // twiddles w_n^(p k): one table read, the powers by (shallow) repeated multiplication
      float2 w1 = tw[p * step];
      if (INV) w1.y = -w1.y;
      float2 w[R];
      w[1] = w1;
      if constexpr (R >= 4) { w[2] = cmul(w1, w1); w[3] = cmul(w[2], w1); }
      if constexpr (R >= 8) {
        w[4] = cmul(w[2], w[2]);
        w[5] = cmul(w[4], w[1]); w[6] = cmul(w[4], w[2]); w[7] = cmul(w[4], w[3]);
      }
      if constexpr (R >= 16) {
        w[8] = cmul(w[4], w[4]);
        w[9] = cmul(w[8], w[1]); w[10] = cmul(w[8], w[2]); w[11] = cmul(w[8], w[3]);
        w[12] = cmul(w[8], w[4]);
        w[13] = cmul(w[12], w[1]); w[14] = cmul(w[12], w[2]); w[15] = cmul(w[12], w[3]);
      }
#pragma unroll
      for (int k = 1; k < R; ++k) v[k] = cmul(v[k], w[k]);
    }
    // outputs k: e_k = (q + s R p) + s k.  s a multiple of 16: PI grows by s + s / 16 per k; s = 1 with R = 16: the base
    // is a multiple of 16 and k < 16, so PI(e_k) = PI(base) + k.
    const int pout = PI(q + s * R * p);
    if ((s & 15) == 0) {
      const int os = s + (s >> 4);
#pragma unroll
      for (int k = 0; k < R; ++k) yb[pout + os * k] = v[k];
    } else if (s == 1 && R == 16) {
#pragma unroll
      for (int k = 0; k < R; ++k) yb[pout + k] = v[k];
    } else {
#pragma unroll
      for (int k = 0; k < R; ++k) yb[PI(q + s * (R * p + k))] = v[k];
    }
  }
  __syncthreads();
}

// Full transform of `nlines` lines (ping-pong between two buffers); returns the buffer holding the result.
// Radix schedule: 16 while >= 4 bits remain, then 8 / 4 / 2 for the rest.
template <int LOG2N, bool INV, bool MUL, int LIN = 0>
__device__ float2* fft_lines(float2* src, float2* dst, const float2* __restrict__ tw, const float2* __restrict__ spec,
                             int nlines) {
  constexpr int N = 1 << LOG2N;
  constexpr int n16 = LOG2N / 4, rem = LOG2N % 4;
  int n = N, ls = 0;
  bool first = true;
  if constexpr (inplace_ok(LOG2N)) dst = src;
  if constexpr (n16 > 0) {
#pragma unroll
    for (int i = 0; i < n16; ++i) {
      if (MUL && first) stockham_pass<LOG2N, 16, INV, true>(src, dst, tw, spec, nlines, n, ls);
      else if (LIN != 0 && first) stockham_pass<LOG2N, 16, INV, false, LIN>(src, dst, tw, spec, nlines, n, ls);
      else stockham_pass<LOG2N, 16, INV, false>(src, dst, tw, spec, nlines, n, ls);
      first = false;
      float2* t = src; src = dst; dst = t;
      n >>= 4; ls += 4;
    }
  }
  if constexpr (rem > 0) {
    constexpr int R = 1 << rem;
    if (MUL && first) stockham_pass<LOG2N, R, INV, true>(src, dst, tw, spec, nlines, n, ls);
    else if (LIN != 0 && first) stockham_pass<LOG2N, R, INV, false, LIN>(src, dst, tw, spec, nlines, n, ls);
    else stockham_pass<LOG2N, R, INV, false>(src, dst, tw, spec, nlines, n, ls);
    float2* t = src; src = dst; dst = t;
  }
  return src;
}

template <int LOG2N>
__device__ __forceinline__ void load_twiddles(float2* tw_s, const float2* __restrict__ tw_g) {
  constexpr int N = 1 << LOG2N;
  for (int i = threadIdx.x; i < N; i += kThreads) tw_s[i] = tw_g[i];
}

template <int LOG2N>
__device__ __forceinline__ void zero_lines(float2* buf, int nlines) {
  constexpr int LS = line_stride(LOG2N);
  for (int i = threadIdx.x; i < nlines * LS; i += kThreads) buf[i] = make_float2(0.f, 0.f);
}

// ---- pass A: x lines built from a pair of real templates (flipped placement) -------------------------
struct TmplPassArgs {
  const float* tmpl;      // [R][tilevol]
  float2* out;            // [npair][nlines][N]
  int tx, nlines;         // nlines = ty*tz
  long long tilevol;
  int R;
  const float2* tw;
};
template <int LOG2N>
__global__ void __launch_bounds__(kThreads) k_fft_x_tmpl(const TmplPassArgs A) {
  constexpr int N = 1 << LOG2N, LS = line_stride(LOG2N), LPB = lines_per_block(LOG2N);
  extern __shared__ __align__(16) float2 sm[];
  float2* b0 = sm;
  float2* b1 = sm + LPB * LS;
  float2* tw = sm + (inplace_ok(LOG2N) ? 1 : 2) * LPB * LS;
  const int pr = blockIdx.y, l0 = blockIdx.x * LPB;
  const int nl = min(LPB, A.nlines - l0);
  load_twiddles<LOG2N>(tw, A.tw);
  zero_lines<LOG2N>(b0, nl);
  __syncthreads();
  const float* t0 = A.tmpl + (long long)(2 * pr) * A.tilevol;
  const bool has1 = (2 * pr + 1) < A.R;
  const float* t1 = A.tmpl + (long long)(2 * pr + 1) * A.tilevol;
  for (int i = threadIdx.x; i < nl * A.tx; i += kThreads) {
    const int line = i / A.tx, qx = i - line * A.tx;
    const long long src = (long long)(l0 + line) * A.tx + qx;
    b0[line * LS + PI((N - qx) & (N - 1))] = make_float2(t0[src], has1 ? t1[src] : 0.f);
  }
  __syncthreads();
  const float2* res = fft_lines<LOG2N, false, false>(b0, b1, tw, nullptr, nl);
  float2* out = A.out + ((long long)pr * A.nlines + l0) * N;
  for (int i = threadIdx.x; i < nl * N; i += kThreads) {
    const int line = i >> LOG2N, e = i & (N - 1);
    out[(long long)line * N + e] = res[line * LS + PI(e)];
  }
}

// ---- pass A': x lines of a real image (spectrum set-up) ----------------------------------------------
struct RealPassArgs {
  const float* img;  // [nlines][nx]
  float2* out;       // [nlines][N]
  int nx, nlines;
  const float2* tw;
};
template <int LOG2N>
__global__ void __launch_bounds__(kThreads) k_fft_x_real(const RealPassArgs A) {
  constexpr int N = 1 << LOG2N, LS = line_stride(LOG2N), LPB = lines_per_block(LOG2N);
  extern __shared__ __align__(16) float2 sm[];
  float2* b0 = sm;
  float2* b1 = sm + LPB * LS;
  float2* tw = sm + (inplace_ok(LOG2N) ? 1 : 2) * LPB * LS;
  const int l0 = blockIdx.x * LPB;
  const int nl = min(LPB, A.nlines - l0);
  load_twiddles<LOG2N>(tw, A.tw);
  zero_lines<LOG2N>(b0, nl);
  __syncthreads();
  for (int i = threadIdx.x; i < nl * A.nx; i += kThreads) {
    const int line = i / A.nx, x = i - line * A.nx;
    b0[line * LS + PI(x)] = make_float2(A.img[(long long)(l0 + line) * A.nx + x], 0.f);
  }
  __syncthreads();
  const float2* res = fft_lines<LOG2N, false, false>(b0, b1, tw, nullptr, nl);
  float2* out = A.out + (long long)l0 * N;
  for (int i = threadIdx.x; i < nl * N; i += kThreads) {
    const int line = i >> LOG2N, e = i & (N - 1);
    out[(long long)line * N + e] = res[line * LS + PI(e)];
  }
}

// ---- pass B: strided lines (y or z axis).  MODE 0 forward, 1 forward * spectrum -> inverse, 2 inverse --
// MODE 1 loops over the `batch` template pairs inside the CTA so that the spectrum tile is fetched once.
struct StridedArgs {
  const float2* in;
  float2* out;
  long long in_sb, in_se, in_batch;    // element e of line (a,b): in[a + b*in_sb + e*in_se]
  long long out_sb, out_se, out_batch;
  int nin, flip, nout;                 // nin input entries (flipped placement if flip), first nout outputs stored
  int na, nb, batch;
  const float2* mul;                   // MODE 1: spectrum, same (a,b,e) addressing with mul_sb / mul_se
  long long mul_sb, mul_se;
  const float2* tw;
  float scale;                         // applied on store
  long long out_sa;                    // 0: lines are 1 apart in the output (default); else stride between lines and the
                                       // store loop runs element-fastest (transposed output for the fused z/y kernel)
};
template <int LOG2N, int MODE>
__global__ void __launch_bounds__(kThreads, 5) k_fft_strided(const StridedArgs A) {
  constexpr int N = 1 << LOG2N, LS = line_stride(LOG2N), LPB = lines_per_block(LOG2N);
  extern __shared__ __align__(16) float2 sm[];
  float2* b0 = sm;
  float2* b1 = sm + LPB * LS;
  float2* tw = sm + (inplace_ok(LOG2N) ? 1 : 2) * LPB * LS;
  float2* spec = tw + N;  // MODE 1 only
  const int a0 = blockIdx.x * LPB, b = blockIdx.y;
  const int nl = min(LPB, A.na - a0);
  load_twiddles<LOG2N>(tw, A.tw);
  if (MODE == 1) {
    const float2* mul = A.mul + (long long)b * A.mul_sb + a0;
    for (int i = threadIdx.x; i < LPB * N; i += kThreads) {
      const int al = i % LPB, e = i / LPB;
      spec[al * LS + PI(e)] = al < nl ? mul[al + (long long)e * A.mul_se] : make_float2(0.f, 0.f);
    }
  }
  const int p0 = (MODE == 1) ? 0 : blockIdx.z, p1 = (MODE == 1) ? A.batch : blockIdx.z + 1;
  for (int pr = p0; pr < p1; ++pr) {
    const float2* in = A.in + (long long)pr * A.in_batch + (long long)b * A.in_sb + a0;
    float2* out = A.out + (long long)pr * A.out_batch + (long long)b * A.out_sb + (long long)a0 * (A.out_sa ? A.out_sa : 1);
    __syncthreads();  // previous pair fully stored / twiddles + spectrum visible
    if (A.nin < N) zero_lines<LOG2N>(b0, LPB);
    __syncthreads();
    if (nl == LPB && (LPB & 1) == 0 && ((A.in_se | A.in_sb | A.in_batch) & 1) == 0) {
      // full tile: two adjacent lines per 16-byte load (half the load instructions, twice the bytes in flight)
      constexpr int H = LPB / 2;
      for (int i = threadIdx.x; i < H * A.nin; i += kThreads) {
        const int ap = i % H, e = i / H;
        const float4 v = *reinterpret_cast<const float4*>(in + 2 * ap + (long long)e * A.in_se);
        const int slot = PI(A.flip ? ((N - e) & (N - 1)) : e);
        b0[(2 * ap) * LS + slot] = make_float2(v.x, v.y);
        b0[(2 * ap + 1) * LS + slot] = make_float2(v.z, v.w);
      }
    } else {
      for (int i = threadIdx.x; i < LPB * A.nin; i += kThreads) {
        const int al = i % LPB, e = i / LPB;  // a fastest: coalesced
        if (al < nl) {
          const int slot = A.flip ? ((N - e) & (N - 1)) : e;
          b0[al * LS + PI(slot)] = in[al + (long long)e * A.in_se];
        }
      }
    }
    __syncthreads();
    float2* res = fft_lines<LOG2N, MODE == 2, false>(b0, b1, tw, nullptr, LPB);
    if (MODE == 1) {
      float2* other = (res == b0) ? b1 : b0;
      res = fft_lines<LOG2N, true, true>(res, other, tw, spec, LPB);
    }
    if (A.out_sa) {
      for (int i = threadIdx.x; i < nl * A.nout; i += kThreads) {
        const int al = i / A.nout, e = i - al * A.nout;  // element fastest: one contiguous run per line
        float2 v = res[al * LS + PI(e)];
        v.x *= A.scale;
        v.y *= A.scale;
        out[(long long)al * A.out_sa + (long long)e * A.out_se] = v;
      }
    } else if (nl == LPB && (LPB & 1) == 0 && ((A.out_se | A.out_sb | A.out_batch) & 1) == 0) {
      constexpr int H = LPB / 2;
      for (int i = threadIdx.x; i < H * A.nout; i += kThreads) {
        const int ap = i % H, e = i / H;
        const int slot = PI(e);
        const float2 v0 = res[(2 * ap) * LS + slot], v1 = res[(2 * ap + 1) * LS + slot];
        *reinterpret_cast<float4*>(out + 2 * ap + (long long)e * A.out_se) =
            make_float4(v0.x * A.scale, v0.y * A.scale, v1.x * A.scale, v1.y * A.scale);
      }
    } else {
      for (int i = threadIdx.x; i < LPB * A.nout; i += kThreads) {
        const int al = i % LPB, e = i / LPB;
        if (al < nl) {
          float2 v = res[al * LS + PI(e)];
          v.x *= A.scale;
          v.y *= A.scale;
          out[al + (long long)e * A.out_se] = v;
        }
      }
    }
  }
}

// ---- pass B': direct correlation along z in the (kx, ky) frequency domain ----------------------------------
// out[pr][z][ky][kx] = sum_{q < tz} Sxy[z + q][ky][kx] * T[pr][q][ky][kx],  z < nzo,
// where Sxy is the 2-D (x, y) spectrum of every image plane and T the 2-D spectrum of the (flipped) template
// planes.  A template is only tz planes deep, so the z axis needs no transform at all: tz complex MACs per output
// beat the padded forward FFT + product + inverse FFT of the fused kernel below for tz <= 24, the image spectrum
// shrinks from Nz to nz planes, and everything is plain FP32 FMAs on coalesced streams (one thread per (kx, ky)
// column, a sliding window of W spectrum planes in registers).
struct ZDirectArgs {
  const float2* sxy;     // [nz][Ny][Nx]
  const float2* tmpl;    // [npair][tz][Ny][Nx]
  float2* out;           // [npair][nzo][Ny][Nx]
  long long plane;       // Ny*Nx
  long long tmpl_batch, out_batch;
  int nz, tz, nzo;
};
#ifndef IQ_ZD_THREADS
#define IQ_ZD_THREADS 128  // 949 vs 990 ms of FFT passes per config-5 simulation with 256 (finer CTA scheduling)
#endif
constexpr int kZdThreads = IQ_ZD_THREADS;
#ifndef IQ_ZPACKED_DEFAULT
#define IQ_ZPACKED_DEFAULT 0  // the packed-FMA z pass is opt-in (IQB200_FFT_ZPACKED=1) until measured
#endif
template <int W>
__global__ void __launch_bounds__(kZdThreads) k_fft_zdirect(const ZDirectArgs A) {
  // blockIdx.x = template pair (fastest): the CTAs that read the same spectrum columns are scheduled together, so the
  // spectrum comes from DRAM once per launch and from L2 for every other pair
  const long long col = (long long)blockIdx.y * kZdThreads + threadIdx.x;
  if (col >= A.plane) return;
  const float2* __restrict__ S = A.sxy + col;
  const float2* __restrict__ T = A.tmpl + (long long)blockIdx.x * A.tmpl_batch + col;
  float2* __restrict__ O = A.out + (long long)blockIdx.x * A.out_batch + col;
  const float2 zero = make_float2(0.f, 0.f);
  float2 t[W], s[W];
#pragma unroll
  for (int q = 0; q < W; ++q) t[q] = q < A.tz ? T[(long long)q * A.plane] : zero;
#pragma unroll
  for (int j = 0; j < W - 1; ++j) s[j] = j < A.nz ? S[(long long)j * A.plane] : zero;
  s[W - 1] = zero;
  const float2* __restrict__ sp = S + (long long)(W - 1) * A.plane;  // next spectrum plane to enter the window
  float2* __restrict__ op = O;
  const long long step = A.plane;
  for (int z0 = 0; z0 < A.nzo; z0 += W) {
    float2 nxt[W];
    const int navail = A.nz - (z0 + W - 1);  // spectrum planes left for this block of W outputs
#pragma unroll
    for (int m = 0; m < W; ++m) {
      nxt[m] = m < navail ? *sp : zero;
      sp += step;
    }
    const int nout = A.nzo - z0;
#pragma unroll
    for (int m = 0; m < W; ++m) {
      s[(m + W - 1) % W] = nxt[m];
      float ax = 0.f, ay = 0.f;
#pragma unroll
      for (int q = 0; q < W; ++q) {
        const float2 sv = s[(m + q) % W], tv = t[q];
        ax = fmaf(sv.x, tv.x, ax);
        ax = fmaf(-sv.y, tv.y, ax);
        ay = fmaf(sv.x, tv.y, ay);
        ay = fmaf(sv.y, tv.x, ay);
      }
      if (m < nout) *op = make_float2(ax, ay);
      op += step;
    }
  }
}

#ifdef IQB200_EXPERIMENTS
// Same pass with packed FMAs (fma.rn.f32x2 -> SASS FFMA2; experiments build, IQB200_FFT_ZPACKED=1).  On sm_100 the packed
// instruction has the FMA throughput of the scalar one but takes half the issue slots, and k_fft_zdirect is bound by
// instruction issue (82 % of the issue slots, FMA pipe 67 %; profiles/r02_resident_ncu_full.txt) -- yet MEASURED SLOWER:
// 984 ms against 949 ms of FFT passes per config-5 simulation (call AN of round 2; 512 FFMA2 instead of 1024 FFMA per 16
// outputs, but 152 registers = 3 CTAs per SM instead of 4 and ~100 extra MOVs for the duplicated window).  Same maps
// within FP32 rounding (the FFT parity tests pass with it).  Two packed accumulators per output,
//   P = sum_q (sx, sx) * (tx, ty) = (sum sx tx, sum sx ty),   Q = sum_q (sy, sy) * (tx, ty) = (sum sy tx, sum sy ty),
//   out = (P.x - Q.y, P.y + Q.x):
// the template values are used as loaded (re, im pairs), only the spectrum window is kept duplicated ((sx, sx) and
// (sy, sy) pairs, built once per plane and used by W outputs).  32 FFMA2 per output instead of 64 FFMA.
typedef unsigned long long u64p;
__device__ __forceinline__ u64p zpack2(float lo, float hi) {
  u64p r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void zunpack2(u64p v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64p zffma2(u64p a, u64p b, u64p c) {
  u64p d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
template <int W>
__global__ void __launch_bounds__(kZdThreads) k_fft_zdirect2(const ZDirectArgs A) {
  const long long col = (long long)blockIdx.y * kZdThreads + threadIdx.x;
  if (col >= A.plane) return;
  const float2* __restrict__ S = A.sxy + col;
  const float2* __restrict__ T = A.tmpl + (long long)blockIdx.x * A.tmpl_batch + col;
  float2* __restrict__ O = A.out + (long long)blockIdx.x * A.out_batch + col;
  const float2 zero = make_float2(0.f, 0.f);
  u64p t[W], sx[W], sy[W];
#pragma unroll
  for (int q = 0; q < W; ++q) {
    const float2 v = q < A.tz ? T[(long long)q * A.plane] : zero;
    t[q] = zpack2(v.x, v.y);
  }
#pragma unroll
  for (int j = 0; j < W - 1; ++j) {
    const float2 v = j < A.nz ? S[(long long)j * A.plane] : zero;
    sx[j] = zpack2(v.x, v.x);
    sy[j] = zpack2(v.y, v.y);
  }
  sx[W - 1] = sy[W - 1] = 0ull;
  const float2* __restrict__ sp = S + (long long)(W - 1) * A.plane;  // next spectrum plane to enter the window
  float2* __restrict__ op = O;
  const long long step = A.plane;
  for (int z0 = 0; z0 < A.nzo; z0 += W) {
    float2 nxt[W];
    const int navail = A.nz - (z0 + W - 1);  // spectrum planes left for this block of W outputs
#pragma unroll
    for (int m = 0; m < W; ++m) {
      nxt[m] = m < navail ? *sp : zero;
      sp += step;
    }
    const int nout = A.nzo - z0;
#pragma unroll
    for (int m = 0; m < W; ++m) {
      sx[(m + W - 1) % W] = zpack2(nxt[m].x, nxt[m].x);
      sy[(m + W - 1) % W] = zpack2(nxt[m].y, nxt[m].y);
      u64p P = 0ull, Q = 0ull;
#pragma unroll
      for (int q = 0; q < W; ++q) {
        P = zffma2(sx[(m + q) % W], t[q], P);
        Q = zffma2(sy[(m + q) % W], t[q], Q);
      }
      float px, py, qx, qy;
      zunpack2(P, px, py);
      zunpack2(Q, qx, qy);
      if (m < nout) *op = make_float2(px - qy, py + qx);
      op += step;
    }
  }
}

#endif  // IQB200_EXPERIMENTS

#ifdef IQB200_EXPERIMENTS
// ---- pass B'': direct z correlation FUSED with the inverse y transform --------------------------------------
// One CTA per (template pair, kx): thread ky keeps the sliding window of its (kx, ky) column in registers exactly as
// k_fft_zdirect does, but 16 consecutive output planes go straight into 16 shared-memory lines (ky along the line),
// are inverse-transformed along y in place and only the valid rows leave the SM.  The (nzo x Ny x Nx) intermediate
// of the separate kernels (44.6 MB per pair on config 5, written once and read once) never exists.  Needs ky on
// the fast axis: transposed image spectrum sxy_t[kx][z][ky], transposed template spectrum [pair][kx][q][ky]
// (k_fft_strided's element-fastest store) and a transposed result [pair][kx][z][y] (read by k_fft_x_final<.,true>).
struct ZYArgs {
  const float2* sxy_t;   // [Nx][nz][Ny]
  const float2* tmpl_t;  // [npair][Nx][tz][Ny]
  float2* out_t;         // [npair][Nx][nzo][nyo]
  long long tmpl_batch, out_batch;
  int nz, tz, nzo, nyo;
  const float2* tw;      // twiddles of the y transform
};
template <int LOG2N>
__global__ void __launch_bounds__(kThreads, 2) k_fft_zy(const ZYArgs A) {
  constexpr int N = 1 << LOG2N, LS = line_stride(LOG2N), LPB = lines_per_block(LOG2N), W = 16;
  static_assert(N == kThreads && LPB == W && inplace_ok(LOG2N), "one thread per ky, 16 planes per transform batch");
  extern __shared__ __align__(16) float2 sm[];
  float2* buf = sm;
  float2* tw = sm + LPB * LS;
  const int pr = blockIdx.x, kx = blockIdx.y, ky = threadIdx.x;
  load_twiddles<LOG2N>(tw, A.tw);
  const float2* __restrict__ S = A.sxy_t + (long long)kx * A.nz * N + ky;
  const float2* __restrict__ T = A.tmpl_t + (long long)pr * A.tmpl_batch + (long long)kx * A.tz * N + ky;
  float2* __restrict__ O = A.out_t + (long long)pr * A.out_batch + (long long)kx * A.nzo * A.nyo;
  const float2 zero = make_float2(0.f, 0.f);
  float2 t[W], s[W];
#pragma unroll
  for (int q = 0; q < W; ++q) t[q] = q < A.tz ? T[q * N] : zero;
#pragma unroll
  for (int j = 0; j < W - 1; ++j) s[j] = j < A.nz ? S[j * N] : zero;
  s[W - 1] = zero;
  const float2* __restrict__ sp = S + (W - 1) * N;
  const int pky = PI(ky);
  __syncthreads();  // twiddles
  for (int z0 = 0; z0 < A.nzo; z0 += W) {
    float2 nxt[W];
    const int navail = A.nz - (z0 + W - 1);
#pragma unroll
    for (int m = 0; m < W; ++m) {
      nxt[m] = m < navail ? *sp : zero;
      sp += N;
    }
#pragma unroll
    for (int m = 0; m < W; ++m) {
      s[(m + W - 1) % W] = nxt[m];
      float ax = 0.f, ay = 0.f;
#pragma unroll
      for (int q = 0; q < W; ++q) {
        const float2 sv = s[(m + q) % W], tv = t[q];
        ax = fmaf(sv.x, tv.x, ax);
        ax = fmaf(-sv.y, tv.y, ax);
        ay = fmaf(sv.x, tv.y, ay);
        ay = fmaf(sv.y, tv.x, ay);
      }
      buf[m * LS + pky] = make_float2(ax, ay);
    }
    __syncthreads();
    const int nl = min(W, A.nzo - z0);
    const float2* res = fft_lines<LOG2N, true, false>(buf, buf, tw, nullptr, nl);
    if (ky < A.nyo) {
      float2* op = O + (long long)z0 * A.nyo + ky;
      for (int line = 0; line < nl; ++line) op[(long long)line * A.nyo] = res[line * LS + pky];
    }
    __syncthreads();  // the lines are refilled by the next batch
  }
}

// out[x][z][y] = in[z][y][x] (set-up only: transposed image spectrum)
__global__ void __launch_bounds__(256) k_transpose_xzy(const float2* __restrict__ in, float2* __restrict__ out, int nx, int ny,
                                                       int nz) {
  __shared__ float2 tile[16][17];
  const int z = blockIdx.z, x0 = blockIdx.x * 16, y0 = blockIdx.y * 16;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  if (x0 + tx < nx && y0 + ty < ny) tile[ty][tx] = in[((long long)z * ny + y0 + ty) * nx + x0 + tx];
  __syncthreads();
  if (x0 + ty < nx && y0 + tx < ny) out[((long long)(x0 + ty) * nz + z) * ny + y0 + tx] = tile[tx][ty];
}

#endif  // IQB200_EXPERIMENTS

// ---- pass C: last inverse pass along x + distance epilogue --------------------------------------------
struct FinalArgs {
  const float2* in;     // [npair][nlines][N], or (transposed) [npair][N][nlines]
  long long in_batch;   // transposed input: float2 per pair
  int nlines, nxo;      // nlines = nyo*nzo; output position p = line*nxo + x
  long long npos;
  int R;
  Epilogue ep;
  const float2* tw;
};
// Distance epilogue of the last pass for the nl lines of tile `blk` of template pair `pr` (result lines in `res`,
// padded layout): |A2 - 2 AB + B2|, disabled -> +Inf, min / max of the map and the minimum of the tile (chunk).
template <int LOG2N>
__device__ __forceinline__ void final_epilogue(const FinalArgs& A, const float2* res, int pr, int blk, int l0, int nl,
                                               unsigned* s_min, unsigned* s_max) {
  constexpr int LS = line_stride(LOG2N);
  const int r0 = 2 * pr, r1 = 2 * pr + 1;
  const bool has1 = r1 < A.R;
  const double b20 = A.ep.b2[r0], b21 = has1 ? A.ep.b2[r1] : 0.0;
  unsigned mn0 = 0x7f800000u, mx0 = 0u, mn1 = 0x7f800000u, mx1 = 0u;
  // one x column per thread, walked down the lines of the tile: no index division, unit-stride pointers (the profile
  // of the first version showed this epilogue, not the transform, saturating the integer pipe)
  const bool rnd = A.ep.round_to_int != 0;
  // uniform 64-bit bases + one 32-bit position per thread (npos < 2^32): the loads / stores use base + u32 addressing
  // instead of a 64-bit add per pointer and line
  const float* __restrict__ a2b = A.ep.a2_list ? A.ep.a2_list[r0] : A.ep.a2;
  const float* __restrict__ a2c = A.ep.a2_list ? (has1 ? A.ep.a2_list[r1] : nullptr) : A.ep.a2;  // second template of the pair
  const bool same_a2 = a2b == a2c;
  const uint8_t* __restrict__ disb = A.ep.disabled;
  float* __restrict__ o0b = A.ep.out + (long long)r0 * A.npos;
  float* __restrict__ o1b = A.ep.out + (long long)r1 * A.npos;
  for (int x = threadIdx.x; x < A.nxo; x += kThreads) {
    const float2* rp = res + PI(x);
    unsigned p = (unsigned)((long long)l0 * A.nxo + x);
#pragma unroll 2
    for (int line = 0; line < nl; ++line, p += (unsigned)A.nxo) {
      float2 ab = rp[line * LS];
      if (rnd) { ab.x = rintf(ab.x); ab.y = rintf(ab.y); }
      const double a2 = a2b ? (double)__ldg(a2b + p) : 0.0;
      const double a2s = same_a2 ? a2 : (a2c ? (double)__ldg(a2c + p) : 0.0);
      const bool dis = disb && disb[p];
      float d0 = (float)fabs(a2 - 2.0 * (double)ab.x + b20);
      if (dis) d0 = CUDART_INF_F;
      o0b[p] = d0;
      if (!dis) { const unsigned u = __float_as_uint(d0); mn0 = min(mn0, u); mx0 = max(mx0, u); }
      if (has1) {
        float d1 = (float)fabs(a2s - 2.0 * (double)ab.y + b21);
        if (dis) d1 = CUDART_INF_F;
        o1b[p] = d1;
        if (!dis) { const unsigned u = __float_as_uint(d1); mn1 = min(mn1, u); mx1 = max(mx1, u); }
      }
    }
  }
  if (A.ep.minbits) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn0 = min(mn0, __shfl_xor_sync(0xffffffffu, mn0, o));
      mx0 = max(mx0, __shfl_xor_sync(0xffffffffu, mx0, o));
      mn1 = min(mn1, __shfl_xor_sync(0xffffffffu, mn1, o));
      mx1 = max(mx1, __shfl_xor_sync(0xffffffffu, mx1, o));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(&s_min[0], mn0); atomicMax(&s_max[0], mx0);
      atomicMin(&s_min[1], mn1); atomicMax(&s_max[1], mx1);
    }
    __syncthreads();
    if (threadIdx.x == 0) { atomicMin(A.ep.minbits + r0, s_min[0]); atomicMax(A.ep.maxbits + r0, s_max[0]); }
    if (threadIdx.x == 1 && has1) { atomicMin(A.ep.minbits + r1, s_min[1]); atomicMax(A.ep.maxbits + r1, s_max[1]); }
    // minimum of this tile's chunk of LPB x-rows: lets the selection kernel skip chunks without candidates
    if (A.ep.chunkmin && threadIdx.x < 2 && (threadIdx.x == 0 || has1))
      A.ep.chunkmin[(long long)(threadIdx.x ? r1 : r0) * A.ep.chunk_pitch + blk] = s_min[threadIdx.x];
  }
}

template <int LOG2N, bool TIN>
__global__ void __launch_bounds__(kThreads, 4) k_fft_x_final(const FinalArgs A) {
  constexpr int N = 1 << LOG2N, LS = line_stride(LOG2N), LPB = lines_per_block(LOG2N);
  extern __shared__ __align__(16) float2 sm[];
  float2* b0 = sm;
  float2* b1 = sm + LPB * LS;
  float2* tw = sm + (inplace_ok(LOG2N) ? 1 : 2) * LPB * LS;
  __shared__ unsigned s_min[2], s_max[2];
  const int pr = blockIdx.y, l0 = blockIdx.x * LPB;
  const int nl = min(LPB, A.nlines - l0);
  load_twiddles<LOG2N>(tw, A.tw);
  if (threadIdx.x < 2) { s_min[threadIdx.x] = 0x7f800000u; s_max[threadIdx.x] = 0u; }
  if (TIN) {
    // transposed input [kx][line]: LPB consecutive lines of one kx are contiguous (128-byte runs for LPB = 16)
    const float2* in = A.in + (long long)pr * A.in_batch + l0;
    for (int i = threadIdx.x; i < LPB * N; i += kThreads) {
      const int al = i % LPB, e = i / LPB;
      if (al < nl) b0[al * LS + PI(e)] = in[(long long)e * A.nlines + al];
    }
  } else {
    const float2* in = A.in + ((long long)pr * A.nlines + l0) * N;
    if (N >= 2) {  // two consecutive elements of a line per 16-byte load (PI(e) and PI(e + 1) are adjacent for even e)
      for (int i = threadIdx.x; i < nl * (N / 2); i += kThreads) {
        const int line = i / (N / 2), e = 2 * (i - line * (N / 2));
        const float4 v = *reinterpret_cast<const float4*>(in + (long long)line * N + e);
        float2* d = b0 + line * LS + PI(e);
        d[0] = make_float2(v.x, v.y);
        d[1] = make_float2(v.z, v.w);
      }
    } else {
      for (int i = threadIdx.x; i < nl * N; i += kThreads) {
        const int line = i >> LOG2N, e = i & (N - 1);
        b0[line * LS + PI(e)] = in[(long long)line * N + e];
      }
    }
  }
  __syncthreads();
  const float2* res = fft_lines<LOG2N, true, false>(b0, b1, tw, nullptr, nl);
  final_epilogue<LOG2N>(A, res, pr, blockIdx.x, l0, nl, s_min, s_max);
}

// TMA helpers (bulk asynchronous copies global -> shared, completion on an mbarrier): iq_tma.cuh
using namespace iqtma;

// ---- pass B, inverse transform along a strided axis with TMA-fed double buffering ------------------------------------
// The inverse y pass is bound by memory latency (ncu: 5 warps per issue slot waiting on the long scoreboard, DRAM at
// 63 %): every CTA first pulls its 16 x N tile through registers (128-byte row chunks, 2 KB apart), then transforms, then
// stores.  Here the CTAs are persistent and the NEXT tile is already in flight while the current one is transformed:
// every thread issues one 128-byte bulk asynchronous copy (cp.async.bulk, the TMA engine; completion counted on an
// mbarrier) of row e of the tile straight into shared memory -- no registers, no LSU issue slots -- in the
// element-major layout [e][16 lines] that the first Stockham pass reads (LIN = 2).  Same arithmetic, same results.
// USE_MAP: ONE tensor-map request per tile (cp.async.bulk.tensor.2d: box of 16 elements x N rows of the work array seen
// as a 2-D tensor [rows][Nx] of 8-byte elements) instead of N row copies.
template <int LOG2N, bool USE_MAP>
__global__ void __launch_bounds__(kThreads, 3) k_fft_strided_inv_tma(const StridedArgs A, int ntile_a, int total,
                                                                     const __grid_constant__ CUtensorMap tmap, int rows_per_b,
                                                                     int rows_per_pair) {
  constexpr int N = 1 << LOG2N, LS = line_stride(LOG2N), LPB = lines_per_block(LOG2N);
  static_assert(inplace_ok(LOG2N) && LPB == 16, "16 lines per tile, in-place transform");
  extern __shared__ __align__(128) unsigned char smraw_[];
  float2* sm = reinterpret_cast<float2*>((reinterpret_cast<uintptr_t>(smraw_) + 127) & ~(uintptr_t)127);  // TMA boxes: 128-byte aligned
  // buffer pitch: the padded lines (LPB * LS float2) rounded up to 128 bytes
  constexpr int kBuf = ((LPB * LS * (int)sizeof(float2) + 127) / 128) * 128 / (int)sizeof(float2);
  float2* const buf0 = sm;
  float2* const buf1 = sm + kBuf;
  float2* tw = sm + 2 * kBuf;
  __shared__ __align__(8) unsigned long long bar[2];
  load_twiddles<LOG2N>(tw, A.tw);
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // tile t -> (pair, plane b, chunk of 16 lines): the chunks of one plane are consecutive, so that concurrently
  // running CTAs read neighbouring 128-byte pieces of the same 2 KB rows
  auto tile_src = [&](int t) -> const float2* {
    const int ac = t % ntile_a, r = t / ntile_a, b = r % A.nb, pr = r / A.nb;
    return A.in + (long long)pr * A.in_batch + (long long)b * A.in_sb + ac * LPB;
  };
  auto issue = [&](int t, int bsel) {  // all threads
    float2* dst = bsel ? buf1 : buf0;
    fence_async_smem();  // the buffer was last touched by ordinary loads / stores
    if (USE_MAP) {
      if (threadIdx.x == 0) {
        const int ac = t % ntile_a, r = t / ntile_a, b = r % A.nb, pr = r / A.nb;
        mbar_expect_tx(&bar[bsel], (unsigned)(N * LPB * sizeof(float2)));
        tma_load_2d(dst, &tmap, ac * LPB, pr * rows_per_pair + b * rows_per_b, &bar[bsel]);
      }
    } else {
      const float2* src = tile_src(t);
      if (threadIdx.x == 0) mbar_expect_tx(&bar[bsel], (unsigned)(N * LPB * sizeof(float2)));
      for (int e = threadIdx.x; e < N; e += kThreads)
        bulk_g2s(dst + e * LPB, src + (long long)e * A.in_se, (unsigned)(LPB * sizeof(float2)), &bar[bsel]);
    }
  };
  int t = blockIdx.x;
  if (t < total) issue(t, 0);
  for (int it = 0; t < total; t += gridDim.x, ++it) {
    const int bsel = it & 1;
    const int tn = t + gridDim.x;
    if (tn < total) issue(tn, bsel ^ 1);  // the other buffer was released by the barrier at the end of the last round
    mbar_wait(&bar[bsel], (unsigned)((it >> 1) & 1));
    float2* cur = bsel ? buf1 : buf0;
    const float2* res = fft_lines<LOG2N, true, false, 2>(cur, cur, tw, nullptr, LPB);
    const int ac = t % ntile_a, r = t / ntile_a, b = r % A.nb, pr = r / A.nb;
    float2* out = A.out + (long long)pr * A.out_batch + (long long)b * A.out_sb + ac * LPB;
    constexpr int H = LPB / 2;
    for (int i = threadIdx.x; i < H * A.nout; i += kThreads) {
      const int ap = i % H, e = i / H;
      const int slot = PI(e);
      const float2 v0 = res[(2 * ap) * LS + slot], v1 = res[(2 * ap + 1) * LS + slot];
      *reinterpret_cast<float4*>(out + 2 * ap + (long long)e * A.out_se) =
          make_float4(v0.x * A.scale, v0.y * A.scale, v1.x * A.scale, v1.y * A.scale);
    }
    __syncthreads();  // everybody is done with this buffer before it is refilled
  }
}

#ifdef IQB200_EXPERIMENTS
// ---- pass C, streaming version: persistent CTAs, the next tile of LPB lines (one contiguous block of global
//      memory) is fetched by ONE bulk TMA copy into the other half of a double buffer while the current tile is
//      transformed in place and written out.  The first pass reads the unpadded lines the copy delivered.
template <int LOG2N>
__global__ void __launch_bounds__(kThreads, 3) k_fft_x_final_tma(const FinalArgs A, int ntile, int total) {
  constexpr int N = 1 << LOG2N, LS = line_stride(LOG2N), LPB = lines_per_block(LOG2N);
  static_assert(inplace_ok(LOG2N), "streaming kernels need the in-place transform");
  extern __shared__ __align__(16) float2 sm[];
  float2* const buf0 = sm;
  float2* const buf1 = sm + LPB * LS;
  float2* tw = sm + 2 * LPB * LS;
  __shared__ unsigned s_min[2], s_max[2];
  __shared__ __align__(8) unsigned long long bar[2];
  load_twiddles<LOG2N>(tw, A.tw);
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int t, int b) {  // thread 0 only
    const int pr = t / ntile, blk = t - pr * ntile;
    const int l0 = blk * LPB, nl = min(LPB, A.nlines - l0);
    const unsigned bytes = (unsigned)(nl * N * sizeof(float2));
    fence_async_smem();
    mbar_expect_tx(&bar[b], bytes);
    bulk_g2s(b ? buf1 : buf0, A.in + ((long long)pr * A.nlines + l0) * N, bytes, &bar[b]);
  };
  int t = blockIdx.x;
  if (threadIdx.x == 0 && t < total) issue(t, 0);
  for (int it = 0; t < total; t += gridDim.x, ++it) {
    const int b = it & 1;
    const int tn = t + gridDim.x;
    if (threadIdx.x == 0 && tn < total) issue(tn, b ^ 1);  // the other buffer was released by the barrier below
    if (threadIdx.x < 2) { s_min[threadIdx.x] = 0x7f800000u; s_max[threadIdx.x] = 0u; }
    mbar_wait(&bar[b], (unsigned)((it >> 1) & 1));
    const int pr = t / ntile, blk = t - pr * ntile;
    const int l0 = blk * LPB, nl = min(LPB, A.nlines - l0);
    float2* cur = b ? buf1 : buf0;
    const float2* res = fft_lines<LOG2N, true, false, 1>(cur, cur, tw, nullptr, nl);
    final_epilogue<LOG2N>(A, res, pr, blk, l0, nl, s_min, s_max);
    __syncthreads();  // everybody is done with buf[b] and s_min before they are reused
  }
}

#endif  // IQB200_EXPERIMENTS

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int ilog2_ceil(int n) {
  int l = 0;
  while ((1 << l) < n) ++l;
  return l;
}
static size_t smem_bytes(int log2n, bool with_spectrum = false) {
  const int N = 1 << log2n;
  const int nbuf = (inplace_ok(log2n) ? 1 : 2) + (with_spectrum ? 1 : 0);
  return (size_t)(nbuf * lines_per_block(log2n) * line_stride(log2n) + N) * sizeof(float2);
}

struct Plan {
  int nx, ny, nz, tx, ty, tz, nxo, nyo, nzo;
  int lx, ly, lz;       // log2 of the padded sizes (lz = 0 for 2-D)
  int Nx, Ny, Nz;
  int max_pairs;
  long long npos;
  float2 *twx = nullptr, *twy = nullptr, *twz = nullptr;
  float2 *w1 = nullptr, *w2 = nullptr, *w3 = nullptr, *w4 = nullptr;
  long long w1_stride = 0, w2_stride = 0, w3_stride = 0, w4_stride = 0;  // float2 per pair
  std::map<int, float2*> spectrum;   // full 3-D (2-D problems: 2-D) spectrum, scaled by 1/N
  std::map<int, float2*> sxy;        // 3-D problems with the direct z pass: (x, y) spectrum of every plane, scaled by 1/(Nx Ny)
  bool zdirect = false;              // direct correlation along z instead of the fused z transforms (tz <= 24)
  bool zyfused = false;              // ... fused with the inverse y transform (k_fft_zy: Ny == 256, tz <= 16)
  std::map<int, float2*> sxy_t;      // transposed (x, y) spectrum [Nx][nz][Ny] of the fused kernel
  size_t workspace = 0;
  CUtensorMap w3_map;                // w3 as a 2-D tensor [max_pairs * nzo * Ny rows][Nx] of 8-byte elements, box 16 x Ny
  bool w3_map_ok = false;
};

#define FFT_DISPATCH(LOG2N, ...)                                   \
  switch (LOG2N) {                                                 \
    case 1: { constexpr int L = 1; __VA_ARGS__; } break;                  \
    case 2: { constexpr int L = 2; __VA_ARGS__; } break;                  \
    case 3: { constexpr int L = 3; __VA_ARGS__; } break;                  \
    case 4: { constexpr int L = 4; __VA_ARGS__; } break;                  \
    case 5: { constexpr int L = 5; __VA_ARGS__; } break;                  \
    case 6: { constexpr int L = 6; __VA_ARGS__; } break;                  \
    case 7: { constexpr int L = 7; __VA_ARGS__; } break;                  \
    case 8: { constexpr int L = 8; __VA_ARGS__; } break;                  \
    case 9: { constexpr int L = 9; __VA_ARGS__; } break;                  \
    case 10: { constexpr int L = 10; __VA_ARGS__; } break;                \
    default: return cudaErrorInvalidValue;                         \
  }

template <typename K>
static cudaError_t set_smem(K kernel, size_t bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

static cudaError_t launch_tmpl(const TmplPassArgs& a, int log2n, int npair, cudaStream_t s) {
  const size_t sm = smem_bytes(log2n);
  const int LPB = lines_per_block(log2n);
  dim3 grid((a.nlines + LPB - 1) / LPB, npair);
  FFT_DISPATCH(log2n, { cudaError_t e = set_smem(k_fft_x_tmpl<L>, sm); if (e != cudaSuccess) return e;
                        k_fft_x_tmpl<L><<<grid, kThreads, sm, s>>>(a); });
  return cudaGetLastError();
}
static cudaError_t launch_real(const RealPassArgs& a, int log2n, cudaStream_t s) {
  const size_t sm = smem_bytes(log2n);
  const int LPB = lines_per_block(log2n);
  dim3 grid((a.nlines + LPB - 1) / LPB);
  FFT_DISPATCH(log2n, { cudaError_t e = set_smem(k_fft_x_real<L>, sm); if (e != cudaSuccess) return e;
                        k_fft_x_real<L><<<grid, kThreads, sm, s>>>(a); });
  return cudaGetLastError();
}
// inverse y pass: 0 = register-staged tiles (k_fft_strided<.,2>), 1 = TMA row copies (cp.async.bulk), 2 = one TMA
// tensor-map request per tile (cp.async.bulk.tensor.2d), both double-buffered in persistent CTAs (IQB200_FFT_INV_TMA).
// Measured on config 5 (FFT passes per simulation): 950 / 1005 / 963 ms -- the pass already moves 5.2 TB/s (ncu: DRAM
// 63 % of its 8 TB/s scale, i.e. ~80 % of the measured copy bandwidth), so hiding the load latency buys nothing and the
// persistent CTAs keep fewer tiles in flight than 5 resident register-staged CTAs.  Default: 0; bit-identical results.
static int inv_tma_mode() {
  const char* ev = std::getenv("IQB200_FFT_INV_TMA");
  return ev ? std::atoi(ev) : 0;
}

// mode: 1 = N row copies per tile (cp.async.bulk), 2 = one tensor-map request per tile (cp.async.bulk.tensor.2d)
template <int L, bool USE_MAP>
static cudaError_t launch_strided_inv_tma_impl(const StridedArgs& a, int batch, const CUtensorMap& map, int rows_per_b,
                                               int rows_per_pair, cudaStream_t s) {
  constexpr int LPB = 16;
  constexpr size_t kBuf = ((LPB * line_stride(L) * sizeof(float2) + 127) / 128) * 128;
  const size_t sm = 2 * kBuf + (size_t)(1 << L) * sizeof(float2) + 128;
  auto kern = k_fft_strided_inv_tma<L, USE_MAP>;
  cudaError_t e = set_smem(kern, sm);
  if (e != cudaSuccess) return e;
  static int nsm = 0;
  int ctas_per_sm = 0;
  if (!nsm) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  }
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, kThreads, sm);
  if (e != cudaSuccess) return e;
  const int ntile_a = a.na / LPB;
  const long long total = (long long)ntile_a * a.nb * batch;
  const int grid = (int)std::max<long long>(1, std::min<long long>(total, (long long)nsm * std::max(ctas_per_sm, 1)));
  kern<<<grid, kThreads, sm, s>>>(a, ntile_a, (int)total, map, rows_per_b, rows_per_pair);
  return cudaGetLastError();
}

template <int L>
static cudaError_t launch_strided_inv_tma(const StridedArgs& a, int batch, int mode, const CUtensorMap* map, cudaStream_t s) {
  if constexpr (inplace_ok(L) && lines_per_block(L) == 16) {
    const int N = 1 << L;
    if (mode == 2 && map) return launch_strided_inv_tma_impl<L, true>(a, batch, *map, (int)(a.in_sb / a.in_se), (int)(a.in_batch / a.in_se), s);
    (void)N;
    CUtensorMap none{};
    return launch_strided_inv_tma_impl<L, false>(a, batch, none, 0, 0, s);
  } else {
    return cudaErrorInvalidValue;
  }
}

template <int MODE>
static cudaError_t launch_strided(const StridedArgs& a_in, int log2n, int batch, cudaStream_t s, const CUtensorMap* map = nullptr) {
  if (MODE == 2 && inv_tma_mode() > 0 && inplace_ok(log2n) && lines_per_block(log2n) == 16 && log2n >= 5 &&
      a_in.nin == (1 << log2n) && a_in.flip == 0 && a_in.out_sa == 0 && a_in.na % 16 == 0 &&
      ((a_in.in_se | a_in.in_sb | a_in.in_batch | a_in.out_se | a_in.out_sb | a_in.out_batch) & 1) == 0 &&
      (long long)(a_in.na / 16) * a_in.nb * batch < (1ll << 31)) {
    StridedArgs a = a_in;
    a.batch = batch;
    const int mode = (inv_tma_mode() == 2 && map && a.in_sb % a.in_se == 0 && a.in_batch % a.in_se == 0) ? 2 : 1;
    FFT_DISPATCH(log2n, { return launch_strided_inv_tma<L>(a, batch, mode, map, s); });
  }
  const size_t sm = smem_bytes(log2n, MODE == 1);
  const int LPB = lines_per_block(log2n);
  StridedArgs a = a_in;
  a.batch = batch;
  dim3 grid((a.na + LPB - 1) / LPB, a.nb, MODE == 1 ? 1 : batch);  // the fused kernel loops over the pairs itself
  FFT_DISPATCH(log2n, { cudaError_t e = set_smem(k_fft_strided<L, MODE>, sm); if (e != cudaSuccess) return e;
                        k_fft_strided<L, MODE><<<grid, kThreads, sm, s>>>(a); });
  return cudaGetLastError();
}
#ifdef IQB200_EXPERIMENTS
static bool tma_enabled() {
  // Experimental (IQB200_FFT_TMA=1): persistent CTAs with a TMA-fed double buffer.  Bit-identical results, but on
  // config 5 it is slower than one tile per CTA (1.19 ms vs 0.77 ms per 32 template pairs): the pass is bound by
  // instruction issue, not by load latency, and the double buffer costs a resident CTA per SM.
  const char* ev = std::getenv("IQB200_FFT_TMA");
  return ev && ev[0] == '1';
}

template <int L>
static cudaError_t launch_final_tma(const FinalArgs& a, int npair, cudaStream_t s) {
  if constexpr (inplace_ok(L)) {
    constexpr int LPB = lines_per_block(L);
    const size_t sm = (size_t)(2 * LPB * line_stride(L) + (1 << L)) * sizeof(float2);
    cudaError_t e = set_smem(k_fft_x_final_tma<L>, sm);
    if (e != cudaSuccess) return e;
    static int ctas_per_sm = 0, nsm = 0;
    if (!nsm) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    }
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, k_fft_x_final_tma<L>, kThreads, sm);
    if (e != cudaSuccess) return e;
    const int ntile = (a.nlines + LPB - 1) / LPB, total = ntile * npair;
    const int grid = std::max(1, std::min(total, nsm * std::max(ctas_per_sm, 1)));
    k_fft_x_final_tma<L><<<grid, kThreads, sm, s>>>(a, ntile, total);
    return cudaGetLastError();
  } else {
    return cudaErrorInvalidValue;
  }
}

#endif  // IQB200_EXPERIMENTS

static cudaError_t launch_final(const FinalArgs& a, int log2n, int npair, cudaStream_t s) {
#ifdef IQB200_EXPERIMENTS
  if (tma_enabled() && inplace_ok(log2n) && log2n >= 4 && a.in_batch == 0) {
    FFT_DISPATCH(log2n, { return launch_final_tma<L>(a, npair, s); });
  }
#endif
  const size_t sm = smem_bytes(log2n);
  const int LPB = lines_per_block(log2n);
  dim3 grid((a.nlines + LPB - 1) / LPB, npair);
  if (a.in_batch > 0) {  // transposed input: only the fused z/y kernel of the experiments build produces it
#ifdef IQB200_EXPERIMENTS
    FFT_DISPATCH(log2n, { cudaError_t e = set_smem(k_fft_x_final<L, true>, sm); if (e != cudaSuccess) return e;
                          k_fft_x_final<L, true><<<grid, kThreads, sm, s>>>(a); });
#else
    return cudaErrorInvalidValue;
#endif
  } else {
    FFT_DISPATCH(log2n, { cudaError_t e = set_smem(k_fft_x_final<L, false>, sm); if (e != cudaSuccess) return e;
                          k_fft_x_final<L, false><<<grid, kThreads, sm, s>>>(a); });
  }
  return cudaGetLastError();
}

#ifdef IQB200_EXPERIMENTS
// z pass with packed FMAs (k_fft_zdirect2) for windows of up to 16 planes: IQB200_FFT_ZPACKED = 1 / 0 overrides the default
static bool zpacked_enabled() {
  static const int mode = [] {
    const char* ev = std::getenv("IQB200_FFT_ZPACKED");
    return ev ? (ev[0] == '1' ? 1 : 0) : IQ_ZPACKED_DEFAULT;
  }();
  return mode != 0;
}
#endif

static cudaError_t launch_zdirect(const ZDirectArgs& a, int npair, cudaStream_t s) {
  dim3 grid(npair, (unsigned)((a.plane + kZdThreads - 1) / kZdThreads));
  const int w = (a.tz + 3) / 4 * 4;
#ifdef IQB200_EXPERIMENTS
  if (w <= 16 && zpacked_enabled()) {
    switch (w) {
      case 4: k_fft_zdirect2<4><<<grid, kZdThreads, 0, s>>>(a); break;
      case 8: k_fft_zdirect2<8><<<grid, kZdThreads, 0, s>>>(a); break;
      case 12: k_fft_zdirect2<12><<<grid, kZdThreads, 0, s>>>(a); break;
      default: k_fft_zdirect2<16><<<grid, kZdThreads, 0, s>>>(a); break;
    }
    return cudaGetLastError();
  }
#endif
  switch (w) {
    case 4: k_fft_zdirect<4><<<grid, kZdThreads, 0, s>>>(a); break;
    case 8: k_fft_zdirect<8><<<grid, kZdThreads, 0, s>>>(a); break;
    case 12: k_fft_zdirect<12><<<grid, kZdThreads, 0, s>>>(a); break;
    case 16: k_fft_zdirect<16><<<grid, kZdThreads, 0, s>>>(a); break;
    case 20: k_fft_zdirect<20><<<grid, kZdThreads, 0, s>>>(a); break;
    case 24: k_fft_zdirect<24><<<grid, kZdThreads, 0, s>>>(a); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

static cudaError_t make_twiddles(float2** out, int N, cudaStream_t s) {
  std::vector<float2> h(N);
  for (int k = 0; k < N; ++k) {
    const double ang = -2.0 * M_PI * (double)k / (double)N;
    h[k] = make_float2((float)std::cos(ang), (float)std::sin(ang));
  }
  cudaError_t e = iq::dmalloc((void**)out, N * sizeof(float2));
  if (e != cudaSuccess) return e;
  e = cudaMemcpyAsync(*out, h.data(), N * sizeof(float2), cudaMemcpyHostToDevice, s);
  if (e != cudaSuccess) return e;
  return cudaStreamSynchronize(s);
}

static bool zdirect_enabled() {
  const char* ev = std::getenv("IQB200_FFT_ZDIRECT");  // experiments: 0 forces the fused z transforms
  return !(ev && ev[0] == '0');
}

#ifdef IQB200_EXPERIMENTS
static bool zyfused_enabled() {
  // Experimental (IQB200_FFT_ZYFUSED=1).  Bit-identical results and 42 % less traffic, but 5 % SLOWER on config 5
  // (FFT passes 1.155 s vs 1.099 s per 512 steps): both halves are bound by instruction issue, and the fused kernel
  // (128 registers, 2 CTAs per SM) overlaps its barrier phases worse than two separate kernels do.
  const char* ev = std::getenv("IQB200_FFT_ZYFUSED");
  return ev && ev[0] == '1';
}
#else
static bool zyfused_enabled() { return false; }  // k_fft_zy is part of the experiments build (make EXTRA=-DIQB200_EXPERIMENTS)
#endif

cudaError_t plan_create(Plan** out, int nx, int ny, int nz, int tx, int ty, int tz, int max_templates, cudaStream_t s) {
  *out = nullptr;
  if (ny < 2 || nx < 2) return cudaErrorInvalidValue;
  Plan* p = new Plan();
  p->nx = nx; p->ny = ny; p->nz = nz; p->tx = tx; p->ty = ty; p->tz = tz;
  p->nxo = nx - tx + 1; p->nyo = ny - ty + 1; p->nzo = nz - tz + 1;
  p->npos = (long long)p->nxo * p->nyo * p->nzo;
  p->lx = ilog2_ceil(nx); p->ly = ilog2_ceil(ny); p->lz = nz > 1 ? ilog2_ceil(nz) : 0;
  p->Nx = 1 << p->lx; p->Ny = 1 << p->ly; p->Nz = 1 << p->lz;
  if (p->lx > 10 || p->ly > 10 || p->lz > 10) { delete p; return cudaErrorInvalidValue; }
  p->max_pairs = (max_templates + 1) / 2;
  p->zdirect = p->lz > 0 && tz >= 2 && tz <= 24 && zdirect_enabled();
  p->zyfused = p->zdirect && tz <= 16 && p->Ny == kThreads && lines_per_block(p->ly) == 16 && zyfused_enabled();
  cudaError_t e;
  if ((e = make_twiddles(&p->twx, p->Nx, s)) != cudaSuccess) { plan_destroy(p); return e; }
  if ((e = make_twiddles(&p->twy, p->Ny, s)) != cudaSuccess) { plan_destroy(p); return e; }
  if (p->lz > 0 && (e = make_twiddles(&p->twz, p->Nz, s)) != cudaSuccess) { plan_destroy(p); return e; }
  const long long Nx = p->Nx, Ny = p->Ny;
  p->w1_stride = (long long)tz * ty * Nx;
  p->w4_stride = (long long)p->nzo * p->nyo * Nx;
  if (p->lz > 0) {
    p->w2_stride = (long long)tz * Ny * Nx;
    p->w3_stride = (long long)p->nzo * Ny * Nx;
  }
  const size_t mp = (size_t)p->max_pairs;
  if ((e = iq::dmalloc((void**)&p->w1, mp * p->w1_stride * sizeof(float2))) != cudaSuccess) { plan_destroy(p); return e; }
  if ((e = iq::dmalloc((void**)&p->w4, mp * p->w4_stride * sizeof(float2))) != cudaSuccess) { plan_destroy(p); return e; }
  if (p->lz > 0) {
    if ((e = iq::dmalloc((void**)&p->w2, mp * p->w2_stride * sizeof(float2))) != cudaSuccess) { plan_destroy(p); return e; }
    if (!p->zyfused && (e = iq::dmalloc((void**)&p->w3, mp * p->w3_stride * sizeof(float2))) != cudaSuccess) { plan_destroy(p); return e; }
  }
  p->workspace = mp * (p->w1_stride + p->w2_stride + p->w3_stride + p->w4_stride) * sizeof(float2);
  // tensor map of w3 for the TMA-fed inverse y pass: rows = (pair, z, ky), 16-element boxes over all Ny rows of a plane
  if (p->w3 && lines_per_block(p->ly) == 16 && p->Ny <= 256 && p->Nx % 16 == 0) {
    if (const iqtma::EncodeTiledFn encode = iqtma::encode_tiled_fn()) {
      const cuuint64_t gdim[2] = {(cuuint64_t)Nx, (cuuint64_t)mp * (cuuint64_t)p->nzo * (cuuint64_t)Ny};
      const cuuint64_t gstr[1] = {(cuuint64_t)Nx * sizeof(float2)};
      const cuuint32_t box[2] = {16u, (cuuint32_t)Ny};
      const cuuint32_t estr[2] = {1u, 1u};
      const CUresult r = encode(&p->w3_map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, p->w3, gdim, gstr, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      p->w3_map_ok = (r == CUDA_SUCCESS);
    }
  }
  *out = p;
  return cudaSuccess;
}

void plan_destroy(Plan* p) {
  if (!p) return;
  cudaFree(p->twx); cudaFree(p->twy); cudaFree(p->twz);
  cudaFree(p->w1); cudaFree(p->w2); cudaFree(p->w3); cudaFree(p->w4);
  for (auto& kv : p->spectrum) cudaFree(kv.second);
  for (auto& kv : p->sxy) cudaFree(kv.second);
  for (auto& kv : p->sxy_t) cudaFree(kv.second);
  delete p;
}

size_t plan_workspace_bytes(const Plan* p) { return p->workspace; }

cudaError_t plan_set_image(Plan* p, int id, const float* d_img, cudaStream_t s) {
  if (p->spectrum.count(id) || p->sxy.count(id)) return cudaSuccess;
  const long long Nx = p->Nx, Ny = p->Ny, Nz = p->Nz;
  float2 *t1 = nullptr, *t2 = nullptr, *spec = nullptr;
  cudaError_t e;
  // x pass: [nz][ny][Nx]
  if ((e = iq::dmalloc((void**)&t1, (size_t)p->nz * p->ny * Nx * sizeof(float2))) != cudaSuccess) return e;
  RealPassArgs ra{d_img, t1, p->nx, p->nz * p->ny, p->twx};
  if ((e = launch_real(ra, p->lx, s)) != cudaSuccess) return e;
  const float scale = (float)(1.0 / ((double)Nx * (double)Ny * (double)Nz));
  if (p->lz == 0) {
    if ((e = iq::dmalloc((void**)&spec, (size_t)Ny * Nx * sizeof(float2))) != cudaSuccess) return e;
    StridedArgs a{};
    a.in = t1; a.out = spec;
    a.in_sb = 0; a.in_se = Nx; a.in_batch = 0;
    a.out_sb = 0; a.out_se = Nx; a.out_batch = 0;
    a.nin = p->ny; a.flip = 0; a.nout = (int)Ny; a.na = (int)Nx; a.nb = 1; a.tw = p->twy; a.scale = scale;
    if ((e = launch_strided<0>(a, p->ly, 1, s)) != cudaSuccess) return e;
  } else {
    if ((e = iq::dmalloc((void**)&t2, (size_t)p->nz * Ny * Nx * sizeof(float2))) != cudaSuccess) return e;
    if ((e = iq::dmalloc((void**)&spec, (size_t)Nz * Ny * Nx * sizeof(float2))) != cudaSuccess) return e;
    StridedArgs a{};
    a.in = t1; a.out = t2;  // y pass: lines (x, z<nz)
    a.in_sb = (long long)p->ny * Nx; a.in_se = Nx;
    a.out_sb = Ny * Nx; a.out_se = Nx;
    a.nin = p->ny; a.flip = 0; a.nout = (int)Ny; a.na = (int)Nx; a.nb = p->nz; a.tw = p->twy; a.scale = 1.f;
    if (p->zdirect) {
      // direct z pass: keep the (x, y) spectrum of every plane, scaled for the two remaining inverse transforms
      cudaFree(spec);
      a.scale = (float)(1.0 / ((double)Nx * (double)Ny));
      if ((e = launch_strided<0>(a, p->ly, 1, s)) != cudaSuccess) return e;
#ifdef IQB200_EXPERIMENTS
      if (p->zyfused) {
        float2* tt = nullptr;
        if ((e = iq::dmalloc((void**)&tt, (size_t)p->nz * Ny * Nx * sizeof(float2))) != cudaSuccess) return e;
        dim3 tg((unsigned)((Nx + 15) / 16), (unsigned)((Ny + 15) / 16), (unsigned)p->nz);
        k_transpose_xzy<<<tg, 256, 0, s>>>(t2, tt, (int)Nx, (int)Ny, p->nz);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        p->sxy_t[id] = tt;
      }
#endif
      if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
      cudaFree(t1);
      p->sxy[id] = t2;
      return cudaSuccess;
    }
    if ((e = launch_strided<0>(a, p->ly, 1, s)) != cudaSuccess) return e;
    StridedArgs b{};
    b.in = t2; b.out = spec;  // z pass: lines (x, y)
    b.in_sb = Nx; b.in_se = Ny * Nx;
    b.out_sb = Nx; b.out_se = Ny * Nx;
    b.nin = p->nz; b.flip = 0; b.nout = (int)Nz; b.na = (int)Nx; b.nb = (int)Ny; b.tw = p->twz; b.scale = scale;
    if ((e = launch_strided<0>(b, p->lz, 1, s)) != cudaSuccess) return e;
  }
  if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
  cudaFree(t1);
  cudaFree(t2);
  p->spectrum[id] = spec;
  return cudaSuccess;
}

cudaError_t correlate(Plan* p, int id, const float* d_tmpl, int R, const Epilogue& ep, cudaStream_t s, int* launches) {
  const float2* spec = nullptr;
  if (p->zdirect) {
    auto it = p->sxy.find(id);
    if (it == p->sxy.end()) return cudaErrorInvalidValue;
    spec = it->second;
  } else {
    auto it = p->spectrum.find(id);
    if (it == p->spectrum.end()) return cudaErrorInvalidValue;
    spec = it->second;
  }
  const int npair = (R + 1) / 2;
  if (npair > p->max_pairs) return cudaErrorInvalidValue;
  const long long Nx = p->Nx, Ny = p->Ny;
  cudaError_t e;
  int nl = 0;
  TmplPassArgs ta{d_tmpl, p->w1, p->tx, p->ty * p->tz, (long long)p->tx * p->ty * p->tz, R, p->twx};
  if ((e = launch_tmpl(ta, p->lx, npair, s)) != cudaSuccess) return e;
  ++nl;
  if (p->lz == 0) {
    StridedArgs a{};  // fused y: lines (x)
    a.in = p->w1; a.out = p->w4;
    a.in_sb = 0; a.in_se = Nx; a.in_batch = p->w1_stride;
    a.out_sb = 0; a.out_se = Nx; a.out_batch = p->w4_stride;
    a.nin = p->ty; a.flip = 1; a.nout = p->nyo; a.na = (int)Nx; a.nb = 1;
    a.mul = spec; a.mul_sb = 0; a.mul_se = Nx; a.tw = p->twy; a.scale = 1.f;
    if ((e = launch_strided<1>(a, p->ly, npair, s)) != cudaSuccess) return e;
    ++nl;
  } else {
    StridedArgs a{};  // forward y: lines (x, qz)
    a.in = p->w1; a.out = p->w2;
    a.in_sb = (long long)p->ty * Nx; a.in_se = Nx; a.in_batch = p->w1_stride;
    a.out_sb = Ny * Nx; a.out_se = Nx; a.out_batch = p->w2_stride;
    a.nin = p->ty; a.flip = 1; a.nout = (int)Ny; a.na = (int)Nx; a.nb = p->tz; a.tw = p->twy; a.scale = 1.f;
    if (p->zyfused) {  // transposed template spectrum [pair][kx][qz][ky]
      a.out_sa = (long long)p->tz * Ny; a.out_sb = Ny; a.out_se = 1;
    }
    if ((e = launch_strided<0>(a, p->ly, npair, s)) != cudaSuccess) return e;
#ifdef IQB200_EXPERIMENTS
    if (p->zyfused) {
      auto itt = p->sxy_t.find(id);
      if (itt == p->sxy_t.end()) return cudaErrorInvalidValue;
      ZYArgs z{};
      z.sxy_t = itt->second; z.tmpl_t = p->w2; z.out_t = p->w4;
      z.tmpl_batch = p->w2_stride; z.out_batch = p->w4_stride;
      z.nz = p->nz; z.tz = p->tz; z.nzo = p->nzo; z.nyo = p->nyo; z.tw = p->twy;
      const size_t zsm = (size_t)(lines_per_block(8) * line_stride(8) + 256) * sizeof(float2);
      if ((e = set_smem(k_fft_zy<8>, zsm)) != cudaSuccess) return e;
      k_fft_zy<8><<<dim3(npair, (unsigned)Nx), kThreads, zsm, s>>>(z);
      if ((e = cudaGetLastError()) != cudaSuccess) return e;
      FinalArgs fz{p->w4, p->w4_stride, p->nyo * p->nzo, p->nxo, p->npos, R, ep, p->twx};
      if ((e = launch_final(fz, p->lx, npair, s)) != cudaSuccess) return e;
      if (launches) *launches = nl + 3;
      return cudaSuccess;
    }
#endif
    if (p->zdirect) {
      ZDirectArgs z{};
      z.sxy = spec; z.tmpl = p->w2; z.out = p->w3;
      z.plane = Ny * Nx; z.tmpl_batch = p->w2_stride; z.out_batch = p->w3_stride;
      z.nz = p->nz; z.tz = p->tz; z.nzo = p->nzo;
      if ((e = launch_zdirect(z, npair, s)) != cudaSuccess) return e;
    } else {
      StridedArgs b{};  // fused z: lines (x, y)
      b.in = p->w2; b.out = p->w3;
      b.in_sb = Nx; b.in_se = Ny * Nx; b.in_batch = p->w2_stride;
      b.out_sb = Nx; b.out_se = Ny * Nx; b.out_batch = p->w3_stride;
      b.nin = p->tz; b.flip = 1; b.nout = p->nzo; b.na = (int)Nx; b.nb = (int)Ny;
      b.mul = spec; b.mul_sb = Nx; b.mul_se = Ny * Nx; b.tw = p->twz; b.scale = 1.f;
      if ((e = launch_strided<1>(b, p->lz, npair, s)) != cudaSuccess) return e;
    }
    StridedArgs c{};  // inverse y: lines (x, z<nzo)
    c.in = p->w3; c.out = p->w4;
    c.in_sb = Ny * Nx; c.in_se = Nx; c.in_batch = p->w3_stride;
    c.out_sb = (long long)p->nyo * Nx; c.out_se = Nx; c.out_batch = p->w4_stride;
    c.nin = (int)Ny; c.flip = 0; c.nout = p->nyo; c.na = (int)Nx; c.nb = p->nzo; c.tw = p->twy; c.scale = 1.f;
    if ((e = launch_strided<2>(c, p->ly, npair, s, p->w3_map_ok ? &p->w3_map : nullptr)) != cudaSuccess) return e;
    nl += 3;
  }
  FinalArgs fa{p->w4, 0, p->nyo * p->nzo, p->nxo, p->npos, R, ep, p->twx};
  if ((e = launch_final(fa, p->lx, npair, s)) != cudaSuccess) return e;
  ++nl;
  if (launches) *launches = nl;
  return cudaSuccess;
}

int final_chunk_rows(const Plan* p) { return lines_per_block(p->lx); }

double correlate_bytes(const Plan* p, int R) {
  const double npair = (R + 1) / 2, c = sizeof(float2);
  const double Nx = p->Nx, Ny = p->Ny, Nz = p->Nz;
  double b = 4.0 * R * p->tx * p->ty * p->tz + npair * c * p->tz * p->ty * Nx;           // pass A
  if (p->lz == 0) {
    b += npair * c * (p->ty * Nx + p->nyo * Nx) + c * Ny * Nx;                             // fused y (+ spectrum once)
  } else {
    b += npair * c * (p->tz * p->ty * Nx + p->tz * Ny * Nx);                               // forward y
    b += npair * c * (p->tz * Ny * Nx + p->nzo * Ny * Nx) + c * (p->zdirect ? p->nz : Nz) * Ny * Nx;  // z pass (+ spectrum once)
    b += npair * c * (p->nzo * Ny * Nx + (double)p->nzo * p->nyo * Nx);                    // inverse y
    if (p->zyfused) b -= 2.0 * npair * c * p->nzo * Ny * Nx;                               // fused: the z-pass output never leaves the SM
  }
  b += npair * c * (double)p->nzo * p->nyo * Nx + (double)R * 4.0 * p->npos + 4.0 * p->npos;  // final + maps + A2
  return b;
}

}  // namespace iqfft

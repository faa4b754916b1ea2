// iq_ctx.cu -- context, per-batch orchestration and the exported C ABI (include/iqb200.h).
//
// Host-side pieces restated here (FP64, exactly as the reference computes them) because they run on a
// few hundred candidates at most and their result must be bit-identical to the reference's host code:
//   taumodel            /root/reference/src/taumodel.jl:5-45
//   StatsBase.sample    /root/reference/src/iqsim.jl:243 (cumulative walk; Base.sum's pairwise order)
//   relaxation's k      /root/reference/src/relaxation.jl:11,20-22,35 (dbsize / frac arithmetic)
// Everything that touches all npos patch positions runs in the kernels of iq_kernels.cu.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/iqb200.h"
#include "iq_internal.h"
#include "iq_tma.cuh"
#include "iq_fft.h"
#include "iq_cut.h"

#include "iq_ctx.h"

namespace iqimpl {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

uint64_t fnv1a(const uint8_t* p, size_t n) {
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
  return h;
}

// Base.sum's pairwise reduction (Base.mapreduce_impl, block 1024): sequential inside a block.
double julia_sum(const double* w, int64_t lo, int64_t hi) {  // inclusive bounds
  if (hi - lo < 1024) {
    double v = w[lo];
    for (int64_t i = lo + 1; i <= hi; ++i) v += w[i];
    return v;
  }
  const int64_t mid = lo + ((hi - lo) >> 1);
  return julia_sum(w, lo, mid) + julia_sum(w, mid + 1, hi);
}
double julia_sum_const(double c, int64_t lo, int64_t hi) {
  if (hi - lo < 1024) {
    double v = c;
    for (int64_t i = lo + 1; i <= hi; ++i) v += c;
    return v;
  }
  const int64_t mid = lo + ((hi - lo) >> 1);
  return julia_sum_const(c, lo, mid) + julia_sum_const(c, mid + 1, hi);
}

// StatsBase.sample(rng, wv): t = u*sum(wv); first i with cumulative >= t, else the last.
int64_t sample_walk(const double* p, int64_t n, double u) {
  const double t = u * julia_sum(p, 0, n - 1);
  int64_t i = 0;
  double cw = p[0];
  while (cw < t && i < n - 1) { ++i; cw += p[i]; }
  return i;
}

// taumodel (src/taumodel.jl:5-45). vals[j*n + i] = distance of candidate i under source j.
void taumodel(int64_t n, int nsrc, const float* vals, std::vector<double>& prob) {
  prob.assign((size_t)n, 1.0);
  if (n == 1) return;
  const double dn = (double)n;
  const double x0 = (1.0 - 1.0 / dn) / (1.0 / dn);
  std::vector<double> prod((size_t)n, 1.0);
  std::vector<uint32_t> order((size_t)n), tmp((size_t)n), rank((size_t)n);
  for (int j = 0; j < nsrc; ++j) {
    const float* v = vals + (size_t)j * n;
    const uint32_t* bits = reinterpret_cast<const uint32_t*>(v);  // values are >= 0 or +Inf: bit order = numeric order
    // LSD radix sort of the candidate slots by value bits (3 passes of 11 bits)
    for (int64_t i = 0; i < n; ++i) order[(size_t)i] = (uint32_t)i;
    for (int pass = 0; pass < 3; ++pass) {
      const int shift = 11 * pass;
      uint32_t hist[2049] = {0};
      for (int64_t i = 0; i < n; ++i) hist[((bits[order[(size_t)i]] >> shift) & 2047u) + 1]++;
      for (int b = 0; b < 2048; ++b) hist[b + 1] += hist[b];
      for (int64_t i = 0; i < n; ++i) tmp[hist[(bits[order[(size_t)i]] >> shift) & 2047u]++] = order[(size_t)i];
      order.swap(tmp);
    }
    // dense ranks: ties share a rank (taumodel.jl:22-31)
    uint32_t r = 0;
    double colsum = 0.0;
    for (int64_t k = 0; k < n; ++k) {
      if (k == 0 || bits[order[(size_t)k]] != bits[order[(size_t)k - 1]]) ++r;
      rank[order[(size_t)k]] = r;
      colsum += (dn - (double)r) + 1.0;  // integer-valued: exact
    }
    for (int64_t i = 0; i < n; ++i) {
      const double P = (dn - (double)rank[(size_t)i]) + 1.0;
      const double Pi = P / colsum;
      const double X = (1.0 - Pi) / Pi;
      const double ratio = X / x0;
      prod[(size_t)i] = (j == 0) ? ratio : prod[(size_t)i] * ratio;
    }
  }
  for (int64_t i = 0; i < n; ++i) prob[(size_t)i] = 1.0 / (1.0 + x0 * prod[(size_t)i]);
}


int stage_reserve(iq_ctx* c, size_t bytes) {
  if (bytes <= c->stage_cap) return IQ_OK;
  size_t cap = std::max(bytes, c->stage_cap * 2);
  cap = (cap + 4095) & ~(size_t)4095;
  CK(cudaStreamSynchronize(c->stream));
  if (c->h_stage) cudaFreeHost(c->h_stage);
  if (c->d_stage) cudaFree(c->d_stage);
  c->h_stage = nullptr;
  c->d_stage = nullptr;
  c->stage_cap = 0;
  CK(cudaMallocHost((void**)&c->h_stage, cap));
  CK(iq::dmalloc((void**)&c->d_stage, cap));
  c->stage_cap = cap;
  return IQ_OK;
}
size_t stage_alloc(iq_ctx* c, size_t bytes) {
  const size_t off = (c->stage_used + 255) & ~(size_t)255;
  c->stage_used = off + bytes;
  return off;
}

// Disjoint box decomposition of an arbitrary tile mask: x-runs -> rectangles in y -> boxes in z.
struct Rect {
  int x0, x1, y0, y1;
  bool operator<(const Rect& o) const {
    if (x0 != o.x0) return x0 < o.x0;
    if (x1 != o.x1) return x1 < o.x1;
    if (y0 != o.y0) return y0 < o.y0;
    return y1 < o.y1;
  }
};

void decompose(const uint8_t* m, int tx, int ty, int tz, std::vector<BoxDesc>& out) {
  out.clear();
  std::map<Rect, int> open;  // rect -> z start
  auto close_box = [&](const Rect& r, int z0, int z1) {
    BoxDesc b;
    b.x0 = r.x0; b.y0 = r.y0; b.z0 = z0;
    b.w = r.x1 - r.x0; b.h = r.y1 - r.y0; b.d = z1 - z0;
    b.pad = b.x0 & 3;
    b.nch = (b.w + b.pad + 7) / 8;
    b.tmpl_off = 0;
    out.push_back(b);
  };
  for (int z = 0; z <= tz; ++z) {
    std::vector<Rect> rects;
    if (z < tz) {
      std::map<std::pair<int, int>, int> runs_open;  // (x0,x1) -> y start
      for (int y = 0; y <= ty; ++y) {
        std::vector<std::pair<int, int>> runs;
        if (y < ty) {
          const uint8_t* row = m + ((size_t)z * ty + y) * tx;
          int x = 0;
          while (x < tx) {
            if (!row[x]) { ++x; continue; }
            int x1 = x;
            while (x1 < tx && row[x1]) ++x1;
            runs.push_back({x, x1});
            x = x1;
          }
        }
        for (auto it = runs_open.begin(); it != runs_open.end();) {
          if (std::find(runs.begin(), runs.end(), it->first) == runs.end()) {
            rects.push_back({it->first.first, it->first.second, it->second, y});
            it = runs_open.erase(it);
          } else {
            ++it;
          }
        }
        for (auto& r : runs)
          if (!runs_open.count(r)) runs_open[r] = y;
      }
    }
    std::sort(rects.begin(), rects.end());
    for (auto it = open.begin(); it != open.end();) {
      if (!std::binary_search(rects.begin(), rects.end(), it->first)) {
        close_box(it->first, it->second, z);
        it = open.erase(it);
      } else {
        ++it;
      }
    }
    for (auto& r : rects)
      if (!open.count(r)) open[r] = z;
  }
}

int get_mask(iq_ctx* c, const uint8_t* mask, MaskEntry** out) {
  const uint64_t h = fnv1a(mask, (size_t)c->tilevol);
  for (auto& e : c->masks)
    if (e->hash == h && std::memcmp(e->mask.data(), mask, (size_t)c->tilevol) == 0) { *out = e.get(); return IQ_OK; }
  auto e = std::make_unique<MaskEntry>();
  e->mask.assign(mask, mask + c->tilevol);
  e->hash = h;
  for (long long i = 0; i < c->tilevol; ++i) e->nnz += mask[i] ? 1 : 0;
  decompose(mask, c->tx, c->ty, c->tz, e->boxes);
  long long off = 0;
  for (auto& b : e->boxes) {
    b.tmpl_off = (int)off;
    off += (long long)b.d * b.h * b.nch * 8;
  }
  e->tmpl_floats = off;
  if (!e->boxes.empty()) {
    CK(iq::dmalloc((void**)&e->d_boxes, e->boxes.size() * sizeof(BoxDesc)));
    CK(cudaMemcpyAsync(e->d_boxes, e->boxes.data(), e->boxes.size() * sizeof(BoxDesc), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
  }
  CK(iq::dmalloc((void**)&e->d_mask, (size_t)c->tilevol));
  CK(cudaMemcpyAsync(e->d_mask, e->mask.data(), (size_t)c->tilevol, cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  *out = e.get();
  c->masks.push_back(std::move(e));
  return IQ_OK;
}

int get_a2(iq_ctx* c, MaskEntry* e, int image, const float** out) {
  if (e->boxes.empty()) { *out = nullptr; return IQ_OK; }
  auto it = e->a2.find(image);
  if (it != e->a2.end()) { *out = it->second; return IQ_OK; }
  float* d = nullptr;
  CK(iq::dmalloc((void**)&d, (size_t)c->npos * sizeof(float)));
  const double* sat = image < 0 ? c->d_sat_ti : c->d_sat_aux[image];
  CK(iq::launch_a2map(sat, c->nx, c->ny, c->nz, e->d_boxes, (int)e->boxes.size(), d, c->nxo, c->nyo, c->nzo, c->stream));
  c->launches++;
  e->a2[image] = d;
  *out = d;
  return IQ_OK;
}

int pick_rb(const iq_ctx* c, int R) {
  if (c->rb_opt == 1 || c->rb_opt == 2 || c->rb_opt == 4) return c->rb_opt;
  if (R >= 4) return 4;
  if (R >= 2) return 2;
  return 1;
}

// Pack R templates (tile-sized arrays `kern[r]`, masked by e->mask) into the kernel layout
// [grp][box][qz][qy][chunk][r(RB)][8] at staging offset; also B2[r] = sum(mask * kern^2) in FP64.
void pack_templates(const iq_ctx* c, const MaskEntry* e, const float* const* kern, int R, int rb, float* dst, double* b2) {
  const int ngrp = (R + rb - 1) / rb;
  const long long gstride = e->tmpl_floats * rb;
  std::memset(dst, 0, (size_t)ngrp * gstride * sizeof(float));
  const uint8_t* m = e->mask.data();
  std::vector<float> dense((size_t)c->tilevol);
  for (int r = 0; r < R; ++r) {
    const int g = r / rb, ri = r % rb;
    const float* k = kern[r];
    for (long long i = 0; i < c->tilevol; ++i) dense[(size_t)i] = m[i] ? k[i] : 0.f;
    for (const BoxDesc& b : e->boxes) {
      float* base = dst + g * gstride + (long long)b.tmpl_off * rb;
      for (int qz = 0; qz < b.d; ++qz)
        for (int qy = 0; qy < b.h; ++qy) {
          const long long trow = ((long long)(b.z0 + qz) * c->ty + (b.y0 + qy)) * c->tx + b.x0;
          float* drow = base + ((long long)qz * b.h + qy) * b.nch * 8 * rb;
          for (int x = 0; x < b.w; ++x) {
            const float v = m[trow + x] ? k[trow + x] : 0.f;
            const int t = x + b.pad;  // tap index in the packed row (b.pad leading zeros)
            drow[(t >> 3) * 8 * rb + ri * 8 + (t & 7)] = v;
          }
        }
    }
    b2[r] = b2_ordered(dense.data(), c->tx, c->ty, c->tz);
  }
}

double b2_ordered(const float* v, int tx, int ty, int tz) {
  // per z plane: 256 strided partial sums (element i goes to partial i mod 256, increasing i), a fixed binary
  // tree over the partials, then the plane sums are added in z order -- exactly what k_sim_templates does
  const int plane = tx * ty;
  double total = 0.0;
  for (int z = 0; z < tz; ++z) {
    double part[256];
    for (int j = 0; j < 256; ++j) part[j] = 0.0;
    const float* p = v + (size_t)z * plane;
    for (int i = 0; i < plane; ++i) part[i & 255] += (double)p[i] * (double)p[i];
    for (int o = 128; o > 0; o >>= 1)
      for (int j = 0; j < o; ++j) part[j] += part[j + o];
    total = (z == 0) ? part[0] : total + part[0];
  }
  return total;
}

bool all_integer(const float* v, long long n) {
  for (long long i = 0; i < n; ++i)
    if (v[i] != std::nearbyint(v[i]) || std::fabs(v[i]) > 4096.f) return false;
  return true;
}

// Direct kernel or FFT path?  Cost model (milliseconds) fitted to the measured sweep profiles/r02_crossover.csv
// (scripts/crossover_sweep.py on one B200: 8 image / tile geometries x 3 masks x R = 1, 8, 64; CUDA-event times of the
// distance kernels alone).  The model picks the measured winner in 70 of the 72 cases; the two misses are 2-D searches
// of < 0.06 ms where both paths are within launch latency of each other (regret 0.04 ms over the whole sweep).
//   direct: 16 us + nnz * npos * R / eff,  eff = 30 TFMA/s * nnz / (nnz + n0) (n0 = 1500 voxels in 3-D, 300 in 2-D: small
//           masks amortise the staging of the image brick badly), x 0.63 / 0.8 for R = 1 / 2-3 (fewer templates per pass);
//   FFT:    24 us + pairs * Nx * Ny * nz * 11 ps in 3-D (x, y padded to powers of two; z direct);
//           24 us + pairs * (4.2 us + Nx * Ny * 17 ps) in 2-D (the fused y pass walks the pairs inside a CTA).
bool want_fft(const iq_ctx* c, const MaskEntry* e, int R) {
  if (c->fft_mode < 0 || c->fft_failed || e->nnz == 0) return false;
  if (c->ny < 2 || c->nx < 2) return false;
  if (c->fft_mode > 0) return true;
  const bool three_d = c->nz > 1;
  const double nnz = (double)e->nnz;
  const double eff = 30e12 * nnz / (nnz + (three_d ? 1500.0 : 300.0)) * (R >= 4 ? 1.0 : (R >= 2 ? 0.8 : 0.63));
  const double direct = 0.016 + nnz * (double)c->npos * R / eff * 1e3;
  auto p2 = [](int n) { double v = 1; while (v < n) v *= 2; return v; };
  const double plane = p2(c->nx) * p2(c->ny);
  const double pairs = (R + 1) / 2;
  const double fft = three_d ? 0.024 + pairs * plane * c->nz * 1.1e-8 : 0.024 + pairs * (0.0042 + plane * 1.7e-8);
  return fft < direct;
}

// Records a start/stop event pair around one distance computation (per-search kernel timing for benchmarks).
static int dist_events(iq_ctx* c, cudaEvent_t* a, cudaEvent_t* b, bool fft) {
  if (c->dist_ev_used + 2 > c->dist_ev.size()) {
    cudaEvent_t x, y;
    CK(cudaEventCreate(&x));
    CK(cudaEventCreate(&y));
    c->dist_ev.push_back(x);
    c->dist_ev.push_back(y);
  }
  *a = c->dist_ev[c->dist_ev_used];
  *b = c->dist_ev[c->dist_ev_used + 1];
  if (c->dist_ev_fft.size() < c->dist_ev.size() / 2) c->dist_ev_fft.resize(c->dist_ev.size() / 2, 0);
  c->dist_ev_fft[c->dist_ev_used / 2] = fft ? 1 : 0;
  c->dist_ev_used += 2;
  return IQ_OK;
}

int ensure_fft(iq_ctx* c, int image) {
  if (!c->fft) {
    cudaError_t ce = iqfft::plan_create(&c->fft, c->nx, c->ny, c->nz, c->tx, c->ty, c->tz, c->max_batch, c->stream);
    if (ce != cudaSuccess) {
      cudaGetLastError();
      c->fft_failed = true;
      return IQ_ERR_STATE;  // caller falls back to the direct kernel (still on the GPU)
    }
  }
  const float* d_img = image < 0 ? c->d_ti : c->d_aux[image];
  CK(iqfft::plan_set_image(c->fft, image, d_img, c->stream));
  return IQ_OK;
}

// FFT correlation of R dense masked templates that already sit in device memory.
int launch_fft(iq_ctx* c, MaskEntry* e, int image, const float* d_tmpl, const double* d_b2, int R, bool tint, float* d_out,
               int kind, int job0, const float* const* a2_list) {
  const float* a2 = nullptr;
  int rc = IQ_OK;
  if (!a2_list) {
    rc = get_a2(c, e, image, &a2);
    if (rc) return rc;
  }
  iqfft::Epilogue ep{};
  ep.a2 = a2;
  ep.a2_list = a2_list;
  ep.b2 = d_b2;
  ep.disabled = c->d_disabled;
  ep.out = d_out;
  ep.minbits = c->d_minmax + (size_t)(kind * 2 + 0) * c->max_batch + job0;
  ep.maxbits = c->d_minmax + (size_t)(kind * 2 + 1) * c->max_batch + job0;
  ep.round_to_int = tint ? 1 : 0;
  if (kind == 0) {
    const int rows = iqfft::final_chunk_rows(c->fft);
    c->chunk_len = rows * c->nxo;
    c->chunk_n = (c->nyo * c->nzo + rows - 1) / rows;
    c->chunk_valid = c->chunk_n <= c->chunk_stride;
    if (c->chunk_valid) { ep.chunkmin = c->d_chunkmin + (size_t)job0 * c->chunk_stride; ep.chunk_pitch = c->chunk_stride; }
  }
  cudaEvent_t ea, eb;
  rc = dist_events(c, &ea, &eb, true);
  if (rc) return rc;
  int nl = 0;
  CK(cudaEventRecord(ea, c->stream));
  CK(iqfft::correlate(c->fft, image, d_tmpl, R, ep, c->stream, &nl));
  CK(cudaEventRecord(eb, c->stream));
  c->launches += nl;
  c->last_fft_bytes += iqfft::correlate_bytes(c->fft, R);
  c->last_fft_searches += R;
  return IQ_OK;
}

int run_fft(iq_ctx* c, MaskEntry* e, int image, const float* const* kern, int R, float* d_out, int kind) {
  int rc = ensure_fft(c, image);
  if (rc) return rc;
  const size_t tbytes = (size_t)R * c->tilevol * sizeof(float);
  const size_t off_t = stage_alloc(c, tbytes);
  const size_t off_b = stage_alloc(c, (size_t)R * sizeof(double));
  if (c->stage_used > c->stage_cap) return fail(IQ_ERR_STATE, "staging overflow (internal)");
  float* wk = (float*)(c->h_stage + off_t);
  double* b2 = (double*)(c->h_stage + off_b);
  const uint8_t* m = e->mask.data();
  bool tint = c->image_is_int[image];
  for (int r = 0; r < R; ++r) {
    float* dst = wk + (size_t)r * c->tilevol;
    const float* k = kern[r];
    for (long long i = 0; i < c->tilevol; ++i) dst[i] = m[i] ? k[i] : 0.f;
    b2[r] = b2_ordered(dst, c->tx, c->ty, c->tz);
    if (tint) tint = all_integer(dst, c->tilevol);
  }
  CK(cudaMemcpyAsync(c->d_stage + off_t, c->h_stage + off_t, (off_b + R * sizeof(double)) - off_t, cudaMemcpyHostToDevice,
                     c->stream));
  return launch_fft(c, e, image, (const float*)(c->d_stage + off_t), (const double*)(c->d_stage + off_b), R, tint, d_out, kind);
}

// Tensor maps of `image` for the TMA-staged direct kernel at panel shape (XT, RS) (one per mask box), cached on the mask.
// ok = false: no tensor-map encoder in this driver, or the encode failed -> the caller uses the register-staged kernel.
static int get_flat_tma(iq_ctx* c, MaskEntry* e, int image, int XT, int RS, const iq::FlatTmaMaps** out, bool* ok) {
  *ok = false;
  const auto key = std::make_pair(image, XT * 16 + RS);
  auto it = e->tma.find(key);
  if (it != e->tma.end()) { *out = &it->second; *ok = true; return IQ_OK; }
  static const iqtma::EncodeTiledFn encode = iqtma::encode_tiled_fn();
  if (!encode) return IQ_OK;
  // TMA needs a row pitch of whole 16-byte units: images whose nx is not a multiple of 4 get a padded copy (once)
  const float* img = image < 0 ? c->d_ti : c->d_aux[image];
  const int nxp = (c->nx + 3) & ~3;
  if (nxp != c->nx) {
    auto ip = c->img_pad.find(image);
    if (ip == c->img_pad.end()) {
      float* d = nullptr;
      CK(iq::dmalloc((void**)&d, (size_t)nxp * c->ny * c->nz * sizeof(float)));
      CK(cudaMemsetAsync(d, 0, (size_t)nxp * c->ny * c->nz * sizeof(float), c->stream));
      CK(cudaMemcpy2DAsync(d, (size_t)nxp * sizeof(float), img, (size_t)c->nx * sizeof(float), (size_t)c->nx * sizeof(float),
                           (size_t)c->ny * c->nz, cudaMemcpyDeviceToDevice, c->stream));
      ip = c->img_pad.emplace(image, d).first;
    }
    img = ip->second;
  }
  iq::FlatTmaMaps maps;
  std::memset(&maps, 0, sizeof(maps));
  for (size_t b = 0; b < e->boxes.size(); ++b) {
    int bw = 0, bh = 0;
    iq::dist_flat_box(e->boxes[b], XT, RS, &bw, &bh);
    const cuuint64_t gdim[3] = {(cuuint64_t)c->nx, (cuuint64_t)c->ny, (cuuint64_t)c->nz};
    const cuuint64_t gstr[2] = {(cuuint64_t)nxp * sizeof(float), (cuuint64_t)nxp * c->ny * sizeof(float)};
    const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUresult r = encode(&maps.m[b], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)img, gdim, gstr, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return IQ_OK;
  }
  *out = &e->tma.emplace(key, maps).first->second;
  *ok = true;
  return IQ_OK;
}

// Direct correlation kernel on R templates already packed in device memory ([grp][box][qz][qy][chunk][r(rb)][8]).
int launch_direct(iq_ctx* c, MaskEntry* e, int image, const float* d_packed, const double* d_b2, int R, int rb, float* d_out,
                  int kind, int job0) {
  const float* a2 = nullptr;
  int rc = get_a2(c, e, image, &a2);
  if (rc) return rc;
  iq::DistParams p{};
  p.img = image < 0 ? c->d_ti : c->d_aux[image];
  p.nx = c->nx; p.ny = c->ny; p.nz = c->nz;
  p.nxo = c->nxo; p.nyo = c->nyo; p.nzo = c->nzo;
  p.npos = c->npos;
  p.nbox = (int)e->boxes.size();
  p.boxes = e->d_boxes;
  p.tmpl = d_packed;
  p.tmpl_grp_stride = e->tmpl_floats * rb;
  p.a2 = a2;
  p.b2 = d_b2;
  p.disabled = c->d_disabled;
  p.out = d_out;
  p.minbits = c->d_minmax + (size_t)(kind * 2 + 0) * c->max_batch + job0;
  p.maxbits = c->d_minmax + (size_t)(kind * 2 + 1) * c->max_batch + job0;
  p.R = R;
  const int nxt = (c->nxo + iq::kT - 1) / iq::kT;
  size_t smem = 0;
  const iq::FlatTmaMaps* maps = nullptr;
  bool tma = false;
  if (c->variant == 0) {
    // TMA kernel: panel width XT, octet shape RS and number of stage buffers NB that minimise the padded work (items
    // are padded to whole octets: XT to a multiple of 8 / RS x-threads, the rows to a multiple of RS), preferring
    // shapes whose shared memory lets two CTAs share an SM (one CTA per SM measured ~1.35x slower on large masks)
    // and double buffering (single: the two CTAs of an SM overlap each other's loads instead)
    const size_t lim2 = 115000, lim1 = 227 * 1024;
    double best = 1e300;
    int bxt = 0, brs = 0, bnb = 0;
    for (int rs = 2; rs <= 8; rs *= 4) {
      const int xs = 8 / rs;
      for (int np = 1; np <= nxt; ++np) {
        const int xt = (((nxt + np - 1) / np) + xs - 1) / xs * xs;
        if (np > 1 && (((nxt + np - 2) / (np - 1)) + xs - 1) / xs * xs == xt) continue;  // same width as with fewer panels
        for (int nb = 2; nb >= 1; --nb) {
          const size_t need = iq::dist_flat_smem(e->boxes.data(), p.nbox, xt, rs, nb, rb, nullptr, nullptr);
          if (!need || need > lim1) continue;
          const int npan = (nxt + xt - 1) / xt;
          double cost = (double)npan * xt * (double)((c->nyo + rs - 1) / rs * rs) * (1.0 + 0.02 * npan);
          if (need > lim2) cost *= 1.35;
          if (nb == 1) cost *= 1.03;
          if (cost < best) { best = cost; bxt = xt; brs = rs; bnb = nb; }
        }
      }
    }
    if (bxt) {
      rc = get_flat_tma(c, e, image, bxt, brs, &maps, &tma);
      if (rc) return rc;
    }
    if (tma) {
      p.XT = bxt; p.RS = brs; p.NB = bnb;
      smem = iq::dist_flat_smem(e->boxes.data(), p.nbox, p.XT, p.RS, p.NB, rb, &p.patch_floats, &p.tmpl_floats);
    }
  }
  if (!tma) {
    // register-staged kernel: fewest column panels whose patch + templates still let two CTAs share an SM
    int best_xt = 0;
    for (int pass = 0; pass < 2 && !best_xt; ++pass) {
      const size_t limit = pass == 0 ? 110 * 1024 : 220 * 1024;
      for (int np = 1; np <= nxt; ++np) {
        const int xt = (nxt + np - 1) / np;
        if (xt > 256) continue;
        if (iq::dist_flat_ldg_smem(e->boxes.data(), p.nbox, xt, rb, nullptr) <= limit) { best_xt = xt; break; }
      }
    }
    if (!best_xt) return fail(IQ_ERR_INVALID, "tile too large for the shared-memory staging of the direct kernel");
    p.XT = best_xt;
    smem = iq::dist_flat_ldg_smem(e->boxes.data(), p.nbox, p.XT, rb, &p.patch_floats);
  }
  cudaEvent_t ea, eb;
  rc = dist_events(c, &ea, &eb, false);
  if (rc) return rc;
  CK(cudaEventRecord(ea, c->stream));
  if (tma) CK(iq::launch_dist_flat(p, *maps, rb, smem, c->stream));
  else CK(iq::launch_dist_flat_ldg(p, rb, std::max<size_t>(smem, 64), c->stream));
  CK(cudaEventRecord(eb, c->stream));
  c->launches++;
  c->last_direct_searches += R;
  (tma ? c->direct_tma_launches : c->direct_ldg_launches)++;
  if (kind == 0) c->chunk_valid = false;
  return IQ_OK;
}

int ensure_chunkmin(iq_ctx* c, int R) {
  if (c->chunk_valid) return IQ_OK;
  c->chunk_len = 4096;
  c->chunk_n = (int)((c->npos + 4095) / 4096);
  CK(iq::launch_chunkmin(c->d_Dovl, R, c->npos, c->chunk_len, c->chunk_n, c->d_chunkmin, c->chunk_stride, c->stream));
  c->launches++;
  c->chunk_valid = true;
  return IQ_OK;
}

int run_dense(iq_ctx* c, MaskEntry* e, int image, const float* const* kern, int R, float* d_out, int kind) {
  if (want_fft(c, e, R)) {
    const int rcf = run_fft(c, e, image, kern, R, d_out, kind);
    if (!(rcf == IQ_ERR_STATE && c->fft_failed)) return rcf;
  }
  const int rb = pick_rb(c, R);
  const int ngrp = (R + rb - 1) / rb;
  const size_t tbytes = (size_t)ngrp * e->tmpl_floats * rb * sizeof(float);
  const size_t off_t = stage_alloc(c, std::max<size_t>(tbytes, 16));
  const size_t off_b = stage_alloc(c, (size_t)R * sizeof(double));
  if (c->stage_used > c->stage_cap) return fail(IQ_ERR_STATE, "staging overflow (internal)");
  pack_templates(c, e, kern, R, rb, (float*)(c->h_stage + off_t), (double*)(c->h_stage + off_b));
  CK(cudaMemcpyAsync(c->d_stage + off_t, c->h_stage + off_t, (off_b + R * sizeof(double)) - off_t, cudaMemcpyHostToDevice,
                     c->stream));
  return launch_direct(c, e, image, (const float*)(c->d_stage + off_t), (const double*)(c->d_stage + off_b), R, rb, d_out, kind);
}

// Sums the event pairs recorded since dist_ev_used was last reset into last_dist_ms / last_fft_ms.
int collect_dist_times(iq_ctx* c) {
  c->last_dist_ms = 0.0;
  c->last_fft_ms = 0.0;
  c->last_dist_launches = (int64_t)(c->dist_ev_used / 2);
  for (size_t i = 0; i + 1 < c->dist_ev_used; i += 2) {
    float dm = 0.f;
    CK(cudaEventElapsedTime(&dm, c->dist_ev[i], c->dist_ev[i + 1]));
    c->last_dist_ms += dm;
    if (c->dist_ev_fft[i / 2]) c->last_fft_ms += dm;
  }
  return IQ_OK;
}

int build_uniform(iq_ctx* c) {
  if (!c->enabled_idx.empty() || c->nenabled == 0) return IQ_OK;
  c->enabled_idx.reserve((size_t)c->nenabled);
  for (long long p = 0; p < c->npos; ++p)
    if (c->h_disabled.empty() || !c->h_disabled[p]) c->enabled_idx.push_back(p);
  const int64_t n = (int64_t)c->enabled_idx.size();
  std::vector<float> zeros((size_t)std::min<int64_t>(n, 2), 0.f);
  std::vector<double> pr;
  double pc = 1.0;
  if (n > 1) {  // all distances equal: every rank is 1 -> every probability is the same number
    taumodel(2, 1, zeros.data(), pr);  // placeholder to keep the code path exercised
    const double dn = (double)n;
    const double x0 = (1.0 - 1.0 / dn) / (1.0 / dn);
    const double colsum = ((dn - 1.0) + 1.0) * dn;  // n entries each (n - 1) + 1
    const double Pi = ((dn - 1.0) + 1.0) / colsum;
    const double X = (1.0 - Pi) / Pi;
    pc = 1.0 / (1.0 + x0 * (X / x0));
  }
  c->uniform_prob.assign((size_t)n, pc);
  c->uniform_sum = julia_sum_const(pc, 0, n - 1);
  c->uniform_cum.resize((size_t)n);
  double cw = 0.0;
  for (int64_t i = 0; i < n; ++i) { cw = (i == 0) ? pc : cw + pc; c->uniform_cum[i] = cw; }
  return IQ_OK;
}

int search_chunk(iq_ctx* c, MaskEntry* e, const iq_tile* tiles, int R, double tol, const double* u, iq_result* results,
                 int res_base) {
  const int S = c->nsoft;
  bool any_hard = false;
  for (int r = 0; r < R; ++r) any_hard |= tiles[r].hard_nnz > 0;

  // ---- fast path: empty overlap mask, no auxiliary information -> every enabled patch, uniform ----
  if (e->nnz == 0 && !any_hard && S == 0) {
    int rc = build_uniform(c);
    if (rc) return rc;
    const int64_t n = (int64_t)c->enabled_idx.size();
    if (n == 0) return fail(IQ_ERR_INVALID, "all patches of the training image are disabled");
    for (int r = 0; r < R; ++r) {
      TileResult& tr = c->res[res_base + r];
      tr.idx_ptr = c->enabled_idx.data();
      tr.prob_ptr = c->uniform_prob.data();
      tr.count = n;
      iq_result& o = results[r];
      o.count = n; o.idx = tr.idx_ptr; o.prob = tr.prob_ptr; o.picked = -1; o.relax_iters = 0; o.dmin = 0.f;
      if (u) {
        const double t = u[r] * c->uniform_sum;
        // first i with cum[i] >= t among i < n-1, else n-1 (cum is non-decreasing)
        const auto it = std::lower_bound(c->uniform_cum.begin(), c->uniform_cum.end() - 1, t);
        o.picked = c->enabled_idx[(size_t)(it - c->uniform_cum.begin())];
      }
    }
    return IQ_OK;
  }

  // ---- staging size ----
  const int rb = pick_rb(c, R);
  const int ngrp = (R + rb - 1) / rb;
  size_t need = 4096 + (size_t)ngrp * rb * e->tmpl_floats * sizeof(float) + R * sizeof(double) + 512;
  need += (size_t)(1 + S) * ((size_t)R * c->tilevol * sizeof(float) + 1024);  // dense templates of the FFT path
  if (S > 0) need += (size_t)S * ((size_t)ngrp * rb * c->full_mask->tmpl_floats * sizeof(float) + R * sizeof(double) + 512);
  long long hard_total = 0;
  for (int r = 0; r < R; ++r) hard_total += tiles[r].hard_nnz;
  need += (size_t)hard_total * (sizeof(long long) + sizeof(float)) + (R + 1) * sizeof(int) + 1024;
  int rc = stage_reserve(c, need);
  if (rc) return rc;
  c->stage_used = 0;

  const int nkind = 2 + S;
  // init min = +Inf bits, max = 0 for every kind
  for (int k = 0; k < nkind; ++k) {
    CK(iq::launch_fill_u32(c->d_minmax + (size_t)(k * 2 + 0) * c->max_batch, 0x7f800000u, c->max_batch, c->stream));
    CK(iq::launch_fill_u32(c->d_minmax + (size_t)(k * 2 + 1) * c->max_batch, 0u, c->max_batch, c->stream));
    c->launches += 2;
  }

  // ---- distances ----
  std::vector<const float*> kern(R);
  for (int r = 0; r < R; ++r) kern[r] = tiles[r].simdev;
  rc = run_dense(c, e, -1, kern.data(), R, c->d_Dovl, 0);
  if (rc) return rc;
  if (any_hard) {
    const size_t off_ptr = stage_alloc(c, (R + 1) * sizeof(int));
    const size_t off_off = stage_alloc(c, (size_t)hard_total * sizeof(long long));
    const size_t off_val = stage_alloc(c, (size_t)hard_total * sizeof(float));
    int* ptr = (int*)(c->h_stage + off_ptr);
    long long* off = (long long*)(c->h_stage + off_off);
    float* val = (float*)(c->h_stage + off_val);
    int n = 0;
    for (int r = 0; r < R; ++r) {
      ptr[r] = n;
      for (int i = 0; i < tiles[r].hard_nnz; ++i) {
        const int o = tiles[r].hard_offset[i];
        if (o < 0 || o >= c->tilevol) return fail(IQ_ERR_INVALID, "hard_offset out of the tile");
        const int qx = o % c->tx, qy = (o / c->tx) % c->ty, qz = o / (c->tx * c->ty);
        off[n] = ((long long)qz * c->ny + qy) * c->nx + qx;
        val[n] = tiles[r].hard_value[i];
        ++n;
      }
    }
    ptr[R] = n;
    CK(cudaMemcpyAsync(c->d_stage + off_ptr, c->h_stage + off_ptr, c->stage_used - off_ptr, cudaMemcpyHostToDevice, c->stream));
    iq::SparseParams sp{};
    sp.img = c->d_ti;
    sp.nx = c->nx; sp.ny = c->ny; sp.nz = c->nz; sp.nxo = c->nxo; sp.nyo = c->nyo; sp.nzo = c->nzo;
    sp.npos = c->npos;
    sp.ptr = (const int*)(c->d_stage + off_ptr);
    sp.ptr_stride = 1;
    sp.off = (const long long*)(c->d_stage + off_off);
    sp.val = (const float*)(c->d_stage + off_val);
    sp.disabled = c->d_disabled;
    sp.out = c->d_Dhard;
    sp.minbits = c->d_minmax + (size_t)(1 * 2 + 0) * c->max_batch;
    sp.maxbits = c->d_minmax + (size_t)(1 * 2 + 1) * c->max_batch;
    sp.R = R;
    CK(iq::launch_dist_sparse(sp, c->stream));
    c->launches++;
  }
  for (int s = 0; s < S; ++s) {
    for (int r = 0; r < R; ++r) {
      if (!tiles[r].softdev || !tiles[r].softdev[s]) return fail(IQ_ERR_INVALID, "tile %d lacks softdev[%d]", r, s);
      kern[r] = tiles[r].softdev[s];
    }
    rc = run_dense(c, c->full_mask, s, kern.data(), R, c->d_Dsoft[s], 2 + s);
    if (rc) return rc;
  }

  // ---- per-tile source lists (src/iqsim.jl:230-234) ----
  std::vector<int> nsrc(R);
  std::vector<int> prim_kind(R);
  bool any_relax = false;
  for (int r = 0; r < R; ++r) {
    iq::PickJob& J = c->h_pick[r];
    std::memset(&J, 0, sizeof J);
    int n = 0;
    const bool hardtile = tiles[r].hard_nnz > 0;
    if (hardtile) J.src[n++] = c->d_Dhard + (size_t)r * c->npos;
    J.src[n++] = c->d_Dovl + (size_t)r * c->npos;
    for (int s = 0; s < S; ++s) J.src[n++] = c->d_Dsoft[s] + (size_t)r * c->npos;
    nsrc[r] = n;
    prim_kind[r] = hardtile ? 1 : 0;
    J.nsrc = n;
    J.mode = n > 1 ? 1 : 0;
    J.tol = tol;
    J.minbits = c->d_minmax + (size_t)(prim_kind[r] * 2 + 0) * c->max_batch + r;
    J.blockcount = c->d_blockcount + (size_t)r * iq::pick_nblk(c->npos);
    J.total = c->d_total + r;
    J.ticket = 0;
    J.cand_idx = c->d_cand_idx + (size_t)r * c->npos;
    J.cand_val = c->d_cand_val + (size_t)r * c->max_src * c->npos;
    J.cap = c->npos;
    J.chunkmin = c->d_chunkmin + (size_t)r * c->chunk_stride;
    any_relax |= n > 1;
  }

  // ---- relaxation bookkeeping (src/relaxation.jl:7-22) ----
  std::vector<double> frac(R, 0.0);
  std::vector<long long> dbsize(R, 0);
  std::vector<int> iters(R, 0);
  std::vector<char> pending(R, 0);
  const long long npatterns = c->nenabled;
  if (any_relax) {
    if (npatterns <= 0) return fail(IQ_ERR_INVALID, "all patches of the training image are disabled");
    CK(cudaMemcpyAsync(c->h_minmax, c->d_minmax, (size_t)nkind * 2 * c->max_batch * sizeof(unsigned), cudaMemcpyDeviceToHost,
                       c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (int r = 0; r < R; ++r) {
      if (nsrc[r] <= 1) continue;
      const unsigned mx = c->h_minmax[(size_t)(prim_kind[r] * 2 + 1) * c->max_batch + r];
      const bool allzero = (mx == 0u);  // all(distance[enabled] .== 0)
      dbsize[r] = allzero ? npatterns : (long long)std::ceil(tol * (double)npatterns);
      frac[r] = 0.1 * ((double)dbsize[r] / (double)npatterns);
      pending[r] = 1;
    }
  }

  // ---- selection ----
  const int maxS = c->max_src;
  bool first_round = true;
  bool by_chunks = !any_relax;  // threshold rule on the overlap distance: scan only chunks that can hold candidates
  if (by_chunks) {
    rc = ensure_chunkmin(c, R);
    if (rc) return rc;
    for (int r = 0; r < R; ++r) c->h_pick[r].sel = c->d_sel + (size_t)r * maxS;
    CK(cudaMemcpyAsync(c->d_pick, c->h_pick, (size_t)R * sizeof(iq::PickJob), cudaMemcpyHostToDevice, c->stream));
    CK(iq::launch_pick_chunks(c->d_pick, R, c->npos, c->chunk_len, c->chunk_n, c->stream));
    c->launches++;
    CK(cudaMemcpyAsync(c->h_total, c->d_total, (size_t)R * sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (int r = 0; r < R; ++r)
      if (c->h_total[r] == iq::kPickOverflow) by_chunks = false;  // too many chunks qualify: generic two-pass path
  }
  for (; !by_chunks;) {
    int njobs = 0;
    if (any_relax) {
      for (int r = 0; r < R; ++r) {
        if (!pending[r]) continue;
        const long long softk = (long long)std::ceil(frac[r] * (double)npatterns);
        for (int s = 0; s < nsrc[r]; ++s) {
          iq::SelJob& sj = c->h_sel[(size_t)r * maxS + s];
          if (s == 0 && !first_round) continue;  // primary threshold already known
          std::memset(&sj, 0, sizeof sj);
          sj.map = c->h_pick[r].src[s];
          sj.k = (unsigned long long)std::max<long long>(1, std::min<long long>(s == 0 ? dbsize[r] : softk, c->npos));
          sj.active = 1;
          sj.cbuf = c->d_selbuf + ((size_t)r * maxS + s) * c->sel_cap;
          sj.ccap = c->sel_cap;
          ++njobs;
        }
        iters[r]++;
      }
      if (njobs > 0) {
        // upload every job of the batch (inactive ones are skipped by the kernel)
        if (first_round) {
          for (int r = 0; r < R; ++r)
            for (int s = 0; s < maxS; ++s)
              if (!(pending[r] && s < nsrc[r])) std::memset(&c->h_sel[(size_t)r * maxS + s], 0, sizeof(iq::SelJob));
          CK(cudaMemcpyAsync(c->d_sel, c->h_sel, (size_t)R * maxS * sizeof(iq::SelJob), cudaMemcpyHostToDevice, c->stream));
        } else {
          for (int r = 0; r < R; ++r) {
            if (!pending[r]) continue;
            CK(cudaMemcpyAsync(c->d_sel + (size_t)r * maxS + 1, c->h_sel + (size_t)r * maxS + 1,
                               (size_t)(nsrc[r] - 1) * sizeof(iq::SelJob), cudaMemcpyHostToDevice, c->stream));
          }
        }
        {
          int nl = 0;
          CK(iq::launch_select_all(c->d_sel, R * maxS, c->npos, c->d_shifts, c->nshift, c->stream, &nl, c->d_sel_list));
          c->launches += nl;
        }
      }
    }
    // pick jobs read the finished SelJob thresholds straight from device memory (no host round trip)
    if (first_round) {
      for (int r = 0; r < R; ++r) c->h_pick[r].sel = c->d_sel + (size_t)r * maxS;
      CK(cudaMemcpyAsync(c->d_pick, c->h_pick, (size_t)R * sizeof(iq::PickJob), cudaMemcpyHostToDevice, c->stream));
    }
    CK(iq::launch_pick_count(c->d_pick, R, c->npos, c->stream));
    c->launches++;
    CK(cudaMemcpyAsync(c->h_total, c->d_total, (size_t)R * sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
    if (!any_relax) break;  // totals are read together with the candidates below
    CK(cudaStreamSynchronize(c->stream));
    bool again = false;
    for (int r = 0; r < R; ++r) {
      if (!pending[r]) continue;
      if (c->h_total[r] > 0) { pending[r] = 0; continue; }
      if (frac[r] >= 1.0) return fail(IQ_ERR_STATE, "relaxation found no candidate at frac = 1 (internal)");
      frac[r] = std::min(frac[r] + 0.1, 1.0);
      again = true;
    }
    first_round = false;
    if (!again) break;
  }

  if (!by_chunks) {
    CK(iq::launch_pick_write(c->d_pick, R, c->npos, c->stream));
    c->launches++;
  }
  if (c->tau_device) {
    CK(iq::launch_tau(c->d_pick, R, maxS, c->d_rank, c->d_colsum, c->d_prob, c->stream));
    c->launches += 2;
  }
  CK(cudaMemcpyAsync(c->h_minmax, c->d_minmax, (size_t)nkind * 2 * c->max_batch * sizeof(unsigned), cudaMemcpyDeviceToHost,
                     c->stream));
  CK(cudaStreamSynchronize(c->stream));  // totals now valid

  // ---- candidates back to the host ----
  size_t tot = 0;
  for (int r = 0; r < R; ++r) tot += c->h_total[r];
  if (tot > c->h_cand_cap) {
    if (c->h_cand_idx) cudaFreeHost(c->h_cand_idx);
    if (c->h_cand_val) cudaFreeHost(c->h_cand_val);
    c->h_cand_idx = nullptr; c->h_cand_val = nullptr;
    const size_t cap = std::max<size_t>(tot * 2, 1 << 16);
    CK(cudaMallocHost((void**)&c->h_cand_idx, cap * sizeof(unsigned)));
    CK(cudaMallocHost((void**)&c->h_cand_val, cap * maxS * sizeof(float)));
    c->h_cand_cap = cap;
  }
  if (c->tau_device && !c->h_prob) CK(cudaMallocHost((void**)&c->h_prob, (size_t)c->max_batch * iq::kTauMax * sizeof(double)));
  {
    size_t o = 0;
    for (int r = 0; r < R; ++r) {
      const size_t n = c->h_total[r];
      if (n == 0) continue;
      CK(cudaMemcpyAsync(c->h_cand_idx + o, c->d_cand_idx + (size_t)r * c->npos, n * sizeof(unsigned), cudaMemcpyDeviceToHost,
                         c->stream));
      const bool on_device = c->tau_device && n >= 2 && n <= (size_t)iq::kTauMax;
      if (on_device) {
        CK(cudaMemcpyAsync(c->h_prob + (size_t)r * iq::kTauMax, c->d_prob + (size_t)r * iq::kTauMax, n * sizeof(double),
                           cudaMemcpyDeviceToHost, c->stream));
      } else {
        for (int s = 0; s < nsrc[r]; ++s)
          CK(cudaMemcpyAsync(c->h_cand_val + (o * maxS) + (size_t)s * n, c->d_cand_val + ((size_t)r * maxS + s) * c->npos,
                             n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
      }
      o += n;
    }
    CK(cudaStreamSynchronize(c->stream));
  }

  // ---- tau model + optional sampling on the host (FP64, candidate-set sized) ----
  {
    size_t o = 0;
    for (int r = 0; r < R; ++r) {
      const int64_t n = c->h_total[r];
      TileResult& tr = c->res[res_base + r];
      tr.idx.resize((size_t)n);
      for (int64_t i = 0; i < n; ++i) tr.idx[(size_t)i] = (int64_t)c->h_cand_idx[o + i];
      const bool on_device = c->tau_device && n >= 2 && n <= (int64_t)iq::kTauMax;
      if (on_device) tr.prob.assign(c->h_prob + (size_t)r * iq::kTauMax, c->h_prob + (size_t)r * iq::kTauMax + n);
      else if (n > 0) taumodel(n, nsrc[r], c->h_cand_val + o * maxS, tr.prob);
      else tr.prob.clear();
      tr.idx_ptr = tr.idx.data();
      tr.prob_ptr = tr.prob.data();
      tr.count = n;
      iq_result& out = results[r];
      out.count = n;
      out.idx = tr.idx_ptr;
      out.prob = tr.prob_ptr;
      out.picked = -1;
      out.relax_iters = iters[r];
      const unsigned mb = c->h_minmax[(size_t)(prim_kind[r] * 2 + 0) * c->max_batch + r];
      std::memcpy(&out.dmin, &mb, 4);
      if (u && n > 0) out.picked = tr.idx[(size_t)sample_walk(tr.prob.data(), n, u[r])];
      o += n;
    }
  }
  return IQ_OK;
}

int do_search(iq_ctx* c, const uint8_t* ovlmask, const iq_tile* tiles, int ntile, double tol, const double* u, iq_result* results) {
  if (!c || !ovlmask || !tiles || !results || ntile <= 0) return fail(IQ_ERR_INVALID, "iq_search: NULL argument or ntile <= 0");
  if (!(tol > 0.0 && tol <= 1.0)) return fail(IQ_ERR_INVALID, "tolerance must be in range (0,1]");
  for (int i = 0; i < ntile; ++i)
    if (!tiles[i].simdev) return fail(IQ_ERR_INVALID, "tile %d: simdev is NULL", i);
  CK(cudaSetDevice(c->device));
  MaskEntry* e = nullptr;
  int rc = get_mask(c, ovlmask, &e);
  if (rc) return rc;
  if ((int)c->res.size() < ntile) c->res.resize(ntile);
  const int64_t l0 = c->launches;
  // a search between the steps of an open resident simulation joins the simulation's accumulated timing (its
  // distance events are summed by iq_sim_sync) instead of resetting it
  const bool in_sim = c->sim != nullptr;
  if (!in_sim) {
    c->dist_ev_used = 0;
    c->last_fft_bytes = 0.0;
    c->last_fft_searches = c->last_direct_searches = 0;
  }
  CK(cudaEventRecord(c->ev0, c->stream));
  for (int base = 0; base < ntile; base += c->max_batch) {
    const int R = std::min(c->max_batch, ntile - base);
    rc = search_chunk(c, e, tiles + base, R, tol, u ? u + base : nullptr, results + base, base);
    if (rc) return rc;
  }
  CK(cudaEventRecord(c->ev1, c->stream));
  CK(cudaEventSynchronize(c->ev1));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  c->last_ms = ms;
  c->last_launches = c->launches - l0;
  if (in_sim) return IQ_OK;
  return collect_dist_times(c);
}

}  // namespace iqimpl

using namespace iqimpl;

// ================================================================================================
// exported ABI
// ================================================================================================
extern "C" {

int32_t iq_abi_version(void) { return IQ_ABI_VERSION; }

int32_t iq_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

const char* iq_last_error(void) { return g_err.c_str(); }
// Host driver (iq_host.cpp): an entry point that failed on a pool thread left its message in THAT thread's string;
// the driver copies it and re-posts it on the thread that returns the error to the caller.
void iq_post_error(const char* msg) { g_err = msg ? msg : ""; }

int32_t iq_ctx_destroy(iq_ctx* c) {
  if (!c) return IQ_OK;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  cudaFree(c->d_ti);
  for (auto p : c->d_aux) cudaFree(p);
  cudaFree(c->d_sat_ti);
  for (auto p : c->d_sat_aux) cudaFree(p);
  cudaFree(c->d_disabled);
  cudaFree(c->d_Dovl);
  cudaFree(c->d_Dhard);
  for (auto p : c->d_Dsoft) cudaFree(p);
  cudaFree(c->d_minmax);
  if (c->h_minmax) cudaFreeHost(c->h_minmax);
  if (c->h_stage) cudaFreeHost(c->h_stage);
  cudaFree(c->d_stage);
  cudaFree(c->d_sel);
  cudaFree(c->d_sel_list);
  cudaFree(c->d_selbuf);
  if (c->h_sel) cudaFreeHost(c->h_sel);
  cudaFree(c->d_pick);
  if (c->h_pick) cudaFreeHost(c->h_pick);
  cudaFree(c->d_blockcount);
  cudaFree(c->d_total);
  cudaFree(c->d_chunkmin);
  if (c->h_total) cudaFreeHost(c->h_total);
  cudaFree(c->d_cand_idx);
  cudaFree(c->d_cand_val);
  if (c->h_cand_idx) cudaFreeHost(c->h_cand_idx);
  if (c->h_cand_val) cudaFreeHost(c->h_cand_val);
  cudaFree(c->d_slice_hist);
  if (c->h_slice_hist) cudaFreeHost(c->h_slice_hist);
  cudaFree(c->d_shifts);
  cudaFree(c->d_fetch);
  if (c->h_cut) cudaFreeHost(c->h_cut);
  cudaFree(c->d_cut);
  cudaFree(c->d_rank);
  cudaFree(c->d_colsum);
  cudaFree(c->d_prob);
  if (c->h_prob) cudaFreeHost(c->h_prob);
  sim_destroy(c);
  for (auto& e : c->masks) {
    cudaFree(e->d_boxes);
    cudaFree(e->d_mask);
    for (auto& kv : e->a2) cudaFree(kv.second);
  }
  for (auto& kv : c->img_pad) cudaFree(kv.second);
  iqfft::plan_destroy(c->fft);
  for (auto ev : c->dist_ev) cudaEventDestroy(ev);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return IQ_OK;
}

static int32_t ctx_create_impl(iq_ctx* c, const iq_ctx_desc* d) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(IQ_ERR_NO_DEVICE, "no CUDA device visible: libiqb200 has no CPU fallback");
  }
  if (d->device < 0 || d->device >= ndev) return fail(IQ_ERR_NO_DEVICE, "device %d out of range (have %d)", d->device, ndev);
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, d->device));
  if (prop.major != 10) return fail(IQ_ERR_NO_DEVICE, "device %d is sm_%d%d; this library carries sm_100a code only", d->device, prop.major, prop.minor);
  CK(cudaSetDevice(d->device));
  c->device = d->device;
  CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CK(cudaEventCreate(&c->ev0));
  CK(cudaEventCreate(&c->ev1));

  const size_t nimg = (size_t)c->nx * c->ny * c->nz;
  const size_t nsat = (size_t)(c->nx + 1) * (c->ny + 1) * (c->nz + 1);
  const size_t slack = 64;  // floats of zeroed slack after each image
  auto upload = [&](const float* src, float** dimg, double** dsat) -> int {
    CK(iq::dmalloc((void**)dimg, (nimg + slack) * sizeof(float)));
    CK(cudaMemsetAsync(*dimg, 0, (nimg + slack) * sizeof(float), c->stream));
    CK(cudaMemcpyAsync(*dimg, src, nimg * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CK(iq::dmalloc((void**)dsat, nsat * sizeof(double)));
    CK(cudaMemsetAsync(*dsat, 0, nsat * sizeof(double), c->stream));
    CK(iq::launch_sat_build(*dimg, *dsat, c->nx, c->ny, c->nz, c->stream));
    c->launches += 3;
    CK(cudaStreamSynchronize(c->stream));
    return IQ_OK;
  };
  int rc = upload(d->ti, &c->d_ti, &c->d_sat_ti);
  if (rc) return rc;
  c->image_is_int[-1] = all_integer(d->ti, (long long)nimg);
  c->d_aux.assign(c->nsoft, nullptr);
  c->d_sat_aux.assign(c->nsoft, nullptr);
  for (int s = 0; s < c->nsoft; ++s) {
    if (!d->auxti || !d->auxti[s]) return fail(IQ_ERR_INVALID, "auxti[%d] is NULL", s);
    rc = upload(d->auxti[s], &c->d_aux[s], &c->d_sat_aux[s]);
    if (rc) return rc;
    c->image_is_int[s] = all_integer(d->auxti[s], (long long)nimg);
  }
  c->nenabled = c->npos;
  if (d->disabled) {
    c->h_disabled.assign(d->disabled, d->disabled + c->npos);
    long long nd = 0;
    for (long long p = 0; p < c->npos; ++p) nd += c->h_disabled[p] ? 1 : 0;
    c->nenabled = c->npos - nd;
    if (nd > 0) {
      CK(iq::dmalloc((void**)&c->d_disabled, (size_t)c->npos));
      CK(cudaMemcpyAsync(c->d_disabled, c->h_disabled.data(), (size_t)c->npos, cudaMemcpyHostToDevice, c->stream));
    } else {
      c->h_disabled.clear();
    }
  }
  const size_t B = (size_t)c->max_batch;
  c->max_src = 2 + c->nsoft;
  CK(iq::dmalloc((void**)&c->d_Dovl, B * c->npos * sizeof(float)));
  CK(iq::dmalloc((void**)&c->d_Dhard, B * c->npos * sizeof(float)));
  c->d_Dsoft.assign(c->nsoft, nullptr);
  for (int s = 0; s < c->nsoft; ++s) CK(iq::dmalloc((void**)&c->d_Dsoft[s], B * c->npos * sizeof(float)));
  const size_t nmm = (size_t)(2 + c->nsoft) * 2 * B;
  CK(iq::dmalloc((void**)&c->d_minmax, nmm * sizeof(unsigned)));
  CK(cudaMallocHost((void**)&c->h_minmax, nmm * sizeof(unsigned)));
  c->sel_cap = (unsigned)std::max<long long>(c->npos / 8, 4096);
  CK(iq::dmalloc((void**)&c->d_selbuf, B * c->max_src * (size_t)c->sel_cap * sizeof(unsigned long long)));
  CK(iq::dmalloc((void**)&c->d_sel, B * c->max_src * sizeof(iq::SelJob)));
  CK(iq::dmalloc((void**)&c->d_sel_list, (2 * B * c->max_src + 3) * sizeof(int)));
  CK(cudaMemsetAsync(c->d_sel_list, 0, (2 * B * c->max_src + 3) * sizeof(int), c->stream));
  CK(cudaMallocHost((void**)&c->h_sel, B * c->max_src * sizeof(iq::SelJob)));
  CK(cudaMemsetAsync(c->d_sel, 0, B * c->max_src * sizeof(iq::SelJob), c->stream));
  CK(iq::dmalloc((void**)&c->d_pick, B * sizeof(iq::PickJob)));
  CK(cudaMallocHost((void**)&c->h_pick, B * sizeof(iq::PickJob)));
  CK(iq::dmalloc((void**)&c->d_blockcount, B * iq::pick_nblk(c->npos) * sizeof(unsigned)));
  c->chunk_stride = c->npos / std::min<long long>(4096, std::max(4 * c->nxo, 1)) + 2;
  CK(iq::dmalloc((void**)&c->d_chunkmin, B * c->chunk_stride * sizeof(unsigned)));
  CK(iq::dmalloc((void**)&c->d_total, B * sizeof(unsigned)));
  CK(cudaMallocHost((void**)&c->h_total, B * sizeof(unsigned)));
  CK(iq::dmalloc((void**)&c->d_cand_idx, B * c->npos * sizeof(unsigned)));
  CK(iq::dmalloc((void**)&c->d_cand_val, B * c->max_src * c->npos * sizeof(float)));
  CK(iq::dmalloc((void**)&c->d_fetch, (size_t)c->tilevol * sizeof(float)));
  CK(iq::dmalloc((void**)&c->d_rank, B * c->max_src * iq::kTauMax * sizeof(unsigned)));
  CK(iq::dmalloc((void**)&c->d_colsum, B * c->max_src * sizeof(unsigned long long)));
  CK(iq::dmalloc((void**)&c->d_prob, B * iq::kTauMax * sizeof(double)));
  // h_prob (page-locked mirror of d_prob: 256 KB per job slot, 64 MB for the 256 slots of a resident context) is
  // allocated by the first iq_search that needs it -- page-locking it cost most of a context's set-up time, and the
  // resident pipeline never reads probabilities back
  // radix-select schedule: 4 value bytes, then only the index bytes that can be non-zero
  std::vector<int> shifts = {56, 48, 40, 32};
  int nib = 1;
  while (nib < 4 && ((unsigned long long)(c->npos - 1) >> (8 * nib)) != 0) ++nib;
  for (int b = nib - 1; b >= 0; --b) shifts.push_back(8 * b);
  c->nshift = (int)shifts.size();
  CK(iq::dmalloc((void**)&c->d_shifts, shifts.size() * sizeof(int)));
  CK(cudaMemcpyAsync(c->d_shifts, shifts.data(), shifts.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  // the all-ones mask of the soft-data distance (fastdistance default weights, src/utils.jl:5)
  if (c->nsoft > 0) {
    std::vector<uint8_t> ones((size_t)c->tilevol, 1);
    rc = get_mask(c, ones.data(), &c->full_mask);
    if (rc) return rc;
  }
  return IQ_OK;
}

int32_t iq_ctx_create(iq_ctx** out, const iq_ctx_desc* d) {
  if (!out || !d) return fail(IQ_ERR_INVALID, "iq_ctx_create: NULL argument");
  *out = nullptr;
  if (d->ndim != 2 && d->ndim != 3) return fail(IQ_ERR_INVALID, "ndim must be 2 or 3");
  if (!d->ti) return fail(IQ_ERR_INVALID, "training image is NULL");
  if (d->nsoft < 0 || d->nsoft > 6) return fail(IQ_ERR_INVALID, "nsoft must be in [0,6]");
  for (int i = 0; i < 3; ++i) {
    const int64_t n = i < d->ndim ? d->ti_size[i] : 1, t = i < d->ndim ? d->tile_size[i] : 1;
    if (!(t > 0 && t <= n)) return fail(IQ_ERR_INVALID, "invalid tile size");
    if (n > (1 << 20)) return fail(IQ_ERR_INVALID, "training image dimension too large");
  }
  iq_ctx* c = new (std::nothrow) iq_ctx();
  if (!c) return fail(IQ_ERR_NOMEM, "out of host memory");
  c->ndim = d->ndim;
  c->nx = (int)d->ti_size[0]; c->ny = (int)d->ti_size[1]; c->nz = d->ndim == 3 ? (int)d->ti_size[2] : 1;
  c->tx = (int)d->tile_size[0]; c->ty = (int)d->tile_size[1]; c->tz = d->ndim == 3 ? (int)d->tile_size[2] : 1;
  c->nxo = c->nx - c->tx + 1; c->nyo = c->ny - c->ty + 1; c->nzo = c->nz - c->tz + 1;
  c->npos = (long long)c->nxo * c->nyo * c->nzo;
  c->tilevol = (long long)c->tx * c->ty * c->tz;
  if (c->npos >= (1ll << 32)) { delete c; return fail(IQ_ERR_INVALID, "more than 2^32 patch positions"); }
  c->nsoft = d->nsoft;
  c->max_batch = std::max(1, d->max_batch);
  if (const char* ev = std::getenv("IQB200_TAU_DEVICE")) c->tau_device = std::atoi(ev) ? 1 : 0;  // experiments only
  const int rc = ctx_create_impl(c, d);
  if (rc != IQ_OK) {
    const std::string keep = g_err;
    iq_ctx_destroy(c);
    g_err = keep;
    return rc;
  }
  *out = c;
  return IQ_OK;
}

// Bitwise comparison of two device arrays: *flag |= 1 when they differ.
__global__ void __launch_bounds__(256) k_differs(const unsigned* __restrict__ a, const unsigned* __restrict__ b, long long n,
                                                 int* __restrict__ flag) {
  const long long i0 = (long long)blockIdx.x * 256 + threadIdx.x;
  bool diff = false;
  for (long long i = i0; i < n; i += (long long)gridDim.x * 256) diff |= a[i] != b[i];
  if (__syncthreads_or(diff) && threadIdx.x == 0) atomicOr(flag, 1);
}

static int32_t ctx_matches_impl(iq_ctx* c, const iq_ctx_desc* d, const iq_ctx* ref, int32_t* same);

int32_t iq_ctx_matches(iq_ctx* c, const iq_ctx_desc* d, int32_t* same) { return ctx_matches_impl(c, d, nullptr, same); }

int32_t iq_ctx_matches_ctx(iq_ctx* c, const iq_ctx_desc* d, const iq_ctx* ref, int32_t* same) {
  if (!ref) return fail(IQ_ERR_INVALID, "iq_ctx_matches_ctx: NULL reference context");
  return ctx_matches_impl(c, d, ref, same);
}

static int32_t ctx_matches_impl(iq_ctx* c, const iq_ctx_desc* d, const iq_ctx* ref, int32_t* same) {
  if (!c || !d || !same || !d->ti) return fail(IQ_ERR_INVALID, "iq_ctx_matches: NULL argument");
  *same = 0;
  if (ref && (ref == c || ref->device != c->device || ref->nsoft != c->nsoft || ref->nx != c->nx || ref->ny != c->ny || ref->nz != c->nz))
    return IQ_OK;
  if (c->sim) return IQ_OK;  // a simulation is open on it
  const int nz = d->ndim == 3 ? (int)d->ti_size[2] : 1, tz = d->ndim == 3 ? (int)d->tile_size[2] : 1;
  if (d->ndim != c->ndim || d->device != c->device || d->nsoft != c->nsoft || std::max(1, d->max_batch) != c->max_batch ||
      d->ti_size[0] != c->nx || d->ti_size[1] != c->ny || nz != c->nz || d->tile_size[0] != c->tx || d->tile_size[1] != c->ty ||
      tz != c->tz)
    return IQ_OK;
  // disabled patches: host-side comparison with the copy the context keeps (empty = none disabled)
  bool any = false;
  if (d->disabled)
    for (long long p = 0; p < c->npos && !any; ++p) any = d->disabled[p] != 0;
  if (any != !c->h_disabled.empty()) return IQ_OK;
  if (any && std::memcmp(d->disabled, c->h_disabled.data(), (size_t)c->npos) != 0) return IQ_OK;
  // images: uploaded again (the caller's arrays may have changed since) and compared with the resident copies
  CK(cudaSetDevice(c->device));
  const size_t nimg = (size_t)c->nx * c->ny * c->nz;
  float* d_new = nullptr;
  int* d_flag = nullptr;
  if (!ref) CK(iq::dmalloc((void**)&d_new, nimg * sizeof(float)));
  CK(iq::dmalloc((void**)&d_flag, sizeof(int)));
  CK(cudaMemsetAsync(d_flag, 0, sizeof(int), c->stream));
  for (int img = -1; img < c->nsoft; ++img) {
    const float* cmp = d_new;
    if (ref) {
      // the reference context was created from / matched against these very host arrays during this call and is idle:
      // compare with its resident copy, no second upload
      cmp = img < 0 ? ref->d_ti : ref->d_aux[img];
    } else {
      const float* src = img < 0 ? d->ti : (d->auxti ? d->auxti[img] : nullptr);
      if (!src) { cudaFree(d_new); cudaFree(d_flag); return fail(IQ_ERR_INVALID, "iq_ctx_matches: auxti[%d] is NULL", img); }
      CK(cudaMemcpyAsync(d_new, src, nimg * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    }
    k_differs<<<1184, 256, 0, c->stream>>>((const unsigned*)cmp, (const unsigned*)(img < 0 ? c->d_ti : c->d_aux[img]),
                                           (long long)nimg, d_flag);
    CK(cudaGetLastError());
  }
  int flag = 1;
  CK(cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  cudaFree(d_new);
  cudaFree(d_flag);
  *same = flag == 0 ? 1 : 0;
  return IQ_OK;
}

int32_t iq_host_alloc(size_t bytes, void** out) {
  if (!out) return fail(IQ_ERR_INVALID, "iq_host_alloc: NULL argument");
  *out = nullptr;
  if (cudaHostAlloc(out, std::max<size_t>(bytes, 1), cudaHostAllocPortable) != cudaSuccess) {
    cudaGetLastError();
    *out = nullptr;
    return fail(IQ_ERR_NOMEM, "iq_host_alloc: cannot allocate %zu bytes of page-locked host memory", bytes);
  }
  return IQ_OK;
}

int32_t iq_host_free(void* p) {
  if (p && cudaFreeHost(p) != cudaSuccess) {
    cudaGetLastError();
    return fail(IQ_ERR_INVALID, "iq_host_free: not a pointer returned by iq_host_alloc");
  }
  return IQ_OK;
}

int32_t iq_ctx_npos(const iq_ctx* c, int64_t* npos, int64_t* nenabled) {
  if (!c) return fail(IQ_ERR_INVALID, "NULL context");
  if (npos) *npos = c->npos;
  if (nenabled) *nenabled = c->nenabled;
  return IQ_OK;
}

int32_t iq_search(iq_ctx* c, const uint8_t* ovlmask, const iq_tile* tiles, int32_t ntile, double tol, iq_result* results) {
  return do_search(c, ovlmask, tiles, ntile, tol, nullptr, results);
}

int32_t iq_search_pick(iq_ctx* c, const uint8_t* ovlmask, const iq_tile* tiles, int32_t ntile, double tol, const double* u,
                       iq_result* results) {
  if (!u) return fail(IQ_ERR_INVALID, "iq_search_pick: u is NULL");
  return do_search(c, ovlmask, tiles, ntile, tol, u, results);
}

int32_t iq_distance(iq_ctx* c, int32_t which, const uint8_t* ovlmask, const iq_tile* tile, float* out_map) {
  if (!c || !tile || !out_map) return fail(IQ_ERR_INVALID, "iq_distance: NULL argument");
  CK(cudaSetDevice(c->device));
  c->stage_used = 0;
  int rc;
  const float* src = nullptr;
  if (which == -1) {
    if (!ovlmask || !tile->simdev) return fail(IQ_ERR_INVALID, "iq_distance: overlap distance needs ovlmask and simdev");
    MaskEntry* e = nullptr;
    rc = get_mask(c, ovlmask, &e);
    if (rc) return rc;
    rc = stage_reserve(c, 8192 + (size_t)4 * e->tmpl_floats * sizeof(float) + (size_t)c->tilevol * sizeof(float));
    if (rc) return rc;
    const float* k = tile->simdev;
    rc = run_dense(c, e, -1, &k, 1, c->d_Dovl, 0);
    if (rc) return rc;
    src = c->d_Dovl;
  } else if (which == -2) {
    if (tile->hard_nnz <= 0) return fail(IQ_ERR_INVALID, "iq_distance: tile has no hard data");
    rc = stage_reserve(c, 8192 + (size_t)tile->hard_nnz * 16);
    if (rc) return rc;
    const size_t off_ptr = stage_alloc(c, 2 * sizeof(int));
    const size_t off_off = stage_alloc(c, (size_t)tile->hard_nnz * sizeof(long long));
    const size_t off_val = stage_alloc(c, (size_t)tile->hard_nnz * sizeof(float));
    int* ptr = (int*)(c->h_stage + off_ptr);
    long long* off = (long long*)(c->h_stage + off_off);
    float* val = (float*)(c->h_stage + off_val);
    ptr[0] = 0; ptr[1] = tile->hard_nnz;
    for (int i = 0; i < tile->hard_nnz; ++i) {
      const int o = tile->hard_offset[i];
      if (o < 0 || o >= c->tilevol) return fail(IQ_ERR_INVALID, "hard_offset out of the tile");
      const int qx = o % c->tx, qy = (o / c->tx) % c->ty, qz = o / (c->tx * c->ty);
      off[i] = ((long long)qz * c->ny + qy) * c->nx + qx;
      val[i] = tile->hard_value[i];
    }
    CK(cudaMemcpyAsync(c->d_stage + off_ptr, c->h_stage + off_ptr, c->stage_used - off_ptr, cudaMemcpyHostToDevice, c->stream));
    iq::SparseParams sp{};
    sp.img = c->d_ti;
    sp.nx = c->nx; sp.ny = c->ny; sp.nz = c->nz; sp.nxo = c->nxo; sp.nyo = c->nyo; sp.nzo = c->nzo;
    sp.npos = c->npos;
    sp.ptr = (const int*)(c->d_stage + off_ptr);
    sp.ptr_stride = 1;
    sp.off = (const long long*)(c->d_stage + off_off);
    sp.val = (const float*)(c->d_stage + off_val);
    sp.disabled = c->d_disabled;
    sp.out = c->d_Dhard;
    sp.minbits = nullptr; sp.maxbits = nullptr;
    sp.R = 1;
    CK(iq::launch_dist_sparse(sp, c->stream));
    c->launches++;
    src = c->d_Dhard;
  } else if (which >= 0 && which < c->nsoft) {
    if (!tile->softdev || !tile->softdev[which]) return fail(IQ_ERR_INVALID, "iq_distance: softdev missing");
    rc = stage_reserve(c, 8192 + (size_t)4 * c->full_mask->tmpl_floats * sizeof(float) + (size_t)c->tilevol * sizeof(float));
    if (rc) return rc;
    const float* k = tile->softdev[which];
    rc = run_dense(c, c->full_mask, which, &k, 1, c->d_Dsoft[which], 2 + which);
    if (rc) return rc;
    src = c->d_Dsoft[which];
  } else {
    return fail(IQ_ERR_INVALID, "iq_distance: bad `which`");
  }
  CK(cudaMemcpyAsync(out_map, src, (size_t)c->npos * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return IQ_OK;
}

int32_t iq_fetch_tile(iq_ctx* c, int64_t pos, float* out_tile) {
  if (!c || !out_tile) return fail(IQ_ERR_INVALID, "iq_fetch_tile: NULL argument");
  if (pos < 0 || pos >= c->npos) return fail(IQ_ERR_INVALID, "iq_fetch_tile: position out of range");
  CK(cudaSetDevice(c->device));
  const long long x0 = pos % c->nxo, y0 = (pos / c->nxo) % c->nyo, z0 = pos / ((long long)c->nxo * c->nyo);
  CK(iq::launch_fetch_tile(c->d_ti, c->nx, c->ny, c->nz, c->tx, c->ty, c->tz, x0, y0, z0, c->d_fetch, c->stream));
  c->launches++;
  CK(cudaMemcpyAsync(out_tile, c->d_fetch, (size_t)c->tilevol * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return IQ_OK;
}

int32_t iq_slice_distance(iq_ctx* c, const uint8_t* ovlmask, const iq_tile* tiles, int32_t ntile, float* dmin_local) {
  if (!c || !ovlmask || !tiles || !dmin_local || ntile <= 0) return fail(IQ_ERR_INVALID, "iq_slice_distance: bad argument");
  if (ntile > c->max_batch) return fail(IQ_ERR_INVALID, "iq_slice_distance: ntile exceeds max_batch");
  const int S = c->nsoft;
  for (int i = 0; i < ntile; ++i) {
    if (!tiles[i].simdev || tiles[i].hard_nnz > 0) return fail(IQ_ERR_INVALID, "position-slice mode: simdev required, hard data unsupported");
    for (int s = 0; s < S; ++s)
      if (!tiles[i].softdev || !tiles[i].softdev[s]) return fail(IQ_ERR_INVALID, "tile %d lacks softdev[%d]", i, s);
  }
  CK(cudaSetDevice(c->device));
  MaskEntry* e = nullptr;
  int rc = get_mask(c, ovlmask, &e);
  if (rc) return rc;
  const int rb = pick_rb(c, ntile);
  const int ngrp = (ntile + rb - 1) / rb;
  size_t need = 8192 + (size_t)ngrp * rb * e->tmpl_floats * sizeof(float) + (size_t)ntile * (c->tilevol * sizeof(float) + 64);
  need += (size_t)(1 + S) * ((size_t)ntile * c->tilevol * sizeof(float) + 1024);
  if (S > 0) need += (size_t)S * ((size_t)ngrp * rb * c->full_mask->tmpl_floats * sizeof(float) + ntile * sizeof(double) + 512);
  rc = stage_reserve(c, need);
  if (rc) return rc;
  c->stage_used = 0;
  c->dist_ev_used = 0;
  const int nkind = 2 + S;
  for (int k = 0; k < nkind; ++k) {
    CK(iq::launch_fill_u32(c->d_minmax + (size_t)(k * 2 + 0) * c->max_batch, 0x7f800000u, c->max_batch, c->stream));
    CK(iq::launch_fill_u32(c->d_minmax + (size_t)(k * 2 + 1) * c->max_batch, 0u, c->max_batch, c->stream));
    c->launches += 2;
  }
  std::vector<const float*> kern(ntile);
  for (int r = 0; r < ntile; ++r) kern[r] = tiles[r].simdev;
  rc = run_dense(c, e, -1, kern.data(), ntile, c->d_Dovl, 0);
  if (rc) return rc;
  for (int s = 0; s < S; ++s) {  // soft-data distances of the slab (src/iqsim.jl:222-227)
    for (int r = 0; r < ntile; ++r) kern[r] = tiles[r].softdev[s];
    rc = run_dense(c, c->full_mask, s, kern.data(), ntile, c->d_Dsoft[s], 2 + s);
    if (rc) return rc;
  }
  CK(cudaMemcpyAsync(c->h_minmax, c->d_minmax, (size_t)nkind * 2 * c->max_batch * sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  for (int r = 0; r < ntile; ++r) std::memcpy(&dmin_local[r], &c->h_minmax[r], 4);
  c->slice_ntile = ntile;
  return IQ_OK;
}

int32_t iq_slice_select(iq_ctx* c, double tol, const float* dmin_global, int64_t* counts) {
  if (!c || !dmin_global || !counts) return fail(IQ_ERR_INVALID, "iq_slice_select: NULL argument");
  if (c->slice_ntile <= 0) return fail(IQ_ERR_STATE, "iq_slice_select without a preceding iq_slice_distance");
  if (!(tol > 0.0 && tol <= 1.0)) return fail(IQ_ERR_INVALID, "tolerance must be in range (0,1]");
  CK(cudaSetDevice(c->device));
  const int R = c->slice_ntile;
  for (int r = 0; r < R; ++r) std::memcpy(&c->h_minmax[r], &dmin_global[r], 4);
  CK(cudaMemcpyAsync(c->d_minmax, c->h_minmax, (size_t)R * sizeof(unsigned), cudaMemcpyHostToDevice, c->stream));
  const int maxS = c->max_src;
  for (int r = 0; r < R; ++r) {
    iq::PickJob& J = c->h_pick[r];
    std::memset(&J, 0, sizeof J);
    J.mode = 0;
    J.nsrc = 1;
    J.src[0] = c->d_Dovl + (size_t)r * c->npos;
    J.sel = c->d_sel + (size_t)r * maxS;
    J.tol = tol;
    J.minbits = c->d_minmax + r;
    J.blockcount = c->d_blockcount + (size_t)r * iq::pick_nblk(c->npos);
    J.total = c->d_total + r;
    J.cand_idx = c->d_cand_idx + (size_t)r * c->npos;
    J.cand_val = c->d_cand_val + (size_t)r * maxS * c->npos;
    J.cap = c->npos;
  }
  CK(cudaMemcpyAsync(c->d_pick, c->h_pick, (size_t)R * sizeof(iq::PickJob), cudaMemcpyHostToDevice, c->stream));
  CK(iq::launch_pick_count(c->d_pick, R, c->npos, c->stream));
  CK(iq::launch_pick_write(c->d_pick, R, c->npos, c->stream));
  c->launches += 2;
  CK(cudaMemcpyAsync(c->h_total, c->d_total, (size_t)R * sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  size_t tot = 0;
  for (int r = 0; r < R; ++r) tot += c->h_total[r];
  if (tot > c->h_cand_cap) {
    if (c->h_cand_idx) cudaFreeHost(c->h_cand_idx);
    if (c->h_cand_val) cudaFreeHost(c->h_cand_val);
    c->h_cand_idx = nullptr; c->h_cand_val = nullptr;
    const size_t cap = std::max<size_t>(tot * 2, 1 << 16);
    CK(cudaMallocHost((void**)&c->h_cand_idx, cap * sizeof(unsigned)));
    CK(cudaMallocHost((void**)&c->h_cand_val, cap * maxS * sizeof(float)));
    c->h_cand_cap = cap;
  }
  size_t o = 0;
  for (int r = 0; r < R; ++r) {
    const size_t n = c->h_total[r];
    if (n) {
      CK(cudaMemcpyAsync(c->h_cand_idx + o, c->d_cand_idx + (size_t)r * c->npos, n * sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
      CK(cudaMemcpyAsync(c->h_cand_val + o, c->d_cand_val + (size_t)r * maxS * c->npos, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    }
    o += n;
  }
  CK(cudaStreamSynchronize(c->stream));
  c->slice_idx.resize(R);
  c->slice_val.resize(R);
  o = 0;
  for (int r = 0; r < R; ++r) {
    const size_t n = c->h_total[r];
    c->slice_idx[r].resize(n);
    c->slice_val[r].assign(c->h_cand_val + o, c->h_cand_val + o + n);
    for (size_t i = 0; i < n; ++i) c->slice_idx[r][i] = (int64_t)c->h_cand_idx[o + i];
    counts[r] = (int64_t)n;
    o += n;
  }
  return IQ_OK;
}

// ---- relaxation path of position-slice mode ----
static const float* slice_src(const iq_ctx* c, int tile, int src) {
  return (src == 0 ? c->d_Dovl : c->d_Dsoft[src - 1]) + (size_t)tile * c->npos;
}

int32_t iq_slice_minmax(iq_ctx* c, int32_t tile, uint32_t* minbits, uint32_t* maxbits) {
  if (!c || !minbits || !maxbits) return fail(IQ_ERR_INVALID, "iq_slice_minmax: NULL argument");
  if (tile < 0 || tile >= c->slice_ntile) return fail(IQ_ERR_STATE, "iq_slice_minmax without a preceding iq_slice_distance");
  for (int s = 0; s <= c->nsoft; ++s) {
    const int kind = s == 0 ? 0 : 1 + s;
    minbits[s] = c->h_minmax[(size_t)(kind * 2 + 0) * c->max_batch + tile];
    maxbits[s] = c->h_minmax[(size_t)(kind * 2 + 1) * c->max_batch + tile];
  }
  return IQ_OK;
}

int32_t iq_slice_hist(iq_ctx* c, int32_t tile, int32_t nreq, const int32_t* src, const int32_t* level, const uint32_t* prefix,
                      int64_t* hist) {
  if (!c || !src || !level || !prefix || !hist) return fail(IQ_ERR_INVALID, "iq_slice_hist: NULL argument");
  if (tile < 0 || tile >= c->slice_ntile) return fail(IQ_ERR_STATE, "iq_slice_hist without a preceding iq_slice_distance");
  if (nreq <= 0 || nreq > 8) return fail(IQ_ERR_INVALID, "iq_slice_hist: 1..8 requests per call");
  CK(cudaSetDevice(c->device));
  if (!c->d_slice_hist) {
    CK(iq::dmalloc((void**)&c->d_slice_hist, 8 * 256 * sizeof(unsigned long long)));
    CK(cudaMallocHost((void**)&c->h_slice_hist, 8 * 256 * sizeof(unsigned long long)));
  }
  iq::SliceHistParams P{};
  for (int i = 0; i < nreq; ++i) {
    if (src[i] < 0 || src[i] > c->nsoft || level[i] < 0 || level[i] > 3) return fail(IQ_ERR_INVALID, "iq_slice_hist: bad source or level");
    P.req[i].map = slice_src(c, tile, src[i]);
    P.req[i].level = level[i];
    P.req[i].prefix = prefix[i];
  }
  P.nreq = nreq;
  P.npos = c->npos;
  P.out = c->d_slice_hist;
  CK(cudaMemsetAsync(c->d_slice_hist, 0, (size_t)nreq * 256 * sizeof(unsigned long long), c->stream));
  CK(iq::launch_slice_hist(P, c->stream));
  c->launches++;
  CK(cudaMemcpyAsync(c->h_slice_hist, c->d_slice_hist, (size_t)nreq * 256 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                     c->stream));
  CK(cudaStreamSynchronize(c->stream));
  for (int i = 0; i < nreq * 256; ++i) hist[i] = (int64_t)c->h_slice_hist[i];
  return IQ_OK;
}

int32_t iq_slice_kth(iq_ctx* c, int32_t tile, int32_t src, int64_t k_local, uint64_t* key) {
  if (!c || !key) return fail(IQ_ERR_INVALID, "iq_slice_kth: NULL argument");
  if (tile < 0 || tile >= c->slice_ntile) return fail(IQ_ERR_STATE, "iq_slice_kth without a preceding iq_slice_distance");
  if (src < 0 || src > c->nsoft || k_local < 1 || k_local > c->npos) return fail(IQ_ERR_INVALID, "iq_slice_kth: bad source or rank");
  CK(cudaSetDevice(c->device));
  iq::SelJob& sj = c->h_sel[0];
  std::memset(&sj, 0, sizeof sj);
  sj.map = slice_src(c, tile, src);
  sj.k = (unsigned long long)k_local;
  sj.active = 1;
  sj.cbuf = c->d_selbuf;
  sj.ccap = c->sel_cap;
  CK(cudaMemcpyAsync(c->d_sel, c->h_sel, sizeof(iq::SelJob), cudaMemcpyHostToDevice, c->stream));
  int nl = 0;
  CK(iq::launch_select_all(c->d_sel, 1, c->npos, c->d_shifts, c->nshift, c->stream, &nl, c->d_sel_list));
  c->launches += nl;
  CK(cudaMemcpyAsync(c->h_sel, c->d_sel, sizeof(iq::SelJob), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  *key = (uint64_t)c->h_sel[0].kth;
  return IQ_OK;
}

int32_t iq_slice_pick(iq_ctx* c, int32_t tile, int32_t nsrc, const uint64_t* kth, int64_t* count) {
  if (!c || !kth || !count) return fail(IQ_ERR_INVALID, "iq_slice_pick: NULL argument");
  if (tile < 0 || tile >= c->slice_ntile) return fail(IQ_ERR_STATE, "iq_slice_pick without a preceding iq_slice_distance");
  if (nsrc < 1 || nsrc > 1 + c->nsoft) return fail(IQ_ERR_INVALID, "iq_slice_pick: bad number of sources");
  CK(cudaSetDevice(c->device));
  const int maxS = c->max_src;
  for (int s = 0; s < nsrc; ++s) {
    iq::SelJob& sj = c->h_sel[(size_t)tile * maxS + s];
    std::memset(&sj, 0, sizeof sj);
    sj.kth = (unsigned long long)kth[s];
  }
  CK(cudaMemcpyAsync(c->d_sel + (size_t)tile * maxS, c->h_sel + (size_t)tile * maxS, (size_t)nsrc * sizeof(iq::SelJob),
                     cudaMemcpyHostToDevice, c->stream));
  iq::PickJob& J = c->h_pick[tile];
  std::memset(&J, 0, sizeof J);
  J.mode = 1;
  J.nsrc = nsrc;
  for (int s = 0; s < nsrc; ++s) J.src[s] = slice_src(c, tile, s);
  J.sel = c->d_sel + (size_t)tile * maxS;
  J.blockcount = c->d_blockcount + (size_t)tile * iq::pick_nblk(c->npos);
  J.total = c->d_total + tile;
  J.cand_idx = c->d_cand_idx + (size_t)tile * c->npos;
  J.cand_val = c->d_cand_val + (size_t)tile * maxS * c->npos;
  J.cap = c->npos;
  CK(cudaMemcpyAsync(c->d_pick + tile, &J, sizeof(iq::PickJob), cudaMemcpyHostToDevice, c->stream));
  CK(iq::launch_pick_count(c->d_pick + tile, 1, c->npos, c->stream));
  CK(iq::launch_pick_write(c->d_pick + tile, 1, c->npos, c->stream));
  c->launches += 2;
  CK(cudaMemcpyAsync(c->h_total + tile, c->d_total + tile, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  const size_t n = c->h_total[tile];
  if (n > c->h_cand_cap) {
    if (c->h_cand_idx) cudaFreeHost(c->h_cand_idx);
    if (c->h_cand_val) cudaFreeHost(c->h_cand_val);
    c->h_cand_idx = nullptr; c->h_cand_val = nullptr;
    const size_t cap = std::max<size_t>(n * 2, 1 << 16);
    CK(cudaMallocHost((void**)&c->h_cand_idx, cap * sizeof(unsigned)));
    CK(cudaMallocHost((void**)&c->h_cand_val, cap * maxS * sizeof(float)));
    c->h_cand_cap = cap;
  }
  if (n) {
    CK(cudaMemcpyAsync(c->h_cand_idx, J.cand_idx, n * sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
    for (int s = 0; s < nsrc; ++s)
      CK(cudaMemcpyAsync(c->h_cand_val + (size_t)s * n, J.cand_val + (size_t)s * c->npos, n * sizeof(float), cudaMemcpyDeviceToHost,
                         c->stream));
    CK(cudaStreamSynchronize(c->stream));
  }
  if ((int)c->slice_idx.size() <= tile) { c->slice_idx.resize(tile + 1); c->slice_val.resize(tile + 1); }
  c->slice_idx[tile].resize(n);
  for (size_t i = 0; i < n; ++i) c->slice_idx[tile][i] = (int64_t)c->h_cand_idx[i];
  c->slice_val[tile].assign(c->h_cand_val, c->h_cand_val + n * (size_t)nsrc);  // [source][candidate]
  *count = (int64_t)n;
  return IQ_OK;
}

int32_t iq_slice_candidates(const iq_ctx* c, int32_t tile, const int64_t** idx, const float** val) {
  if (!c || tile < 0 || tile >= (int)c->slice_idx.size()) return fail(IQ_ERR_INVALID, "iq_slice_candidates: bad tile index");
  if (idx) *idx = c->slice_idx[tile].data();
  if (val) *val = c->slice_val[tile].data();
  return IQ_OK;
}

int32_t iq_taumodel(int64_t n, int32_t nsrc, const float* vals, double* prob) {
  if (n <= 0 || nsrc <= 0 || !vals || !prob) return fail(IQ_ERR_INVALID, "iq_taumodel: bad argument");
  std::vector<double> p;
  taumodel(n, nsrc, vals, p);
  std::memcpy(prob, p.data(), (size_t)n * sizeof(double));
  return IQ_OK;
}

int32_t iq_sample(const double* prob, int64_t n, double u, int64_t* pos) {
  if (!prob || n <= 0 || !pos) return fail(IQ_ERR_INVALID, "iq_sample: bad argument");
  *pos = sample_walk(prob, n, u);
  return IQ_OK;
}

int32_t iq_cut_batch(iq_ctx* c, const iq_cut_task* tasks, int32_t ntask, int32_t* iters) {
  if (!c || !tasks || ntask < 0) return fail(IQ_ERR_INVALID, "iq_cut_batch: bad argument");
  if (ntask == 0) return IQ_OK;
  CK(cudaSetDevice(c->device));
  const size_t smem_limit = 220 * 1024;
  // layout of the staging block: [task records][iters][per device task: A, B (doubles), keep (bytes)]
  struct Plan { int dev; int n0, n1, L, da, db; size_t offA, offB, offK; long long nv; };
  std::vector<Plan> plan(ntask);
  size_t off = ((size_t)ntask * (sizeof(iq::CutTask) + sizeof(int)) + 255) & ~(size_t)255;
  size_t smem = 0;
  int ndev = 0;
  for (int t = 0; t < ntask; ++t) {
    const iq_cut_task& T = tasks[t];
    if (!T.A || !T.B || !T.keep || T.dim < 0 || T.dim > 2 || T.sz[T.dim] < 2)
      return fail(IQ_ERR_INVALID, "iq_cut_batch: task %d is malformed", t);
    Plan& P = plan[t];
    P.da = T.dim == 0 ? 1 : 0;
    P.db = T.dim == 2 ? 1 : 2;
    P.n0 = T.sz[P.da]; P.n1 = T.sz[P.db]; P.L = T.sz[T.dim];
    P.nv = (long long)T.sz[0] * T.sz[1] * T.sz[2];
    const size_t need = P.L >= 3 ? iq::graphcut_smem(P.n0, P.n1, P.L, c->cut_exact != 0) : 0;
    P.dev = (P.L >= 3 && need <= smem_limit && (long long)(P.L - 2) * P.n0 * P.n1 <= 4096) ? 1 : 0;
    if (P.dev) {
      smem = std::max(smem, need);
      P.offA = off; off += (size_t)P.nv * 8;
      P.offB = off; off += (size_t)P.nv * 8;
      P.offK = off; off += ((size_t)P.nv + 255) & ~(size_t)255;
      ++ndev;
    }
  }
  if (ndev > 0) {
    if (off > c->cut_cap) {
      if (c->h_cut) cudaFreeHost(c->h_cut);
      cudaFree(c->d_cut);
      c->h_cut = nullptr; c->d_cut = nullptr; c->cut_cap = 0;
      const size_t cap = off * 2;
      CK(cudaMallocHost((void**)&c->h_cut, cap));
      CK(iq::dmalloc((void**)&c->d_cut, cap));
      c->cut_cap = cap;
    }
    iq::CutTask* recs = (iq::CutTask*)c->h_cut;
    int* d_iters = (int*)(c->d_cut + (size_t)ntask * sizeof(iq::CutTask));
    int k = 0;
    for (int t = 0; t < ntask; ++t) {
      const Plan& P = plan[t];
      if (!P.dev) continue;
      const iq_cut_task& T = tasks[t];
      double* a = (double*)(c->h_cut + P.offA);
      double* b = (double*)(c->h_cut + P.offB);
      // re-layout: cut dimension slowest
      const long long st[3] = {1, T.sz[0], (long long)T.sz[0] * T.sz[1]};
      long long o = 0;
      for (int kk = 0; kk < P.L; ++kk)
        for (int bb = 0; bb < P.n1; ++bb)
          for (int aa = 0; aa < P.n0; ++aa, ++o) {
            const long long src = kk * st[T.dim] + bb * st[P.db] + aa * st[P.da];
            a[o] = T.A[src];
            b[o] = T.B[src];
          }
      iq::CutTask& R = recs[k];
      R.A = (const double*)(c->d_cut + P.offA);
      R.B = (const double*)(c->d_cut + P.offB);
      R.keep = (unsigned char*)(c->d_cut + P.offK);
      R.n0 = P.n0; R.n1 = P.n1; R.L = P.L;
      R.iters = d_iters + k;
      ++k;
    }
    CK(cudaMemcpyAsync(c->d_cut, c->h_cut, off, cudaMemcpyHostToDevice, c->stream));
    CK(iq::launch_graphcut((const iq::CutTask*)c->d_cut, ndev, smem, c->stream, c->cut_exact != 0));
    c->launches++;
    CK(cudaMemcpyAsync(c->h_cut, c->d_cut, off, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    const int* h_iters = (const int*)(c->h_cut + (size_t)ntask * sizeof(iq::CutTask));
    k = 0;
    for (int t = 0; t < ntask; ++t) {
      Plan& P = plan[t];
      if (!P.dev) continue;
      const iq_cut_task& T = tasks[t];
      if (h_iters[k] < 0) {
        P.dev = 0;  // iteration cap hit (never observed) / exact range exceeded: recompute on the host below
      } else {
        const unsigned char* kp = (const unsigned char*)(c->h_cut + P.offK);
        const long long st[3] = {1, T.sz[0], (long long)T.sz[0] * T.sz[1]};
        long long o = 0;
        for (int kk = 0; kk < P.L; ++kk)
          for (int bb = 0; bb < P.n1; ++bb)
            for (int aa = 0; aa < P.n0; ++aa, ++o) T.keep[kk * st[T.dim] + bb * st[P.db] + aa * st[P.da]] = kp[o];
        if (iters) iters[t] = h_iters[k];
      }
      ++k;
    }
  }
  for (int t = 0; t < ntask; ++t) {
    if (plan[t].dev) continue;
    const iq_cut_task& T = tasks[t];
    const int sz[3] = {T.sz[0], T.sz[1], T.sz[2]};
    if (!(c->cut_exact && iqcut::graphcut_exact(T.A, T.B, sz, T.dim, T.keep, c->cut_work)))
      iqcut::graphcut(T.A, T.B, sz, T.dim, T.keep, c->cut_work);
    if (iters) iters[t] = 0;
  }
  return IQ_OK;
}

int32_t iq_last_search_stats(const iq_ctx* c, double* device_ms, int64_t* kernel_launches) {
  if (!c) return fail(IQ_ERR_INVALID, "NULL context");
  if (device_ms) *device_ms = c->last_ms;
  if (kernel_launches) *kernel_launches = c->last_launches;
  return IQ_OK;
}

int32_t iq_last_search_path(const iq_ctx* c, int64_t* direct_searches, int64_t* fft_searches, double* fft_bytes, double* fft_ms) {
  if (!c) return fail(IQ_ERR_INVALID, "NULL context");
  if (direct_searches) *direct_searches = c->last_direct_searches;
  if (fft_searches) *fft_searches = c->last_fft_searches;
  if (fft_bytes) *fft_bytes = c->last_fft_bytes;
  if (fft_ms) *fft_ms = c->last_fft_ms;
  return IQ_OK;
}

int32_t iq_ctx_direct_kernel_launches(const iq_ctx* c, int64_t* tma_staged, int64_t* register_staged) {
  if (!c) return fail(IQ_ERR_INVALID, "NULL context");
  if (tma_staged) *tma_staged = c->direct_tma_launches;
  if (register_staged) *register_staged = c->direct_ldg_launches;
  return IQ_OK;
}

int32_t iq_last_search_kernel_ms(const iq_ctx* c, double* dist_ms, int64_t* dist_launches) {
  if (!c) return fail(IQ_ERR_INVALID, "NULL context");
  if (dist_ms) *dist_ms = c->last_dist_ms;
  if (dist_launches) *dist_launches = c->last_dist_launches;
  return IQ_OK;
}

static int32_t bench_fma_impl(int32_t device, int packed, double* tfma);
int32_t iq_bench_fma_peak(int32_t device, double* tfma) { return bench_fma_impl(device, 0, tfma); }
int32_t iq_bench_fma2_peak(int32_t device, double* tfma) { return bench_fma_impl(device, 1, tfma); }
static int32_t bench_fma_impl(int32_t device, int packed, double* tfma) {
  if (!tfma) return fail(IQ_ERR_INVALID, "NULL argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    cudaGetLastError();
    return fail(IQ_ERR_NO_DEVICE, "no such CUDA device");
  }
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  float* d = nullptr;
  CK(iq::dmalloc((void**)&d, 64));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int blocks = prop.multiProcessorCount * 8, iters = 4096;
  double best = 0.0;
  for (int rep = 0; rep < 6; ++rep) {
    CK(cudaEventRecord(e0, 0));
    CK(iq::launch_fma_peak(blocks, packed ? -iters : iters, d, 0));
    CK(cudaEventRecord(e1, 0));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double fma = (double)blocks * 256.0 * iters * 128.0;
    if (rep > 0) best = std::max(best, fma / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  *tfma = best;
  return IQ_OK;
}

int32_t iq_release_device_memory(int32_t device) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    cudaGetLastError();
    return fail(IQ_ERR_NO_DEVICE, "no such CUDA device");
  }
  CK(cudaSetDevice(device));
  CK(cudaDeviceSynchronize());
  cudaMemPool_t pool;
  CK(cudaDeviceGetDefaultMemPool(&pool, device));
  CK(cudaMemPoolTrimTo(pool, 0));
  return IQ_OK;
}

int32_t iq_device_free_memory(int32_t device, size_t* free_bytes, size_t* total_bytes) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    cudaGetLastError();
    return fail(IQ_ERR_NO_DEVICE, "no such CUDA device");
  }
  CK(cudaSetDevice(device));
  size_t f = 0, t = 0;
  CK(cudaMemGetInfo(&f, &t));
  // memory parked in the pool is handed out again by iq::dmalloc: count it as free
  cudaMemPool_t pool;
  unsigned long long reserved = 0, used = 0;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess &&
      cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved) == cudaSuccess &&
      cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used) == cudaSuccess && reserved > used)
    f += (size_t)(reserved - used);
  if (free_bytes) *free_bytes = f;
  if (total_bytes) *total_bytes = t;
  return IQ_OK;
}

int32_t iq_ctx_set_option(iq_ctx* c, const char* key, int64_t value) {
  if (!c || !key) return fail(IQ_ERR_INVALID, "NULL argument");
  if (std::strcmp(key, "rb") == 0) {
    if (value != 0 && value != 1 && value != 2 && value != 4) return fail(IQ_ERR_INVALID, "rb must be 0 (auto), 1, 2 or 4");
    c->rb_opt = (int)value;
    return IQ_OK;
  }
  if (std::strcmp(key, "tau_device") == 0) {
    c->tau_device = value ? 1 : 0;
    return IQ_OK;
  }
  if (std::strcmp(key, "fft") == 0) {
    if (value < -1 || value > 1) return fail(IQ_ERR_INVALID, "fft must be -1 (never), 0 (auto) or 1 (always)");
    c->fft_mode = (int)value;
    return IQ_OK;
  }
  if (std::strcmp(key, "cut_exact") == 0) {
    c->cut_exact = value ? 1 : 0;
    return IQ_OK;
  }
  if (std::strcmp(key, "variant") == 0) {
    if (value < 0 || value > 1) return fail(IQ_ERR_INVALID, "variant must be 0 (TMA-staged direct kernel) or 1 (register-staged direct kernel)");
    c->variant = (int)value;
    return IQ_OK;
  }
  return fail(IQ_ERR_INVALID, "unknown option '%s'", key);
}

}  // extern "C"

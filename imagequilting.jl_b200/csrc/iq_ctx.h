// iq_ctx.h -- private declarations shared by the context TU (iq_ctx.cu) and the device-resident simulation
// TU (iq_sim.cu).  Not part of the public ABI (include/iqb200.h is).
#pragma once
#include <cstdarg>
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/iqb200.h"
#include "iq_internal.h"
#include "iq_fft.h"
#include "iq_cut.h"

namespace iqimpl {

int fail(int code, const char* fmt, ...);

#define CK(call)                                                                                      \
  do {                                                                                                \
    cudaError_t e__ = (call);                                                                         \
    if (e__ != cudaSuccess)                                                                           \
      return iqimpl::fail(IQ_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e__)); \
  } while (0)

using iq::BoxDesc;

struct MaskEntry {
  std::vector<uint8_t> mask;
  uint64_t hash = 0;
  std::vector<BoxDesc> boxes;
  BoxDesc* d_boxes = nullptr;
  uint8_t* d_mask = nullptr;           // device copy of the mask bytes (resident simulation)
  long long nnz = 0;
  long long tmpl_floats = 0;           // packed template floats per tile
  std::map<int, float*> a2;            // image id (-1 = TI, s = aux s) -> A2 map
  // tensor maps of the TMA-staged direct kernel, keyed by (image id, panel shape XT * 16 + RS): the TMA box of every mask box
  std::map<std::pair<int, int>, iq::FlatTmaMaps> tma;
};

struct TileResult {
  std::vector<int64_t> idx;
  std::vector<double> prob;
  const int64_t* idx_ptr = nullptr;
  const double* prob_ptr = nullptr;
  int64_t count = 0;
};

struct SimState;
void sim_destroy(iq_ctx* c);

}  // namespace iqimpl

struct iq_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int ndim = 3;
  int nx = 1, ny = 1, nz = 1, tx = 1, ty = 1, tz = 1, nxo = 1, nyo = 1, nzo = 1;
  long long npos = 0, tilevol = 0, nenabled = 0;
  int nsoft = 0, max_batch = 1;
  int rb_opt = 0;  // 0 = auto
  int fft_mode = 0;   // 0 = auto (estimated crossover), 1 = always FFT for non-empty masks, -1 = never
  iqfft::Plan* fft = nullptr;
  bool fft_failed = false;
  std::map<int, bool> image_is_int;  // image id -> every voxel is integer-valued (exact rounding of AB)
  double last_fft_bytes = 0.0;
  int64_t last_fft_searches = 0, last_direct_searches = 0;
  int64_t direct_tma_launches = 0, direct_ldg_launches = 0;  // direct-kernel launches since the context was created
  int variant = 0; // direct kernel: 0 = TMA-staged, double-buffered (default), 1 = register-staged (also the fallback)
  std::map<int, float*> img_pad;  // image id -> copy with the row pitch rounded up to 16 bytes (TMA needs it; only when nx % 4 != 0)

  float* d_ti = nullptr;
  std::vector<float*> d_aux;
  double* d_sat_ti = nullptr;
  std::vector<double*> d_sat_aux;
  uint8_t* d_disabled = nullptr;
  std::vector<uint8_t> h_disabled;

  float* d_Dovl = nullptr;
  float* d_Dhard = nullptr;
  std::vector<float*> d_Dsoft;
  unsigned* d_minmax = nullptr;  // [kind][2][max_batch], kind 0 = ovl, 1 = hard, 2+s = soft
  unsigned* h_minmax = nullptr;

  // bump-allocated staging (pinned host + device mirror)
  char* h_stage = nullptr;
  char* d_stage = nullptr;
  size_t stage_cap = 0, stage_used = 0;

  unsigned long long* d_selbuf = nullptr;  // [max_batch][max_src][sel_cap] survivor lists of the radix select
  unsigned sel_cap = 0;
  iq::SelJob* d_sel = nullptr;
  int* d_sel_list = nullptr;               // [2 * max_batch * max_src + 3] job lists + counts of the select pass being launched
  iq::SelJob* h_sel = nullptr;
  iq::PickJob* d_pick = nullptr;
  iq::PickJob* h_pick = nullptr;
  unsigned* d_blockcount = nullptr;
  unsigned* d_chunkmin = nullptr;  // [max_batch][chunk_stride] chunk minima of the overlap distance maps
  long long chunk_stride = 0;
  int chunk_len = 0, chunk_n = 0;  // chunking of the most recent overlap-distance computation
  bool chunk_valid = false;        // d_chunkmin already filled (by the FFT epilogue)
  unsigned* d_total = nullptr;
  unsigned* h_total = nullptr;
  unsigned* d_cand_idx = nullptr;
  float* d_cand_val = nullptr;
  int max_src = 1;
  unsigned* h_cand_idx = nullptr;
  float* h_cand_val = nullptr;
  size_t h_cand_cap = 0;
  unsigned* d_rank = nullptr;             // [max_batch][max_src][kTauMax] dense ranks (device tau model)
  unsigned long long* d_colsum = nullptr; // [max_batch][max_src]
  double* d_prob = nullptr;               // [max_batch][kTauMax]
  double* h_prob = nullptr;               // pinned mirror
  int tau_device = 1;                     // 0 = always evaluate the tau model on the host
  int cut_exact = 0;                      // iq_cut_batch: exact integer arithmetic (option "cut_exact")
  // position-slice mode (iq_slice_*): candidates of the last select call
  std::vector<std::vector<int64_t>> slice_idx;
  std::vector<std::vector<float>> slice_val;
  int slice_ntile = 0;
  unsigned long long* d_slice_hist = nullptr;  // [8][256] digit histograms of iq_slice_hist
  unsigned long long* h_slice_hist = nullptr;  // page-locked
  char* h_cut = nullptr;  // pinned staging of the device boundary cut (slabs, masks, task records)
  char* d_cut = nullptr;
  size_t cut_cap = 0;
  iqcut::Work cut_work;   // host fallback scratch
  int* d_shifts = nullptr;
  int nshift = 0;
  float* d_fetch = nullptr;

  std::vector<std::unique_ptr<iqimpl::MaskEntry>> masks;
  iqimpl::MaskEntry* full_mask = nullptr;

  std::vector<iqimpl::TileResult> res;
  // cached "every enabled patch, uniform weights" answer of an empty overlap mask
  std::vector<int64_t> enabled_idx;
  std::vector<double> uniform_prob, uniform_cum;
  double uniform_sum = 0.0;

  double last_ms = 0.0;
  int64_t last_launches = 0;
  std::vector<cudaEvent_t> dist_ev;  // start/stop pairs around every distance computation of a search
  std::vector<char> dist_ev_fft;     // per pair: 1 = FFT path
  double last_fft_ms = 0.0;
  size_t dist_ev_used = 0;
  double last_dist_ms = 0.0;
  int64_t last_dist_launches = 0;
  int64_t launches = 0;
  iqimpl::SimState* sim = nullptr;  // device-resident simulation (iq_sim.cu), or nullptr
};

namespace iqimpl {

// helpers of iq_ctx.cu used by the resident simulation
int get_mask(iq_ctx* c, const uint8_t* mask, MaskEntry** out);
int get_a2(iq_ctx* c, MaskEntry* e, int image, const float** out);
int pick_rb(const iq_ctx* c, int R);
bool want_fft(const iq_ctx* c, const MaskEntry* e, int R);
int build_uniform(iq_ctx* c);
int ensure_fft(iq_ctx* c, int image);
// job0: first batch slot the R templates occupy (offset into the per-slot min/max and chunk-minimum arrays; d_b2 and
// d_out already point at that slot).  a2_list (FFT only): device array of one A2 map per template for launches that
// mix masks (e may then be nullptr).
int launch_fft(iq_ctx* c, MaskEntry* e, int image, const float* d_tmpl, const double* d_b2, int R, bool tint, float* d_out,
               int kind, int job0 = 0, const float* const* a2_list = nullptr);
int launch_direct(iq_ctx* c, MaskEntry* e, int image, const float* d_packed, const double* d_b2, int R, int rb, float* d_out,
                  int kind, int job0 = 0);
int collect_dist_times(iq_ctx* c);
// Makes sure d_chunkmin describes the R overlap-distance maps in d_Dovl (FFT epilogue, or one extra pass).
int ensure_chunkmin(iq_ctx* c, int R);
// B2 = sum of the squared (masked) template values in the library's fixed summation order, shared by the
// host-staged path and the device kernel k_sim_templates so that both produce the same bits.
double b2_ordered(const float* v, int tx, int ty, int tz);

}  // namespace iqimpl

// iq_internal.h -- shared declarations between the kernel TU (iq_kernels.cu) and the context /
// orchestration TU (iq_ctx.cu).  Not part of the public ABI (include/iqb200.h is).
#pragma once
#include <cuda.h>  // CUtensorMap (type only)
#include <cuda_runtime.h>
#include <stdint.h>

namespace iq {

constexpr int kTauMax = 32768; // largest candidate set ranked on the device (shared-memory bitonic sort)
constexpr int kT = 8;           // outputs per thread along x (register tile)
constexpr int kFlatTmaMaxBox = 8;  // mask boxes the TMA-staged direct kernel carries tensor maps for

// One axis-aligned box of the (disjoint) mask decomposition, in tile coordinates.
struct BoxDesc {
  int x0, y0, z0;   // origin inside the tile
  int w, h, d;      // extent
  int nch;          // ceil((pad + w) / 8): template rows are zero-padded to nch*8
  int tmpl_off;     // offset (floats, for ONE tile) of this box's packed template
  int pad;          // x0 & 3: leading zero taps of every packed template row.  The kernels read the image patch from
                    // x0 - pad, a multiple of 4 floats: a tensor-map box must start on a 16-byte boundary in its
                    // innermost dimension (a misaligned start coordinate traps with "illegal instruction" on sm_100)
};

// Parameters of the dense (box) correlation kernel.
struct DistParams {
  const float* img;            // training image / auxiliary TI, dense column-major
  int nx, ny, nz;              // image size
  int nxo, nyo, nzo;           // distance-map size
  long long npos;              // nxo*nyo*nzo
  int nbox;
  const BoxDesc* boxes;        // device array [nbox]
  const float* tmpl;           // packed templates [grp][box][qz][qy][chunk][r(RB)][8]
  long long tmpl_grp_stride;   // floats per tile group
  const float* a2;             // sum of img^2 over the mask per position (npos) or nullptr (=0)
  const double* b2;            // [R] sum of w*kern^2 per tile
  const uint8_t* disabled;     // [npos] or nullptr
  float* out;                  // [R][npos]
  unsigned* minbits;           // [R] atomicMin of float bits over enabled positions (or nullptr)
  unsigned* maxbits;           // [R] atomicMax of float bits over enabled positions (or nullptr)
  int R;                       // tiles in this launch
  int patch_floats;            // smem floats reserved for the image patch (per stage buffer)
  int tmpl_floats;             // TMA kernel: smem floats reserved for the template plane (per stage buffer)
  int XT;                      // x-threads (8 outputs each) per panel row
  int RS;                      // TMA kernel: rows per octet of work items (2: 4 x-threads x 2 rows, 8: 1 x-thread x 8 rows)
  int NB;                      // TMA kernel: stage buffers (2 = double-buffered, 1 = single)
};

// Tensor maps of the image for the TMA-staged direct kernel, one per mask box: 3-D (x, y, z) FP32, box =
// (XT + nch) * 8 + 4 columns x (rows of the CTA + h - 1) rows x 1 plane, no swizzle, out-of-bounds elements read as 0.
struct FlatTmaMaps {
  CUtensorMap m[kFlatTmaMaxBox];
};

struct SparseParams {
  const float* img;
  int nx, ny, nz, nxo, nyo, nzo;
  long long npos;
  const int* ptr;              // row r holds entries ptr[r * ptr_stride] .. ptr[r * ptr_stride + 1] - 1
  int ptr_stride;              // 1 = CSR row pointers [R+1]; 2 = one (begin, end) pair per row
  const long long* off;        // image offsets of the data voxels relative to the patch origin
  const float* val;            // data values
  const uint8_t* disabled;
  float* out;                  // [R][npos]
  unsigned* minbits;
  unsigned* maxbits;
  int R;
};

// Radix-select job: k-th smallest (value,index) key of one distance map.
struct SelJob {
  const float* map;              // npos floats
  unsigned long long k;          // 1-based rank still wanted inside the current prefix
  unsigned long long prefix;     // decided high bits of the key
  unsigned long long mask;       // which bits are decided
  unsigned long long kth;        // result: all keys <= kth are selected
  int pass;                      // index into the shift schedule
  int active;                    // 0 = done / skip
  unsigned ticket;               // last-block detection
  unsigned hist[256];
  // After the first two passes (16 value bits decided) the keys that still match the prefix -- typically ~1 % of the
  // map -- are copied to `cbuf`, and the remaining passes scan that list instead of the whole map.
  unsigned long long* cbuf;      // [ccap] scratch of this job (or nullptr: never compact)
  unsigned ccap;
  unsigned ccount;               // keys in cbuf
  int compact;                   // 1 = cbuf holds the surviving keys (set by the pass that decides it, filled by k_select_compact)
  // Range-adaptive digits (optional, nv > 0): the map's finite values lie in [sub, sub + 2^t) as float bits (known from
  // the distance epilogue), so the select runs on keys (bits - sub) << 32 | position and its FIRST digit is the top 8
  // bits of that range instead of the float's sign / exponent byte, which hardly discriminates: the k-th key's bin then
  // holds ~1 % of the map and the survivors are gathered after ONE full pass instead of two.  vshift = the value-digit
  // shifts (the position digits follow from the context's schedule); 0 everywhere = the plain schedule.
  unsigned sub;
  int nv;
  int vshift[4];
};

// Candidate predicate + outputs for one tile.
struct PickJob {
  int mode;                      // 0 = threshold on source 0, 1 = key <= kth for all sources
  int nsrc;
  const float* src[8];           // source maps (primary first)
  const SelJob* sel;             // mode 1: device array [nsrc] of finished radix-select jobs
  double tol;                    // mode 0: thr = (1+tol)*min
  const unsigned* minbits;       // mode 0: pointer to the tile's min bits
  unsigned* blockcount;          // [nblk] scratch
  unsigned* total;               // [1] out: number of candidates
  unsigned ticket;
  unsigned* cand_idx;            // [cap]
  float* cand_val;               // [nsrc][cap]
  long long cap;
  const unsigned* chunkmin;      // mode 0, optional: float bits of the minimum of every chunk of src[0] (k_pick_chunks)
  const int* pending;            // optional: the count / write kernels skip the job unless *pending != 0
};

// One boundary cut for the device kernel (iq_cutgpu.cu): slabs laid out with the cut dimension slowest.
struct CutTask {
  const double* A;      // already pasted content, [L][n1][n0]
  const double* B;      // new patch, same layout
  unsigned char* keep;  // out: 1 = keep the already pasted voxel
  int n0, n1, L;
  int* iters;           // out (optional): push-relabel sweeps used, -1 = iteration cap hit
};
size_t graphcut_smem(int n0, int n1, int L, bool exact = false);
cudaError_t launch_graphcut(const CutTask* d_tasks, int ntask, size_t smem, cudaStream_t s, bool exact = false);
// Implicit task grid (resident simulation): nslab * njobs CTAs, task (job, k) at (job * maxslabs + k) * maxslab, skipped
// when dims[task].z (L) < 2.
cudaError_t launch_graphcut_grid(const double* A, const double* B, unsigned char* keep, int* iters, const int4* dims,
                                 long long maxslab, int maxslabs, int nslab, int njobs, size_t smem, bool exact, cudaStream_t s);

// Device allocation from the stream-ordered default memory pool with the release threshold lifted: memory freed by
// one simulation (cudaFree returns it to the pool) is handed to the next without a trip to the driver's slow
// cudaMalloc path.  The pointer is ready for every stream on return.
cudaError_t dmalloc(void** p, size_t bytes);

// ---- launch wrappers (iq_kernels.cu) -------------------------------------------------
cudaError_t launch_dist_flat(const DistParams& p, const FlatTmaMaps& maps, int rb, size_t smem_bytes, cudaStream_t s);
cudaError_t launch_dist_flat_ldg(const DistParams& p, int rb, size_t smem_bytes, cudaStream_t s);
// shared memory of the TMA kernel for panel width XT (0 = the box list exceeds TMA's limits) / of the register-staged one
size_t dist_flat_smem(const BoxDesc* boxes, int nbox, int XT, int RS, int NB, int rb, int* patch_floats, int* tmpl_floats);
size_t dist_flat_ldg_smem(const BoxDesc* boxes, int nbox, int XT, int rb, int* patch_floats);
void dist_flat_box(const BoxDesc& b, int XT, int RS, int* width, int* rows);  // TMA box of mask box b at panel shape (XT, RS)
cudaError_t launch_dist_sparse(const SparseParams& p, cudaStream_t s);
cudaError_t launch_sat_build(const float* img, double* sat, int nx, int ny, int nz, cudaStream_t s);
cudaError_t launch_a2map(const double* sat, int nx, int ny, int nz, const BoxDesc* boxes, int nbox,
                         float* a2, int nxo, int nyo, int nzo, cudaStream_t s);
cudaError_t launch_fill_u32(unsigned* p, unsigned v, long long n, cudaStream_t s);
// All passes of a radix select (k_select_pass x nshift, with the compaction of the survivors after the second pass).
// scratch: 2 * njobs + 3 ints of device memory (job lists + counts of the pass being launched, zero-initialised)
cudaError_t launch_select_all(SelJob* jobs, int njobs, long long npos, const int* shifts, int nshift, cudaStream_t s,
                              int* launches, int* scratch);
// position-slice mode: digit histograms of the value bits of up to 8 local maps (k_slice_hist)
struct SliceHistReq {
  const float* map;
  unsigned prefix;  // value of the `level` higher digits (bits >> (32 - 8 level))
  int level;        // 0..3: digit = (bits >> (24 - 8 level)) & 255
};
struct SliceHistParams {
  SliceHistReq req[8];
  int nreq;
  long long npos;
  unsigned long long* out;  // [nreq][256], zeroed by the caller
};
cudaError_t launch_slice_hist(const SliceHistParams& P, cudaStream_t s);
cudaError_t launch_pick_count(PickJob* jobs, int njobs, long long npos, cudaStream_t s);
cudaError_t launch_pick_write(PickJob* jobs, int njobs, long long npos, cudaStream_t s);
// Threshold selection driven by chunk minima (chunk = chunklen consecutive positions): only chunks whose minimum
// passes the threshold are scanned.  total = kPickOverflow when more than kPickMaxChunks chunks qualify (the caller
// falls back to count/write).
constexpr int kPickMaxChunks = 2048;
constexpr unsigned kPickOverflow = 0xffffffffu;
cudaError_t launch_pick_chunks(PickJob* jobs, int njobs, long long npos, int chunklen, int nchunk, cudaStream_t s);
cudaError_t launch_chunkmin(const float* maps, int njobs, long long npos, int chunklen, int nchunk, unsigned* chunkmin,
                            long long pitch, cudaStream_t s);
cudaError_t launch_fetch_tile(const float* img, int nx, int ny, int nz, int tx, int ty, int tz,
                              long long x0, long long y0, long long z0, float* out, cudaStream_t s);
int pick_nblk(long long npos);
cudaError_t launch_tau(const PickJob* jobs, int njobs, int maxS, unsigned* rank, unsigned long long* colsum, double* prob,
                       cudaStream_t s);
cudaError_t launch_fma_peak(int blocks, int iters, float* out, cudaStream_t s);

}  // namespace iq

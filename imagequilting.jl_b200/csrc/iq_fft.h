// iq_fft.h -- hand-written shared-memory FFT cross-correlation (the path above the template-size
// crossover).  Internal API used by iq_ctx.cu; not part of the public ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace iqfft {

struct Plan;

// Epilogue arguments: the last inverse pass writes the distance maps directly.
struct Epilogue {
  const float* a2;          // [npos] sum of img^2 over the mask, or nullptr (= 0)
  const float* const* a2_list;  // optional [R] device array: one a2 map (or nullptr) per template -- templates of
                                // different masks in one call (mixed-shape launches of the resident pipeline)
  const double* b2;         // [R] sum of mask*kern^2
  const uint8_t* disabled;  // [npos] or nullptr
  float* out;               // [R][npos]
  unsigned* minbits;        // [R]
  unsigned* maxbits;        // [R]
  unsigned* chunkmin;       // [R][nchunk] float bits of the minimum over enabled positions of every chunk of
                            // final_chunk_rows(plan) consecutive x-rows (nchunk = ceil(nyo*nzo / rows)), or nullptr
  long long chunk_pitch;    // entries between the chunk minima of consecutive templates (>= nchunk)
  int round_to_int;         // 1: image and templates are integer-valued -> AB is rounded (exact result)
};

// nx,ny,nz: image size; tx,ty,tz: tile size; max_templates: largest R per correlate call.
cudaError_t plan_create(Plan** out, int nx, int ny, int nz, int tx, int ty, int tz, int max_templates, cudaStream_t s);
void plan_destroy(Plan* p);
// Computes (once) and caches the spectrum of the zero-padded image `id`.
cudaError_t plan_set_image(Plan* p, int id, const float* d_img, cudaStream_t s);
// d_tmpl: [R][tx*ty*tz] dense masked templates (zeros outside the mask), column-major.
// Writes |A2 - 2*AB + B2| (disabled -> +Inf) for every valid patch position and the min/max bits.
cudaError_t correlate(Plan* p, int id, const float* d_tmpl, int R, const Epilogue& ep, cudaStream_t s, int* launches);
// Bytes moved through global memory by one correlate call with R templates (algorithmic, for the roofline).
double correlate_bytes(const Plan* p, int R);
size_t plan_workspace_bytes(const Plan* p);
// x-rows of the distance map handled by one CTA of the last pass (granularity of Epilogue::chunkmin).
int final_chunk_rows(const Plan* p);

}  // namespace iqfft

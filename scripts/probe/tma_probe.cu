// tma_probe.cu -- which (tensor, box) geometries does cp.async.bulk.tensor.3d accept on sm_100?  One load of a box
// (bw x bh x 1) at (cx, cy, 0) of an nx x ny x nz FP32 tensor into shared memory, checked against the host image
// (out-of-bounds elements must read 0).  Prints OK / MISMATCH / the CUDA error; run every case in its own process
// (an illegal instruction poisons the context).  nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../imagequilting.jl_b200/csrc/iq_tma.cuh"

__global__ void k_probe(const __grid_constant__ CUtensorMap map, int cx, int cy, int n, float* out) {
  using namespace iqtma;
  extern __shared__ __align__(128) unsigned char raw[];
  float* sm = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(raw) + 127) & ~(uintptr_t)127);
  __shared__ __align__(8) unsigned long long bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_init_fence();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, (unsigned)n * 4u);
    tma_load_3d(sm, &map, cx, cy, 0, &bar);
  }
  mbar_wait_bounded(&bar, 0);
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = sm[i];
}

int main(int argc, char** argv) {
  if (argc < 8) { printf("usage: tma_probe nx ny nz bw bh cx cy\n"); return 2; }
  const int nx = atoi(argv[1]), ny = atoi(argv[2]), nz = atoi(argv[3]), bw = atoi(argv[4]), bh = atoi(argv[5]), cx = atoi(argv[6]),
            cy = atoi(argv[7]);
  const int nxp = (nx + 3) & ~3;
  std::vector<float> h((size_t)nxp * ny * nz, 0.f);
  for (int z = 0; z < nz; ++z)
    for (int y = 0; y < ny; ++y)
      for (int x = 0; x < nx; ++x) h[((size_t)z * ny + y) * nxp + x] = 1.f + x + 1000.f * y;
  float *d = nullptr, *o = nullptr;
  cudaMalloc(&d, h.size() * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  const int n = bw * bh;
  cudaMalloc(&o, (size_t)n * 4);
  CUtensorMap map;
  const iqtma::EncodeTiledFn enc = iqtma::encode_tiled_fn();
  if (!enc) { printf("no encoder\n"); return 1; }
  const cuuint64_t gdim[3] = {(cuuint64_t)nx, (cuuint64_t)ny, (cuuint64_t)nz};
  const cuuint64_t gstr[2] = {(cuuint64_t)nxp * 4, (cuuint64_t)nxp * ny * 4};
  const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1u};
  const cuuint32_t es[3] = {1, 1, 1};
  const CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("tensor %dx%dx%d box %dx%d at (%d,%d): encode %d ", nx, ny, nz, bw, bh, cx, cy, (int)r);
  if (r != CUDA_SUCCESS) { printf("\n"); return 0; }
  const size_t smem = (size_t)n * 4 + 256;
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_probe<<<1, 128, smem>>>(map, cx, cy, n, o);
  const cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("-> %s\n", cudaGetErrorString(e)); return 0; }
  std::vector<float> got((size_t)n);
  cudaMemcpy(got.data(), o, (size_t)n * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int j = 0; j < bh; ++j)
    for (int i = 0; i < bw; ++i) {
      const int x = cx + i, y = cy + j;
      const float want = (x < nx && y < ny) ? 1.f + x + 1000.f * y : 0.f;
      bad += got[(size_t)j * bw + i] != want;
    }
  printf("-> %s\n", bad ? "MISMATCH" : "OK");
  return 0;
}

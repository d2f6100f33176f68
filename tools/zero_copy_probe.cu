// Device -> page-locked host memory by SM stores (zero copy) vs the copy engine: GB/s by grid size and alignment.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_variants/zero_copy_probe tools/zero_copy_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

// rows of `row` doubles per problem, gathered from [row][ld] (like the engine's X) and written as [problem][row]
__global__ void gather_rows(const double* __restrict__ src, double* dst, int problems, int ld, int row, int aligned) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < problems; q += warps) {
    const size_t base = static_cast<size_t>(q) * row;
    // aligned: lane l owns the elements whose host index is = l (mod 32): every store instruction is one 256-byte line pair
    const int shift = aligned ? static_cast<int>(base & 31) : 0;
    for (int e = lane - shift; e < row; e += 32) {
      if (e >= 0) dst[base + e] = __ldcs(src + static_cast<size_t>(e) * ld + q);
    }
  }
}
__global__ void plain_copy(const double* __restrict__ src, double* dst, size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) dst[i] = __ldcs(src + i);
}
int main() {
  const int problems = 65536, ld = 65536, row = 324;
  const size_t n = static_cast<size_t>(problems) * row;
  double *d, *h;
  CK(cudaMalloc(&d, n * 8));
  CK(cudaMemset(d, 1, n * 8));
  CK(cudaHostAlloc(&h, n * 8, cudaHostAllocDefault));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  auto time = [&](auto fn) {
    fn();
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    for (int i = 0; i < 3; ++i) fn();
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return n * 8.0 * 3 / (ms * 1e-3) / 1e9;
  };
  printf("copy engine            %.1f GB/s\n", time([&] { cudaMemcpyAsync(h, d, n * 8, cudaMemcpyDeviceToHost); }));
  for (int g : {4, 8, 16, 32, 64, 148, 296}) {
    printf("grid %3d x 256: plain %.1f  gather %.1f  gather aligned %.1f GB/s\n", g, time([&] { plain_copy<<<g, 256>>>(d, h, n); }),
           time([&] { gather_rows<<<g, 256>>>(d, h, problems, ld, row, 0); }), time([&] { gather_rows<<<g, 256>>>(d, h, problems, ld, row, 1); }));
  }
  return 0;
}

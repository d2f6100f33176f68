// DMMA (fp64 tensor-core mma.sync) latency and throughput on sm_100a, next to DFMA (tools/fp64_latency_probe.cu).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_variants/dmma_probe tools/dmma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int SHAPE, int C>
__global__ void chain(double* out, int iters, long long* cycles) {
  double acc[C][4];
  for (int c = 0; c < C; ++c)
    for (int i = 0; i < 4; ++i) acc[c][i] = threadIdx.x * 1e-9 + c + i;
  double a0 = 1.0 + threadIdx.x * 1e-9, a1 = 1.0000001, a2 = 0.9999999, a3 = 1.0000002, b0 = 1e-3, b1 = 2e-3;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < C; ++c) {
      if (SHAPE == 0)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(acc[c][0]), "+d"(acc[c][1]) : "d"(a0), "d"(b0));
      if (SHAPE == 1)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+d"(acc[c][0]), "+d"(acc[c][1]), "+d"(acc[c][2]), "+d"(acc[c][3])
                     : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
    }
  }
  long long t1 = clock64();
  double s = 0;
  for (int c = 0; c < C; ++c)
    for (int i = 0; i < 4; ++i) s += acc[c][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int SHAPE, int C>
void run(int warps) {
  double* out;
  long long* cyc;
  cudaMalloc(&out, 1 << 22);
  cudaMalloc(&cyc, 8);
  const int iters = 2048;
  chain<SHAPE, C><<<148, 32 * warps>>>(out, iters, cyc);
  chain<SHAPE, C><<<148, 32 * warps>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double flop = (SHAPE == 0 ? 512.0 : 2048.0) * C * iters * warps;
  printf("%-8s warps/SM %2d chains %d : %7.1f cycles per dependent mma, %6.1f flop/cycle/SM (%5.1f TFLOP/s at 1.965 GHz x 148 SMs); cuda: %s\n",
         SHAPE == 0 ? "m8n8k4" : "m16n8k8", warps, C, (double)h / iters, flop / h, flop / h * 1.965e9 * 148 / 1e12, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<0, 1>(1); run<0, 2>(1); run<0, 4>(1); run<0, 2>(4); run<0, 2>(8); run<0, 4>(8); run<0, 4>(16);
  run<1, 1>(1); run<1, 2>(1); run<1, 2>(4); run<1, 2>(8); run<1, 4>(8); run<1, 4>(16);
  return 0;
}

// fp64 pipe microbenchmark for sm_100a: dependent-issue latency of DFMA / DADD / DMUL and the throughput one SM reaches with
// W warps per scheduler and C independent chains per thread (what bounds the iLQR kernels: profiles/README.md).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_variants/fp64_latency_probe tools/fp64_latency_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int C, int OP>
__global__ void chain(double* out, int iters, long long* cycles) {
  double a[C];
  for (int c = 0; c < C; ++c) a[c] = 1.0 + threadIdx.x * 1e-9 + c;
  const double m = 1.0000001, b = 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < C; ++c) {
      if (OP == 0) a[c] = fma(a[c], m, b);
      if (OP == 1) a[c] = __dadd_rn(a[c], b);
      if (OP == 2) a[c] = __dmul_rn(a[c], m);
      if (OP == 3) a[c] = __dadd_rn(__dmul_rn(a[c], m), b);  // unfused multiply-add: two dependent instructions
    }
  }
  long long t1 = clock64();
  double s = 0;
  for (int c = 0; c < C; ++c) s += a[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int C, int OP>
void run(const char* name, int warps_per_sm) {
  double* out;
  long long* cyc;
  cudaMalloc(&out, 1 << 22);
  cudaMalloc(&cyc, 8);
  const int iters = 4096;
  const int threads = 32 * warps_per_sm;
  chain<C, OP><<<148, threads>>>(out, iters, cyc);
  chain<C, OP><<<148, threads>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double ops = (OP == 3 ? 2.0 : 1.0) * C * iters;
  printf("%-6s warps/SM %2d chains/thread %d : %6.2f cycles per chain-op, %5.2f warp-instr / cycle / SM\n", name, warps_per_sm, C, (double)h / (ops / C),
         ops * warps_per_sm / (double)h);
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<1, 0>("DFMA", 1);
  run<1, 1>("DADD", 1);
  run<1, 2>("DMUL", 1);
  run<1, 3>("MUL+ADD", 1);
  run<2, 0>("DFMA", 1);
  run<4, 0>("DFMA", 1);
  run<8, 0>("DFMA", 1);
  run<1, 0>("DFMA", 4);
  run<2, 0>("DFMA", 4);
  run<4, 0>("DFMA", 4);
  run<1, 0>("DFMA", 8);
  run<2, 0>("DFMA", 8);
  run<1, 0>("DFMA", 16);
  run<2, 0>("DFMA", 16);
  run<4, 0>("DFMA", 16);
  run<2, 3>("MUL+ADD", 16);
  run<1, 0>("DFMA", 32);
  run<2, 0>("DFMA", 32);
  return 0;
}

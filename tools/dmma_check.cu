// Checks dmma_product (centralized.cuh, opt-in tensor-core path) against plain host sums on random operands.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Iinclude -Imulti_agent_solver_b200/csrc -o tools/_variants/dmma_check tools/dmma_check.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "centralized.cuh"
using namespace mas_b200;

__global__ void k_vxx(const double* fK, const double* fQ, const double* KtQ, const double* Qxx, double* Vxx, int ns, int ms, int ldk) {
  const DmmaTerm tv[3] = {{fK, (size_t)ldk, 1, 1.0, fQ, 1, (size_t)ldk}, {fQ, (size_t)ldk, 1, 1.0, fK, 1, (size_t)ldk}, {KtQ, 1, (size_t)ns, 1.0, fK, 1, (size_t)ldk}};
  dmma_product<3>(tv, ns, ns, ms, Qxx, ns, Vxx, ns, nullptr, 0, threadIdx.x, blockDim.x);
}
__global__ void k_gain(const double* inv, const double* fQ, double* Kt, double* fK, int ns, int ms, int ldk) {
  const DmmaTerm tk[1] = {{inv, 1, (size_t)ldk, -1.0, fQ, 1, (size_t)ldk}};
  dmma_product<1>(tk, ms, ns, ms, nullptr, 0, Kt, ms, fK, ldk, threadIdx.x, blockDim.x);
}
int main() {
  const int ns = 128, ms = 64, ldk = ms + 1;
  std::vector<double> fK((size_t)ns * ldk), fQ((size_t)ns * ldk), KtQ((size_t)ns * ms), Qxx((size_t)ns * ns), inv((size_t)ms * ldk);
  srand(1);
  auto rnd = [] { return rand() / (double)RAND_MAX - 0.5; };
  for (auto& v : fK) v = rnd();
  for (auto& v : fQ) v = rnd();
  for (auto& v : KtQ) v = rnd();
  for (auto& v : Qxx) v = rnd();
  for (auto& v : inv) v = rnd();
  double *dK, *dQ, *dW, *dX, *dV, *dI, *dKt, *dK2;
  cudaMalloc(&dK, fK.size() * 8); cudaMalloc(&dQ, fQ.size() * 8); cudaMalloc(&dW, KtQ.size() * 8); cudaMalloc(&dX, Qxx.size() * 8);
  cudaMalloc(&dV, Qxx.size() * 8); cudaMalloc(&dI, inv.size() * 8); cudaMalloc(&dKt, (size_t)ms * ns * 8); cudaMalloc(&dK2, fK.size() * 8);
  cudaMemcpy(dK, fK.data(), fK.size() * 8, cudaMemcpyHostToDevice); cudaMemcpy(dQ, fQ.data(), fQ.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dW, KtQ.data(), KtQ.size() * 8, cudaMemcpyHostToDevice); cudaMemcpy(dX, Qxx.data(), Qxx.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dI, inv.data(), inv.size() * 8, cudaMemcpyHostToDevice);
  k_vxx<<<1, 256>>>(dK, dQ, dW, dX, dV, ns, ms, ldk);
  k_gain<<<1, 256>>>(dI, dQ, dKt, dK2, ns, ms, ldk);
  std::vector<double> V(Qxx.size()), Kt((size_t)ms * ns);
  cudaMemcpy(V.data(), dV, V.size() * 8, cudaMemcpyDeviceToHost); cudaMemcpy(Kt.data(), dKt, Kt.size() * 8, cudaMemcpyDeviceToHost);
  printf("cuda: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  double e1 = 0, e2 = 0;
  for (int j = 0; j < ns; ++j)
    for (int i = 0; i < ns; ++i) {
      double s = Qxx[i + (size_t)j * ns];
      for (int k = 0; k < ms; ++k) s += fK[k + (size_t)i * ldk] * fQ[k + (size_t)j * ldk];
      for (int k = 0; k < ms; ++k) s += fQ[k + (size_t)i * ldk] * fK[k + (size_t)j * ldk];
      for (int k = 0; k < ms; ++k) s += KtQ[i + (size_t)k * ns] * fK[k + (size_t)j * ldk];
      e1 = fmax(e1, fabs(s - V[i + (size_t)j * ns]));
    }
  for (int j = 0; j < ns; ++j)
    for (int i = 0; i < ms; ++i) {
      double s = 0;
      for (int k = 0; k < ms; ++k) s += -inv[i + (size_t)k * ldk] * fQ[k + (size_t)j * ldk];
      e2 = fmax(e2, fabs(s - Kt[i + (size_t)j * ms]));
    }
  printf("max |V_xx(dmma) - V_xx(host)| = %.3e   max |K(dmma) - K(host)| = %.3e  (operands in [-0.5, 0.5], inner dimension 64)\n", e1, e2);
  return (e1 < 1e-12 && e2 < 1e-12) ? 0 : 1;
}

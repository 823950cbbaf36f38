// Standalone (no Python) check + timing of the experimental systolic Jacobi against the product
// direct-mode solver, both called through libkfac_b200.so.
//   nvcc -O2 -std=c++17 -I include tests/host/jacobi_systolic_gpu.cu -L kfac-pytorch_b200/csrc -lkfac_b200 -o tests/host/bin/jsys_gpu
//   LD_LIBRARY_PATH=kfac-pytorch_b200/csrc tests/host/bin/jsys_gpu [n] [count]
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "kfac_b200.h"

extern "C" int kfac_experimental_jacobi_systolic(const float* F, int n, int count, float* Q, float* d, int max_sweeps, int flags,
                                                 void* stream);

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { std::printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 2; } } while (0)

static void verify(const char* tag, const std::vector<float>& F, const std::vector<float>& Q, const std::vector<float>& d, int n,
                   int count) {
  double worst_orth = 0, worst_rec = 0;
  for (int m = 0; m < count; ++m) {
    const float* q = &Q[(size_t)m * n * n];
    const float* f = &F[(size_t)m * n * n];
    double scale = 0;
    for (int i = 0; i < n * n; ++i) scale = std::fmax(scale, std::fabs((double)f[i]));
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        double o = 0, r = 0;
        for (int k = 0; k < n; ++k) { o += (double)q[k * n + i] * q[k * n + j]; r += (double)q[i * n + k] * d[(size_t)m * n + k] * q[j * n + k]; }
        worst_orth = std::fmax(worst_orth, std::fabs(o - (i == j)));
        worst_rec = std::fmax(worst_rec, std::fabs(r - f[i * n + j]) / scale);
      }
  }
  std::printf("%-10s orth=%.2e reconstruction=%.2e %s\n", tag, worst_orth, worst_rec, (worst_orth < 1e-4 && worst_rec < 5e-5) ? "OK" : "BAD");
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? std::atoi(argv[1]) : 64, count = argc > 2 ? std::atoi(argv[2]) : 72;
  std::mt19937 rng(1);
  std::normal_distribution<float> nd;
  std::vector<float> F((size_t)count * n * n);
  for (int m = 0; m < count; ++m) {   // graded SPD: X^T X with column scales 10^(-3 j / n)
    std::vector<double> X((size_t)2 * n * n);
    for (int i = 0; i < 2 * n; ++i)
      for (int j = 0; j < n; ++j) X[(size_t)i * n + j] = nd(rng) * std::pow(10.0, -3.0 * j / n * (m % 2));
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) { double s = 0; for (int r = 0; r < 2 * n; ++r) s += X[(size_t)r * n + i] * X[(size_t)r * n + j]; F[(size_t)m * n * n + i * n + j] = (float)s; }
  }
  float *dF, *dQ, *dD;
  CK(cudaMalloc(&dF, F.size() * 4)); CK(cudaMalloc(&dQ, F.size() * 4)); CK(cudaMalloc(&dD, (size_t)count * n * 4));
  CK(cudaMemcpy(dF, F.data(), F.size() * 4, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  std::vector<float> Q(F.size()), d((size_t)count * n);
  float ms;
  // ---- experimental (flags 0: IEEE rotation chain, 1: fast chain)
  for (int flags = 0; flags < 8; ++flags) {
    if (n <= 64 && (flags & 2)) continue;   // bit 1 only matters for n > 64
    if ((flags & 5) == 5) continue;          // bits 0 and 2 are alternatives
    for (int it = 0; it < 3; ++it)
      if (kfac_experimental_jacobi_systolic(dF, n, count, dQ, dD, 0, flags, nullptr)) { std::printf("systolic: %s\n", kfac_last_error()); return 1; }
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int it = 0; it < 20; ++it) kfac_experimental_jacobi_systolic(dF, n, count, dQ, dD, 0, flags, nullptr);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    std::printf("systolic/%d n=%d x%d: %.1f us per launch\n", flags, n, count, ms * 50);
    CK(cudaMemcpy(Q.data(), dQ, Q.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(d.data(), dD, d.size() * 4, cudaMemcpyDeviceToHost));
    char tag[32]; std::snprintf(tag, sizeof tag, "systolic/%d", flags); verify(tag, F, Q, d, n, count);
  }
  // ---- product direct mode (jacobi_smem_kernel through kfac_eigh_batched)
  std::vector<kfac_eigh_item> items(count);
  std::vector<int> ns(count, n);
  for (int m = 0; m < count; ++m) items[m] = kfac_eigh_item{dF + (size_t)m * n * n, dQ + (size_t)m * n * n, nullptr, dD + (size_t)m * n, n, n, nullptr};
  const size_t wsb = kfac_eigh_workspace_bytes(ns.data(), count);
  void* ws; CK(cudaMalloc(&ws, wsb + 256));
  for (int it = 0; it < 3; ++it)
    if (kfac_eigh_batched(items.data(), count, ws, wsb, 0, 0.f, nullptr)) { std::printf("eigh: %s\n", kfac_last_error()); return 1; }
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int it = 0; it < 20; ++it) kfac_eigh_batched(items.data(), count, ws, wsb, 0, 0.f, nullptr);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
  std::printf("product    n=%d x%d: %.1f us per call (incl. descriptor upload)\n", n, count, ms * 50);
  CK(cudaMemcpy(Q.data(), dQ, Q.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(d.data(), dD, d.size() * 4, cudaMemcpyDeviceToHost));
  verify("product", F, Q, d, n, count);
  return 0;
}

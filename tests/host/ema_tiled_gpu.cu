// Standalone (no Python) check + timing of the experimental tiled EMA kernel against kfac_factor_ema,
// on the ResNet-50 factor dimensions.
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I include tests/host/ema_tiled_gpu.cu \
//        -L kfac-pytorch_b200/csrc -lkfac_b200 -o tests/host/bin/ema_gpu
//   LD_LIBRARY_PATH=kfac-pytorch_b200/csrc tests/host/bin/ema_gpu
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include "kfac_b200.h"

extern "C" int kfac_experimental_factor_ema_tiled(const kfac_ema_item* items, int count, float alpha, void* stream);

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { std::printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 2; } } while (0)

int main() {
  // (a, g) x count of torchvision ResNet-50
  const int layers[][3] = {{147, 64, 1}, {64, 64, 1}, {64, 256, 4}, {256, 64, 2}, {576, 64, 3}, {128, 512, 4}, {256, 128, 1},
                           {256, 512, 1}, {512, 128, 3}, {1152, 128, 4}, {256, 1024, 6}, {512, 256, 1}, {512, 1024, 1},
                           {1024, 256, 5}, {2304, 256, 6}, {512, 2048, 3}, {1024, 512, 1}, {1024, 2048, 1}, {2048, 512, 2},
                           {4608, 512, 3}, {2049, 1000, 1}};
  std::vector<int> dims;
  for (auto& l : layers) for (int c = 0; c < l[2]; ++c) { dims.push_back(l[0]); dims.push_back(l[1]); }
  size_t total = 0; for (int d : dims) total += (size_t)d * d;
  std::printf("%zu factors, %.1f MB\n", dims.size(), total * 4e-6);
  std::vector<float> hf(total), hb(total);
  std::mt19937 rng(3); std::normal_distribution<float> nd;
  for (size_t i = 0; i < total; ++i) { hf[i] = nd(rng); hb[i] = nd(rng); }
  float *F[2], *B[2];
  for (int v = 0; v < 2; ++v) { CK(cudaMalloc(&F[v], total * 4)); CK(cudaMalloc(&B[v], total * 4)); }
  std::vector<kfac_ema_item> items[2];
  for (int v = 0; v < 2; ++v) {
    size_t off = 0;
    for (size_t k = 0; k < dims.size(); ++k) { items[v].push_back(kfac_ema_item{F[v] + off, B[v] + off, dims[k], (int)(k % 5 == 0), 0.5f}); off += (size_t)dims[k] * dims[k]; }
  }
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float ms[2] = {0, 0};
  for (int rep = 0; rep < 6; ++rep)
    for (int v = 0; v < 2; ++v) {
      CK(cudaMemcpy(F[v], hf.data(), total * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(B[v], hb.data(), total * 4, cudaMemcpyHostToDevice));
      CK(cudaEventRecord(e0));
      const int rc = v ? kfac_experimental_factor_ema_tiled(items[v].data(), (int)items[v].size(), 0.95f, nullptr)
                       : kfac_factor_ema(items[v].data(), (int)items[v].size(), 0.95f, nullptr);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      if (rc) { std::printf("error: %s\n", kfac_last_error()); return 1; }
      float t; CK(cudaEventElapsedTime(&t, e0, e1));
      if (rep >= 2) ms[v] += t / 4;
    }
  std::vector<float> r0(total), r1(total), z0(total), z1(total);
  CK(cudaMemcpy(r0.data(), F[0], total * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(r1.data(), F[1], total * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(z0.data(), B[0], total * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(z1.data(), B[1], total * 4, cudaMemcpyDeviceToHost));
  const bool same = std::memcmp(r0.data(), r1.data(), total * 4) == 0 && std::memcmp(z0.data(), z1.data(), total * 4) == 0;
  std::printf("product %.3f ms, tiled %.3f ms (5 x %.0f MB at 7.7 TB/s = %.3f ms); results %s\n", ms[0], ms[1], total * 4e-6,
              5 * total * 4 / 7.7e9, same ? "bit-identical" : "DIFFER");
  return same ? 0 : 1;
}

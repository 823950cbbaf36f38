// EXPERIMENTAL (round-2 candidate, not on the product path): tiled EMA update.
// kfac_factor_ema's kernel (factor.cu) touches the mirror element of every (i, j) with stride d
// (3.6 ms per ResNet-50 step where moving the 5 x 615 MB at HBM speed takes 0.4 ms).  Here a CTA
// owns the pair of 32x32 tiles (ti, tj) / (tj, ti): both are read and written row-wise, the
// transposition happens in shared memory.  Same arithmetic, bit-identical results.
#include "common.cuh"

namespace kfac {

struct EmaTiledBatch {
  kfac_ema_item it[40];
  int first_block[41];   // prefix sums of the number of tile pairs T (T + 1) / 2 per item
  int count;
  float alpha;
};

__global__ void __launch_bounds__(256) ema_tiled_kernel(const __grid_constant__ EmaTiledBatch eb) {
  __shared__ float A[32][33], B[32][33];
  int item = 0;
  while (item + 1 < eb.count && (int)blockIdx.x >= eb.first_block[item + 1]) ++item;
  const kfac_ema_item it = eb.it[item];
  const int p = blockIdx.x - eb.first_block[item];
  // unrank the tile pair (ti <= tj) from p = tj (tj + 1) / 2 + ti
  int tj = (int)((sqrtf(8.f * (float)p + 1.f) - 1.f) * 0.5f);
  while ((tj + 1) * (tj + 2) / 2 <= p) ++tj;
  while (tj * (tj + 1) / 2 > p) --tj;
  const int ti = p - tj * (tj + 1) / 2;
  const int d = it.d, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float alpha = eb.alpha, beta = (1.f - eb.alpha) * it.inv_count * 0.5f;
  const int i0 = ti * 32, j0 = tj * 32;
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int iu = i0 + r, ju = j0 + tx;   // upper tile (ti, tj)
    A[r][tx] = (iu < d && ju < d) ? it.batch[(int64_t)iu * d + ju] : 0.f;
    const int il = j0 + r, jl = i0 + tx;   // mirror tile (tj, ti)
    B[r][tx] = (il < d && jl < d) ? it.batch[(int64_t)il * d + jl] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    {
      const int i = i0 + r, j = j0 + tx;
      if (i < d && j < d) {
        const int64_t idx = (int64_t)i * d + j;
        const float s = beta * (A[r][tx] + B[tx][r]);
        const float f = it.first ? (i == j ? 1.f : 0.f) : it.factor[idx];
        it.factor[idx] = alpha * f + s;
        it.batch[idx] = 0.f;
      }
    }
    if (ti != tj) {
      const int i = j0 + r, j = i0 + tx;
      if (i < d && j < d) {
        const int64_t idx = (int64_t)i * d + j;
        const float s = beta * (A[tx][r] + B[r][tx]);
        const float f = it.first ? 0.f : it.factor[idx];
        it.factor[idx] = alpha * f + s;
        it.batch[idx] = 0.f;
      }
    }
  }
}

}  // namespace kfac

// Not declared in include/kfac_b200.h on purpose (experimental); same contract as kfac_factor_ema.
extern "C" int kfac_experimental_factor_ema_tiled(const kfac_ema_item* items, int count, float alpha, void* stream) {
  using namespace kfac;
  KFAC_CHECK_ARG(count >= 0 && (items || count == 0), "items");
  cudaStream_t s = (cudaStream_t)stream;
  for (int base = 0; base < count; base += 40) {
    EmaTiledBatch eb{};
    eb.count = min(40, count - base);
    eb.alpha = alpha;
    int blocks = 0;
    for (int i = 0; i < eb.count; ++i) {
      eb.it[i] = items[base + i];
      KFAC_CHECK_ARG(eb.it[i].factor && eb.it[i].batch && eb.it[i].d > 0, "ema item");
      eb.first_block[i] = blocks;
      const int T = (eb.it[i].d + 31) / 32;
      blocks += T * (T + 1) / 2;
    }
    eb.first_block[eb.count] = blocks;
    ema_tiled_kernel<<<blocks, 256, 0, s>>>(eb);
    KFAC_LAUNCH_CHECK();
  }
  return KFAC_OK;
}

// EXPERIMENTAL: device side of jacobi_systolic.cuh (see there).  Not used by the product path;
// exported only through kfac_experimental_jacobi_systolic() so that round 2 can validate and time
// it against jacobi_smem_kernel before wiring it into the block solver of eigh.cu.
#include "common.cuh"
#include "eigh_common.cuh"
#include "jacobi_systolic.cuh"

namespace kfac {

// One CTA per matrix: TB bulk threads + N/2 crit threads.  F: count matrices n x n (n <= N),
// Q: eigenvectors in columns (ld n), d: eigenvalues (unsorted, clamped at 0).
template <int N, int TB>
__global__ void __launch_bounds__(TB + N / 2) jacobi_systolic_kernel(const float* F, int n, float* Q, float* d,
                                                                    int max_sweeps, int fast) {
  using namespace sysj;
  constexpr int h = N / 2, LD = N;
  extern __shared__ float sm[];
  float* M = sm;
  float* W = sm + N * N;
  Rot* rot = reinterpret_cast<Rot*>(W + N * N);
  int* flag = reinterpret_cast<int*>(rot + 2 * h);   // rot: 2 x h (double buffered)
  float* colscale = reinterpret_cast<float*>(flag + 4);
  const int tid = threadIdx.x, T = TB + h;
  const bool is_crit = tid >= TB;
  const int k = tid - TB;
  const float* Fm = F + (size_t)blockIdx.x * n * n;
  for (int idx = tid; idx < N * N; idx += T) {
    const int i = idx / N, j = idx % N;
    M[idx] = (i < n && j < n) ? Fm[(size_t)i * n + j] : 0.f;
    W[idx] = (i == j) ? 1.f : 0.f;
  }
  if (tid == 0) { flag[0] = 0; flag[1] = 0; }
  __syncthreads();
  const Criteria cr{0, 1e-7f, 0.f, 0.f, fast};
  // rot is double buffered by step parity: the crit threads publish the rotations of step t+1
  // while the bulk threads may still be reading those of step t
  int par = 0;
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    int* fl = flag + (sweep & 1);              // "somebody rotated / rotated a lot" of this sweep
    if (tid == 0) flag[(sweep + 1) & 1] = 0;   // the next sweep's word: last read one sweep ago
    if (is_crit) {
      int f = 0;
      rot[par * h + k] = first_rotation<N, LD>(k, cr, M, f);
      if (f) atomicOr(fl, f);
    }
    __syncthreads();
    for (int st = 0; st < N - 1; ++st) {
      const bool more = st < N - 2;
      const Rot* rcur = rot + par * h;
      if (!is_crit) {
        BulkRegs<N, TB> regs;
        bulk_load<N, LD, TB>(tid, rcur, M, W, regs);
        // every load of the step (bulk and crit) is done before anybody stores in place
        asm volatile("bar.sync 1, %0;" ::"r"(T) : "memory");
        bulk_store<N, LD, TB>(tid, M, W, regs);
      } else {
        CritRegs cregs;
        if (more) crit_load<N, LD>(k, rcur, M, cregs);
        // arrive without waiting: the div/sqrt chain of the next rotation runs while the bulk
        // threads finish loading and then store
        asm volatile("bar.arrive 1, %0;" ::"r"(T) : "memory");
        if (more) {
          int f = 0;
          rot[(par ^ 1) * h + k] = crit_rotation(cr, cregs, f);
          if (f) atomicOr(fl, f);
        }
      }
      __syncthreads();
      par ^= 1;
    }
    const int f = *fl;   // complete: the last step ended with a barrier
    __syncthreads();     // everybody has read it before thread 0 recycles the word (two sweeps later)
    // all rotations of the sweep tiny (or none): the leftovers are second order
    if (!(f & 1) || !(f & 2)) break;
  }
  __syncthreads();
  // rescale by the (rounding-drifted) column norms of W
  for (int j = tid; j < N; j += T) {
    float ww = 0.f;
    for (int i = 0; i < N; ++i) ww = fmaf(W[i * LD + j], W[i * LD + j], ww);
    colscale[j] = ww > 0.f ? 1.f / sqrtf(ww) : 0.f;
  }
  __syncthreads();
  float* Qm = Q + (size_t)blockIdx.x * n * n;
  for (int idx = tid; idx < n * n; idx += T) {
    const int i = idx / n, j = idx % n;
    Qm[idx] = W[i * LD + j] * colscale[j];
  }
  for (int j = tid; j < n; j += T) d[(size_t)blockIdx.x * n + j] = fmaxf(M[j * LD + j] * colscale[j] * colscale[j], 0.f);
}

template <int N, int TB>
static int launch_systolic(const float* F, int n, int count, float* Q, float* d, int max_sweeps, int fast,
                           cudaStream_t s) {
  const size_t smem = ((size_t)2 * N * N + 2 * N /*rot x2*/ + 4 /*flag*/ + N /*colscale*/) * sizeof(float);
  KFAC_CUDA(cudaFuncSetAttribute(jacobi_systolic_kernel<N, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  jacobi_systolic_kernel<N, TB><<<count, TB + N / 2, smem, s>>>(F, n, Q, d, max_sweeps, fast);
  KFAC_LAUNCH_CHECK();
  return KFAC_OK;
}

// ---------------------------------------------------------------------------------------------
// Block mode: same contract as jacobi_smem_kernel<N> with mode_block = 1 in eigh.cu (one CTA per
// block pair: Gram M read from and zeroed in the pair buffer, convergence bookkeeping of the
// matrix, W (mode 3: W^T) written to the pair buffer), with the systolic sweeps in the middle.
template <int N, int TB>
__global__ void __launch_bounds__(TB + N / 2) jacobi_systolic_block_kernel(EighMat* mats, const int* list, int max_inner,
                                                                          int tc_first, int* active_list,
                                                                          int* active_count, int pair_shift, int opts) {
  using namespace sysj;
  constexpr int h = N / 2, LD = N;
  const int fast = opts & 1;
  extern __shared__ float sm[];
  float* M = sm;
  float* W = sm + N * N;
  Rot* rot = reinterpret_cast<Rot*>(W + N * N);     // 2 x h, double buffered by step parity
  int* flag = reinterpret_cast<int*>(rot + 2 * h);  // 2 words, alternating by sweep
  __shared__ float redmax[32];
  const int tid = threadIdx.x, T = TB + h;
  const bool is_crit = tid >= TB;
  const int k = tid - TB;
  EighMat& mt = mats[list[blockIdx.x]];
  if (mt.done) return;
  const int local = blockIdx.x + pair_shift - mt.inner_base;
  float* Mg = mt.M + (int64_t)local * N * N;
  for (int idx = tid; idx < N * N; idx += T) {
    M[idx] = Mg[idx];
    Mg[idx] = 0.f;
    W[idx] = (idx / N == idx % N) ? 1.f : 0.f;
  }
  if (tid == 0) { flag[0] = 0; flag[1] = 0; }
  __syncthreads();
  const float tol = mt.tol;
  {
    // largest relative off-diagonal of this pair's Gram (convergence measure), as in eigh.cu
    float dmax = 0.f;
    for (int j = tid; j < N; j += T) dmax = fmaxf(dmax, M[j * LD + j]);
    if (dmax > 0.f) atomicMax(&mt.max_diag, __float_as_uint(dmax));
    const float max_diag = __uint_as_float(mt.max_diag);
    const float nw_ratio = mt.nw_ratio;
    float mx = 0.f, ss = 0.f;
    for (int idx = tid; idx < N * N; idx += T) {
      const int i = idx / N, j = idx % N;
      if (j > i) {
        const float r = rel_off(M[idx], M[i * LD + i], M[j * LD + j], max_diag, nw_ratio);
        mx = fmaxf(mx, r);
        ss = fmaf(fminf(r, 1.f), fminf(r, 1.f), ss);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((tid & 31) == 0 && ss > 0.f) atomicAdd(&mt.sweep_sumsq, ss);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((tid & 31) == 0) redmax[tid >> 5] = mx;
    __syncthreads();
    if (tid < 32) {
      float v = (tid < T / 32) ? redmax[tid] : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
      if (tid == 0) redmax[0] = v;
    }
    __syncthreads();
    mx = redmax[0];
    if (tid == 0) {
      atomicMax(&mt.sweep_off, __float_as_uint(mx));
      mt.pair_skip[local] = (mx < tol) ? 1 : 0;
      if (!(mx < tol) && mt.mode == 3 && active_list) active_list[atomicAdd(active_count, 1)] = blockIdx.x - tc_first;
    }
    if (mx < tol) return;
  }
  const Criteria cr{1, fminf(tol * 0.125f, 1e-6f), __uint_as_float(mt.max_diag), mt.nw_ratio, fast};
  int par = 0;
  for (int sweep = 0; sweep < max_inner; ++sweep) {
    int* fl = flag + (sweep & 1);
    if (tid == 0) flag[(sweep + 1) & 1] = 0;
    if (is_crit) {
      int f = 0;
      rot[par * h + k] = first_rotation<N, LD>(k, cr, M, f);
      if (f) atomicOr(fl, f);
    }
    __syncthreads();
    for (int st = 0; st < N - 1; ++st) {
      const bool more = st < N - 2;
      const Rot* rcur = rot + par * h;
      if (!is_crit) {
        BulkRegs<N, TB> regs;
        bulk_load<N, LD, TB>(tid, rcur, M, W, regs);
        asm volatile("bar.sync 1, %0;" ::"r"(T) : "memory");
        bulk_store<N, LD, TB>(tid, M, W, regs);
      } else {
        CritRegs cregs;
        if (more) crit_load<N, LD>(k, rcur, M, cregs);
        asm volatile("bar.arrive 1, %0;" ::"r"(T) : "memory");
        if (more) {
          int f = 0;
          rot[(par ^ 1) * h + k] = crit_rotation(cr, cregs, f);
          if (f) atomicOr(fl, f);
        }
      }
      __syncthreads();
      par ^= 1;
    }
    const int f = *fl;
    __syncthreads();
    if (!(f & 1) || !(f & 2)) break;
  }
  __syncthreads();
  float* Wg = mt.W + (int64_t)local * N * N;
  if (opts & 2) {   // columns in the order of decreasing new squared norm (see jacobi_smem_kernel SORT)
    __shared__ int sperm[N];
    for (int j = tid; j < N; j += T) {
      const float dj = M[j * LD + j];
      int r = 0;
      for (int k2 = 0; k2 < N; ++k2) { const float dk = M[k2 * LD + k2]; r += (dk > dj) || (dk == dj && k2 < j); }
      sperm[r] = j;
    }
    __syncthreads();
    if (mt.mode == 3) { for (int idx = tid; idx < N * N; idx += T) Wg[idx] = W[(idx % N) * LD + sperm[idx / N]]; }
    else { for (int idx = tid; idx < N * N; idx += T) Wg[idx] = W[(idx / N) * LD + sperm[idx % N]]; }
    return;
  }
  if (mt.mode == 3) { for (int idx = tid; idx < N * N; idx += T) Wg[idx] = W[(idx % N) * LD + idx / N]; }   // W^T
  else { for (int idx = tid; idx < N * N; idx += T) Wg[idx] = W[idx]; }
}

int launch_systolic_block64(EighMat* mats, const int* list, int nblk, int max_inner, int tc_first, int* active_list,
                            int* active_count, int pair_shift, int opts, cudaStream_t s) {
  constexpr int N = 64, TB = 512;
  const size_t smem = ((size_t)2 * N * N + 2 * N /*rot x2*/ + 4 /*flags*/) * sizeof(float);   // 33.3 KB
  jacobi_systolic_block_kernel<N, TB><<<nblk, TB + N / 2, smem, s>>>(mats, list, max_inner, tc_first, active_list,
                                                                    active_count, pair_shift, opts);
  KFAC_CUDA(cudaGetLastError());   // the caller counts the launch
  return KFAC_OK;
}

}  // namespace kfac

// Not declared in include/kfac_b200.h on purpose (experimental, test-only).
// flags: bits 0 and 2 = rotation chain (0: IEEE, 1: sysj::Criteria::fast = 1, 4: fast = 2, half-angle form);
// bit 1 = n > 64 with 480 + 64 threads
// (no register spills, more work per thread) instead of 960 + 64
extern "C" int kfac_experimental_jacobi_systolic(const float* F, int n, int count, float* Q, float* d, int max_sweeps,
                                                 int flags, void* stream) {
  using namespace kfac;
  KFAC_CHECK_ARG(F && Q && d && n > 0 && n <= 128 && count > 0, "jacobi_systolic arguments");
  if (max_sweeps <= 0) max_sweeps = 24;
  cudaStream_t s = (cudaStream_t)stream;
  if (n <= 64) return launch_systolic<64, 512>(F, n, count, Q, d, max_sweeps, (flags & 4) ? 2 : (flags & 1), s);
  if (flags & 2) return launch_systolic<128, 480>(F, n, count, Q, d, max_sweeps, (flags & 4) ? 2 : (flags & 1), s);
  return launch_systolic<128, 960>(F, n, count, Q, d, max_sweeps, (flags & 4) ? 2 : (flags & 1), s);
}

// K7-K12: damped inverse from the eigendecomposition, two-sided Kronecker
// precondition, kl-clip scale and in-place gradient write-back.
#include <string.h>

#include "common.cuh"

namespace kfac {

// grad matrix (g x a) fp32 = [wgrad (g x a-hb) | bgrad]   (modules.py:56-69)
__global__ void gather_grad_kernel(const void* w, const void* b, int dtype, int g, int a, float* out,
                                   int ldo) {
  const int aw = b ? a - 1 : a;
  const int64_t total = (int64_t)g * ldo;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx / ldo;
    const int j = (int)(idx % ldo);
    float v = 0.f;   // padding columns are zero (they feed the K dimension of the next GEMM)
    if (j < aw) v = load_as_float(w, dtype, i * aw + j);
    else if (j < a) v = load_as_float(b, dtype, i);
    out[idx] = v;
  }
}

// vg += sum(P * grad)  (double accumulation, one atomic per block)
__global__ void vg_kernel(const float* P, int ldp, const void* w, const void* b, int dtype, int g, int a,
                          double* vg) {
  const int aw = b ? a - 1 : a;
  const int64_t total = (int64_t)g * a;
  double acc = 0.0;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx / a;
    const int j = (int)(idx % a);
    const float gr = (j < aw) ? load_as_float(w, dtype, i * aw + j) : load_as_float(b, dtype, i);
    acc += (double)P[i * ldp + j] * (double)gr;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += red[k];
    atomicAdd(vg, t);
  }
}

__global__ void nu_kernel(const double* vg, float kl_clip, float lr, float* out) {
  const double v = (*vg) * (double)lr * (double)lr;
  float nu = 1.f;
  if (v != 0.0) nu = (float)fmin(1.0, sqrt((double)kl_clip / fabs(v)));
  *out = nu;
}

__global__ void update_kernel(const float* P, int ldp, void* w, void* b, int dtype, int g, int a,
                              const float* scale) {
  const int aw = b ? a - 1 : a;
  const float s = scale ? *scale : 1.f;
  const int64_t total = (int64_t)g * a;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx / a;
    const int j = (int)(idx % a);
    const float v = s * P[i * ldp + j];
    if (j < aw) store_from_float(w, dtype, i * aw + j, v);
    else store_from_float(b, dtype, i, v);
  }
}

// out[i][j] = Q[i][j] / (d[j] + damping), zero in the padding columns
__global__ void scale_cols_kernel(const float* Q, int ldq, const float* d, int n, float damping, float* out) {
  const int64_t total = (int64_t)n * ldq;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(idx % ldq);
    out[idx] = j < n ? Q[idx] / (d[j] + damping) : 0.f;
  }
}

__global__ void transpose_kernel(const float* src, int lds, float* dst, int ldd, int rows, int cols) {
  __shared__ float tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int i = by + r, j = bx + threadIdx.x;
    tile[r][threadIdx.x] = (i < rows && j < cols) ? src[(int64_t)i * lds + j] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int j = bx + r, i = by + threadIdx.x;    // dst[j][i] = src[i][j]
    if (j < cols && i < rows) dst[(int64_t)j * ldd + i] = tile[threadIdx.x][r];
  }
}

static inline int grid_for(int64_t total) {
  int64_t b = (total + 255) / 256;
  return (int)max((int64_t)1, min(b, (int64_t)148 * 8));
}

struct Epi { int kind = EPI_NONE; const float* E = nullptr; int64_t lde = 0;
             const float* dg = nullptr; const float* da = nullptr; float damping = 0.f; };

// D (M x N, ldd) = epi(A B^T), A (M x K, lda), B (N x K, ldb): tensor-core engine for
// aligned operands with at least one full tile worth of work, SIMT engine otherwise.
// row-major (rows x ld) copy of D into the peers' buffers (layers that did not go through
// the tensor-core epilogue)
__global__ void peer_scatter_kernel(const float* D, int64_t total, float* p0, float* p1, float* p2, float* p3,
                                    float* p4, float* p5, float* p6, int npeer) {
  float* peers[7] = {p0, p1, p2, p3, p4, p5, p6};
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const float v = D[idx];
    for (int q = 0; q < npeer; ++q) peers[q][idx] = v;
  }
}

int gemm_tn(const float* A, int64_t lda, const float* B, int64_t ldb, float* D, int64_t ldd, int M, int N,
            int K, const Epi& e, cudaStream_t s, float* const* peers = nullptr, int npeer = 0) {
  TcGemmArgs t{};
  t.npeer = 0;
  t.A = A; t.lda = lda; t.B = B; t.ldb = ldb; t.D = D; t.ldd = ldd; t.M = M; t.N = N; t.K = K;
  t.kbatch = 1; t.alpha = 1.f; t.splits = 1;
  t.epi = e.kind; t.E = e.E; t.lde = e.lde; t.dg = e.dg; t.da = e.da; t.damping = e.damping;
  if (M >= 64 && N >= 64 && K >= 32 && tc_gemm_supported(t)) {
    if (npeer > 0) {   // fused broadcast: plain-store epilogue, no split-K
      t.npeer = npeer;
      for (int q = 0; q < npeer; ++q) t.peerD[q] = peers[q];
      return launch_tc_gemm(t, s);
    }
    if (e.kind == EPI_NONE && K > 1024) {
      // long reductions: split K across CTAs and add the partial tiles with (round-to-
      // nearest) L2 atomics -- the tensor-core accumulator truncates, so short chains
      // keep the 3xTF32 result at fp32 accuracy, and small tile counts fill the SMs.
      KFAC_CUDA(cudaMemset2DAsync(D, (size_t)ldd * 4, 0, (size_t)N * 4, (size_t)M, s));
      t.atomic = 1;
      t.splits = ceil_div(K, 512);
    }
    return launch_tc_gemm(t, s);
  }
  GemmArgs g{};
  g.A = A; g.sa_m = lda; g.sa_k = 1; g.B = B; g.sb_k = 1; g.sb_n = ldb;
  g.C = D; g.ldc = ldd; g.M = M; g.N = N; g.K = K; g.batch = 1; g.splitk = 1;
  g.alpha = 1.f; g.beta = 0.f;
  g.epi = e.kind; g.E = e.E; g.lde = e.lde; g.dg = e.dg; g.da = e.da; g.damping = e.damping;
  const int rc = launch_gemm(g, s);
  if (rc || npeer == 0) return rc;
  float* pp[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  for (int q = 0; q < npeer; ++q) pp[q] = peers[q];
  const int64_t total = (int64_t)M * ldd;
  peer_scatter_kernel<<<grid_for(total), 256, 0, s>>>(D, total, pp[0], pp[1], pp[2], pp[3], pp[4], pp[5], pp[6], npeer);
  KFAC_LAUNCH_CHECK();
  return KFAC_OK;
}

int gemm_tn_plain(const float* A, int64_t lda, const float* B, int64_t ldb, float* D, int64_t ldd, int M, int N,
                  int K, cudaStream_t s) {
  return gemm_tn(A, lda, B, ldb, D, ldd, M, N, K, Epi{}, s);
}

}  // namespace kfac

using namespace kfac;

static inline int ld4(int x) { return (x + 3) & ~3; }

extern "C" int kfac_inverse_from_eigh(const float* Q, int ldq, const float* d, int n, float damping,
                                      float* inv, int ld_inv, void* ws, size_t ws_bytes, void* stream) {
  KFAC_CHECK_ARG(Q && d && inv && n > 0 && ldq >= n && ld_inv >= n, "inverse args");
  const size_t need = (size_t)n * ldq * sizeof(float);
  if (!ws || ws_bytes < need) {
    set_error("inverse_from_eigh: workspace too small (%zu < %zu)", ws_bytes, need);
    return KFAC_ERR_WORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  float* T = (float*)ws;
  scale_cols_kernel<<<grid_for((int64_t)n * ldq), 256, 0, s>>>(Q, ldq, d, n, damping, T);
  KFAC_LAUNCH_CHECK();
  // inv[i][j] = sum_k T[i][k] Q[j][k]
  return gemm_tn(T, ldq, Q, ldq, inv, ld_inv, n, n, n, Epi{}, s);
}

extern "C" int kfac_transpose(const float* src, int ld_src, float* dst, int ld_dst, int rows, int cols,
                              void* stream) {
  KFAC_CHECK_ARG(src && dst && rows > 0 && cols > 0 && ld_src >= cols && ld_dst >= rows, "transpose args");
  dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32));
  transpose_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(src, ld_src, dst, ld_dst, rows, cols);
  KFAC_LAUNCH_CHECK();
  return KFAC_OK;
}

extern "C" size_t kfac_precondition_workspace_bytes(const kfac_precond_item* items, int count) {
  size_t mx = 0;
  for (int i = 0; i < count; ++i)
    mx = std::max(mx, std::max((size_t)items[i].g * ld4(items[i].a), (size_t)items[i].a * ld4(items[i].g)));
  return align_up(mx * sizeof(float), 256) * 3;
}

extern "C" int kfac_precondition(const kfac_precond_item* items, int count, int method, float damping,
                                 void* ws, size_t ws_bytes, void* stream) {
  KFAC_CHECK_ARG(count >= 0 && (items || count == 0), "items");
  KFAC_CHECK_ARG(method == KFAC_EIGEN || method == KFAC_INVERSE, "method");
  if (count == 0) return KFAC_OK;
  const size_t need = kfac_precondition_workspace_bytes(items, count);
  if (!ws || ws_bytes < need) {
    set_error("precondition: workspace too small (%zu < %zu)", ws_bytes, need);
    return KFAC_ERR_WORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const size_t slab = need / 3;
  float* GR = (float*)ws;                          // grad      g x a   (ld4(a))
  float* T1 = (float*)((char*)ws + slab);          // T^T / U^T a x g   (ld4(g))
  float* T2 = (float*)((char*)ws + 2 * slab);      // V2        g x a   (ld4(a))
  for (int i = 0; i < count; ++i) {
    const kfac_precond_item& it = items[i];
    KFAC_CHECK_ARG(it.wgrad && it.P && it.g > 0 && it.a > 0 && it.ldp >= it.a, "precond item");
    KFAC_CHECK_ARG(it.n_peers >= 0 && it.n_peers <= 7 && (it.n_peers == 0 || it.peer_P), "precond peers");
    const int g = it.g, a = it.a, lga = ld4(a), lag = ld4(g);
    gather_grad_kernel<<<grid_for((int64_t)g * lga), 256, 0, s>>>(it.wgrad, it.bgrad, it.grad_dtype, g, a, GR, lga);
    KFAC_LAUNCH_CHECK();
    int rc;
    if (method == KFAC_EIGEN) {
      if (!(it.qa && it.qg && it.qaT && it.qgT && (it.dgda || (it.da && it.dg)))) {
        set_error("precondition: eigendecompositions for both A and G have not been computed");
        return KFAC_ERR_NOT_READY;
      }
      KFAC_CHECK_ARG(it.ldqa >= a && it.ldqg >= g, "eigenbasis leading dims");
      // (1) T^T[a', r] = sum_k QaT[a', k] grad[r, k]                     (a x g)
      if ((rc = gemm_tn(it.qaT, it.ldqa, GR, lga, T1, lag, a, g, a, Epi{}, s))) return rc;
      // (2) V2[g', a'] = (sum_k QgT[g', k] T^T[a', k]) * dgda[g', a']     (g x a)
      Epi e;
      if (it.dgda) { e.kind = EPI_MUL; e.E = it.dgda; e.lde = it.ld_dgda; }
      else { e.kind = EPI_DIV_OUTER; e.dg = it.dg; e.da = it.da; e.damping = damping; }
      if ((rc = gemm_tn(it.qgT, it.ldqg, T1, lag, T2, lga, g, a, g, e, s))) return rc;
      // (3) U^T[c, g'] = sum_k Qa[c, k] V2[g', k]                         (a x g)
      if ((rc = gemm_tn(it.qa, it.ldqa, T2, lga, T1, lag, a, g, a, Epi{}, s))) return rc;
      // (4) P[r, c] = sum_k Qg[r, k] U^T[c, k]                            (g x a)
      if ((rc = gemm_tn(it.qg, it.ldqg, T1, lag, it.P, it.ldp, g, a, g, Epi{}, s, it.peer_P, it.n_peers))) return rc;
    } else {
      if (!(it.a_inv && it.g_inv)) {
        set_error("precondition: A and G have not been inverted");
        return KFAC_ERR_NOT_READY;
      }
      KFAC_CHECK_ARG(it.ldqa >= a && it.ldqg >= g, "inverse leading dims");
      // the damped inverses are symmetric: T^T = Ainv grad^T, P = Ginv T
      if ((rc = gemm_tn(it.a_inv, it.ldqa, GR, lga, T1, lag, a, g, a, Epi{}, s))) return rc;
      if ((rc = gemm_tn(it.g_inv, it.ldqg, T1, lag, it.P, it.ldp, g, a, g, Epi{}, s, it.peer_P, it.n_peers))) return rc;
    }
  }
  return KFAC_OK;
}

extern "C" int kfac_grad_scale(const kfac_grad_item* items, int count, float kl_clip, float lr,
                               double* scratch, float* scale_out, void* stream) {
  KFAC_CHECK_ARG(count >= 0 && (items || count == 0) && scratch && scale_out, "grad_scale args");
  cudaStream_t s = (cudaStream_t)stream;
  KFAC_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double), s));
  for (int i = 0; i < count; ++i) {
    const kfac_grad_item& it = items[i];
    KFAC_CHECK_ARG(it.P && it.wgrad && it.g > 0 && it.a > 0 && it.ldp >= it.a, "grad item");
    vg_kernel<<<grid_for((int64_t)it.g * it.a), 256, 0, s>>>(it.P, it.ldp, it.wgrad, it.bgrad, it.grad_dtype,
                                                            it.g, it.a, scratch);
  }
  count_launch(count - 1);
  KFAC_LAUNCH_CHECK();
  nu_kernel<<<1, 1, 0, s>>>(scratch, kl_clip, lr, scale_out);
  KFAC_LAUNCH_CHECK();
  return KFAC_OK;
}

extern "C" int kfac_grad_update(const kfac_grad_item* items, int count, const float* scale, void* stream) {
  KFAC_CHECK_ARG(count >= 0 && (items || count == 0), "grad_update args");
  cudaStream_t s = (cudaStream_t)stream;
  for (int i = 0; i < count; ++i) {
    const kfac_grad_item& it = items[i];
    KFAC_CHECK_ARG(it.P && it.wgrad && it.g > 0 && it.a > 0 && it.ldp >= it.a, "grad item");
    update_kernel<<<grid_for((int64_t)it.g * it.a), 256, 0, s>>>(it.P, it.ldp, it.wgrad, it.bgrad, it.grad_dtype,
                                                                it.g, it.a, scale);
  }
  count_launch(count - 1);
  KFAC_LAUNCH_CHECK();
  return KFAC_OK;
}

// ------------------------------------------------------------ peer memory (CUDA IPC)
extern "C" int kfac_peer_alloc(size_t bytes, void** dev_ptr, void* handle64) {
  KFAC_CHECK_ARG(bytes > 0 && dev_ptr && handle64, "peer_alloc args");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
  void* p = nullptr;
  KFAC_CUDA(cudaMalloc(&p, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) { cudaFree(p); set_error("cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); return KFAC_ERR_CUDA; }
  KFAC_CUDA(cudaMemset(p, 0, bytes));
  memcpy(handle64, &h, 64);
  *dev_ptr = p;
  return KFAC_OK;
}
extern "C" int kfac_peer_open(const void* handle64, void** dev_ptr) {
  KFAC_CHECK_ARG(handle64 && dev_ptr, "peer_open args");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  KFAC_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *dev_ptr = p;
  return KFAC_OK;
}
extern "C" int kfac_peer_close(void* dev_ptr) {
  if (dev_ptr) KFAC_CUDA(cudaIpcCloseMemHandle(dev_ptr));
  return KFAC_OK;
}
extern "C" int kfac_peer_free(void* dev_ptr) {
  if (dev_ptr) KFAC_CUDA(cudaFree(dev_ptr));
  return KFAC_OK;
}

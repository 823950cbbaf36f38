// K7-K12: damped inverse from the eigendecomposition, two-sided Kronecker
// precondition, kl-clip scale and in-place gradient write-back.
#include "common.cuh"

namespace kfac {

// grad matrix (g x a) fp32 = [wgrad (g x a-hb) | bgrad]   (modules.py:56-69)
__global__ void gather_grad_kernel(const void* w, const void* b, int dtype, int g, int a, float* out) {
  const int aw = b ? a - 1 : a;
  const int64_t total = (int64_t)g * a;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx / a;
    const int j = (int)(idx % a);
    out[idx] = (j < aw) ? load_as_float(w, dtype, i * aw + j) : load_as_float(b, dtype, i);
  }
}

// vg += sum(P * grad)  (double accumulation, one atomic per block)
__global__ void vg_kernel(const float* P, const void* w, const void* b, int dtype, int g, int a,
                          double* vg) {
  const int aw = b ? a - 1 : a;
  const int64_t total = (int64_t)g * a;
  double acc = 0.0;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx / a;
    const int j = (int)(idx % a);
    const float gr = (j < aw) ? load_as_float(w, dtype, i * aw + j) : load_as_float(b, dtype, i);
    acc += (double)P[idx] * (double)gr;
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

__global__ void update_kernel(const float* P, void* w, void* b, int dtype, int g, int a,
                              const float* scale) {
  const int aw = b ? a - 1 : a;
  const float s = scale ? *scale : 1.f;
  const int64_t total = (int64_t)g * a;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx / a;
    const int j = (int)(idx % a);
    const float v = s * P[idx];
    if (j < aw) store_from_float(w, dtype, i * aw + j, v);
    else store_from_float(b, dtype, i, v);
  }
}

// out[i][j] = Q[i][j] / (d[j] + damping)
__global__ void scale_cols_kernel(const float* Q, const float* d, int n, float damping, float* out) {
  const int64_t total = (int64_t)n * n;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x)
    out[idx] = Q[idx] / (d[idx % n] + damping);
}

static inline int grid_for(int64_t total) {
  int64_t b = (total + 255) / 256;
  return (int)max((int64_t)1, min(b, (int64_t)148 * 8));
}

static GemmArgs mk(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn,
                   float* C, int M, int N, int K) {
  GemmArgs g{};
  g.A = A; g.sa_m = sam; g.sa_k = sak; g.B = B; g.sb_k = sbk; g.sb_n = sbn;
  g.C = C; g.ldc = N; g.M = M; g.N = N; g.K = K; g.batch = 1; g.splitk = 1;
  g.alpha = 1.f; g.beta = 0.f;
  return g;
}

}  // namespace kfac

using namespace kfac;

extern "C" int kfac_inverse_from_eigh(const float* Q, const float* d, int n, float damping, float* inv,
                                      void* ws, size_t ws_bytes, void* stream) {
  KFAC_CHECK_ARG(Q && d && inv && n > 0, "inverse args");
  const size_t need = (size_t)n * n * sizeof(float);
  if (!ws || ws_bytes < need) {
    set_error("inverse_from_eigh: workspace too small (%zu < %zu)", ws_bytes, need);
    return KFAC_ERR_WORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  float* T = (float*)ws;
  scale_cols_kernel<<<grid_for((int64_t)n * n), 256, 0, s>>>(Q, d, n, damping, T);
  KFAC_LAUNCH_CHECK();
  // inv = T Q^T : A(m,k)=T[m*n+k], B(k,j)=Q[j*n+k]
  GemmArgs g = mk(T, n, 1, Q, 1, n, inv, n, n, n);
  return launch_gemm(g, s);
}

extern "C" size_t kfac_precondition_workspace_bytes(const kfac_precond_item* items, int count) {
  size_t mx = 0;
  for (int i = 0; i < count; ++i) mx = std::max(mx, (size_t)items[i].g * items[i].a);
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
  float* GR = (float*)ws;
  float* T1 = (float*)((char*)ws + slab);
  float* T2 = (float*)((char*)ws + 2 * slab);
  for (int i = 0; i < count; ++i) {
    const kfac_precond_item& it = items[i];
    KFAC_CHECK_ARG(it.wgrad && it.P && it.g > 0 && it.a > 0, "precond item");
    const int g = it.g, a = it.a;
    gather_grad_kernel<<<grid_for((int64_t)g * a), 256, 0, s>>>(it.wgrad, it.bgrad, it.grad_dtype, g, a, GR);
    KFAC_LAUNCH_CHECK();
    int rc;
    if (method == KFAC_EIGEN) {
      if (!(it.qa && it.qg && (it.dgda || (it.da && it.dg)))) {
        set_error("precondition: eigendecompositions for both A and G have not been computed");
        return KFAC_ERR_NOT_READY;
      }
      // T1 = grad Qa            (g x a)
      GemmArgs g1 = mk(GR, a, 1, it.qa, a, 1, T1, g, a, a);
      if ((rc = launch_gemm(g1, s))) return rc;
      // T2 = (Qg^T T1) * dgda   A(m,k)=Qg[k*g+m]
      GemmArgs g2 = mk(it.qg, 1, g, T1, a, 1, T2, g, a, g);
      if (it.dgda) { g2.epi = EPI_MUL; g2.E = it.dgda; g2.lde = a; }
      else { g2.epi = EPI_DIV_OUTER; g2.dg = it.dg; g2.da = it.da; g2.damping = damping; }
      if ((rc = launch_gemm(g2, s))) return rc;
      // T1 = T2 Qa^T            B(k,n)=Qa[n*a+k]
      GemmArgs g3 = mk(T2, a, 1, it.qa, 1, a, T1, g, a, a);
      if ((rc = launch_gemm(g3, s))) return rc;
      // P = Qg T1
      GemmArgs g4 = mk(it.qg, g, 1, T1, a, 1, it.P, g, a, g);
      if ((rc = launch_gemm(g4, s))) return rc;
    } else {
      if (!(it.a_inv && it.g_inv)) {
        set_error("precondition: A and G have not been inverted");
        return KFAC_ERR_NOT_READY;
      }
      GemmArgs g1 = mk(GR, a, 1, it.a_inv, a, 1, T1, g, a, a);
      if ((rc = launch_gemm(g1, s))) return rc;
      GemmArgs g2 = mk(it.g_inv, g, 1, T1, a, 1, it.P, g, a, g);
      if ((rc = launch_gemm(g2, s))) return rc;
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
    KFAC_CHECK_ARG(it.P && it.wgrad && it.g > 0 && it.a > 0, "grad item");
    vg_kernel<<<grid_for((int64_t)it.g * it.a), 256, 0, s>>>(it.P, it.wgrad, it.bgrad, it.grad_dtype,
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
    KFAC_CHECK_ARG(it.P && it.wgrad && it.g > 0 && it.a > 0, "grad item");
    update_kernel<<<grid_for((int64_t)it.g * it.a), 256, 0, s>>>(it.P, it.wgrad, it.bgrad, it.grad_dtype,
                                                                it.g, it.a, scale);
  }
  count_launch(count - 1);
  KFAC_LAUNCH_CHECK();
  return KFAC_OK;
}

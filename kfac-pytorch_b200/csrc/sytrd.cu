// Householder tridiagonalisation F = H T H^T of symmetric fp32 matrices (first stage of the direct
// eigensolver, eigh_direct.cuh) as ONE persistent kernel: a matrix is worked on by a group of CTAs
// that meet at two group barriers per column.
//
// Blocked formulation (LAPACK ssytrd/slatrd): inside a panel of TRD_NB columns the trailing matrix is
// NOT updated; the effective matrix is A - V W^T - W V^T with the panel's Householder vectors V and
// w-vectors W.  Per column s (x = effective column s below the diagonal):
//   phase C   beta, tau, v from x (every CTA redundantly; x and |x|^2 partials were published before)
//   phase A   symmetric product with the panel-start matrix: only the lower-triangle 64 x 64 tiles
//             (I >= J) are read -- tile (I,J) yields a partial of y_I (rows) and of y_J (columns) --
//             plus the per-CTA partials of W^T v, V^T v and v^T A v            -> group barrier 1
//   phase B   row owners: y = A v - V (W^T v) - W (V^T v), w = tau y - tau^2/2 (y^T v) v, then the next
//             effective column x' = A[:, s+1] - V W[s+1,:]^T - W V[s+1,:]^T and |x'|^2 partials -> barrier 2
// After TRD_NB columns the lower tiles get the rank-2*NB update A -= V W^T + W V^T (fp32 SIMT: the
// tensor core's 3xTF32 accumulate would cost the tridiagonal ~1e-5 of accuracy) -> one more barrier.
// The matrix (<= 42 MB of lower tiles at n = 4608) stays resident in the 126 MB L2; algorithmic traffic
// is 2/3 n^3 * 4 bytes of L2 reads for the products.  Everything is deterministic (no atomics on data).
#include "eigh_direct.cuh"

namespace kfac {

namespace {

constexpr int NB = TRD_NB, T = TRD_T, CP = TRD_CP;
constexpr int UPAD = T + 4;                 // row length of the transposed update staging
constexpr int SUB_STAGE = 4 * NB * UPAD;    // floats per sub-group: VI^T, WI^T, VJ^T, WJ^T as [NB][UPAD]

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void group_barrier(unsigned* bar, unsigned& epoch, int ncta) {
  __syncthreads();
  if (ncta > 1) {
    if (threadIdx.x == 0) {
      __threadfence();
      epoch += 1;
      atomicAdd(bar, 1u);
      const unsigned target = epoch * (unsigned)ncta;
      while (ld_acquire(bar) < target) {}
      __threadfence();
    }
    __syncthreads();
  }
}

__device__ __forceinline__ void sub_sync(int sg) {
  asm volatile("bar.sync %0, %1;" ::"r"(sg + 1), "r"(256) : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct Ctx {
  const TrdMat* m;
  int cta, ncta;
  // shared memory
  float* vs;        // np
  float* stage;     // 4 * SUB_STAGE  (update staging; phase A reuses it for the column partial sums)
  float* red;       // 32 x 66 cross-warp reduction scratch
  float* sc;        // scalars: [0,32) p1, [32,64) p2, 64 vAv, 65 sigma, 66 tau, 67 beta, 68 ytv, 69 w_{s+1}, 70 dnext
  float* vrow;      // NB  V[s+1][:]
  float* wrow;      // NB  W[s+1][:]
};

// x' for column c with P panel columns; also the diagonal d[c] and the |x'|^2 partial (rows >= c+2)
__device__ void prep_column(const Ctx& cx, int c, int P) {
  const TrdMat& mt = *cx.m;
  const int n = mt.n, np = mt.np;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int GW = cx.ncta * 32, gw = cx.cta * 32 + warp;
  float sig = 0.f;
  int r0 = c + 1;
  r0 += ((gw - r0) % GW + GW) % GW;        // first owned row >= c+1
  for (int r = r0; r < n; r += GW) {
    float t = 0.f;
    if (lane < P) {
      const float vr = __ldcg(&mt.Vp[(int64_t)r * NB + lane]);
      const float wr = __ldcg(&mt.Wp[(int64_t)r * NB + lane]);
      t = vr * cx.wrow[lane] + wr * cx.vrow[lane];
    }
    t = warp_sum(t);
    if (lane == 0) {
      const float x = __ldcg(&mt.A[(int64_t)r * np + c]) - t;
      mt.col[r] = x;
      if (r >= c + 2) sig = fmaf(x, x, sig);
    }
  }
  if (lane == 0) cx.red[warp] = sig;
  __syncthreads();
  if (warp == 0) {
    float t = warp_sum(cx.red[lane]);
    if (lane == 0) mt.cpart[cx.cta * CP + 65] = t;
  }
  if (cx.cta == 0 && warp == 1) {
    float t = 0.f;
    if (lane < P) t = cx.vrow[lane] * cx.wrow[lane];
    t = warp_sum(t);
    if (lane == 0) mt.d[c] = __ldcg(&mt.A[(int64_t)c * np + c]) - 2.f * t;
  }
}

__device__ void tridiagonalise(const TrdMat& mt, int cta, int ncta, float* smem) {
  const int n = mt.n, np = mt.np, nblk = mt.nblk;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int sg = warp >> 3, sw = warp & 7;           // sub-group (tile worker) and warp within it
  const int G = ncta * 4, sgid = cta * 4 + sg;       // tile owners
  const int GW = ncta * 32, gw = cta * 32 + warp;    // row owners
  Ctx cx;
  cx.m = &mt; cx.cta = cta; cx.ncta = ncta;
  cx.vs = smem;
  cx.stage = cx.vs + np;
  cx.red = cx.stage + 4 * SUB_STAGE;
  cx.sc = cx.red + 32 * 66;
  cx.vrow = cx.sc + 80;
  cx.wrow = cx.vrow + NB;
  unsigned epoch = 0;

  if (n == 1) {
    if (cta == 0 && tid == 0) { mt.d[0] = mt.A[0]; mt.tau[0] = 0.f; }
    return;
  }
  for (int i = tid; i < np; i += blockDim.x) cx.vs[i] = 0.f;
  if (tid < NB) { cx.vrow[tid] = 0.f; cx.wrow[tid] = 0.f; }
  __syncthreads();
  int P = 0;                 // columns in the current panel
  bool pending_update = false;
  prep_column(cx, 0, 0);
  group_barrier(mt.bar, epoch, ncta);

  for (int s = 0; s <= n - 2; ++s) {
    // ------------------------------------------------------------ phase C: Householder vector of column s
    if (warp == 0) {
      float t = 0.f;
      for (int c = lane; c < ncta; c += 32) t += __ldcg(&mt.cpart[c * CP + 65]);
      t = warp_sum(t);
      if (lane == 0) {
        const float alpha = __ldcg(&mt.col[s + 1]);
        float beta, tau, scal;
        if (t == 0.f) { beta = alpha; tau = 0.f; scal = 0.f; }
        else {
          beta = -copysignf(sqrtf(fmaf(alpha, alpha, t)), alpha);
          tau = (beta - alpha) / beta;
          scal = 1.f / (alpha - beta);
        }
        cx.sc[66] = tau; cx.sc[67] = beta; cx.sc[71] = scal;
        if (cta == 0) { mt.e[s] = beta; mt.tau[s] = tau; }
      }
    }
    __syncthreads();
    {
      const float scal = cx.sc[71];
      for (int r = tid; r < np; r += blockDim.x) {
        float v = 0.f;
        if (r == s + 1) v = 1.f;
        else if (r > s + 1 && r < n) v = __ldcg(&mt.col[r]) * scal;
        cx.vs[r] = v;
      }
    }
    __syncthreads();
    // the Householder vector is kept for the back-transformation (each CTA writes a slice of the row)
    {
      const int per = (n + ncta - 1) / ncta;
      const int a = cta * per, b = min(n, a + per);
      float* vt = mt.VT + (int64_t)s * mt.ldv;
      for (int r = a + tid; r < b; r += blockDim.x) vt[r] = cx.vs[r];
    }
    // ------------------------------------------------------------ pending rank-2NB update of the lower tiles
    if (pending_update) {
      const int b0 = (s + 1) / T;
      float* st = cx.stage + sg * SUB_STAGE;
      float* VIt = st, *WIt = st + NB * UPAD, *VJt = st + 2 * NB * UPAD, *WJt = st + 3 * NB * UPAD;
      const int st_tid = tid & 255;
      for (int I = b0; I < nblk; ++I) {
        const int tri = (int)(((int64_t)I * (I + 1) / 2) % G);
        for (int J = ((sgid - tri) % G + G) % G; J <= I; J += G) {     // uniform within the sub-group
        if (J < b0) continue;
        sub_sync(sg);
        // stage the four 64 x NB operand blocks transposed ([k][row])
#pragma unroll
        for (int it = 0; it < 2; ++it) {
          const int idx = st_tid + it * 256;          // 0..511: row = idx / 8, k4 = idx % 8
          const int row = idx >> 3, k4 = (idx & 7) * 4;
          const float4 vi = __ldcg(reinterpret_cast<const float4*>(&mt.Vp[(int64_t)(I * T + row) * NB + k4]));
          const float4 wi = __ldcg(reinterpret_cast<const float4*>(&mt.Wp[(int64_t)(I * T + row) * NB + k4]));
          const float4 vj = __ldcg(reinterpret_cast<const float4*>(&mt.Vp[(int64_t)(J * T + row) * NB + k4]));
          const float4 wj = __ldcg(reinterpret_cast<const float4*>(&mt.Wp[(int64_t)(J * T + row) * NB + k4]));
          VIt[(k4 + 0) * UPAD + row] = vi.x; VIt[(k4 + 1) * UPAD + row] = vi.y; VIt[(k4 + 2) * UPAD + row] = vi.z; VIt[(k4 + 3) * UPAD + row] = vi.w;
          WIt[(k4 + 0) * UPAD + row] = wi.x; WIt[(k4 + 1) * UPAD + row] = wi.y; WIt[(k4 + 2) * UPAD + row] = wi.z; WIt[(k4 + 3) * UPAD + row] = wi.w;
          VJt[(k4 + 0) * UPAD + row] = vj.x; VJt[(k4 + 1) * UPAD + row] = vj.y; VJt[(k4 + 2) * UPAD + row] = vj.z; VJt[(k4 + 3) * UPAD + row] = vj.w;
          WJt[(k4 + 0) * UPAD + row] = wj.x; WJt[(k4 + 1) * UPAD + row] = wj.y; WJt[(k4 + 2) * UPAD + row] = wj.z; WJt[(k4 + 3) * UPAD + row] = wj.w;
        }
        sub_sync(sg);
        const int tx = st_tid & 15, ty = st_tid >> 4;
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 8
        for (int k = 0; k < NB; ++k) {
          const float4 a1 = *reinterpret_cast<const float4*>(&VIt[k * UPAD + ty * 4]);
          const float4 b1 = *reinterpret_cast<const float4*>(&WJt[k * UPAD + tx * 4]);
          const float4 a2 = *reinterpret_cast<const float4*>(&WIt[k * UPAD + ty * 4]);
          const float4 b2 = *reinterpret_cast<const float4*>(&VJt[k * UPAD + tx * 4]);
          const float av1[4] = {a1.x, a1.y, a1.z, a1.w}, bv1[4] = {b1.x, b1.y, b1.z, b1.w};
          const float av2[4] = {a2.x, a2.y, a2.z, a2.w}, bv2[4] = {b2.x, b2.y, b2.z, b2.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av1[i], bv1[j], fmaf(av2[i], bv2[j], acc[i][j]));
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float4* p = reinterpret_cast<float4*>(&mt.A[(int64_t)(I * T + ty * 4 + i) * np + J * T + tx * 4]);
          float4 a = __ldcg(p);
          a.x -= acc[i][0]; a.y -= acc[i][1]; a.z -= acc[i][2]; a.w -= acc[i][3];
          *p = a;
        }
        }
      }
      pending_update = false;
      P = 0;
      group_barrier(mt.bar, epoch, ncta);
    }
    // ------------------------------------------------------------ phase A: symmetric product, lower tiles
    {
      const int b0 = (s + 1) / T;
      float vav = 0.f;
      float* cs = cx.stage + sg * (8 * T);          // column partial sums of this sub-group: [8 warps][64]
      for (int I = b0; I < nblk; ++I) {
        const int tri = (int)(((int64_t)I * (I + 1) / 2) % G);
        for (int J = ((sgid - tri) % G + G) % G; J <= I; J += G) {
        if (J < b0) continue;
        const int rb = I * T + sw * 8, c0 = J * T + 2 * lane;
        float2 a[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = __ldcg(reinterpret_cast<const float2*>(&mt.A[(int64_t)(rb + k) * np + c0]));
        const float vj0 = cx.vs[c0], vj1 = cx.vs[c0 + 1];
        float c0acc = 0.f, c1acc = 0.f, vloc = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float vi = cx.vs[rb + k];
          float rs = fmaf(a[k].x, vj0, a[k].y * vj1);
          c0acc = fmaf(a[k].x, vi, c0acc);
          c1acc = fmaf(a[k].y, vi, c1acc);
          rs = warp_sum(rs);
          if (lane == 0) {
            mt.part[(int64_t)J * np + rb + k] = rs;
            vloc = fmaf(rs, vi, vloc);
          }
        }
        vav += (I == J) ? vloc : 2.f * vloc;
        if (I != J) {
          sub_sync(sg);                       // previous tile's readers are done with cs
          cs[sw * T + 2 * lane] = c0acc;
          cs[sw * T + 2 * lane + 1] = c1acc;
          sub_sync(sg);
          const int st_tid = tid & 255;
          if (st_tid < T) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) t += cs[w * T + st_tid];
            mt.part[(int64_t)I * np + J * T + st_tid] = t;
          }
        }
        }
      }
      // per-CTA partials of p1 = W^T v, p2 = V^T v (lane = panel column) over the owned rows
      float p1 = 0.f, p2 = 0.f;
      if (P > 0) {
        int r0 = s + 1;
        r0 += ((gw - r0) % GW + GW) % GW;
        for (int r = r0; r < n; r += GW) {
          if (lane < P) {
            const float v = cx.vs[r];
            p1 = fmaf(__ldcg(&mt.Wp[(int64_t)r * NB + lane]), v, p1);
            p2 = fmaf(__ldcg(&mt.Vp[(int64_t)r * NB + lane]), v, p2);
          }
        }
      }
      __syncthreads();                        // all sub-groups are done with `stage`
      cx.red[warp * 66 + lane] = p1;
      cx.red[warp * 66 + 32 + lane] = p2;
      if (lane == 0) cx.red[warp * 66 + 64] = vav;
      __syncthreads();
      if (tid < 65) {
        float t = 0.f;
#pragma unroll 8
        for (int w = 0; w < 32; ++w) t += cx.red[w * 66 + tid];
        mt.cpart[cta * CP + tid] = t;
      }
    }
    group_barrier(mt.bar, epoch, ncta);
    // ------------------------------------------------------------ phase B: w column, next effective column
    {
      // reduce the per-CTA partials (outputs 0..64), every CTA redundantly
      for (int o = warp; o < 65; o += 32) {
        float t = 0.f;
        const bool need = (o == 64) || ((o & 31) < P);
        if (need)
          for (int c = lane; c < ncta; c += 32) t += __ldcg(&mt.cpart[c * CP + o]);
        t = warp_sum(t);
        if (lane == 0) cx.sc[o] = t;
      }
      __syncthreads();
      const float tau = cx.sc[66];
      const int b0 = (s + 1) / T;
      if (warp == 0) {
        float t = (lane < P) ? cx.sc[lane] * cx.sc[32 + lane] : 0.f;
        t = warp_sum(t);
        const float ytv = cx.sc[64] - 2.f * t;
        // row s+1: y, w and the panel rows V[s+1][:], W[s+1][:]
        const int r = s + 1;
        float yr = 0.f;
        for (int X = b0 + lane; X < nblk; X += 32) yr += __ldcg(&mt.part[(int64_t)X * np + r]);
        float vr = 0.f, wr = 0.f, corr = 0.f;
        if (lane < P) {
          vr = __ldcg(&mt.Vp[(int64_t)r * NB + lane]);
          wr = __ldcg(&mt.Wp[(int64_t)r * NB + lane]);
          corr = vr * cx.sc[lane] + wr * cx.sc[32 + lane];
        }
        yr = warp_sum(yr) - warp_sum(corr);
        const float w1 = tau * (yr - 0.5f * tau * ytv);          // v[s+1] = 1
        if (lane == P) { vr = 1.f; wr = w1; }
        cx.vrow[lane] = vr; cx.wrow[lane] = wr;
        if (lane == 0) cx.sc[68] = ytv;
      }
      __syncthreads();
      const float ytv = cx.sc[68];
      int r0 = s + 1;
      r0 += ((gw - r0) % GW + GW) % GW;
      for (int r = r0; r < n; r += GW) {
        float yr = 0.f;
        for (int X = b0 + lane; X < nblk; X += 32) yr += __ldcg(&mt.part[(int64_t)X * np + r]);
        float corr = 0.f;
        if (lane < P)
          corr = __ldcg(&mt.Vp[(int64_t)r * NB + lane]) * cx.sc[lane] + __ldcg(&mt.Wp[(int64_t)r * NB + lane]) * cx.sc[32 + lane];
        yr = warp_sum(yr) - warp_sum(corr);
        if (lane == 0) {
          const float v = cx.vs[r];
          mt.Wp[(int64_t)r * NB + P] = tau * (yr - 0.5f * tau * ytv * v);
          mt.Vp[(int64_t)r * NB + P] = v;
        }
      }
      __syncthreads();        // own rows' panel column P is written (re-read below by the same warps)
      prep_column(cx, s + 1, P + 1);
      P += 1;
      if (P == NB) pending_update = true;
    }
    group_barrier(mt.bar, epoch, ncta);
  }
  if (cta == 0 && tid == 0) { mt.e[n - 1] = 0.f; mt.tau[n - 1] = 0.f; }
}

__global__ void __launch_bounds__(TRD_THREADS, 1) sytrd_kernel(const TrdMat* mats, const TrdJob* jobs, int njobs) {
  extern __shared__ __align__(16) float trd_smem[];
  for (int j = 0; j < njobs; ++j) {
    const TrdJob job = jobs[j];
    const int b = (int)blockIdx.x;
    if (b < job.cta0 || b >= job.cta0 + job.ncta) continue;
    tridiagonalise(mats[job.mat], b - job.cta0, job.ncta, trd_smem);
    __syncthreads();
  }
}

size_t trd_smem_bytes(int np_max) {
  return sizeof(float) * ((size_t)np_max + 4 * SUB_STAGE + 32 * 66 + 80 + 2 * NB + 16);
}

}  // namespace

int sytrd_max_grid() { return tc_num_sms(); }

int launch_sytrd(const TrdMat* d_mats, const TrdJob* d_jobs, int njobs, int np_max, int grid, cudaStream_t s) {
  if (njobs <= 0) return KFAC_OK;
  const size_t smem = trd_smem_bytes(np_max);
  if (smem > 227 * 1024) { set_error("sytrd: matrix too large for the shared-memory vector (np = %d)", np_max); return KFAC_ERR_UNSUPPORTED; }
  static size_t attr_set = 0;
  if (smem > attr_set) {
    KFAC_CUDA(cudaFuncSetAttribute(sytrd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = smem;
  }
  void* args[] = {(void*)&d_mats, (void*)&d_jobs, (void*)&njobs};
  // cooperative launch: every CTA of the grid is resident (the group barriers spin)
  KFAC_CUDA(cudaLaunchCooperativeKernel((const void*)sytrd_kernel, dim3(grid), dim3(TRD_THREADS), args, smem, s));
  count_launch(1);
  return KFAC_OK;
}

}  // namespace kfac

// K5: batched symmetric (PSD) eigendecomposition by Jacobi rotations.
//
// Size classes
//   n <= 64 / n <= 128 : the whole matrix lives in shared memory of ONE CTA;
//                        two-sided cyclic Jacobi with a round-robin ordering
//                        (N/2 disjoint rotations per step, N-1 steps / sweep).
//   n  > 128           : block ONE-SIDED Jacobi (Hestenes) on G = F V:
//                        per round, for every disjoint pair of 32-column blocks
//                        (I,J): M = [G_I G_J]^T [G_I G_J] (64x64 Gram, GEMM) ->
//                        W = eigvecs(M) (shared-memory Jacobi, as above) ->
//                        [G_I G_J] <- [G_I G_J] W, [V_I V_J] <- [V_I V_J] W
//                        (GEMM).  At convergence the columns of G are
//                        orthogonal: lambda_j = |g_j|, Q = V.
// All matrices of a batch advance through the rounds concurrently (one launch
// per phase for the whole batch); per-matrix `done` flags computed on the
// device make converged matrices drop out without a host sync.
#include "common.cuh"
#include "tc_pipeline.cuh"
#include "eigh_common.cuh"
#include "eigh_direct.cuh"

#include <stdlib.h>

#include <algorithm>
#include <thread>
#include <vector>

namespace kfac {

constexpr int JB = 32;        // block width
constexpr int JP = 2 * JB;    // pair width
constexpr int GR = 256;       // rows of G handled per CTA in gram/apply

constexpr int TC_MIN_N = 768;
// 1: the Gram reads G itself as an MN-major operand (no transposed copy of G is kept);
// 0: the Gram reads a K-major G^T that the apply kernel refreshes.
// (runtime switch: env KFAC_GRAM_MNMAJOR=1 selects the experimental MN-major variant)
static bool gram_mn_major() {
  static int v = -1;
  if (v < 0) v = getenv("KFAC_GRAM_MNMAJOR") ? 1 : 0;   // experimental (descriptor not validated): off by default
  return v != 0;
} // block matrices at least this large use the tcgen05 Gram/apply kernels

// Tensor-core class with 64-column blocks (128-wide pairs): half the rounds per sweep, one pair per
// 128-row Gram tile (no discarded quadrants), 128x128 pair problems in the shared-memory Jacobi.
static bool eigh_wide() {
  static int v = -1;
  if (v < 0) v = getenv("KFAC_EIGH_WIDE") ? atoi(getenv("KFAC_EIGH_WIDE")) : 0;
  return v != 0;
}

// ------------------------------------------------------------ init / final
__global__ void eigh_init_kernel(EighMat* mats, const int* block_list) {
  EighMat& mt = mats[block_list[blockIdx.y]];
  const int np = mt.np, n = mt.n;
  const int64_t total = (int64_t)np * np;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx / np), j = (int)(idx % np);
    const bool in = (i < n && j < n);
    if (mt.V0T) {
      // warm start: V = V0; G = F V0 (and G^T) are produced by GEMMs right after this kernel
      mt.V[idx] = in ? mt.V0T[(int64_t)j * mt.ldq + i] : (i == j ? 1.f : 0.f);
      if (!in) { mt.G[idx] = 0.f; if (mt.Gt) mt.Gt[idx] = 0.f; }
    } else {
      mt.G[idx] = in ? mt.F[(int64_t)i * n + j] : 0.f;
      if (mt.Gt) mt.Gt[idx] = in ? mt.F[(int64_t)j * n + i] : 0.f;
      mt.V[idx] = (i == j) ? 1.f : 0.f;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) { mt.sweep_off = 0u; mt.done = 0; mt.sweeps = 0; mt.prev_off = 1e30f; mt.max_diag = 0u; mt.sweep_sumsq = 0.f; }
}

__global__ void eigh_final_kernel(EighMat* mats, const int* block_list) {
  EighMat& mt = mats[block_list[blockIdx.y]];
  const int np = mt.np, n = mt.n;
  __shared__ float red[8][33];
  __shared__ float redv[8][33];
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;  // 32 x 8
  for (int j0 = blockIdx.x * 32; j0 < n; j0 += gridDim.x * 32) {
    const int j = j0 + tx;
    // G = F V holds column-wise even if rounding let |v_j| drift from 1:
    // lambda_j = |g_j| / |v_j|, q_j = v_j / |v_j|.
    float ss = 0.f, vv = 0.f;
    for (int i = ty; i < n; i += 8) {
      if (j < n) {
        const float g = mt.G[(int64_t)i * np + j];
        const float v = mt.V[(int64_t)i * np + j];
        ss = fmaf(g, g, ss);
        vv = fmaf(v, v, vv);
      }
    }
    red[ty][tx] = ss;
    redv[ty][tx] = vv;
    __syncthreads();
    float tg = 0.f, tv = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { tg += red[k][tx]; tv += redv[k][tx]; }
    const float inv_v = tv > 0.f ? 1.f / sqrtf(tv) : 0.f;
    if (ty == 0 && j < n) mt.d[j] = sqrtf(tg) * inv_v;
    for (int i = ty; i < n; i += 8)
      if (j < n) {
        const float q = mt.V[(int64_t)i * np + j] * inv_v;
        mt.Q[(int64_t)i * mt.ldq + j] = q;
        if (mt.QT) mt.QT[(int64_t)j * mt.ldq + i] = q;
      }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ gram
__global__ void __launch_bounds__(256) eigh_gram_kernel(EighMat* mats, const int* pair_mat, int round) {
  const int p = blockIdx.x;
  EighMat& mt = mats[pair_mat[p]];
  if (mt.done) return;
  const int np = mt.np, n = mt.n;
  const int r0c = blockIdx.y * GR;
  if (r0c >= n) return;
  const int r1c = min(n, r0c + GR);
  const int local = p - mt.pair_base;
  int I, J;
  tournament(round % (mt.nb - 1), local, mt.nb, I, J);
  __shared__ __align__(16) float Xs[16][JP + 4];
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const float* __restrict__ G = mt.G;
  for (int r0 = r0c; r0 < r1c; r0 += 16) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = tid % 64, r = tid / 64 + 4 * i;
      const int col = (c < JB) ? (I * JB + c) : (J * JB + c - JB);
      Xs[r][c] = (r0 + r < r1c) ? G[(int64_t)(r0 + r) * np + col] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&Xs[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Xs[k][tx * 4]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w};
      const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* M = mt.M + (int64_t)local * JP * JP;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) atomicAdd(&M[(ty * 4 + i) * JP + tx * 4 + j], acc[i][j]);
}

// --------------------------------------------------- shared-memory Jacobi
// mode_block = 1: blockIdx.x = global pair index, M read from (and zeroed in)
//                 the pair buffer, W written to the pair buffer.
// mode_block = 0: blockIdx.x indexes `list`; the matrix itself is solved,
//                 Q and d written directly.
// OPT (compile-time, experimental, selected by KFAC_EIGH_JOPT; 0 = the validated default):
//   bit 0  rotation parameters from MUFU rsqrt/rcp + one Newton step instead of IEEE div/sqrt chains
//   bit 1  block mode: the first inner sweep visits only the N/2 x N/2 cross pairs (the two diagonal
//          blocks of a pair Gram are already diagonal from earlier rounds) -- N/2 steps instead of N-1
//   bit 2  fewer shared-memory instructions per step (the kernel is bound by them: ~56 LDS/STS per
//          thread and step): (c, s, p, q) of a rotation packed into one float4, and only the upper
//          triangle of M is kept up to date (block pairs a <= b)
//   SORT   block mode: the columns of W are written in the order of decreasing new squared column norm
//          (de Rijk's ordering carried through the pairs: the block solver's columns become sorted by
//          eigenvalue, which brings the quadratic phase forward -- numpy model: 7-9 -> 5-7 sweeps)
template <int N, int OPT = 0, bool SORT = false>
__global__ void __launch_bounds__(N * 8) jacobi_smem_kernel(EighMat* mats, const int* list,
                                                           int mode_block, int max_inner,
                                                           int tc_first = 0, int* active_list = nullptr,
                                                           int* active_count = nullptr, int pair_shift = 0) {
  extern __shared__ float sm[];
  float (*M)[N + 1] = reinterpret_cast<float (*)[N + 1]>(sm);
  float (*W)[N + 1] = reinterpret_cast<float (*)[N + 1]>(sm + N * (N + 1));
  float* cs = sm + 2 * N * (N + 1);       // c[N/2], s[N/2]
  __shared__ float redmax[32];
  __shared__ int sh_big;
  __shared__ int sh_within;
  __shared__ float4 rot4[(OPT & 4) ? N / 2 : 1];
  constexpr int T = N * 8;   // 512 / 1024 threads: the 2x2-block update is latency bound, more threads = fewer serial LDS/STS
  const int tid = threadIdx.x;
  EighMat& mt = mats[list[blockIdx.x]];
  int local = 0;
  float* Mg = nullptr;
  if (mode_block) {
    if (mt.done) return;
    local = blockIdx.x + pair_shift - mt.inner_base;
    Mg = mt.M + (int64_t)local * N * N;
    for (int idx = tid; idx < N * N; idx += T) {
      const int i = idx / N, j = idx % N;
      M[i][j] = Mg[idx];
      Mg[idx] = 0.f;
      W[i][j] = (i == j) ? 1.f : 0.f;
    }
  } else {
    const int n = mt.n;
    for (int idx = tid; idx < N * N; idx += T) {
      const int i = idx / N, j = idx % N;
      M[i][j] = (i < n && j < n) ? mt.F[(int64_t)i * n + j] : 0.f;
      W[i][j] = (i == j) ? 1.f : 0.f;
    }
  }
  __syncthreads();
  const float tol = mt.tol;
  if (mode_block) {
    // largest relative off-diagonal of this pair's Gram (convergence measure)
    float dmax = 0.f;
    for (int j = tid; j < N; j += T) dmax = fmaxf(dmax, M[j][j]);
    if (dmax > 0.f) atomicMax(&mt.max_diag, __float_as_uint(dmax));
    const float max_diag = __uint_as_float(mt.max_diag);   // running maximum over all rounds
    const float nw_ratio = mt.nw_ratio;
    float mx = 0.f, ss = 0.f;
    int within = 0;   // OPT bit 1: some pair INSIDE one of the two blocks is above tol
    for (int idx = tid; idx < N * N; idx += T) {
      const int i = idx / N, j = idx % N;
      if (j > i) {
        const float r = rel_off(M[i][j], M[i][i], M[j][j], max_diag, nw_ratio);
        mx = fmaxf(mx, r);
        ss = fmaf(fminf(r, 1.f), fminf(r, 1.f), ss);
        if ((OPT & 2) && (i < N / 2) == (j < N / 2) && !(r < mt.tol)) within = 1;
      }
    }
    if (OPT & 2) { within = __syncthreads_or(within); if (tid == 0) sh_within = within; }
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
      // compact list of tensor-core-class pairs that will be applied this round
      if (!(mx < tol) && mt.mode == 3 && active_list) active_list[atomicAdd(active_count, 1)] = blockIdx.x - tc_first;
    }
    if (mx < tol) return;
  }
  // block mode: the leftover below this threshold is what limits the final accuracy
  // (0.5*tol was measured to double/triple the error of the damped inverse on graded spectra)
  const float tol_in = mode_block ? fminf(tol * 0.125f, 1e-6f) : 1e-7f;
  const float blk_max_diag = mode_block ? __uint_as_float(mt.max_diag) : 0.f;
  const float blk_nw_ratio = mode_block ? mt.nw_ratio : 0.f;

  int* pq = reinterpret_cast<int*>(cs + N);   // (p,q) of every pair of the current step
  if (OPT & 2) __syncthreads();
  const bool bip_first = (OPT & 2) && mode_block && !sh_within;
  for (int sweep = 0; sweep < max_inner; ++sweep) {   // OPT bit 1: the bipartite sweep counts as one of them
    int rotated_sweep = 0;
    if (tid == 0) sh_big = 0;   // set when some |sin| >= 2e-3 in this sweep
    __syncthreads();
    const bool bip = bip_first && sweep == 0;
    for (int st = 0; st < (bip ? N / 2 : N - 1); ++st) {
      int rotated = 0;
      if (tid < N / 2) {
        int p, q;
        if (bip) { p = tid; q = N / 2 + (tid + st) % (N / 2); }
        else tournament(st, tid, N, p, q);
        const float apq = M[p][q], app = M[p][p], aqq = M[q][q];
        float c = 1.f, s = 0.f;
        const float thr = mode_block ? pair_den(app, aqq, blk_max_diag, blk_nw_ratio) : sqrtf(fabsf(app * aqq));
        if (fabsf(apq) > tol_in * thr) {
          if (OPT & 1) {
            // t = sign(a b) |b| / (|a| + sqrt(a^2 + b^2)), a = aqq - app, b = 2 apq; operands scaled
            // by 1/max(|a|,|b|) so the squares cannot overflow; rsqrt + one Newton step is
            // accurate to ~1 ulp without the bias of the raw MUFU value
            float a = aqq - app, b = 2.f * apq;
            const float sc = __frcp_rn(fmaxf(fabsf(a), fabsf(b)));
            a *= sc; b *= sc;
            const float x = fmaf(a, a, b * b);
            float r = rsqrtf(x);
            r = r * fmaf(-0.5f * x * r, r, 1.5f);
            const float t = copysignf(fabsf(b), (a < 0.f) != (b < 0.f) ? -1.f : 1.f) * __frcp_rn(fabsf(a) + x * r);
            const float y = fmaf(t, t, 1.f);
            float rc = rsqrtf(y);
            rc = rc * fmaf(-0.5f * y * rc, rc, 1.5f);
            c = rc;
            s = t * c;
          } else {
            const float tau = (aqq - app) / (2.f * apq);
            const float t = copysignf(1.f, tau) / (fabsf(tau) + sqrtf(1.f + tau * tau));
            c = 1.f / sqrtf(1.f + t * t);   // IEEE sqrt/div: rsqrtf's bias makes column norms drift
            s = t * c;
          }
          if (s != 0.f) { rotated = 1; if (fabsf(s) >= 2e-3f) sh_big = 1; }
        }
        if (OPT & 4) rot4[tid] = make_float4(c, s, __int_as_float(p), __int_as_float(q));
        else {
          cs[tid] = c;
          cs[N / 2 + tid] = s;
          pq[tid] = p | (q << 16);
        }
      }
      // barrier + "did anybody rotate": a step without rotations is skipped entirely
      if (!__syncthreads_or(rotated)) continue;
      rotated_sweep = 1;
      if (OPT & 4) {
        auto MS = [&](int i, int j) -> float& { return i <= j ? M[i][j] : M[j][i]; };
        for (int idx = tid; idx < (N / 2) * (N / 2); idx += T) {
          const int b = idx % (N / 2), a = idx / (N / 2);
          if (a > b) continue;                       // the mirror block is never read
          const float4 ra = rot4[a], rb = rot4[b];
          const float ca = ra.x, sa = ra.y, cb = rb.x, sb = rb.y;
          if (sa == 0.f && sb == 0.f) continue;
          const int pa = __float_as_int(ra.z), qa = __float_as_int(ra.w);
          const int pb = __float_as_int(rb.z), qb = __float_as_int(rb.w);
          if (a == b) {                              // diagonal 2x2 block: symmetric, 3 entries
            const float m00 = M[pa][pa], m01 = M[pa][qa], m11 = M[qa][qa];
            const float n00 = ca * m00 - sa * m01, n01 = sa * m00 + ca * m01;
            const float n10 = ca * m01 - sa * m11, n11 = sa * m01 + ca * m11;
            M[pa][pa] = ca * n00 - sa * n10;
            M[pa][qa] = ca * n01 - sa * n11;
            M[qa][qa] = sa * n01 + ca * n11;
          } else {
            float &r00 = MS(pa, pb), &r01 = MS(pa, qb), &r10 = MS(qa, pb), &r11 = MS(qa, qb);
            const float m00 = r00, m01 = r01, m10 = r10, m11 = r11;
            const float n00 = cb * m00 - sb * m01, n01 = sb * m00 + cb * m01;
            const float n10 = cb * m10 - sb * m11, n11 = sb * m10 + cb * m11;
            r00 = ca * n00 - sa * n10;
            r01 = ca * n01 - sa * n11;
            r10 = sa * n00 + ca * n10;
            r11 = sa * n01 + ca * n11;
          }
        }
        for (int idx = tid; idx < N * (N / 2); idx += T) {
          const int i = idx % N, k = idx / N;
          const float4 r = rot4[k];
          if (r.y == 0.f) continue;
          const int p = __float_as_int(r.z), q = __float_as_int(r.w);
          const float u = W[i][p], v = W[i][q];
          W[i][p] = r.x * u - r.y * v;
          W[i][q] = r.y * u + r.x * v;
        }
      } else {
      // M <- J^T M J on independent 2x2 blocks (pair a rows) x (pair b columns)
      for (int idx = tid; idx < (N / 2) * (N / 2); idx += T) {
        const int b = idx % (N / 2), a = idx / (N / 2);
        const float ca = cs[a], sa = cs[N / 2 + a], cb = cs[b], sb = cs[N / 2 + b];
        if (sa == 0.f && sb == 0.f) continue;
        const int pa = pq[a] & 0xffff, qa = pq[a] >> 16, pb = pq[b] & 0xffff, qb = pq[b] >> 16;
        const float m00 = M[pa][pb], m01 = M[pa][qb], m10 = M[qa][pb], m11 = M[qa][qb];
        const float n00 = cb * m00 - sb * m01, n01 = sb * m00 + cb * m01;
        const float n10 = cb * m10 - sb * m11, n11 = sb * m10 + cb * m11;
        M[pa][pb] = ca * n00 - sa * n10;
        M[pa][qb] = ca * n01 - sa * n11;
        M[qa][pb] = sa * n00 + ca * n10;
        M[qa][qb] = sa * n01 + ca * n11;
      }
      // W <- W J
      for (int idx = tid; idx < N * (N / 2); idx += T) {
        const int i = idx % N, k = idx / N;
        const float c = cs[k], s = cs[N / 2 + k];
        if (s == 0.f) continue;
        const int p = pq[k] & 0xffff, q = pq[k] >> 16;
        const float u = W[i][p], v = W[i][q];
        W[i][p] = c * u - s * v;
        W[i][q] = s * u + c * v;
      }
      }
      __syncthreads();
    }
    // all rotations tiny: the next sweep would only find second-order leftovers
    const int big = sh_big;
    __syncthreads();
    if (!rotated_sweep || !big) break;
  }

  if (mode_block) {
    float* Wg = mt.W + (int64_t)local * N * N;
    if (SORT) {
      // perm[r] = column with the r-th largest rotated diagonal entry (ties by index)
      __shared__ int sperm[N];
      for (int j = tid; j < N; j += T) {
        const float dj = M[j][j];
        int r = 0;
        for (int k2 = 0; k2 < N; ++k2) { const float dk = M[k2][k2]; r += (dk > dj) || (dk == dj && k2 < j); }
        sperm[r] = j;
      }
      __syncthreads();
      if (mt.mode == 3) { for (int idx = tid; idx < N * N; idx += T) Wg[idx] = W[idx % N][sperm[idx / N]]; }   // W^T
      else { for (int idx = tid; idx < N * N; idx += T) Wg[idx] = W[idx / N][sperm[idx % N]]; }
    } else
    if (mt.mode == 3) { for (int idx = tid; idx < N * N; idx += T) Wg[idx] = W[idx % N][idx / N]; }   // W^T
    else { for (int idx = tid; idx < N * N; idx += T) Wg[idx] = W[idx / N][idx % N]; }
  } else {
    const int n = mt.n;
    // M = W^T F W: rescale by the (rounding-drifted) column norms of W
    for (int j = tid; j < N; j += T) {
      float ww = 0.f;
      for (int i = 0; i < N; ++i) ww = fmaf(W[i][j], W[i][j], ww);
      cs[j] = ww > 0.f ? 1.f / sqrtf(ww) : 0.f;
    }
    __syncthreads();
    for (int idx = tid; idx < n * n; idx += T) {
      const int i = idx / n, j = idx % n;
      const float q = W[i][j] * cs[j];
      mt.Q[(int64_t)i * mt.ldq + j] = q;
      if (mt.QT) mt.QT[(int64_t)j * mt.ldq + i] = q;
    }
    for (int j = tid; j < n; j += T) mt.d[j] = fmaxf(M[j][j] * cs[j] * cs[j], 0.f);
  }
}

// ------------------------------------------------------------------ apply
__global__ void __launch_bounds__(256) eigh_apply_kernel(EighMat* mats, const int* pair_mat, int round) {
  const int p = blockIdx.x;
  EighMat& mt = mats[pair_mat[p]];
  if (mt.done) return;
  const int local = p - mt.pair_base;
  if (mt.pair_skip[local]) return;
  const int np = mt.np, n = mt.n;
  const int r0c = blockIdx.y * GR;
  if (r0c >= n) return;
  const int r1c = min(n, r0c + GR);
  int I, J;
  tournament(round % (mt.nb - 1), local, mt.nb, I, J);
  float* X = blockIdx.z ? mt.V : mt.G;
  __shared__ __align__(16) float Ws[JP][JP + 4];
  __shared__ float Xs[64][JP + 1];
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const float* Wg = mt.W + (int64_t)local * JP * JP;
  for (int idx = tid; idx < JP * JP; idx += 256) Ws[idx / JP][idx % JP] = Wg[idx];
  for (int r0 = r0c; r0 < r1c; r0 += 64) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int c = tid % 64, r = tid / 64 + 4 * i;
      const int col = (c < JB) ? (I * JB + c) : (J * JB + c - JB);
      Xs[r][c] = (r0 + r < r1c) ? X[(int64_t)(r0 + r) * np + col] : 0.f;
    }
    __syncthreads();
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 8
    for (int k = 0; k < JP; ++k) {
      const float4 b4 = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
      const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float a = Xs[ty * 4 + i][k];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a, bv[j], acc[i][j]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = r0 + ty * 4 + i;
      if (r >= r1c) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = tx * 4 + j;
        const int col = (c < JB) ? (I * JB + c) : (J * JB + c - JB);
        X[(int64_t)r * np + col] = acc[i][j];
      }
    }
  }
}

__global__ void eigh_ctl_kernel(EighMat* mats, const int* block_list, int nblock, int round, int* all_done,
                                int* active_count) {
  if (threadIdx.x == 0) *active_count = 0;
  int pending = 0;
  for (int i = threadIdx.x; i < nblock; i += blockDim.x) {
    EighMat& mt = mats[block_list[i]];
    if (mt.done) continue;
    if ((round + 1) % (mt.nb - 1) == 0) {
      mt.sweeps += 1;
      const float off = __uint_as_float(mt.sweep_off);
      // root-mean-square contamination per eigenvector: the error of the preconditioned
      // gradient follows this, not the single worst pair.  (Within-block pairs are seen in
      // every round of the sweep: divide their share out approximately with nb - 1.)
      const float rms = sqrtf(2.f * mt.sweep_sumsq / (float)mt.n);
      if (mt.sweeps <= 12) { mt.off_hist[mt.sweeps - 1] = off; mt.rms_hist[mt.sweeps - 1] = rms; }
      if (rms < mt.rms_tol) mt.done = 1;
      mt.sweep_sumsq = 0.f;
      // `off` is measured BEFORE this sweep's rotations; every pair above tol has just been
      // re-diagonalised, so a sweep that started below conv_tol ends at the rounding floor
      // (quadratic convergence) -- no verification sweep needed.
      if (off < mt.conv_tol) mt.done = 1;
      // safety net: stalled at the rounding floor (no longer shrinking, already small)
      else if (mt.sweeps >= 4 && off < 10.f * mt.tol && off > 0.9f * mt.prev_off) mt.done = 1;
      mt.prev_off = off;
      mt.sweep_off = 0u;
    }
    if (!mt.done) pending = 1;
  }
  pending = __syncthreads_or(pending);
  if (threadIdx.x == 0) *all_done = pending ? 0 : 1;
}


// ------------------------------------------------ tcgen05 Gram / apply (mode 3)
// Gram: two block pairs share one 128-row MMA tile.  A = B = rows {I1,J1,I2,J2} of
// G^T (K-major, K = row index of G), D[128x128] = A A^T; the two diagonal 64x64
// quadrants are the pair Grams (the off-diagonal quadrants are discarded).
struct GramParams { EighMat* mats; const int* item_mat; int ksplits; int round; };
template <bool MN>
struct GramPolicyT {
  using Params = GramParams;
  struct Item { EighMat* mt; int p0, I1, J1, I2, J2, kb0, kb1; };
  static constexpr int BN = 128;
  static constexpr bool B_IS_A = true;
  static constexpr bool MN_MAJOR = MN;
  static constexpr int CHUNK = 1;
  __device__ static void reset(Item&) {}
  __device__ static int total_work(const Params&, int t) { return t; }
  static constexpr uint32_t TX_BYTES = tc::PTILE;
  __device__ static bool decode(const Params& p, int w, Item& it) {
    const int gi = w / p.ksplits, sp = w % p.ksplits;
    EighMat* mt = &p.mats[p.item_mat[gi]];
    if (mt->done) return false;
    it.mt = mt;
    it.p0 = 2 * (gi - mt->gram_base);
    const int r = p.round % (mt->nb - 1);
    tournament(r, it.p0, mt->nb, it.I1, it.J1);
    tournament(r, it.p0 + 1, mt->nb, it.I2, it.J2);
    const int kb_total = (mt->n + 31) / 32;
    const int per = (kb_total + p.ksplits - 1) / p.ksplits;
    it.kb0 = sp * per; it.kb1 = min(kb_total, it.kb0 + per);
    return it.kb1 > it.kb0;
  }
  __device__ static int num_kb(const Params&, const Item& it) { return it.kb1 - it.kb0; }
  __device__ static void load(const Params&, const Item& it, int kbi, uint8_t* a, uint8_t*, uint64_t* bar) {
    const int kc = (it.kb0 + kbi) * 32;
    if (MN_MAJOR) {   // boxes {32 columns of a block, 32 rows kc..kc+31} of G
      tc::tma_load_3d(a, &it.mt->tmGt, bar, it.I1 * JB, kc, 0);
      tc::tma_load_3d(a + 4096, &it.mt->tmGt, bar, it.J1 * JB, kc, 0);
      tc::tma_load_3d(a + 8192, &it.mt->tmGt, bar, it.I2 * JB, kc, 0);
      tc::tma_load_3d(a + 12288, &it.mt->tmGt, bar, it.J2 * JB, kc, 0);
    } else {          // boxes {32 reduction indices, 32 rows of G^T}
      tc::tma_load_3d(a, &it.mt->tmGt, bar, kc, it.I1 * JB, 0);
      tc::tma_load_3d(a + 4096, &it.mt->tmGt, bar, kc, it.J1 * JB, 0);
      tc::tma_load_3d(a + 8192, &it.mt->tmGt, bar, kc, it.I2 * JB, 0);
      tc::tma_load_3d(a + 12288, &it.mt->tmGt, bar, kc, it.J2 * JB, 0);
    }
  }
  __device__ static void store(const Params&, const Item& it, int row, int col0, float (&v)[32]) {
    float* M;
    if (row < 64) { if (col0 >= 64) return; M = it.mt->M + (int64_t)it.p0 * JP * JP + row * JP + col0; }
    else { if (col0 < 64) return; M = it.mt->M + (int64_t)(it.p0 + 1) * JP * JP + (row - 64) * JP + (col0 - 64); }
#pragma unroll
    for (int j = 0; j < 32; ++j) atomicAdd(M + j, v[j]);
  }
};

// Apply: X[rows, I u J] <- X[rows, I u J] W for X in {G, V}; G also refreshes G^T.
// active_list/active_count: tensor-core-class pairs that the shared-memory Jacobi actually
// rotated this round (compacted on the device), so converged pairs cost nothing here
struct ApplyParams { EighMat* mats; const int* pair_mat; int max_tiles; int round; const int* active_list; const int* active_count; };
struct ApplyPolicy {
  using Params = ApplyParams;
  // pair-level state (pr .. Gt) is cached in the Item: consecutive work items of a CTA
  // belong to the same pair, so the dependent global loads happen once per pair, not per tile
  struct Item { EighMat* mt; int local, I, J, m0, which; int pr, active, n, np; float* G; float* V; float* Gt; };
  static constexpr int BN = 64;
  static constexpr bool B_IS_A = false;
  static constexpr bool MN_MAJOR = false;
  static constexpr int CHUNK = 16;
  static constexpr uint32_t TX_BYTES = tc::PTILE + 64 * 32 * 4;
  __device__ static void reset(Item& it) { it.pr = -1; it.active = 0; }
  __device__ static int total_work(const Params& p, int) { return *p.active_count * p.max_tiles * 2; }
  __device__ static bool decode(const Params& p, int w, Item& it) {
    it.which = w & 1;
    const int t = (w >> 1) % p.max_tiles, slot = (w >> 1) / p.max_tiles;
    if (slot != it.pr) {
      it.pr = slot;
      const int pr = p.active_list[slot];
      EighMat* mt = &p.mats[p.pair_mat[pr]];
      it.mt = mt;
      it.local = pr - mt->pair_base;
      it.active = 1;
      it.n = mt->n; it.np = mt->np; it.G = mt->G; it.V = mt->V; it.Gt = mt->Gt;
      tournament(p.round % (mt->nb - 1), it.local, mt->nb, it.I, it.J);
    }
    if (!it.active) return false;
    it.m0 = t * 128;
    return it.m0 < it.n;
  }
  __device__ static int num_kb(const Params&, const Item&) { return 2; }
  __device__ static void load(const Params&, const Item& it, int kbi, uint8_t* a, uint8_t* b, uint64_t* bar) {
    tc::tma_load_3d(a, it.which ? &it.mt->tmV : &it.mt->tmG, bar, (kbi == 0 ? it.I : it.J) * JB, it.m0, 0);
    tc::tma_load_3d(b, &it.mt->tmW, bar, kbi * 32, 0, it.local);
  }
  __device__ static void store(const Params&, const Item& it, int row, int col0, float (&v)[32]) {
    const int r = it.m0 + row;
    if (r >= it.n) return;
    const int np = it.np, cb = (col0 == 0 ? it.I : it.J) * JB;
    float* X = (it.which ? it.V : it.G) + (int64_t)r * np + cb;
#pragma unroll
    for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(X + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    if (!it.which && it.Gt) {
      float* T = it.Gt + (int64_t)cb * np + r;   // lanes hold consecutive rows r: coalesced columns of G^T
#pragma unroll
      for (int j = 0; j < 32; ++j) T[(int64_t)j * np] = v[j];
    }
  }
};

// ---- 128-wide pairs (EighMat::wide): one pair per Gram tile, apply with N = 128 ----
struct GramWidePolicy {
  using Params = GramParams;
  struct Item { EighMat* mt; int local, I, J, kb0, kb1; };
  static constexpr int BN = 128;
  static constexpr bool B_IS_A = true;
  static constexpr bool MN_MAJOR = false;
  static constexpr int CHUNK = 1;
  __device__ static void reset(Item&) {}
  __device__ static int total_work(const Params&, int t) { return t; }
  static constexpr uint32_t TX_BYTES = tc::PTILE;
  __device__ static bool decode(const Params& p, int w, Item& it) {
    const int gi = w / p.ksplits, sp = w % p.ksplits;
    EighMat* mt = &p.mats[p.item_mat[gi]];
    if (mt->done) return false;
    it.mt = mt;
    it.local = gi - mt->gram_base;
    tournament(p.round % (mt->nb - 1), it.local, mt->nb, it.I, it.J);
    const int kb_total = (mt->n + 31) / 32;
    const int per = (kb_total + p.ksplits - 1) / p.ksplits;
    it.kb0 = sp * per; it.kb1 = min(kb_total, it.kb0 + per);
    return it.kb1 > it.kb0;
  }
  __device__ static int num_kb(const Params&, const Item& it) { return it.kb1 - it.kb0; }
  __device__ static void load(const Params&, const Item& it, int kbi, uint8_t* a, uint8_t*, uint64_t* bar) {
    const int kc = (it.kb0 + kbi) * 32;   // boxes {32 reduction indices, 64 rows of G^T}
    tc::tma_load_3d(a, &it.mt->tmGt, bar, kc, it.I * 64, 0);
    tc::tma_load_3d(a + 8192, &it.mt->tmGt, bar, kc, it.J * 64, 0);
  }
  __device__ static void store(const Params&, const Item& it, int row, int col0, float (&v)[32]) {
    float* M = it.mt->M + (int64_t)it.local * 128 * 128 + row * 128 + col0;
#pragma unroll
    for (int j = 0; j < 32; ++j) atomicAdd(M + j, v[j]);
  }
};

struct ApplyWidePolicy {
  using Params = ApplyParams;
  struct Item { EighMat* mt; int local, I, J, m0, which; int pr, active, n, np; float* G; float* V; float* Gt; };
  static constexpr int BN = 128;
  static constexpr bool B_IS_A = false;
  static constexpr bool MN_MAJOR = false;
  static constexpr int CHUNK = 8;
  static constexpr uint32_t TX_BYTES = 2 * tc::PTILE;
  __device__ static void reset(Item& it) { it.pr = -1; it.active = 0; }
  __device__ static int total_work(const Params& p, int) { return *p.active_count * p.max_tiles * 2; }
  __device__ static bool decode(const Params& p, int w, Item& it) {
    it.which = w & 1;
    const int t = (w >> 1) % p.max_tiles, slot = (w >> 1) / p.max_tiles;
    if (slot != it.pr) {
      it.pr = slot;
      const int pr = p.active_list[slot];
      EighMat* mt = &p.mats[p.pair_mat[pr]];
      it.mt = mt;
      it.local = pr - mt->pair_base;
      it.active = 1;
      it.n = mt->n; it.np = mt->np; it.G = mt->G; it.V = mt->V; it.Gt = mt->Gt;
      tournament(p.round % (mt->nb - 1), it.local, mt->nb, it.I, it.J);
    }
    if (!it.active) return false;
    it.m0 = t * 128;
    return it.m0 < it.n;
  }
  __device__ static int num_kb(const Params&, const Item&) { return 4; }
  __device__ static void load(const Params&, const Item& it, int kbi, uint8_t* a, uint8_t* b, uint64_t* bar) {
    const int col = (kbi < 2 ? it.I : it.J) * 64 + (kbi & 1) * 32;
    tc::tma_load_3d(a, it.which ? &it.mt->tmV : &it.mt->tmG, bar, col, it.m0, 0);
    tc::tma_load_3d(b, &it.mt->tmW, bar, kbi * 32, 0, it.local);
  }
  __device__ static void store(const Params&, const Item& it, int row, int col0, float (&v)[32]) {
    const int r = it.m0 + row;
    if (r >= it.n) return;
    const int np = it.np, cb = (col0 < 64 ? it.I : it.J) * 64 + (col0 & 32);
    float* X = (it.which ? it.V : it.G) + (int64_t)r * np + cb;
#pragma unroll
    for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(X + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    if (!it.which && it.Gt) {
      float* T = it.Gt + (int64_t)cb * np + r;
#pragma unroll
      for (int j = 0; j < 32; ++j) T[(int64_t)j * np] = v[j];
    }
  }
};

int gemm_tn_plain(const float* A, int64_t lda, const float* B, int64_t ldb, float* D, int64_t ldd, int M, int N,
                  int K, cudaStream_t s);
int make_tmap_3d(CUtensorMap* tm, const float* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                 uint64_t stride2_bytes, uint32_t box_rows);

// --------------------------------------------------------------- host side
struct EighPlan {
  std::vector<EighMat> mats;
  std::vector<int> pair_mat, block_list, d64_list, d128_list;
  std::vector<int> tc_pair_mat, tc_gram_mat, simt_list;
  size_t off_mats, off_pair_mat, off_block, off_d64, off_d128, off_flag, off_data, total;
  size_t off_tc_pair, off_tc_gram, off_simt, off_all_pair, off_active;
  int total_pairs, max_nb, max_rows;
  int tc_pairs, tc_gram_items, tc_max_rows, simt_max_rows;
};

static void build_plan(const int* n, int count, EighPlan& pl) {
  pl.mats.resize(count);
  pl.total_pairs = 0; pl.max_nb = 0; pl.max_rows = 0;
  pl.tc_pairs = 0; pl.tc_gram_items = 0; pl.tc_max_rows = 0; pl.simt_max_rows = 0;
  for (int i = 0; i < count; ++i) {
    EighMat& m = pl.mats[i];
    m = EighMat{};
    m.n = n[i];
    if (n[i] <= 64) { m.mode = 0; pl.d64_list.push_back(i); }
    else if (n[i] <= 128) { m.mode = 1; pl.d128_list.push_back(i); }
    else if (n[i] < TC_MIN_N) {
      m.mode = 2;
      m.np = (n[i] + JP - 1) / JP * JP;
      m.nb = m.np / JB;
      m.pair_base = pl.total_pairs;
      for (int k = 0; k < m.nb / 2; ++k) pl.pair_mat.push_back(i);
      pl.total_pairs += m.nb / 2;
      pl.max_nb = std::max(pl.max_nb, m.nb);
      pl.max_rows = std::max(pl.max_rows, n[i]);
      pl.simt_max_rows = std::max(pl.simt_max_rows, n[i]);
      pl.block_list.push_back(i);
      pl.simt_list.push_back(i);
    } else {
      m.mode = 3;
      m.wide = eigh_wide() ? 1 : 0;
      m.np = (n[i] + 127) / 128 * 128;      // 128-row MMA tiles; nb % 4 == 0 -> pairs come in twos
      m.nb = m.np / (m.wide ? 64 : JB);
      m.pair_base = pl.tc_pairs;
      m.gram_base = pl.tc_gram_items;
      const int gitems = m.wide ? m.nb / 2 : m.nb / 4;   // wide: one pair per Gram tile
      for (int k = 0; k < m.nb / 2; ++k) pl.tc_pair_mat.push_back(i);
      for (int k = 0; k < gitems; ++k) pl.tc_gram_mat.push_back(i);
      pl.tc_pairs += m.nb / 2;
      pl.tc_gram_items += gitems;
      pl.max_nb = std::max(pl.max_nb, m.nb);
      pl.max_rows = std::max(pl.max_rows, n[i]);
      pl.tc_max_rows = std::max(pl.tc_max_rows, n[i]);
      pl.block_list.push_back(i);
    }
  }
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  pl.off_mats = take(sizeof(EighMat) * count);
  pl.off_pair_mat = take(sizeof(int) * std::max<size_t>(1, pl.pair_mat.size()));
  pl.off_block = take(sizeof(int) * std::max<size_t>(1, pl.block_list.size()));
  pl.off_d64 = take(sizeof(int) * std::max<size_t>(1, pl.d64_list.size()));
  pl.off_d128 = take(sizeof(int) * std::max<size_t>(1, pl.d128_list.size()));
  // unified pair list for the shared-memory Jacobi: SIMT-class pairs, then tensor-core-class pairs
  for (int i = 0; i < count; ++i) {
    EighMat& m = pl.mats[i];
    if (m.mode == 2) m.inner_base = m.pair_base;
    else if (m.mode == 3) m.inner_base = pl.total_pairs + m.pair_base;
  }
  pl.off_flag = take(sizeof(int) * 4);
  pl.off_active = take(sizeof(int) * std::max<size_t>(1, pl.tc_pair_mat.size()));
  pl.off_all_pair = take(sizeof(int) * std::max<size_t>(1, pl.pair_mat.size() + pl.tc_pair_mat.size()));
  pl.off_tc_pair = take(sizeof(int) * std::max<size_t>(1, pl.tc_pair_mat.size()));
  pl.off_tc_gram = take(sizeof(int) * std::max<size_t>(1, pl.tc_gram_mat.size()));
  pl.off_simt = take(sizeof(int) * std::max<size_t>(1, pl.simt_list.size()));
  pl.off_data = off;
  for (int i = 0; i < count; ++i) {
    EighMat& m = pl.mats[i];
    if (m.mode < 2) continue;
    const size_t sq = (size_t)m.np * m.np * sizeof(float);
    const size_t pw = m.wide ? 128 : JP;
    const size_t pb = (size_t)(m.nb / 2) * pw * pw * sizeof(float);
    m.G = (float*)take(sq); m.V = (float*)take(sq);      // offsets, rebased later
    if (m.mode == 3 && !gram_mn_major()) m.Gt = (float*)take(sq);
    m.M = (float*)take(pb); m.W = (float*)take(pb);
    m.pair_skip = (int*)take(sizeof(int) * (m.nb / 2));
  }
  pl.total = off;
}

}  // namespace kfac

using namespace kfac;

static const size_t SMEM64 = (2 * 64 * 65 + 64 + 32) * sizeof(float);
static const size_t SMEM128 = (2 * 128 * 129 + 128 + 64) * sizeof(float);
static int eigh_set_attrs() {   // once, from the calling thread, before any worker starts
  static bool done = false;
  if (done) return KFAC_OK;
  KFAC_CUDA(cudaFuncSetAttribute(jacobi_smem_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM128));
  KFAC_CUDA(cudaFuncSetAttribute(jacobi_smem_kernel<128, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM128));
  KFAC_CUDA(cudaFuncSetAttribute(jacobi_smem_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM64));
  KFAC_CUDA(cudaFuncSetAttribute(tc::pipeline_kernel<GramPolicyT<true>>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::PSMEM));
  KFAC_CUDA(cudaFuncSetAttribute(tc::pipeline_kernel<GramPolicyT<false>>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::PSMEM));
  KFAC_CUDA(cudaFuncSetAttribute(tc::pipeline_kernel<ApplyPolicy>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::PSMEM));
  KFAC_CUDA(cudaFuncSetAttribute(tc::pipeline_kernel<GramWidePolicy>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::PSMEM));
  KFAC_CUDA(cudaFuncSetAttribute(tc::pipeline_kernel<ApplyWidePolicy>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::PSMEM));
  (void)tc_num_sms();
  done = true;
  return KFAC_OK;
}

// The batch is split into two groups of similar cost that run their round loops on
// two streams from two host threads: the latency-bound shared-memory Jacobi of one group
// overlaps the tensor-core Gram/apply kernels of the other.
static void split_groups(const int* n, int count, std::vector<int> (&idx)[2]) {
  std::vector<int> order(count);
  for (int i = 0; i < count; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return n[a] > n[b]; });
  double load[2] = {0, 0};
  int nbig = 0;
  for (int i = 0; i < count; ++i) nbig += (n[i] >= TC_MIN_N);
  for (int i : order) {
    const int g = (nbig >= 2 && load[1] < load[0]) ? 1 : 0;
    idx[g].push_back(i);
    load[g] += (double)n[i] * n[i] * n[i];
  }
  std::sort(idx[0].begin(), idx[0].end());
  std::sort(idx[1].begin(), idx[1].end());
}

static size_t group_ws_bytes(const int* n, const std::vector<int>& idx) {
  if (idx.empty()) return 0;
  std::vector<int> ns(idx.size());
  for (size_t k = 0; k < idx.size(); ++k) ns[k] = n[idx[k]];
  EighPlan pl;
  build_plan(ns.data(), (int)ns.size(), pl);
  return align_up(pl.total, 1024);
}

// Which solver handles the factors with n > 128: the direct one (Householder tridiagonalisation + divide and
// conquer + block-reflector back-transformation, eigh_direct.cu) or the round-1 block Jacobi of this file.
static bool use_direct_solver() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("KFAC_EIGH_SOLVER"); v = (e && e[0] == 'j') ? 0 : 1; }
  return v != 0;
}

static size_t jacobi_workspace_bytes(const int* n, int count) {
  if (!n || count <= 0) return 0;
  std::vector<int> idx[2];
  split_groups(n, count, idx);
  return group_ws_bytes(n, idx[0]) + group_ws_bytes(n, idx[1]);
}

extern "C" size_t kfac_eigh_workspace_bytes(const int* n, int count) {
  if (!n || count <= 0) return 0;
  if (!use_direct_solver()) return jacobi_workspace_bytes(n, count);
  std::vector<int> big, small;
  for (int i = 0; i < count; ++i) (n[i] > 128 ? big : small).push_back(n[i]);
  return align_up(eigh_direct_workspace_bytes(big.data(), (int)big.size()), 1024) +
         jacobi_workspace_bytes(small.data(), (int)small.size());
}

static int eigh_run(const kfac_eigh_item* items, int count, void* ws, size_t ws_bytes, int max_sweeps, float tol,
                    cudaStream_t s, int slot) {
  if (count == 0) return KFAC_OK;
  std::vector<int> ns(count);
  for (int i = 0; i < count; ++i) {
    KFAC_CHECK_ARG(items[i].F && items[i].Q && items[i].d && items[i].n > 0 &&
                       (items[i].ldq == 0 || items[i].ldq >= items[i].n), "eigh item");
    ns[i] = items[i].n;
  }
  EighPlan pl;
  build_plan(ns.data(), count, pl);
  if (!ws || ws_bytes < pl.total) {
    set_error("eigh: workspace too small (%zu < %zu)", ws_bytes, pl.total);
    return KFAC_ERR_WORKSPACE;
  }
  if (max_sweeps <= 0) max_sweeps = 40;
  static const int env_inner = getenv("KFAC_EIGH_INNER") ? atoi(getenv("KFAC_EIGH_INNER")) : 0;
  const int inner_sweeps = env_inner > 0 ? env_inner : 2;   // per block pair and round (early exit when nothing rotates)
  char* base = (char*)ws;
  for (int i = 0; i < count; ++i) {
    EighMat& m = pl.mats[i];
    m.F = items[i].F; m.Q = items[i].Q; m.QT = items[i].QT; m.d = items[i].d;
    m.V0T = (m.mode >= 2) ? items[i].V0T : nullptr;
    m.ldq = items[i].ldq > 0 ? items[i].ldq : items[i].n;
    // tuning knobs (diagnostics): KFAC_EIGH_TOL / KFAC_EIGH_CONV_TOL / KFAC_EIGH_RMS_TOL
    static const float env_tol = getenv("KFAC_EIGH_TOL") ? (float)atof(getenv("KFAC_EIGH_TOL")) : 0.f;
    static const float env_conv = getenv("KFAC_EIGH_CONV_TOL") ? (float)atof(getenv("KFAC_EIGH_CONV_TOL")) : 0.f;
    static const float env_rms = getenv("KFAC_EIGH_RMS_TOL") ? (float)atof(getenv("KFAC_EIGH_RMS_TOL")) : 0.f;
    m.tol = tol > 0.f ? tol : (env_tol > 0.f ? env_tol : 3e-6f);   // pairs above this are re-diagonalised (see rel_off); ~ fp32 Gram noise floor
    m.conv_tol = fmaxf(m.tol, tol > 0.f ? tol : (env_conv > 0.f ? env_conv : 2e-5f));
    m.rms_tol = tol > 0.f ? tol : (env_rms > 0.f ? env_rms : 3e-5f);
    m.nw_ratio = (3e-6f / sqrtf((float)m.n)) / m.tol;
    if (m.mode >= 2) {
      m.G = (float*)(base + (size_t)m.G); m.V = (float*)(base + (size_t)m.V);
      if (m.mode == 3) {
        m.Gt = gram_mn_major() ? nullptr : (float*)(base + (size_t)m.Gt);
        const uint64_t np = m.np, rowb = np * 4;
        int rc;
        // G / V: 128-row x 32-column boxes; G^T: 32-row boxes; W^T pair buffers as {k, j, pair}
        if ((rc = make_tmap_3d(&m.tmG, m.G, np, np, 1, rowb, rowb * np, 128))) return rc;
        if ((rc = make_tmap_3d(&m.tmV, m.V, np, np, 1, rowb, rowb * np, 128))) return rc;
        if ((rc = make_tmap_3d(&m.tmGt, gram_mn_major() ? m.G : m.Gt, np, np, 1, rowb, rowb * np, m.wide ? 64 : 32))) return rc;
      }
      m.M = (float*)(base + (size_t)m.M); m.W = (float*)(base + (size_t)m.W);
      m.pair_skip = (int*)(base + (size_t)m.pair_skip);
      const uint64_t pw = m.wide ? 128 : JP;   // pair width
      KFAC_CUDA(cudaMemsetAsync(m.M, 0, (size_t)(m.nb / 2) * pw * pw * sizeof(float), s));
      if (m.mode == 3) {
        int rc;
        if ((rc = make_tmap_3d(&m.tmW, m.W, pw, pw, m.nb / 2, pw * 4, pw * pw * 4, (uint32_t)pw))) return rc;
      }
    }
  }
  EighMat* d_mats = (EighMat*)(base + pl.off_mats);
  int* d_pair_mat = (int*)(base + pl.off_pair_mat);
  int* d_block = (int*)(base + pl.off_block);
  int* d_d64 = (int*)(base + pl.off_d64);
  int* d_d128 = (int*)(base + pl.off_d128);
  KFAC_CUDA(cudaMemcpyAsync(d_mats, pl.mats.data(), sizeof(EighMat) * count, cudaMemcpyHostToDevice, s));
  if (!pl.pair_mat.empty())
    KFAC_CUDA(cudaMemcpyAsync(d_pair_mat, pl.pair_mat.data(), sizeof(int) * pl.pair_mat.size(), cudaMemcpyHostToDevice, s));
  if (!pl.block_list.empty())
    KFAC_CUDA(cudaMemcpyAsync(d_block, pl.block_list.data(), sizeof(int) * pl.block_list.size(), cudaMemcpyHostToDevice, s));
  if (!pl.d64_list.empty())
    KFAC_CUDA(cudaMemcpyAsync(d_d64, pl.d64_list.data(), sizeof(int) * pl.d64_list.size(), cudaMemcpyHostToDevice, s));
  if (!pl.d128_list.empty())
    KFAC_CUDA(cudaMemcpyAsync(d_d128, pl.d128_list.data(), sizeof(int) * pl.d128_list.size(), cudaMemcpyHostToDevice, s));
  int* d_all_pair = (int*)(base + pl.off_all_pair);
  if (!pl.pair_mat.empty())
    KFAC_CUDA(cudaMemcpyAsync(d_all_pair, pl.pair_mat.data(), sizeof(int) * pl.pair_mat.size(), cudaMemcpyHostToDevice, s));
  if (!pl.tc_pair_mat.empty())
    KFAC_CUDA(cudaMemcpyAsync(d_all_pair + pl.pair_mat.size(), pl.tc_pair_mat.data(), sizeof(int) * pl.tc_pair_mat.size(),
                              cudaMemcpyHostToDevice, s));
  int* d_tc_pair = (int*)(base + pl.off_tc_pair);
  int* d_tc_gram = (int*)(base + pl.off_tc_gram);
  if (!pl.tc_pair_mat.empty()) {
    KFAC_CUDA(cudaMemcpyAsync(d_tc_pair, pl.tc_pair_mat.data(), sizeof(int) * pl.tc_pair_mat.size(), cudaMemcpyHostToDevice, s));
    KFAC_CUDA(cudaMemcpyAsync(d_tc_gram, pl.tc_gram_mat.data(), sizeof(int) * pl.tc_gram_mat.size(), cudaMemcpyHostToDevice, s));
  }

  // the host vectors are pageable: cudaMemcpyAsync stages them before returning.

  const size_t SMEM64 = (2 * 64 * 65 + 64 + 32) * sizeof(float);
  const size_t SMEM128 = (2 * 128 * 129 + 128 + 64) * sizeof(float);
  const int nblock = (int)pl.block_list.size();
  // The small (n <= 128) matrices are solved by single latency-bound CTAs: run them on a
  // side stream so they overlap with the block-Jacobi rounds of the larger matrices.
  static cudaStream_t side_[2] = {nullptr, nullptr};
  static cudaEvent_t ev_fork_[2] = {nullptr, nullptr}, ev_join_[2] = {nullptr, nullptr};
  cudaStream_t& side = side_[slot];
  cudaEvent_t& ev_fork = ev_fork_[slot];
  cudaEvent_t& ev_join = ev_join_[slot];
  const bool have_direct = !pl.d64_list.empty() || !pl.d128_list.empty();
  cudaStream_t ds = s;
  if (have_direct && nblock > 0) {
    if (!side) {
      KFAC_CUDA(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
      KFAC_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
      KFAC_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
    }
    KFAC_CUDA(cudaEventRecord(ev_fork, s));          // descriptors are uploaded on s
    KFAC_CUDA(cudaStreamWaitEvent(side, ev_fork, 0));
    ds = side;
  }
  if (!pl.d64_list.empty()) {
    jacobi_smem_kernel<64><<<(int)pl.d64_list.size(), 512, SMEM64, ds>>>(d_mats, d_d64, 0, 24);
    KFAC_LAUNCH_CHECK();
  }
  if (!pl.d128_list.empty()) {
    jacobi_smem_kernel<128><<<(int)pl.d128_list.size(), 1024, SMEM128, ds>>>(d_mats, d_d128, 0, 24);
    KFAC_LAUNCH_CHECK();
  }
  if (ds != s) KFAC_CUDA(cudaEventRecord(ev_join, side));
  if (nblock > 0) {
    eigh_init_kernel<<<dim3(256, nblock), 256, 0, s>>>(d_mats, d_block);
    KFAC_LAUNCH_CHECK();
    for (int i = 0; i < count; ++i) {
      const EighMat& m = pl.mats[i];
      if (m.mode < 2 || !m.V0T) continue;
      // G[i][j] = sum_k F[i][k] V0[k][j]  (TN: A = F, B = V0^T);  G^T[i][j] = sum_k V0^T[i][k] F[j][k]
      int rc;
      if ((rc = gemm_tn_plain(m.F, m.n, m.V0T, m.ldq, m.G, m.np, m.n, m.n, m.n, s))) return rc;
      if (m.Gt && (rc = gemm_tn_plain(m.V0T, m.ldq, m.F, m.n, m.Gt, m.np, m.n, m.n, m.n, s))) return rc;
    }
    const int rps = pl.max_nb - 1;                 // rounds per sweep of the largest matrix
    const bool wide = pl.tc_pairs > 0 && eigh_wide();
    const int chunks = ceil_div(std::max(1, pl.simt_max_rows), GR);
    // tcgen05 class launch geometry
    const int sms = tc_num_sms();
    GramParams gp{d_mats, d_tc_gram, 1, 0};
    ApplyParams ap{d_mats, d_tc_pair, std::max(1, ceil_div(std::max(1, pl.tc_max_rows), 128)), 0};
    int gram_total = 0, apply_total = 0;
    if (pl.tc_pairs > 0) {
      const int kb_max = ceil_div(pl.tc_max_rows, 32);
      gp.ksplits = std::max(1, std::min(ceil_div(2 * sms, pl.tc_gram_items), std::max(1, kb_max / 4)));
      gram_total = pl.tc_gram_items * gp.ksplits;
      apply_total = pl.tc_pairs * ap.max_tiles * 2;
    }
    int* d_flag = (int*)(base + pl.off_flag);
    int* d_active_count = d_flag + 1;
    int* d_active_list = (int*)(base + pl.off_active);
    KFAC_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int) * 4, s));
    ap.active_list = d_active_list; ap.active_count = d_active_count;
    // Early exit without draining the GPU: the host enqueues sweep s+1, then waits
    // for the "all matrices converged" flag of sweep s (pinned read-back + event).
    static int* h_flag_[2] = {nullptr, nullptr};
    static cudaEvent_t ev_[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    int*& h_flag = h_flag_[slot];
    cudaEvent_t* ev = ev_[slot];
    if (!h_flag) {
      KFAC_CUDA(cudaMallocHost(&h_flag, 2 * sizeof(int)));
      KFAC_CUDA(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
      KFAC_CUDA(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
    }
    for (int sw = 0; sw < max_sweeps; ++sw) {
      for (int rr = 0; rr < rps; ++rr) {
        const int r = sw * rps + rr;
        // Gram of both classes, ONE shared-memory Jacobi launch over all pairs, then the applies
        if (pl.total_pairs > 0) {
          eigh_gram_kernel<<<dim3(pl.total_pairs, chunks), 256, 0, s>>>(d_mats, d_pair_mat, r);
          count_launch(1);
        }
        if (pl.tc_pairs > 0) {
          gp.round = r; ap.round = r;
          if (wide)
            tc::pipeline_kernel<GramWidePolicy><<<std::min(gram_total, sms), tc::PTHREADS, tc::PSMEM, s>>>(gp, gram_total);
          else if (gram_mn_major())
            tc::pipeline_kernel<GramPolicyT<true>><<<std::min(gram_total, sms), tc::PTHREADS, tc::PSMEM, s>>>(gp, gram_total);
          else
            tc::pipeline_kernel<GramPolicyT<false>><<<std::min(gram_total, sms), tc::PTHREADS, tc::PSMEM, s>>>(gp, gram_total);
          count_launch(1);
        }
        if (!wide) {
          // experimental inner-solver variants (see jacobi_smem_kernel): KFAC_EIGH_JOPT = bit mask 1 | 2 | 4;
          // 8 (+1) = the data-moving solver of jacobi_systolic.cu (with the fast rotation chain);
          // 16 = pair columns written sorted by their new norm (combines with 0, 1, 8, 9)
          static const int jopt_all = getenv("KFAC_EIGH_JOPT") ? atoi(getenv("KFAC_EIGH_JOPT")) : 0;
          const int jopt = jopt_all & 7;
          const bool jopt8 = (jopt_all & 8) != 0;
          const bool jsort = (jopt_all & 16) != 0;   // sorted pair columns (with jopt 0 / 1, or with the systolic solver)
          const int nblk = pl.total_pairs + pl.tc_pairs;
#define KFAC_JACOBI_BLOCK(OPT)                                                                              \
  jacobi_smem_kernel<64, OPT><<<nblk, 512, SMEM64, s>>>(d_mats, d_all_pair, 1, inner_sweeps, pl.total_pairs, \
                                                        d_active_list, d_active_count, 0)
          if (jopt8) {   // experimental data-moving pair solver (jacobi_systolic.cu)
            const int rc = launch_systolic_block64(d_mats, d_all_pair, nblk, inner_sweeps, pl.total_pairs, d_active_list,
                                                   d_active_count, 0, (jopt & 1) | (jsort ? 2 : 0), s);
            if (rc) return rc;
          } else if (jsort) {
            if (jopt & 1)
              jacobi_smem_kernel<64, 1, true><<<nblk, 512, SMEM64, s>>>(d_mats, d_all_pair, 1, inner_sweeps, pl.total_pairs,
                                                                      d_active_list, d_active_count, 0);
            else
              jacobi_smem_kernel<64, 0, true><<<nblk, 512, SMEM64, s>>>(d_mats, d_all_pair, 1, inner_sweeps, pl.total_pairs,
                                                                      d_active_list, d_active_count, 0);
          } else
          switch (jopt) {
            case 1: KFAC_JACOBI_BLOCK(1); break;
            case 2: KFAC_JACOBI_BLOCK(2); break;
            case 3: KFAC_JACOBI_BLOCK(3); break;
            case 4: KFAC_JACOBI_BLOCK(4); break;
            case 5: KFAC_JACOBI_BLOCK(5); break;
            case 6: KFAC_JACOBI_BLOCK(6); break;
            case 7: KFAC_JACOBI_BLOCK(7); break;
            default: KFAC_JACOBI_BLOCK(0); break;
          }
#undef KFAC_JACOBI_BLOCK
          count_launch(1);
        } else {   // 64x64 problems of the SIMT class, 128x128 problems of the tensor-core class
          static const bool wsort = getenv("KFAC_EIGH_JOPT") && (atoi(getenv("KFAC_EIGH_JOPT")) & 16);   // sorted pair columns
          if (pl.total_pairs > 0) {
            if (wsort) jacobi_smem_kernel<64, 0, true><<<pl.total_pairs, 512, SMEM64, s>>>(d_mats, d_all_pair, 1, inner_sweeps);
            else jacobi_smem_kernel<64><<<pl.total_pairs, 512, SMEM64, s>>>(d_mats, d_all_pair, 1, inner_sweeps);
            count_launch(1);
          }
          if (wsort)
            jacobi_smem_kernel<128, 0, true><<<pl.tc_pairs, 1024, SMEM128, s>>>(d_mats, d_all_pair + pl.total_pairs, 1, inner_sweeps,
                                                                                0, d_active_list, d_active_count, pl.total_pairs);
          else
          jacobi_smem_kernel<128><<<pl.tc_pairs, 1024, SMEM128, s>>>(d_mats, d_all_pair + pl.total_pairs, 1, inner_sweeps, 0,
                                                                     d_active_list, d_active_count, pl.total_pairs);
          count_launch(1);
        }
        if (pl.total_pairs > 0) {
          eigh_apply_kernel<<<dim3(pl.total_pairs, chunks, 2), 256, 0, s>>>(d_mats, d_pair_mat, r);
          count_launch(1);
        }
        if (pl.tc_pairs > 0) {
          if (wide)
            tc::pipeline_kernel<ApplyWidePolicy><<<std::min(apply_total, sms), tc::PTHREADS, tc::PSMEM, s>>>(ap, apply_total);
          else
            tc::pipeline_kernel<ApplyPolicy><<<std::min(apply_total, sms), tc::PTHREADS, tc::PSMEM, s>>>(ap, apply_total);
          count_launch(1);
        }
        eigh_ctl_kernel<<<1, 128, 0, s>>>(d_mats, d_block, nblock, r, d_flag, d_active_count);
      }
      KFAC_LAUNCH_CHECK();
      h_flag[sw & 1] = 0;
      KFAC_CUDA(cudaMemcpyAsync(&h_flag[sw & 1], d_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
      KFAC_CUDA(cudaEventRecord(ev[sw & 1], s));
      if (sw >= 1) {
        KFAC_CUDA(cudaEventSynchronize(ev[(sw - 1) & 1]));
        if (h_flag[(sw - 1) & 1]) break;
      }
    }
    eigh_final_kernel<<<dim3(64, nblock), 256, 0, s>>>(d_mats, d_block);
    KFAC_LAUNCH_CHECK();
    if (ds != s) KFAC_CUDA(cudaStreamWaitEvent(s, ev_join, 0));
    if (getenv("KFAC_EIGH_DEBUG")) {   // diagnostics only: per-matrix sweep counts (synchronises)
      std::vector<EighMat> back(count);
      KFAC_CUDA(cudaStreamSynchronize(s));
      KFAC_CUDA(cudaMemcpy(back.data(), d_mats, sizeof(EighMat) * count, cudaMemcpyDeviceToHost));
      for (int i = 0; i < count; ++i)
        if (back[i].mode >= 2)
        {
          fprintf(stderr, "[kfac eigh] n=%d mode=%d warm=%d sweeps=%d done=%d off:", back[i].n, back[i].mode,
                  back[i].V0T != nullptr, back[i].sweeps, back[i].done);
          for (int k = 0; k < back[i].sweeps && k < 12; ++k) fprintf(stderr, " %.1e/%.1e", back[i].off_hist[k], back[i].rms_hist[k]);
          fprintf(stderr, "\n");
        }
    }
  }
  return KFAC_OK;
}

extern "C" int kfac_eigh_batched(const kfac_eigh_item* items, int count, void* ws, size_t ws_bytes,
                                 int max_sweeps, float tol, void* stream) {
  KFAC_CHECK_ARG(count >= 0 && (items || count == 0), "items");
  if (count == 0) return KFAC_OK;
  cudaStream_t s = (cudaStream_t)stream;
  { const int rc = eigh_set_attrs(); if (rc) return rc; }
  std::vector<int> ns(count);
  for (int i = 0; i < count; ++i) {
    KFAC_CHECK_ARG(items[i].n > 0, "eigh item");
    ns[i] = items[i].n;
  }
  if (use_direct_solver()) {
    // n > 128: direct solver on `s`; n <= 128: shared-memory Jacobi on a side stream (overlaps)
    std::vector<kfac_eigh_item> big, small;
    std::vector<int> nbig, nsmall;
    for (int i = 0; i < count; ++i) {
      if (items[i].n > 128) { big.push_back(items[i]); nbig.push_back(items[i].n); }
      else { small.push_back(items[i]); nsmall.push_back(items[i].n); }
    }
    const size_t bd = align_up(eigh_direct_workspace_bytes(nbig.data(), (int)nbig.size()), 1024);
    const size_t bs = jacobi_workspace_bytes(nsmall.data(), (int)nsmall.size());
    if (!ws || ws_bytes < bd + bs) { set_error("eigh: workspace too small (%zu < %zu)", ws_bytes, bd + bs); return KFAC_ERR_WORKSPACE; }
    static cudaStream_t sd = nullptr;
    static cudaEvent_t ed_fork = nullptr, ed_join = nullptr;
    cudaStream_t ss = s;
    if (!big.empty() && !small.empty()) {
      if (!sd) {
        KFAC_CUDA(cudaStreamCreateWithFlags(&sd, cudaStreamNonBlocking));
        KFAC_CUDA(cudaEventCreateWithFlags(&ed_fork, cudaEventDisableTiming));
        KFAC_CUDA(cudaEventCreateWithFlags(&ed_join, cudaEventDisableTiming));
      }
      KFAC_CUDA(cudaEventRecord(ed_fork, s));
      KFAC_CUDA(cudaStreamWaitEvent(sd, ed_fork, 0));
      ss = sd;
    }
    if (!small.empty()) {
      const int rc = eigh_run(small.data(), (int)small.size(), (char*)ws + bd, bs, max_sweeps, tol, ss, 0);
      if (rc) return rc;
    }
    if (ss != s) KFAC_CUDA(cudaEventRecord(ed_join, ss));
    if (!big.empty()) {
      const int rc = eigh_direct_run(big.data(), (int)big.size(), ws, bd, s);
      if (rc) return rc;
    }
    if (ss != s) KFAC_CUDA(cudaStreamWaitEvent(s, ed_join, 0));
    return KFAC_OK;
  }
  std::vector<int> idx[2];
  split_groups(ns.data(), count, idx);
  const size_t b0 = group_ws_bytes(ns.data(), idx[0]), b1 = group_ws_bytes(ns.data(), idx[1]);
  if (!ws || ws_bytes < b0 + b1) {
    set_error("eigh: workspace too small (%zu < %zu)", ws_bytes, b0 + b1);
    return KFAC_ERR_WORKSPACE;
  }
  std::vector<kfac_eigh_item> g[2];
  for (int k = 0; k < 2; ++k)
    for (int i : idx[k]) g[k].push_back(items[i]);
  if (g[1].empty()) return eigh_run(g[0].data(), (int)g[0].size(), ws, b0, max_sweeps, tol, s, 0);

  // group 1 runs on a second stream, driven by a helper thread (each group's early-exit
  // logic blocks its own host thread only)
  static cudaStream_t s1 = nullptr;
  static cudaEvent_t e_fork = nullptr, e_join = nullptr;
  if (!s1) {
    KFAC_CUDA(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
    KFAC_CUDA(cudaEventCreateWithFlags(&e_fork, cudaEventDisableTiming));
    KFAC_CUDA(cudaEventCreateWithFlags(&e_join, cudaEventDisableTiming));
  }
  int dev = 0;
  KFAC_CUDA(cudaGetDevice(&dev));
  KFAC_CUDA(cudaEventRecord(e_fork, s));
  KFAC_CUDA(cudaStreamWaitEvent(s1, e_fork, 0));
  int rc1 = KFAC_OK;
  char err1[256] = "";
  std::thread worker([&]() {
    if (cudaSetDevice(dev) != cudaSuccess) { rc1 = KFAC_ERR_CUDA; return; }
    rc1 = eigh_run(g[1].data(), (int)g[1].size(), (char*)ws + b0, b1, max_sweeps, tol, s1, 1);
    if (rc1 != KFAC_OK) snprintf(err1, sizeof(err1), "%s", kfac_last_error());
  });
  const int rc0 = eigh_run(g[0].data(), (int)g[0].size(), ws, b0, max_sweeps, tol, s, 0);
  worker.join();
  KFAC_CUDA(cudaEventRecord(e_join, s1));
  KFAC_CUDA(cudaStreamWaitEvent(s, e_join, 0));
  if (rc0 != KFAC_OK) return rc0;
  if (rc1 != KFAC_OK) { set_error("%s", err1); return rc1; }
  return KFAC_OK;
}

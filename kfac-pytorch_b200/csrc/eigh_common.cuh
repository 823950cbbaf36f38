// Definitions shared by the block eigensolver (eigh.cu) and its experimental pair solver
// (jacobi_systolic.cu): the per-matrix descriptor and the convergence measure.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace kfac {

struct alignas(64) EighMat {
  CUtensorMap tmG, tmGt, tmV, tmW;   // mode 3 only (TMA views of G, G^T, V and the W^T pair buffers)
  const float* F; float* Q; float* QT; float* d;
  const float* V0T;            // optional warm start: transposed previous eigenbasis (ld = ldq)
  int ldq;
  float* G; float* V;          // np x np
  float* Gt;                   // mode 3: transposed copy of G (K-major operand of the Gram)
  float* M; float* W;          // pairs x JP x JP   (mode 3: W holds W^T)
  int* pair_skip;              // pairs
  int n, np, nb, pair_base, mode;
  int gram_base;               // mode 3: first Gram item (two pairs per 128-row MMA tile)
  int inner_base;              // first pair of this matrix in the unified (SIMT + TC) pair list
  float tol;
  unsigned int sweep_off;      // float bits, atomicMax
  float prev_off;              // convergence measure of the previous sweep
  unsigned int max_diag;       // float bits: largest |g_j|^2 seen (lambda_max^2), atomicMax
  float nw_ratio;              // normwise relaxation, see pair_den()
  float conv_tol;              // matrix is done when a sweep STARTS below this (>= tol)
  float sweep_sumsq;           // sum over column pairs of rel_off^2 in this sweep (atomicAdd)
  float rms_tol;               // ... or when sqrt(2*sumsq/n) starts below this
  float rms_hist[12];
  float off_hist[12];          // diagnostics: convergence measure after each of the first sweeps
  int done;
  int sweeps;
  int wide;                    // mode 3: block width 64 (pair width 128) instead of 32 (64)
};

// round-robin tournament: pair k of round r among nb (even) players
__device__ __forceinline__ void tournament(int r, int k, int nb, int& p, int& q) {
  const int m = nb - 1;
  int a, b;
  if (k == 0) { a = r % m; b = m; }
  else { a = (r + k) % m; b = (r - k + m) % m; }
  p = min(a, b); q = max(a, b);
}

// Convergence measure of a column pair in Gram space (M = G^T G, M_ii = lambda_i^2):
// |g_p . g_q| / max(|g_p|^2, |g_q|^2) ~ the amplitude with which the two eigenvectors
// still contaminate each other.  (The classical |.|/sqrt(M_pp M_qq) asks for RELATIVE
// accuracy of tiny eigenvalues, which fp32 Grams of graded factors cannot deliver --
// the sweeps would chase rounding noise forever -- and which K-FAC does not need:
// everything below the damping is flattened by 1/(dg*da + damping).)
//
// Pairs of columns that are BOTH small next to the largest column carry absolute
// rounding noise ~eps*lambda_max from their history of rotations, so their mutual
// measure can never reach the tolerance; for them the test relaxes to the normwise
// backward-stable form |g_p.g_q| <= tol_n * lambda_max * max(|g_p|,|g_q|) with
// tol_n = nw_ratio*tol ~ 3e-6/sqrt(n) (what LAPACK-class fp32 solvers deliver).
__device__ __forceinline__ float pair_den(float app, float aqq, float max_diag, float nw_ratio) {
  const float den = fmaxf(fabsf(app), fabsf(aqq));
  return fmaxf(den, nw_ratio * sqrtf(max_diag * den));
}
__device__ __forceinline__ float rel_off(float apq, float app, float aqq, float max_diag, float nw_ratio) {
  const float den = pair_den(app, aqq, max_diag, nw_ratio);
  const float x = fabsf(apq);
  if (x == 0.f) return 0.f;
  return den > 0.f ? x / den : 1e30f;
}

// experimental block-mode pair solver (jacobi_systolic.cu), selected by KFAC_EIGH_JOPT bit 3;
// same contract as jacobi_smem_kernel<64> in block mode (opts bit 0: rsqrt/rcp rotation chain, bit 1: sorted columns)
int launch_systolic_block64(EighMat* mats, const int* list, int nblk, int max_inner, int tc_first, int* active_list,
                            int* active_count, int pair_shift, int opts, cudaStream_t s);

}  // namespace kfac

// EXPERIMENTAL (round-2 candidate, not on the product path yet): data-moving ("systolic")
// two-sided Jacobi for the N x N pair problems of the block eigensolver.
//
// Why: the shared-memory Jacobi of eigh.cu is half of every round (127 us of 267 us for a lone
// n = 4608 factor, profiles/r01_launches_probe_n4608.md).  It addresses the rotation partners
// (p, q) of a step through an index table, which costs ~56 LDS/STS per thread and step (rotation
// parameters + pair table + scattered M / W entries) and exposes the ~600-cycle rotation-parameter
// chain between two barriers.  Here the round-robin tournament is realised by MOVING the data:
//
//   * positions 0..h-1 are the "tops", h..N-1 the "bottoms" (h = N/2); pair k = (k, h + k),
//     always.  After the rotations of a step every row/column moves to dest(pos) (the classic
//     rotation with top 0 fixed), folded into the store addresses of the update.  After N-1
//     steps everything is back where it started, so whole sweeps need no bookkeeping at all.
//   * the 2x2 block (a, b) of the update reads M[a][b], M[a][h+b], M[h+a][b], M[h+a][h+b]:
//     consecutive lanes (b) touch consecutive words, loads and stores are conflict free and no
//     index table is read: 4 LDS + 4 STS + 2 LDS.64 per block instead of 14 scattered accesses.
//   * a step is  [load + rotate into registers] -> barrier -> [permuted store] -> barrier  (in
//     place, no ping-pong buffers: 32 KB of shared memory for N = 64, 128 KB for N = 128).
//   * the rotation parameters of step t+1 are computed by N/2 extra ("crit") threads DURING the
//     store phase of step t: they fetch the three 2x2 blocks their next pair depends on in the
//     load phase, redo those updates in registers and run the div/sqrt chain while the bulk
//     threads store -- the chain is off the critical path.
//
// The functions below are plain inline code on explicit worker indices, so the same source is
// exercised by a host emulation (tests/host/jacobi_systolic_host.cpp, run by
// tests/test_host_logic.py): index maps, update formulas and convergence are checked on the
// CPU; only bank behaviour and timing need the GPU.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define KFAC_SYS_HD __host__ __device__ __forceinline__
#else
#define KFAC_SYS_HD inline
#endif

namespace kfac {
namespace sysj {

// where the row/column at `pos` sits after one tournament rotation
template <int N>
KFAC_SYS_HD int dest(int pos) {
  constexpr int h = N / 2;
  if (pos < h) return pos == 0 ? 0 : (pos == h - 1 ? N - 1 : pos + 1);
  return pos == h ? 1 : pos - 1;
}
// inverse of dest
template <int N>
KFAC_SYS_HD int src(int pos) {
  constexpr int h = N / 2;
  if (pos < h) return pos == 0 ? 0 : (pos == 1 ? h : pos - 1);
  return pos == N - 1 ? h - 1 : pos + 1;
}

struct Rot { float c, s; };

struct Criteria {
  int mode_block;      // 1: pair Gram of the block solver (relative + normwise threshold), 0: whole matrix
  float tol_in;
  float max_diag;      // block mode: largest squared column norm seen so far
  float nw_ratio;      // block mode: normwise relaxation (see pair_den in eigh.cu)
  int fast;            // 1: rotation parameters from rsqrt/rcp + one Newton step; 2: half-angle form (2 rsqrt)
};

// 1/sqrt(x) and 1/x to ~1 ulp from the hardware approximations (device) / exact (host emulation)
KFAC_SYS_HD float rsqrt_refined(float x) {
#ifdef __CUDA_ARCH__
  float r = rsqrtf(x);
  return r * fmaf(-0.5f * x * r, r, 1.5f);
#else
  return 1.f / sqrtf(x);
#endif
}
KFAC_SYS_HD float rcp_rn(float x) {
#ifdef __CUDA_ARCH__
  return __frcp_rn(x);
#else
  return 1.f / x;
#endif
}

KFAC_SYS_HD float threshold(const Criteria& cr, float app, float aqq) {
  if (!cr.mode_block) return sqrtf(fabsf(app * aqq));
  const float den = fmaxf(fabsf(app), fabsf(aqq));
  return fmaxf(den, cr.nw_ratio * sqrtf(cr.max_diag * den));
}

// Jacobi rotation that annihilates apq; flags: bit 0 = rotated, bit 1 = |sin| >= 2e-3
KFAC_SYS_HD Rot rotation(const Criteria& cr, float app, float aqq, float apq, int& flags) {
  Rot r{1.f, 0.f};
  if (fabsf(apq) > cr.tol_in * threshold(cr, app, aqq)) {
    if (cr.fast == 2) {
      // half-angle form, two dependent special-function ops: with a = aqq - app, b = 2 apq and
      // h = sqrt(a^2 + b^2):  cos 2t = |a| / h,  c = sqrt((1 + cos 2t) / 2),  s = sign(a b) |b| / (2 h c)
      // (no cancellation: small angles give s ~ b / (2 |a|) by multiplication).  The operands are
      // scaled by a power of two taken from the exponent of max(|a|, |b|) -- no reciprocal.
      float a = aqq - app, b = 2.f * apq;
      const float mxv = fmaxf(fabsf(a), fabsf(b));
      int e;
      (void)frexpf(mxv, &e);
      a = ldexpf(a, -e); b = ldexpf(b, -e);
      const float rh = rsqrt_refined(fmaf(a, a, b * b));        // 1 / h
      const float cc = fmaf(0.5f * fabsf(a), rh, 0.5f);         // c^2 in [0.5, 1]
      const float rc = rsqrt_refined(cc);                        // 1 / c
      r.c = cc * rc;
      r.s = copysignf(0.5f * fabsf(b) * rh * rc, (a < 0.f) != (b < 0.f) ? -1.f : 1.f);
    } else if (cr.fast) {
      // t = sign(a b) |b| / (|a| + sqrt(a^2 + b^2)), a = aqq - app, b = 2 apq (scaled by
      // 1 / max(|a|, |b|) so the squares stay in range): 4 short dependent special-function ops
      // instead of 5 IEEE div/sqrt -- the chain is the critical path of a step (measured)
      float a = aqq - app, b = 2.f * apq;
      const float sc = rcp_rn(fmaxf(fabsf(a), fabsf(b)));
      a *= sc; b *= sc;
      const float x = fmaf(a, a, b * b);
      const float t = copysignf(fabsf(b), (a < 0.f) != (b < 0.f) ? -1.f : 1.f) * rcp_rn(fabsf(a) + x * rsqrt_refined(x));
      r.c = rsqrt_refined(fmaf(t, t, 1.f));
      r.s = t * r.c;
    } else {
      const float tau = (aqq - app) / (2.f * apq);
      const float t = copysignf(1.f, tau) / (fabsf(tau) + sqrtf(1.f + tau * tau));
      r.c = 1.f / sqrtf(1.f + t * t);
      r.s = t * r.c;
    }
    if (r.s != 0.f) flags |= 1 | (fabsf(r.s) >= 2e-3f ? 2 : 0);
  }
  return r;
}

// [o00 o01; o10 o11] = Ja^T [m00 m01; m10 m11] Jb with J = [c s; -s c]
KFAC_SYS_HD void block_update(Rot a, Rot b, float m00, float m01, float m10, float m11, float& o00, float& o01,
                              float& o10, float& o11) {
  const float n00 = b.c * m00 - b.s * m01, n01 = b.s * m00 + b.c * m01;
  const float n10 = b.c * m10 - b.s * m11, n11 = b.s * m10 + b.c * m11;
  o00 = a.c * n00 - a.s * n10;
  o01 = a.c * n01 - a.s * n11;
  o10 = a.s * n00 + a.c * n10;
  o11 = a.s * n01 + a.c * n11;
}

// per-worker register file of the bulk update: TB bulk workers share h*h blocks of M and
// N*h row/pair items of W
template <int N, int TB>
struct BulkRegs {
  static constexpr int h = N / 2;
  static constexpr int MB = (h * h + TB - 1) / TB;
  static constexpr int WB = (N * h + TB - 1) / TB;
  float m[MB][4];
  float w[WB][2];
};

// phase 1 of a step (bulk worker w of TB): read the current M / W (row pitch LD) and rotate
template <int N, int LD, int TB>
KFAC_SYS_HD void bulk_load(int w, const Rot* rot, const float* M, const float* W, BulkRegs<N, TB>& r) {
  constexpr int h = N / 2;
#pragma unroll
  for (int j = 0; j < BulkRegs<N, TB>::MB; ++j) {
    const int idx = w + j * TB;
    if (idx < h * h) {   // (guard, not break: keeps the loop unrollable and r.m in registers)
      const int b = idx % h, a = idx / h;
      block_update(rot[a], rot[b], M[a * LD + b], M[a * LD + h + b], M[(h + a) * LD + b], M[(h + a) * LD + h + b],
                   r.m[j][0], r.m[j][1], r.m[j][2], r.m[j][3]);
    }
  }
#pragma unroll
  for (int j = 0; j < BulkRegs<N, TB>::WB; ++j) {
    const int idx = w + j * TB;
    if (idx < N * h) {
      const int k = idx % h, i = idx / h;
      const Rot q = rot[k];
      const float u = W[i * LD + k], v = W[i * LD + h + k];
      r.w[j][0] = q.c * u - q.s * v;
      r.w[j][1] = q.s * u + q.c * v;
    }
  }
}

// phase 2 of a step: permuted store (rows of W keep their place)
template <int N, int LD, int TB>
KFAC_SYS_HD void bulk_store(int w, float* M, float* W, const BulkRegs<N, TB>& r) {
  constexpr int h = N / 2;
#pragma unroll
  for (int j = 0; j < BulkRegs<N, TB>::MB; ++j) {
    const int idx = w + j * TB;
    if (idx < h * h) {
      const int b = idx % h, a = idx / h;
      const int r0 = dest<N>(a), r1 = dest<N>(h + a), c0 = dest<N>(b), c1 = dest<N>(h + b);
      M[r0 * LD + c0] = r.m[j][0];
      M[r0 * LD + c1] = r.m[j][1];
      M[r1 * LD + c0] = r.m[j][2];
      M[r1 * LD + c1] = r.m[j][3];
    }
  }
#pragma unroll
  for (int j = 0; j < BulkRegs<N, TB>::WB; ++j) {
    const int idx = w + j * TB;
    if (idx < N * h) {
      const int k = idx % h, i = idx / h;
      W[i * LD + dest<N>(k)] = r.w[j][0];
      W[i * LD + dest<N>(h + k)] = r.w[j][1];
    }
  }
}

// What crit worker k needs to know the rotation of pair k in the NEXT step: that pair will hold
// the items now at x = src(k) and y = src(h + k); their three Gram entries after this step's
// update come from the diagonal 2x2 blocks of their current pairs and from the cross block.
struct CritRegs {
  float dx[4];   // diagonal block of x's pair (all four entries: M is symmetric only up to rounding)
  float dy[4];
  float xy[4];   // cross block (pair of x, pair of y)
  Rot rx, ry;
  int bx, by;    // bottom members?
};

template <int N, int LD>
KFAC_SYS_HD void crit_load(int k, const Rot* rot, const float* M, CritRegs& c) {
  constexpr int h = N / 2;
  const int x = src<N>(k), y = src<N>(h + k);
  const int ax = x % h, ay = y % h;
  c.bx = x >= h; c.by = y >= h;
  c.rx = rot[ax]; c.ry = rot[ay];
  c.dx[0] = M[ax * LD + ax]; c.dx[1] = M[ax * LD + h + ax];
  c.dx[2] = M[(h + ax) * LD + ax]; c.dx[3] = M[(h + ax) * LD + h + ax];
  c.dy[0] = M[ay * LD + ay]; c.dy[1] = M[ay * LD + h + ay];
  c.dy[2] = M[(h + ay) * LD + ay]; c.dy[3] = M[(h + ay) * LD + h + ay];
  c.xy[0] = M[ax * LD + ay]; c.xy[1] = M[ax * LD + h + ay];
  c.xy[2] = M[(h + ax) * LD + ay]; c.xy[3] = M[(h + ax) * LD + h + ay];
}

KFAC_SYS_HD Rot crit_rotation(const Criteria& cr, const CritRegs& c, int& flags) {
  float o00, o01, o10, o11;
  block_update(c.rx, c.rx, c.dx[0], c.dx[1], c.dx[2], c.dx[3], o00, o01, o10, o11);
  const float app = c.bx ? o11 : o00;
  block_update(c.ry, c.ry, c.dy[0], c.dy[1], c.dy[2], c.dy[3], o00, o01, o10, o11);
  const float aqq = c.by ? o11 : o00;
  block_update(c.rx, c.ry, c.xy[0], c.xy[1], c.xy[2], c.xy[3], o00, o01, o10, o11);
  const float apq = c.bx ? (c.by ? o11 : o10) : (c.by ? o01 : o00);
  return rotation(cr, app, aqq, apq, flags);
}

// rotation of pair k from data that is in place (first step of a sweep)
template <int N, int LD>
KFAC_SYS_HD Rot first_rotation(int k, const Criteria& cr, const float* M, int& flags) {
  constexpr int h = N / 2;
  return rotation(cr, M[k * LD + k], M[(h + k) * LD + h + k], M[k * LD + h + k], flags);
}

}  // namespace sysj
}  // namespace kfac

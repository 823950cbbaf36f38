// Host emulation of kfac-pytorch_b200/csrc/jacobi_systolic.cuh (experimental data-moving Jacobi):
// the CTA is replayed serially, one phase at a time (phase 1: every worker loads + rotates into its
// register file, phase 2: every worker stores; a barrier separates them on the device).  Checks
//   * dest()/src() are inverse permutations, N-1 moves restore the arrangement, and every pair
//     of items meets exactly once per sweep;
//   * the rotation the crit workers predict for the next step equals, bit for bit, the rotation
//     computed from the stored matrix;
//   * the sweeps converge: W^T M0 W is diagonal, W orthogonal, eigenvalues match a plain cyclic
//     Jacobi in double precision.
// Build: g++ -O1 -ffp-contract=off -std=c++17 -I kfac-pytorch_b200/csrc tests/host/jacobi_systolic_host.cpp
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <random>
#include <set>
#include <vector>

#include "jacobi_systolic.cuh"

using namespace kfac::sysj;

static int fails = 0;
#define CHECK(cond, ...) do { if (!(cond)) { ++fails; std::printf("FAIL %s:%d: ", __FILE__, __LINE__); std::printf(__VA_ARGS__); std::printf("\n"); } } while (0)

template <int N>
static void check_schedule() {
  constexpr int h = N / 2;
  std::vector<int> item(N), nxt(N);
  for (int i = 0; i < N; ++i) { item[i] = i; CHECK(src<N>(dest<N>(i)) == i && dest<N>(src<N>(i)) == i, "inverse N=%d pos=%d", N, i); }
  std::set<std::pair<int, int>> met;
  for (int st = 0; st < N - 1; ++st) {
    for (int k = 0; k < h; ++k) {
      auto pr = std::minmax(item[k], item[h + k]);
      CHECK(met.insert({pr.first, pr.second}).second, "pair met twice N=%d step=%d", N, st);
    }
    for (int p = 0; p < N; ++p) nxt[dest<N>(p)] = item[p];
    item = nxt;
  }
  CHECK((int)met.size() == N * (N - 1) / 2, "pairs covered %zu of %d", met.size(), N * (N - 1) / 2);
  for (int i = 0; i < N; ++i) CHECK(item[i] == i, "arrangement not restored N=%d pos=%d", N, i);
}

// reference: cyclic Jacobi in double, eigenvalues only
static std::vector<double> ref_eigs(std::vector<double> A, int n) {
  for (int sw = 0; sw < 60; ++sw) {
    double off = 0;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = A[p * n + q];
        off += apq * apq;
        if (apq == 0) continue;
        const double tau = (A[q * n + q] - A[p * n + p]) / (2 * apq);
        const double t = (tau >= 0 ? 1 : -1) / (std::fabs(tau) + std::sqrt(1 + tau * tau));
        const double c = 1 / std::sqrt(1 + t * t), s = t * c;
        for (int k = 0; k < n; ++k) { const double u = A[k * n + p], v = A[k * n + q]; A[k * n + p] = c * u - s * v; A[k * n + q] = s * u + c * v; }
        for (int k = 0; k < n; ++k) { const double u = A[p * n + k], v = A[q * n + k]; A[p * n + k] = c * u - s * v; A[q * n + k] = s * u + c * v; }
      }
    if (off < 1e-30) break;
  }
  std::vector<double> d(n);
  for (int i = 0; i < n; ++i) d[i] = A[i * n + i];
  std::sort(d.begin(), d.end());
  return d;
}

template <int N, int TB>
static void check_solve(unsigned seed, bool graded, int fast = 0, int block_mode = 0) {
  constexpr int h = N / 2, LD = N;
  std::mt19937 rng(seed);
  std::normal_distribution<float> nd;
  // SPD test matrix M0 = X^T X (graded column scales on request)
  const int rows = 2 * N;
  std::vector<double> X(rows * N);
  for (int i = 0; i < rows; ++i)
    for (int j = 0; j < N; ++j) X[i * N + j] = nd(rng) * (graded ? std::pow(10.0, -3.0 * j / N) : 1.0);
  std::vector<double> M0(N * N);
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) { double s = 0; for (int r = 0; r < rows; ++r) s += X[r * N + i] * X[r * N + j]; M0[i * N + j] = s; }
  std::vector<float> M(N * N), W(N * N, 0.f);
  for (int i = 0; i < N * N; ++i) M[i] = (float)M0[i];
  for (int i = 0; i < N; ++i) W[i * LD + i] = 1.f;
  Criteria cr{0, 1e-7f, 0.f, 0.f, fast};
  if (block_mode) {   // thresholds of the block solver: tol 3e-6, tol_in = tol / 8, normwise relaxation for n = 4608
    float md = 0.f; for (int i = 0; i < N; ++i) md = std::fmax(md, M[i * LD + i]);
    cr = Criteria{1, 3.75e-7f, md, (3e-6f / std::sqrt(4608.f)) / 3e-6f, fast};
  }
  Rot rot[h], nrot[h];
  std::vector<BulkRegs<N, TB>> regs(TB);
  std::vector<CritRegs> crit(h);
  int sweeps = 0;
  for (; sweeps < 40; ++sweeps) {
    int flags = 0;
    for (int k = 0; k < h; ++k) rot[k] = first_rotation<N, LD>(k, cr, M.data(), flags);
    for (int st = 0; st < N - 1; ++st) {
      const bool more = st < N - 2;
      for (int w = 0; w < TB; ++w) bulk_load<N, LD, TB>(w, rot, M.data(), W.data(), regs[w]);        // phase 1
      if (more) for (int k = 0; k < h; ++k) crit_load<N, LD>(k, rot, M.data(), crit[k]);
      for (int w = 0; w < TB; ++w) bulk_store<N, LD, TB>(w, M.data(), W.data(), regs[w]);            // phase 2
      if (more) {
        for (int k = 0; k < h; ++k) nrot[k] = crit_rotation(cr, crit[k], flags);
        for (int k = 0; k < h; ++k) {
          int f2 = 0;
          const Rot chk = first_rotation<N, LD>(k, cr, M.data(), f2);
          CHECK(std::memcmp(&chk, &nrot[k], sizeof(Rot)) == 0, "crit prediction N=%d sweep=%d step=%d pair=%d: (%g,%g) vs (%g,%g)", N,
                sweeps, st, k, nrot[k].c, nrot[k].s, chk.c, chk.s);
          rot[k] = nrot[k];
        }
      }
    }
    if (!(flags & 1)) break;
  }
  CHECK(sweeps < 40, "no convergence N=%d", N);
  // symmetry of the maintained M, diagonalisation, orthogonality
  double mx = 0; for (int i = 0; i < N; ++i) mx = std::max(mx, (double)std::fabs(M[i * LD + i]));
  double worst_off = 0, worst_orth = 0, worst_res = 0;
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) {
      double wtw = 0;
      for (int k = 0; k < N; ++k) wtw += (double)W[k * LD + i] * W[k * LD + j];
      worst_orth = std::max(worst_orth, std::fabs(wtw - (i == j)));
      if (i != j) worst_off = std::max(worst_off, std::fabs((double)M[i * LD + j]) / std::sqrt(std::fabs((double)M[i * LD + i] * M[j * LD + j]) + 1e-300));
    }
  // residual: W^T M0 W vs diag(M)
  std::vector<double> T(N * N);
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) { double s = 0; for (int k = 0; k < N; ++k) s += M0[i * N + k] * W[k * LD + j]; T[i * N + j] = s; }
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) {
      double s = 0; for (int k = 0; k < N; ++k) s += (double)W[k * LD + i] * T[k * N + j];
      worst_res = std::max(worst_res, std::fabs(s - (i == j ? (double)M[i * LD + i] : 0.0)) / mx);
    }
  std::vector<double> d(N); for (int i = 0; i < N; ++i) d[i] = M[i * LD + i];
  std::sort(d.begin(), d.end());
  const std::vector<double> ref = ref_eigs(M0, N);
  double worst_eig = 0; for (int i = 0; i < N; ++i) worst_eig = std::max(worst_eig, std::fabs(d[i] - ref[i]) / mx);
  std::printf("N=%d TB=%d graded=%d fast=%d sweeps=%d off=%.2e orth=%.2e resid=%.2e eig=%.2e\n", N, TB, (int)graded, fast, sweeps + 1, worst_off,
              worst_orth, worst_res, worst_eig);
  if (!block_mode) CHECK(worst_off < 5e-6, "off-diagonal %.3e", worst_off);   // (block mode measures against max, not the geometric mean)
  CHECK(worst_orth < 5e-5, "orthogonality %.3e", worst_orth);
  CHECK(worst_res < 2e-5, "residual %.3e", worst_res);
  CHECK(worst_eig < 2e-5, "eigenvalues %.3e", worst_eig);
}

int main() {
  check_schedule<4>(); check_schedule<8>(); check_schedule<64>(); check_schedule<128>();
  check_solve<8, 5>(1, false);
  check_solve<8, 16>(2, true);
  check_solve<64, 512>(3, false);
  check_solve<64, 512>(4, true);
  check_solve<128, 960>(5, true);
  check_solve<128, 480>(11, true);
  check_solve<64, 512>(9, true, 0, 1);
  check_solve<64, 512>(10, true, 1, 1);
  check_solve<8, 5>(6, true, 1);
  check_solve<64, 512>(12, false, 2);
  check_solve<64, 512>(13, true, 2);
  check_solve<64, 512>(14, true, 2, 1);
  check_solve<64, 512>(7, false, 1);
  check_solve<64, 512>(8, true, 1);
  if (fails) { std::printf("%d failure(s)\n", fails); return 1; }
  std::printf("OK\n");
  return 0;
}

// tcgen05 GEMM engine: D = epilogue(alpha * A B^T), A (M x K) and B (N x K)
// row-major fp32 ("TN", both operands K-major), fp32-equivalent accuracy from a
// 3xTF32 split:  a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi  with fp32 TMEM
// accumulation (single-pass TF32 fails the 1e-3 parity bar).
//
// One persistent CTA per SM, warp-specialised:
//   warp 0      TMA producer   (cp.async.bulk.tensor 3-D tiles, SWIZZLE_128B)
//   warp 1      MMA issuer     (one lane: 12 x tcgen05.mma 128x128x8 per k-block)
//   warps 2-5   splitter       (smem -> smem: lo = x - trunc_tf32(x))
//   warps 6-9   epilogue       (tcgen05.ld TMEM -> regs -> fused epilogue -> HBM)
// 3 smem stages x {A_hi, B_hi, A_lo, B_lo} x 16 KB = 192 KB; 2 TMEM accumulators
// (2 x 128 columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
#include "common.cuh"
#include "tc_common.cuh"

namespace kfac {

constexpr int TBM = 128, TBN = 128, TBK = 32, TSTAGES = 3;
constexpr int TILE_BYTES = TBM * TBK * 4;     // 16 KB
constexpr int STAGE_BYTES = 4 * TILE_BYTES;   // 64 KB
constexpr int TC_THREADS = 320;
constexpr size_t TC_SMEM = (size_t)TSTAGES * STAGE_BYTES + 1024 + 256;

struct TcParams {
  float* D; int64_t ldd;
  int M, N, K, kbatch;
  int tiles_m, tiles_n, splits, upper_only, atomic;
  float alpha;
  int epi; const float* E; int64_t lde; const float* dg; const float* da; float damping;
};

struct WorkItem { int m0, n0, kb0, kb1, diag; };

__device__ __forceinline__ WorkItem decode_work(const TcParams& p, int w) {
  WorkItem it;
  const int tile = w / p.splits, sp = w % p.splits;
  int tm, tn;
  if (p.upper_only) {
    int t = tile; tm = 0;
    while (t >= p.tiles_n - tm) { t -= p.tiles_n - tm; ++tm; }
    tn = tm + t;
  } else { tm = tile / p.tiles_n; tn = tile % p.tiles_n; }
  it.m0 = tm * TBM; it.n0 = tn * TBN; it.diag = (tm == tn);
  const int kblocks = (p.K + TBK - 1) / TBK;
  const int kb_total = kblocks * p.kbatch;
  const int per = (kb_total + p.splits - 1) / p.splits;
  it.kb0 = sp * per; it.kb1 = min(kb_total, it.kb0 + per);
  return it;
}

__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, TcParams p,
               int total_work) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)TSTAGES * STAGE_BYTES);
  uint64_t* full = bars;                   // [TSTAGES]  TMA landed
  uint64_t* conv = bars + TSTAGES;         // [TSTAGES]  lo tiles written
  uint64_t* empty = bars + 2 * TSTAGES;    // [TSTAGES]  MMAs retired
  uint64_t* tfull = bars + 3 * TSTAGES;    // [2] accumulator ready
  uint64_t* tempty = tfull + 2;            // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kblocks = (p.K + TBK - 1) / TBK;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmA);
    tc::prefetch_tmap(&tmB);
    for (int s = 0; s < TSTAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&conv[s], 4); tc::mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { tc::mbar_init(&tfull[a], 1); tc::mbar_init(&tempty[a], 4); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, 256);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto tile_ptr = [&](int s, int which) { return smem + (size_t)s * STAGE_BYTES + (size_t)which * TILE_BYTES; };

  if (warp == 0) {
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        const WorkItem it = decode_work(p, w);
        for (int kb = it.kb0; kb < it.kb1; ++kb) {
          const int b = kb / kblocks, kc = (kb % kblocks) * TBK;
          tc::mbar_wait(&empty[s], ph ^ 1);
          tc::mbar_arrive_expect_tx(&full[s], 2 * TILE_BYTES);
          tc::tma_load_3d(tile_ptr(s, 0), &tmA, &full[s], kc, it.m0, b);
          tc::tma_load_3d(tile_ptr(s, 1), &tmB, &full[s], kc, it.n0, b);
          if (++s == TSTAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = tc::make_idesc_tf32(TBM, TBN);
      int s = 0; uint32_t ph = 0; int acc = 0; uint32_t aph = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        const WorkItem it = decode_work(p, w);
        tc::mbar_wait(&tempty[acc], aph ^ 1);
        tc::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * TBN;
        uint32_t accum = 0;
        for (int kb = it.kb0; kb < it.kb1; ++kb) {
          tc::mbar_wait(&conv[s], ph);
          tc::tc_fence_after();
          const uint64_t ahi = tc::make_kmajor_sw128_desc(tc::smem_u32(tile_ptr(s, 0)));
          const uint64_t bhi = tc::make_kmajor_sw128_desc(tc::smem_u32(tile_ptr(s, 1)));
          const uint64_t alo = tc::make_kmajor_sw128_desc(tc::smem_u32(tile_ptr(s, 2)));
          const uint64_t blo = tc::make_kmajor_sw128_desc(tc::smem_u32(tile_ptr(s, 3)));
#pragma unroll
          for (int kk = 0; kk < TBK / 8; ++kk) {
            const uint64_t adv = (uint64_t)(kk * 2);   // 8 fp32 = 32 B = 2 x 16 B
            tc::mma_tf32(d_tmem, alo + adv, bhi + adv, idesc, accum);
            tc::mma_tf32(d_tmem, ahi + adv, blo + adv, idesc, 1u);
            tc::mma_tf32(d_tmem, ahi + adv, bhi + adv, idesc, 1u);
            accum = 1u;
          }
          tc::tc_commit(&empty[s]);
          if (++s == TSTAGES) { s = 0; ph ^= 1; }
        }
        tc::tc_commit(&tfull[acc]);
        acc ^= 1; if (acc == 0) aph ^= 1;
      }
    }
  } else if (warp < 6) {
    const int t = threadIdx.x - 64;   // 0..127
    int s = 0; uint32_t ph = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      const WorkItem it = decode_work(p, w);
      for (int kb = it.kb0; kb < it.kb1; ++kb) {
        tc::mbar_wait(&full[s], ph);
#pragma unroll
        for (int which = 0; which < 2; ++which) {
          const float4* src = reinterpret_cast<const float4*>(tile_ptr(s, which));
          float4* dst = reinterpret_cast<float4*>(tile_ptr(s, which + 2));
#pragma unroll
          for (int i = 0; i < TILE_BYTES / 16 / 128; ++i) {
            const float4 v = src[t + i * 128];
            float4 lo;
            lo.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
            lo.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
            lo.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
            lo.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
            dst[t + i * 128] = lo;
          }
        }
        tc::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&conv[s]);
        if (++s == TSTAGES) { s = 0; ph ^= 1; }
      }
    }
  } else {
    const int q = warp & 3;   // TMEM lane quarter this warp may access
    int acc = 0; uint32_t aph = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      const WorkItem it = decode_work(p, w);
      tc::mbar_wait(&tfull[acc], aph);
      tc::tc_fence_after();
      const int m = it.m0 + q * 32 + lane;
      const bool mirror = p.upper_only && !it.diag;
#pragma unroll 1
      for (int c = 0; c < TBN / 32; ++c) {
        float v[32];
        tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * TBN + c * 32), v);
        tc::tmem_ld_wait();
        const int nb = it.n0 + c * 32;
        if (m < p.M && nb < p.N) {
          float* drow = p.D + (int64_t)m * p.ldd + nb;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = p.alpha * v[j];
            if (nb + j < p.N) {
              if (p.epi == EPI_MUL) x *= p.E[(int64_t)m * p.lde + nb + j];
              else if (p.epi == EPI_DIV_OUTER) x = x / (p.dg[m] * p.da[nb + j] + p.damping);
            }
            v[j] = x;
          }
          if (!p.atomic && !mirror && (p.ldd & 3) == 0 && nb + 32 <= p.N) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(drow + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (nb + j >= p.N) continue;
              if (p.atomic) {
                atomicAdd(drow + j, v[j]);
                if (mirror) atomicAdd(p.D + (int64_t)(nb + j) * p.ldd + m, v[j]);
              } else {
                drow[j] = v[j];
                if (mirror) p.D[(int64_t)(nb + j) * p.ldd + m] = v[j];
              }
            }
          }
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tempty[acc]);
      acc ^= 1; if (acc == 0) aph ^= 1;
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, 256);
}

// ------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 3-D map {K, rows, kbatch} over a row-major fp32 matrix (ld elements) with box {32, 128, 1}
static int make_tmap(CUtensorMap* tm, const float* base, int64_t rows, int64_t K, int64_t ld, int kbatch,
                     int64_t kb_stride) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return KFAC_ERR_CUDA; }
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)kbatch};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)(kbatch > 1 ? kb_stride : ld * rows) * 4};
  cuuint32_t box[3] = {TBK, TBM, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return KFAC_ERR_CUDA; }
  return KFAC_OK;
}

bool tc_gemm_supported(const TcGemmArgs& a) {
  auto al16 = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
  if (!(al16(a.A) && al16(a.B))) return false;
  if ((a.lda & 3) || (a.ldb & 3)) return false;
  if (a.kbatch > 1 && ((a.a_kb_stride & 3) || (a.b_kb_stride & 3))) return false;
  if (a.M < 1 || a.N < 1 || a.K < 1) return false;
  return true;
}

int launch_tc_gemm(const TcGemmArgs& a, cudaStream_t stream) {
  if (!tc_gemm_supported(a)) { set_error("tc_gemm: unsupported alignment/shape"); return KFAC_ERR_UNSUPPORTED; }
  static int num_sms = 0;
  static bool attr = false;
  if (!attr) {
    int dev = 0;
    KFAC_CUDA(cudaGetDevice(&dev));
    KFAC_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    KFAC_CUDA(cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
    attr = true;
  }
  const int kbatch = a.kbatch > 0 ? a.kbatch : 1;
  CUtensorMap tmA, tmB;
  int rc;
  if ((rc = make_tmap(&tmA, a.A, a.M, a.K, a.lda, kbatch, a.a_kb_stride))) return rc;
  if ((rc = make_tmap(&tmB, a.B, a.N, a.K, a.ldb, kbatch, a.b_kb_stride))) return rc;
  TcParams p{};
  p.D = a.D; p.ldd = a.ldd; p.M = a.M; p.N = a.N; p.K = a.K; p.kbatch = kbatch;
  p.tiles_m = ceil_div(a.M, TBM); p.tiles_n = ceil_div(a.N, TBN);
  p.upper_only = a.upper_only; p.atomic = a.atomic; p.alpha = a.alpha;
  p.epi = a.epi; p.E = a.E; p.lde = a.lde; p.dg = a.dg; p.da = a.da; p.damping = a.damping;
  const int ntiles = a.upper_only ? p.tiles_m * (p.tiles_m + 1) / 2 : p.tiles_m * p.tiles_n;
  const int kb_total = ceil_div(a.K, TBK) * kbatch;
  int splits = a.splits;
  if (splits <= 0) {
    splits = 1;
    if (a.atomic && ntiles < num_sms) splits = std::max(1, std::min(num_sms / ntiles, kb_total / 8));
  }
  splits = std::max(1, std::min(splits, kb_total));
  if (splits > 1 && !a.atomic) { set_error("tc_gemm: split-K needs atomic accumulation"); return KFAC_ERR_BAD_ARG; }
  // make every split non-empty
  const int per = ceil_div(kb_total, splits);
  splits = ceil_div(kb_total, per);
  p.splits = splits;
  const int total = ntiles * splits;
  const int grid = std::min(total, num_sms);
  tc_gemm_kernel<<<grid, TC_THREADS, TC_SMEM, stream>>>(tmA, tmB, p, total);
  KFAC_LAUNCH_CHECK();
  return KFAC_OK;
}

}  // namespace kfac

extern "C" int kfac_gemm_tn_tc(const float* A, int64_t lda, const float* B, int64_t ldb, float* D, int64_t ldd,
                               int M, int N, int K, float alpha, int atomic, int splits, void* stream) {
  using namespace kfac;
  KFAC_CHECK_ARG(A && B && D, "null pointer");
  TcGemmArgs a{};
  a.A = A; a.lda = lda; a.B = B; a.ldb = ldb; a.D = D; a.ldd = ldd;
  a.M = M; a.N = N; a.K = K; a.kbatch = 1; a.alpha = alpha; a.atomic = atomic; a.splits = splits;
  return launch_tc_gemm(a, (cudaStream_t)stream);
}

// C[M,N] (+)= alpha * A[M,K] . B[N,K]^T + bias[N]  — bf16 operands (both K-major), fp32 accumulate in TMEM.
// One CTA per 128 x BN output tile (x split-K slice): warp 0 = TMA producer, warp 1 = tcgen05.mma issuer,
// warp 2 = TMEM allocator, warps 4-7 = epilogue (tcgen05.ld -> registers -> global).
// Replaces ATen addmm / matmul in the tower MLPs (reference src/models/two_tower.py:62,70,129,276) and the
// chunked logits of the in-batch loss backward (:470).  "fp32 mode" is obtained by the caller feeding the
// split-bf16 expansions produced by prep.cu (K grows 6x), so this file only knows bf16.
#include "host_util.h"
#include "tc_common.cuh"
#include "../../include/b200rec.h"

namespace b200 {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 256;

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN <= 64) ? 8 : 6;
  // tcgen05 accumulates in fp32 with truncation, so the error of one accumulator chain grows linearly with the
  // number of MMAs added into it.  K-blocks are dealt round-robin onto NACC TMEM accumulators that the epilogue sums
  // with round-to-nearest adds: 4x shorter chains for the fp32-grade (split-bf16) products.
  static constexpr int NACC = 4;
  static constexpr int TMEM_COLS = NACC * BN;  // 512 or 256
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    float* __restrict__ C, long long ldc, int M, int N, int kb_total, int kb_per_split,
                    const float* __restrict__ bias, float alpha, int atomic) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space (LDS / STS)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* acc_bar = empty_bar + Cfg::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * GEMM_BM;
  const int n0 = blockIdx.y * BN;
  const int kb_begin = blockIdx.z * kb_per_split;
  int kb_end = kb_begin + kb_per_split;
  if (kb_end > kb_total) kb_end = kb_total;
  const int nkb = kb_end - kb_begin;  // may be <= 0 for a trailing split: then this CTA contributes nothing

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (nkb > 0) {
    // Producer and MMA issuer run warp-uniform loops and let one ELECTED lane issue (elect.sync): under a `lane == 0`
    // branch ptxas wraps every TMA / tcgen05 instruction in an ELECT + R2UR waterfall loop (~19 SASS instructions per
    // MMA), which for the 32-64 cycle MMAs of these tiles made the issuing thread the bottleneck (ncu: tensor pipe
    // 22 % active on the in-batch backward GEMMs).
    if (warp == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % Cfg::STAGES;
        const uint32_t ph = (i / Cfg::STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        if (elect_one()) {
          uint8_t* a_dst = smem + s * Cfg::STAGE_BYTES;
          uint8_t* b_dst = a_dst + Cfg::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
          tma_load_2d(a_dst, &tmap_a, &full_bar[s], (kb_begin + i) * GEMM_BK, m0);
          tma_load_2d(b_dst, &tmap_b, &full_bar[s], (kb_begin + i) * GEMM_BK, n0);
        }
        __syncwarp();
      }
    } else if (warp == 1) {
      const uint32_t idesc = umma_idesc_bf16(GEMM_BM, BN);
      const uint64_t desc_base = umma_desc_k_sw128(0);
      const uint32_t smem_lo = (smem_u32(smem) & 0x3FFFFu) >> 4;
      for (int i = 0; i < nkb; ++i) {
        const int s = i % Cfg::STAGES;
        const uint32_t ph = (i / Cfg::STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t a_desc = desc_base + (smem_lo + ((s * Cfg::STAGE_BYTES) >> 4));
          const uint64_t b_desc = a_desc + (Cfg::A_BYTES >> 4);
          const uint32_t d_addr = tmem_base + (i % Cfg::NACC) * BN;
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k)
            umma_bf16(d_addr, a_desc + 2 * k, b_desc + 2 * k, idesc, (i >= Cfg::NACC) || (k != 0));
          umma_commit(&empty_bar[s]);
          if (i == nkb - 1) umma_commit(acc_bar);
        }
        __syncwarp();
      }
    } else if (warp >= 4) {
      const int quarter = warp & 3;
      mbar_wait(acc_bar, 0);
      tc_fence_after();
      const int row = m0 + quarter * 32 + lane;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
      const bool vec_ok = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + c, v);
        tmem_ld_wait();
        const int nacc = nkb < Cfg::NACC ? nkb : Cfg::NACC;
        for (int a = 1; a < nacc; ++a) {
          uint32_t w[32];
          tmem_ld_32x32(taddr + a * BN + c, w);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
        }
        if (row < M) {
          float* crow = C + static_cast<long long>(row) * ldc;
          const int col0 = n0 + c;
          const bool add_bias = (bias != nullptr) && (blockIdx.z == 0);
          if (!atomic && vec_ok && col0 + 32 <= N) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 o;
              o.x = __uint_as_float(v[j + 0]) * alpha;
              o.y = __uint_as_float(v[j + 1]) * alpha;
              o.z = __uint_as_float(v[j + 2]) * alpha;
              o.w = __uint_as_float(v[j + 3]) * alpha;
              if (add_bias) {
                o.x += __ldg(bias + col0 + j + 0);
                o.y += __ldg(bias + col0 + j + 1);
                o.z += __ldg(bias + col0 + j + 2);
                o.w += __ldg(bias + col0 + j + 3);
              }
              *reinterpret_cast<float4*>(crow + col0 + j) = o;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int col = col0 + j;
              if (col < N) {
                float o = __uint_as_float(v[j]) * alpha;
                if (add_bias) o += __ldg(bias + col);
                if (atomic)
                  atomicAdd(crow + col, o);
                else
                  crow[col] = o;
              }
            }
          }
        }
      }
      tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN>
static int launch_gemm(const void* A, int64_t lda, int64_t M, const void* B, int64_t ldb, int64_t N, int64_t K,
                       float* C, int64_t ldc, const float* bias, float alpha, int k_splits, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  CUtensorMap ta, tb;
  if (make_tmap_bf16_2d(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, GEMM_BM)) return 1;
  if (make_tmap_bf16_2d(&tb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, BN)) return 1;
  const int kb_total = (int)((K + GEMM_BK - 1) / GEMM_BK);
  const bool force_atomic = k_splits < 0;  // negative: accumulate into C even when a single split remains
  if (force_atomic) k_splits = -k_splits;
  if (k_splits < 1) k_splits = 1;
  if (k_splits > kb_total) k_splits = kb_total;
  const int kb_per_split = (kb_total + k_splits - 1) / k_splits;
  k_splits = (kb_total + kb_per_split - 1) / kb_per_split;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA_OK(cudaFuncSetAttribute(gemm_bf16_tn_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      Cfg::SMEM_BYTES));
    attr_set = true;
  }
  dim3 grid((unsigned)((M + GEMM_BM - 1) / GEMM_BM), (unsigned)((N + BN - 1) / BN), (unsigned)k_splits);
  gemm_bf16_tn_kernel<BN><<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, st>>>(ta, tb, C, (long long)ldc, (int)M, (int)N,
                                                                       kb_total, kb_per_split, bias, alpha,
                                                                       (k_splits > 1 || force_atomic) ? 1 : 0);
  B200_LAUNCH_OK("gemm_bf16_tn_kernel");
  return 0;
}

}  // namespace b200

extern "C" int b200rec_gemm_bf16_tn(const void* A, int64_t lda, int64_t M, const void* B, int64_t ldb, int64_t N,
                                    int64_t K, float* C, int64_t ldc, const float* bias, float alpha, int k_splits,
                                    void* stream) {
  using namespace b200;
  if (!A || !B || !C) return fail("gemm: null pointer");
  if (M <= 0 || N <= 0 || K <= 0) return fail("gemm: empty problem M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
  if (M > INT32_MAX || N > INT32_MAX || K > INT32_MAX) return fail("gemm: dimension exceeds int32");
  if ((lda % 8) || (ldb % 8)) return fail("gemm: lda/ldb must be multiples of 8 elements (16 bytes)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (N <= 64) return launch_gemm<64>(A, lda, M, B, ldb, N, K, C, ldc, bias, alpha, k_splits, st);
  return launch_gemm<128>(A, lda, M, B, ldb, N, K, C, ldc, bias, alpha, k_splits, st);
}

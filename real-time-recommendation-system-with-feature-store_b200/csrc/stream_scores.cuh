// Streaming score kernel skeleton shared by exact top-K retrieval (topk.cu) and the fused in-batch
// log-sum-exp (inbatch_ce.cu):  S = Qm . X^T  where a block of 128*NQ "query" rows stays resident in shared
// memory and the N "catalogue" rows stream through a TMA ring; scores live only in TMEM (double buffered) and are
// consumed by the epilogue warps — they never reach HBM.
//
// Work decomposition (persistent, static, perfectly balanced): the (supertile, X-tile) space is linearised and cut
// into gridDim.x equal ranges, so a CTA handles at most a few "segments" (one supertile, a contiguous tile range).
// Warp roles: 0 = TMA producer, 1 = tcgen05.mma issuer, 2 = TMEM allocator, 3 = idle, 4..11 = epilogue (consume TMEM),
// 12..15 = helper warps owned by the epilogue policy (asynchronous candidate compaction for top-K; idle for LSE).
#pragma once
#include <cstdlib>
#include "host_util.h"
#include "tc_common.cuh"

namespace b200 {

constexpr int ST_EPI_WARP0 = 4;
constexpr int ST_HELP_WARP0 = 12;
constexpr int ST_HELP_WARPS = 4;
constexpr int ST_THREADS = 16 * 32;
constexpr int ST_QTILE_BYTES = 128 * 64 * 2;  // one 128-row x 64-col bf16 SW128 block
constexpr int ST_MAX_STAGES = 16;
constexpr int ST_SMEM_LIMIT = 232448;  // 227 KB
constexpr int ST_BAR_BYTES = 512;  // barriers + tmem slot; the policy-owned scratch (Epi::SCRATCH_BYTES) follows

struct StreamGeom {
  long long N;      // streamed rows
  long long T;      // X tiles per sweep = ceil(N / BN)
  long long total;  // S * T
  long long W;      // tiles per CTA
  int Q;            // resident-side rows
  int KB;           // 64-column k-blocks per row
  int S;            // supertiles = ceil(Q / (128*NQ))
  int grid;
  int max_parts;    // CTAs that can touch one supertile
  int stages;
  int smem_bytes;
  int dbg_nofeed;   // development knob (2-CTA kernel): issue MMAs without waiting for / loading operands
  int ks;           // 2-CTA kernel: 64-column atoms per ring stage
  int dbg_stats;    // development knob (2-CTA kernel): accumulate pipeline wait cycles in g_stream_stats
  int mma_order;    // development knob (2-CTA kernel): MMA issue order / accumulator hand-off variant
  // 2-CTA kernel work decomposition (see geom2_segment): units [0, n_main) sweep tiles r, r+R, r+2R, ... < Tmain of ONE
  // supertile each; the remaining units share the tail tiles [Tmain, T) of all supertiles as contiguous ranges.
  int n_main, R;
  long long Tmain, Tt, We;
};

// development knobs of the geometry (read from the environment once; b200rec_debug_reload_env re-reads them)
struct StreamKnobs {
  int mma_order = 0, sched = -1, ks = 0;
};
inline StreamKnobs read_stream_knobs() {
  StreamKnobs k;
  if (const char* e = getenv("B200REC_MMA_ORDER")) k.mma_order = atoi(e);
  if (const char* e = getenv("B200REC_SCHED")) k.sched = atoi(e);
  if (const char* e = getenv("B200REC_KS")) k.ks = atoi(e);
  return k;
}
inline StreamKnobs& stream_knobs() {
  static StreamKnobs k = read_stream_knobs();
  return k;
}

// development counters of the 2-CTA kernel (one copy per translation unit): [0] MMA warp cycles waiting for a free
// accumulator, [1] waiting for operands, [2] epilogue-warp cycles waiting for an accumulator, [3] consuming it, [4] tiles
static __device__ unsigned long long g_stream_stats[8];

template <int NQ, int BN>
inline bool stream_geom(StreamGeom& g, long long N, int Q, int KB, int sms, int scratch_bytes) {
  const int ST_TAIL_BYTES = ST_BAR_BYTES + scratch_bytes;
  g.N = N;
  g.Q = Q;
  g.KB = KB;
  g.dbg_nofeed = 0;
  g.dbg_stats = 0;
  g.mma_order = 0;
  g.ks = 1;
  const int rows_per_super = 128 * NQ;
  g.S = (Q + rows_per_super - 1) / rows_per_super;
  g.T = (N + BN - 1) / BN;
  g.total = g.S * g.T;
  long long grid = g.total < sms ? g.total : sms;
  g.W = (g.total + grid - 1) / grid;
  g.grid = (int)((g.total + g.W - 1) / g.W);
  long long mp = (g.T + g.W - 1) / g.W + 1;
  g.max_parts = (int)(mp < g.grid ? mp : g.grid);
  const int q_bytes = NQ * KB * ST_QTILE_BYTES;
  const int stage_bytes = BN * 128;
  const int avail = ST_SMEM_LIMIT - 1024 - ST_TAIL_BYTES - q_bytes;
  int stages = avail / stage_bytes;
  if (stages > ST_MAX_STAGES) stages = ST_MAX_STAGES;
  g.stages = stages;
  g.smem_bytes = 1024 + q_bytes + stages * stage_bytes + ST_TAIL_BYTES;
  return stages >= 2;
}

__device__ __forceinline__ int geom_first_cta(const StreamGeom& g, int s) { return (int)((s * g.T) / g.W); }
__device__ __forceinline__ int geom_last_cta(const StreamGeom& g, int s) { return (int)(((s + 1) * g.T - 1) / g.W); }

// Epilogue policy interface (all methods are called by the epilogue warps only, warp-uniformly):
//   struct Epi {
//     struct Args;                                                   // POD kernel arguments
//     template<...> __device__ void begin_segment(...);              // reset per-query state
//     __device__ void tile(...);                                     // consume one accumulator buffer
//     __device__ void end_segment(...);                              // publish the segment's partial result
//   };

template <int NQ, int BN, class Epi>
__global__ void __launch_bounds__(ST_THREADS, 1)
stream_scores_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_x,
                     const StreamGeom g, const typename Epi::Args ea) {
  constexpr int QPT = NQ >= 2 ? NQ / 2 : 1;          // queries owned by one epilogue thread
  constexpr int EPI_WARPS = NQ >= 2 ? 8 : 4;         // warps that actually consume TMEM
  constexpr int STAGE_BYTES = BN * 128;
  static_assert(NQ <= 2, "helper-warp mailboxes are sized for 256 resident rows");
  static_assert(2 * NQ * BN <= 512, "two accumulator buffers must fit the 512 TMEM columns");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space (LDS / STS)
  uint8_t* q_smem = smem;
  uint8_t* ring = smem + NQ * g.KB * ST_QTILE_BYTES;
  uint8_t* tail = ring + g.stages * STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + ST_MAX_STAGES;
  uint64_t* acc_full = empty_bar + ST_MAX_STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* q_full = acc_empty + 2;
  uint64_t* q_empty = q_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_empty + 1);
  uint32_t* scratch = reinterpret_cast<uint32_t*>(tail + ST_BAR_BYTES);  // policy-owned scratch (Epi::SCRATCH_BYTES)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_x);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < g.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], EPI_WARPS);
    }
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  if (warp == 3) Epi::init_scratch(scratch, lane);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const long long w_begin = (long long)blockIdx.x * g.W;
  long long w_end = w_begin + g.W;
  if (w_end > g.total) w_end = g.total;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // Producer and MMA issuer walk their loops with the whole warp (uniform control flow) and let one ELECTED lane
    // issue: under a `lane == 0` branch ptxas wraps each TMA / tcgen05 instruction in an ELECT + R2UR waterfall loop
    // (~19 SASS instructions per MMA), which makes the single issuing thread the bottleneck of the pipeline.
    {
      const uint64_t pol_q = l2_policy_evict_last();
      uint32_t it = 0, seg = 0;
      for (long long w = w_begin; w < w_end; ++seg) {
        const int s = (int)(w / g.T);
        const long long t0 = w - (long long)s * g.T;
        long long t1 = t0 + (w_end - w);
        if (t1 > g.T) t1 = g.T;
        mbar_wait(q_empty, (seg & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(q_full, NQ * g.KB * ST_QTILE_BYTES);
          for (int qs = 0; qs < NQ; ++qs)
            for (int kb = 0; kb < g.KB; ++kb)
              tma_load_2d_hint(q_smem + (qs * g.KB + kb) * ST_QTILE_BYTES, &tmap_q, q_full, kb * 64,
                               (s * NQ + qs) * 128, pol_q);
        }
        __syncwarp();
        for (long long t = t0; t < t1; ++t) {
          for (int kb = 0; kb < g.KB; ++kb, ++it) {
            const int st = it % g.stages;
            const uint32_t ph = (it / g.stages) & 1;
            mbar_wait(&empty_bar[st], ph ^ 1);
            if (elect_one()) {
              mbar_arrive_expect_tx(&full_bar[st], STAGE_BYTES);
              tma_load_2d(ring + st * STAGE_BYTES, &tmap_x, &full_bar[st], kb * 64, (int)(t * BN));
            }
            __syncwarp();
          }
        }
        w += t1 - t0;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    {
      const uint32_t idesc = umma_idesc_bf16(128, BN);
      const uint64_t desc_base = umma_desc_k_sw128(0);
      const uint32_t q_lo = (smem_u32(q_smem) & 0x3FFFFu) >> 4;
      const uint32_t ring_lo = (smem_u32(ring) & 0x3FFFFu) >> 4;
      uint32_t it = 0, seg = 0, tc = 0;
      for (long long w = w_begin; w < w_end; ++seg) {
        const int s = (int)(w / g.T);
        const long long t0 = w - (long long)s * g.T;
        long long t1 = t0 + (w_end - w);
        if (t1 > g.T) t1 = g.T;
        mbar_wait(q_full, seg & 1);
        tc_fence_after();
        for (long long t = t0; t < t1; ++t, ++tc) {
          const uint32_t buf = tc & 1;
          mbar_wait(&acc_empty[buf], ((tc >> 1) & 1) ^ 1);
          tc_fence_after();
          for (int kb = 0; kb < g.KB; ++kb, ++it) {
            const int st = it % g.stages;
            const uint32_t ph = (it / g.stages) & 1;
            mbar_wait(&full_bar[st], ph);
            tc_fence_after();
            if (elect_one()) {
              const uint64_t b_desc = desc_base + (ring_lo + ((st * STAGE_BYTES) >> 4));
#pragma unroll
              for (int qs = 0; qs < NQ; ++qs) {
                const uint64_t a_desc = desc_base + (q_lo + (((qs * g.KB + kb) * ST_QTILE_BYTES) >> 4));
                const uint32_t d_addr = tmem_base + (buf * NQ + qs) * BN;
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(d_addr, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
              }
              umma_commit(&empty_bar[st]);
              if (kb == g.KB - 1) umma_commit(&acc_full[buf]);
            }
            __syncwarp();
          }
        }
        if (elect_one()) umma_commit(q_empty);
        __syncwarp();
        w += t1 - t0;
      }
    }
  } else if (warp >= ST_EPI_WARP0 && warp < ST_EPI_WARP0 + EPI_WARPS) {
    // ------------------------------------------------------------------ epilogue
    const int ew = warp - ST_EPI_WARP0;
    const int quarter = warp & 3;  // TMEM lane quarter this warp may read
    const int half = ew >> 2;
    Epi epi;
    uint32_t tc = 0;
    for (long long w = w_begin; w < w_end;) {
      const int s = (int)(w / g.T);
      const long long t0 = w - (long long)s * g.T;
      long long t1 = t0 + (w_end - w);
      if (t1 > g.T) t1 = g.T;
      const int part = (int)blockIdx.x - geom_first_cta(g, s);
      int qslot[QPT];
      long long qrow[QPT];
#pragma unroll
      for (int a = 0; a < QPT; ++a) {
        qslot[a] = (half * QPT + a) * 128 + quarter * 32 + lane;
        const long long q = (long long)s * (128 * NQ) + qslot[a];
        qrow[a] = q < g.Q ? q : -1;
      }
      epi.template begin_segment<128 * NQ, QPT>(ea, g, s, part, qrow, qslot, lane, scratch);
      for (long long t = t0; t < t1; ++t, ++tc) {
        const uint32_t buf = tc & 1;
        epi.template pre_tile<BN, QPT>(ea, g, qslot, lane, scratch);
        mbar_wait(&acc_full[buf], (tc >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int a = 0; a < QPT; ++a) {
          const uint32_t taddr =
              tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + (buf * NQ + half * QPT + a) * BN;
          epi.template tile<BN>(ea, g, a, taddr, (unsigned long long)t * BN);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);
      }
      epi.template end_segment<128 * NQ, QPT>(ea, g, s, part, qslot, lane, scratch, /*last=*/(w + (t1 - t0)) >= w_end);
      w += t1 - t0;
    }
    Epi::epilogue_exit(scratch, lane);
  } else if (warp >= ST_HELP_WARP0) {
    Epi::template helper<128 * NQ, EPI_WARPS>(ea, g, warp - ST_HELP_WARP0, lane, scratch);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace b200

// 2-CTA variant of the streaming score kernel: a thread-block cluster of two CTAs (one SM pair) computes a
// 256-query x 256-row score tile per step with tcgen05.mma.cta_group::2 (M = 256: each CTA supplies its 128 query
// rows; N = 256: each CTA supplies 128 catalogue rows), so every operand byte staged in shared memory is used by two
// SMs — the single-CTA 128x128 form is bound by shared-memory operand bandwidth (128 B/cycle at the nominal MMA rate),
// this form needs 64 B/cycle.  Each CTA ends up with its 128 queries x 256 columns in its own TMEM (double buffered,
// 2 x 256 = 512 columns) and runs the same epilogue policies as stream_scores.cuh.
//
// Roles per CTA: warp 0 TMA producer (its half of every tile), warp 1 MMA issuer (leader CTA only), warp 2 TMEM
// allocator, warps 4..11 epilogue (thread = TMEM lane x 128-column half), warps 12..15 policy helper warps.
// Barriers: TMA of both CTAs completes on the LEADER's full barrier; tcgen05.commit multicasts to both CTAs' empty /
// accumulator-full barriers; epilogue warps of both CTAs arrive on the leader's accumulator-empty barrier.
#pragma once
#include "stream_scores.cuh"

namespace b200 {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA tile load of this CTA's half; completion bytes are credited to the barrier at `bar_cluster_addr` (leader's)
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// arrive (once all prior MMAs of this thread are complete) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"((uint16_t)3)
      : "memory");
}

// Two shapes of the pair tile (NQ2 = query tiles per CTA):
//   NQ2 = 1: 256 queries x 256 rows per step; an epilogue thread owns (TMEM lane, 128-column half)
//   NQ2 = 2: 512 queries x 128 rows per step (two M=256,N=128 MMAs share every catalogue tile: half the L2 -> SM
//            traffic per score, which is what bounds the NQ2 = 1 form at ~4.4 TB/s); a thread owns (lane, query tile)
constexpr int S2_SLOTS = 256;  // epilogue slots per CTA in both shapes
template <int NQ2> struct S2Shape {
  static constexpr int ROWS = NQ2 == 1 ? 256 : 128;      // catalogue rows per step for the pair
  static constexpr int QUERIES = 256 * NQ2;              // queries per supertile
  static constexpr int STAGE_BYTES = (ROWS / 2) * 128;   // this CTA's half of one 64-column k-block
  static constexpr int UMMA_N = ROWS;
};

// A ring stage holds g.ks 64-column atoms (ks = 2 when the row length allows it): one full-barrier wait and one
// commit per 8 MMAs instead of per 4 — the barrier handshake on the single MMA-issuing thread is what separates the
// fed pipeline (~1000 TFLOP/s) from the unfed issue rate (~1500 TFLOP/s).
template <int NQ2>
inline bool stream_geom2(StreamGeom& g, long long N, int Q, int KB, int sms, int scratch_bytes) {
  constexpr int S2_QUERIES = S2Shape<NQ2>::QUERIES, S2_ROWS = S2Shape<NQ2>::ROWS, S2_STAGE_BYTES = S2Shape<NQ2>::STAGE_BYTES;
  g.N = N;
  g.Q = Q;
  g.KB = KB;
  g.dbg_nofeed = 0;
  g.dbg_stats = 0;
  const StreamKnobs& skn = stream_knobs();
  g.mma_order = skn.mma_order;
  g.S = (Q + S2_QUERIES - 1) / S2_QUERIES;
  g.T = (N + S2_ROWS - 1) / S2_ROWS;
  g.total = g.S * g.T;
  long long units = sms / 2;
  if (units > g.total) units = g.total;
  g.W = (g.total + units - 1) / units;
  // Time-aligned schedule: with S supertiles and U units, R = U / S units sweep each supertile with stride R, so at any
  // moment all units read the same few catalogue tiles (one per stride slot) for different queries and every tile
  // comes from HBM once instead of once per supertile (ncu: 11.3 GB of DRAM reads per launch for a 2.56 GB catalogue
  // with the contiguous split).  The U - R*S leftover units share the last Tt tiles of every supertile so that all
  // units still do W tiles.  Falls back to the contiguous split when there are more supertiles than units.
  const int sch = skn.sched;
  const bool aligned = sch != 0 && g.S <= units && g.T >= 4 * (units / g.S);
  long long used;
  if (aligned) {
    g.R = (int)(units / g.S);
    long long extra = units - (long long)g.R * g.S;
    // A unit that changes supertile hands 256 candidate lists to its helper warps at once (hundreds of microseconds of
    // merge backlog), so when only a few units are left over they stay idle rather than become stragglers.
    // (~600 us per unit, measured); on long scans the extra 2.7 % of capacity is worth more than that tail (10 M rows:
    // 9.93 ms with the leftover units working vs 10.05 ms idle; 1.25 M rows: 2.49 vs 1.93 ms).
    const bool long_scan = g.W >= 6000 && sch != 4;
    if (extra * 16 <= units && !long_scan && sch != 3) extra = 0;
    g.n_main = g.R * g.S;
    g.Tmain = extra == 0 ? g.T : (g.R * g.W < g.T ? g.R * g.W : g.T);
    g.Tt = g.T - g.Tmain;
    g.We = (extra > 0 && g.Tt > 0) ? ((long long)g.S * g.Tt + extra - 1) / extra : 1;
    used = g.n_main + (g.Tt > 0 ? ((long long)g.S * g.Tt + g.We - 1) / g.We : 0);
  } else {
    g.R = 0;
    g.n_main = 0;
    g.Tmain = 0;
    g.Tt = g.T;
    g.We = g.W;
    used = (g.total + g.W - 1) / g.W;
  }
  g.grid = (int)(2 * used);
  {
    const long long lin = (g.T + g.W - 1) / g.W + 1;
    g.max_parts = (int)(aligned ? g.R + (units - g.n_main) + 1 : (lin < used ? lin : used));
  }
  const int q_bytes = NQ2 * KB * ST_QTILE_BYTES;
  const int avail = ST_SMEM_LIMIT - 1024 - ST_BAR_BYTES - scratch_bytes - q_bytes;
  g.ks = (KB % 2 == 0) ? 2 : 1;
  if (skn.ks == 1) g.ks = 1;
  int stages = avail / (S2_STAGE_BYTES * g.ks);
  if (stages > ST_MAX_STAGES) stages = ST_MAX_STAGES;
  g.stages = stages;
  g.smem_bytes = 1024 + q_bytes + stages * S2_STAGE_BYTES * g.ks + ST_BAR_BYTES + scratch_bytes;
  return stages >= 3;
}

// k-th segment (one supertile `s`, tiles t0, t0 + stride, ... (cnt of them)) of a work unit; false when there is none
__device__ __forceinline__ bool geom2_segment(const StreamGeom& g, long long unit, int k, int& s, long long& t0,
                                              long long& cnt, long long& stride, bool& last) {
  if (unit < g.n_main) {
    if (k > 0) return false;
    s = (int)(unit % g.S);
    const long long r = unit / g.S;
    t0 = r;
    stride = g.R;
    cnt = r < g.Tmain ? (g.Tmain - r + g.R - 1) / g.R : 0;
    last = true;
    return cnt > 0;
  }
  const long long x0 = (unit - g.n_main) * g.We, xt = (long long)g.S * g.Tt;
  long long x1 = x0 + g.We;
  if (x1 > xt) x1 = xt;
  if (x0 >= x1) return false;
  const long long sk = x0 / g.Tt + k;
  const long long a = x0 > sk * g.Tt ? x0 : sk * g.Tt;
  const long long b = x1 < (sk + 1) * g.Tt ? x1 : (sk + 1) * g.Tt;
  if (a >= b) return false;
  s = (int)sk;
  t0 = g.Tmain + (a - sk * g.Tt);
  cnt = b - a;
  stride = 1;
  last = b >= x1;
  return true;
}
// supertile of a unit's LAST segment (-1: the unit has no work); the final per-query kernel collects leftovers with it
__device__ __forceinline__ int geom2_last_super(const StreamGeom& g, long long unit) {
  if (unit < g.n_main) return unit / g.S < g.Tmain ? (int)(unit % g.S) : -1;
  const long long x0 = (unit - g.n_main) * g.We, xt = (long long)g.S * g.Tt;
  long long x1 = x0 + g.We;
  if (x1 > xt) x1 = xt;
  return x0 < x1 ? (int)((x1 - 1) / g.Tt) : -1;
}

template <int NQ2, class Epi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(ST_THREADS, 1)
stream_scores2_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_x,
                      const StreamGeom g, const typename Epi::Args ea) {
  constexpr int EPI_WARPS = 8;
  constexpr int S2_ROWS = S2Shape<NQ2>::ROWS, S2_QUERIES = S2Shape<NQ2>::QUERIES;
  constexpr int S2_STAGE_BYTES = S2Shape<NQ2>::STAGE_BYTES, UMMA_N = S2Shape<NQ2>::UMMA_N;
  constexpr int COLS = 128;  // accumulator columns one epilogue thread consumes per step
  // Accumulator hand-off granularity.  NQ2 = 2: the two query tiles of a step are separate accumulators; with one
  // ring stage per catalogue tile the MMAs are issued query-tile-major and each accumulator has its own full/empty
  // barrier pair (4 epilogue warps per CTA each), so the MMA of (tile t+2, query tile 0) only waits for the 8 warps
  // that read (t, 0) and the epilogue of (t, 0) starts half a tile earlier.  NQ2 = 1: one 256-column accumulator.
  constexpr int NSUB = NQ2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space (LDS / STS)
  uint8_t* q_smem = smem;
  uint8_t* ring = smem + NQ2 * g.KB * ST_QTILE_BYTES;
  const int ks = g.ks;                       // 64-column atoms per ring stage
  const int stage_bytes = S2_STAGE_BYTES * ks;
  const int nsteps = g.KB / ks;              // ring stages per catalogue tile
  uint8_t* tail = ring + g.stages * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + ST_MAX_STAGES;
  uint64_t* acc_full = empty_bar + ST_MAX_STAGES;
  uint64_t* acc_empty = acc_full + 4;   // [buffer][sub]: NQ2 = 2 hands each query tile's accumulator over separately
  uint64_t* q_full = acc_empty + 4;
  uint64_t* q_empty = q_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_empty + 1);
  uint32_t* scratch = reinterpret_cast<uint32_t*>(tail + ST_BAR_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const long long unit = blockIdx.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_x);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < g.stages; ++s) {
      mbar_init(&full_bar[s], 1);    // leader's: one arrive.expect_tx by the leader's producer, bytes from both CTAs
      mbar_init(&empty_bar[s], 1);   // one multicast commit
    }
    for (int b = 0; b < 4; ++b) {
      mbar_init(&acc_full[b], 1);                       // one multicast commit
      mbar_init(&acc_empty[b], 2 * EPI_WARPS / NSUB);   // leader's: the sub-buffer's epilogue warps of both CTAs
    }
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc2(tmem_slot, 512);
  if (warp == 3) Epi::init_scratch(scratch, lane);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // peer barriers are initialised before any remote arrive / TMA completion can target them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // development paths (pipeline counters, operand-less MMA issue, alternative MMA orders) exist only in the
  // instantiations of policies with kDbg = true; release kernels fold all of this to constants
  constexpr bool kDbg = Epi::kDbg;
  const bool dbg_st = kDbg && (g.dbg_stats == 1 || g.dbg_stats == 2 + (int)unit);
  const bool nofeed = kDbg && g.dbg_nofeed != 0;
  const int mma_order = kDbg ? g.mma_order : 0;
  int s;                       // segment: supertile, first tile, tile count, tile stride, last segment of the unit
  long long t0, ntile, tstep;
  bool last_seg;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (this CTA's halves)
    // warp-uniform loop, one elected lane issues (same reason as the MMA issuer below)
    {
      const uint32_t q_full_leader = mapa_u32(smem_u32(q_full), 0);
      uint32_t it = 0;
      for (int seg = 0; geom2_segment(g, unit, seg, s, t0, ntile, tstep, last_seg); ++seg) {
        mbar_wait(q_empty, (seg & 1) ^ 1);
        if (elect_one()) {
          if (leader) mbar_arrive_expect_tx(q_full, 2 * NQ2 * g.KB * ST_QTILE_BYTES);
          for (int qt = 0; qt < NQ2; ++qt)
            for (int kb = 0; kb < g.KB; ++kb)
              tma_load_2d_2sm(q_smem + (qt * g.KB + kb) * ST_QTILE_BYTES, &tmap_q, q_full_leader, kb * 64,
                              s * S2_QUERIES + qt * 256 + (int)rank * 128);
        }
        __syncwarp();
        if (!nofeed) {  // development knob off: MMA issue rate without any operand traffic
          for (long long ti = 0, t = t0; ti < ntile; ++ti, t += tstep) {
            for (int step = 0; step < nsteps; ++step, ++it) {
              const int st = it % g.stages;
              const uint32_t ph = (it / g.stages) & 1;
              mbar_wait(&empty_bar[st], ph ^ 1);
              if (elect_one()) {
                if (leader) mbar_arrive_expect_tx(&full_bar[st], 2 * stage_bytes);
                const uint32_t bar = mapa_u32(smem_u32(&full_bar[st]), 0);
                for (int a = 0; a < ks; ++a)
                  tma_load_2d_2sm(ring + st * stage_bytes + a * S2_STAGE_BYTES, &tmap_x, bar, (step * ks + a) * 64,
                                  (int)(t * S2_ROWS + rank * (S2_ROWS / 2)));
              }
              __syncwarp();
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA)
    // The whole warp walks the loop (warp-uniform control flow) and one ELECTED lane issues: with a `lane == 0`
    // branch ptxas wraps every tcgen05 instruction in an ELECT / R2UR waterfall loop (~19 SASS instructions per MMA,
    // more than the 64 cycles one M=256,N=128,K=16 MMA takes once the warp shares its scheduler with epilogue warps).
    // Descriptors are a constant high word plus (smem address >> 4): one add per operand per MMA.
    if (leader) {
      const uint32_t idesc = umma_idesc_bf16(256, UMMA_N);
      const uint64_t desc_base = umma_desc_k_sw128(0);
      const uint32_t q_lo = (smem_u32(q_smem) & 0x3FFFFu) >> 4;
      const uint32_t ring_lo = (smem_u32(ring) & 0x3FFFFu) >> 4;
      uint32_t it = 0, tc = 0;
      long long c_acc = 0, c_full = 0;
      for (int seg = 0; geom2_segment(g, unit, seg, s, t0, ntile, tstep, last_seg); ++seg) {
        mbar_wait(q_full, seg & 1);
        tc_fence_after();
        for (long long ti = 0; ti < ntile; ++ti, ++tc) {
          const uint32_t buf = tc & 1;
          const uint32_t acc_ph = ((tc >> 1) & 1) ^ 1;
          if (NSUB == 2 && nsteps == 1 && mma_order != 1 && mma_order != 3) {
            // query-tile-major: [wait operands] { wait acc(sub) ; 4*ks MMAs ; commit acc_full(sub) } x 2 ; release stage
            const int st = it % g.stages;
            const uint32_t ph = (it / g.stages) & 1;
            ++it;
            if (!nofeed) {
              const long long c2 = dbg_st ? clock64() : 0;
              mbar_wait(&full_bar[st], ph);
              if (dbg_st) c_full += clock64() - c2;
            }
#pragma unroll
            for (int qt = 0; qt < NSUB; ++qt) {
              const long long c0 = dbg_st ? clock64() : 0;
              if (mma_order == 2) {
                if (qt == 0) {
                  mbar_wait(&acc_empty[buf * 2 + 0], acc_ph);
                  mbar_wait(&acc_empty[buf * 2 + 1], acc_ph);
                }
              } else {
                mbar_wait(&acc_empty[buf * 2 + qt], acc_ph);
              }
              tc_fence_after();
              if (dbg_st) c_acc += clock64() - c0;
              if (elect_one()) {
                const uint32_t d_addr = tmem_base + buf * 256 + qt * 128;
                for (int a = 0; a < ks; ++a) {
                  const uint64_t b_desc = desc_base + (ring_lo + ((st * stage_bytes + a * S2_STAGE_BYTES) >> 4));
                  const uint64_t a_desc = desc_base + (q_lo + (((qt * g.KB + a) * ST_QTILE_BYTES) >> 4));
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    umma_bf16_2sm(d_addr, a_desc + 2 * k, b_desc + 2 * k, idesc, (a | k) != 0);
                }
                umma_commit_2sm(&acc_full[buf * 2 + qt]);
                if (qt == NSUB - 1 && !nofeed) umma_commit_2sm(&empty_bar[st]);
              }
              __syncwarp();
            }
            continue;
          }
          const long long c0 = dbg_st ? clock64() : 0;
          for (int sub = 0; sub < NSUB; ++sub) mbar_wait(&acc_empty[buf * 2 + sub], acc_ph);
          tc_fence_after();
          const long long c1 = dbg_st ? clock64() : 0;
          c_acc += c1 - c0;
          for (int step = 0; step < nsteps; ++step, ++it) {
            const int st = it % g.stages;
            const uint32_t ph = (it / g.stages) & 1;
            if (!nofeed) {
              const long long c2 = dbg_st ? clock64() : 0;
              mbar_wait(&full_bar[st], ph);
              tc_fence_after();
              if (dbg_st) c_full += clock64() - c2;
            }
            if (elect_one()) {
              for (int a = 0; a < ks; ++a) {
                const int kb = step * ks + a;
                const uint64_t b_desc = desc_base + (ring_lo + ((st * stage_bytes + a * S2_STAGE_BYTES) >> 4));
                if (mma_order == 3) {  // development variant: alternate the query tiles every MMA
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
#pragma unroll
                    for (int qt = 0; qt < NQ2; ++qt) {
                      const uint64_t a_desc = desc_base + (q_lo + (((qt * g.KB + kb) * ST_QTILE_BYTES) >> 4));
                      umma_bf16_2sm(tmem_base + buf * 256 + qt * 128, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
                    }
                  }
                  continue;
                }
#pragma unroll
                for (int qt = 0; qt < NQ2; ++qt) {
                  const uint64_t a_desc = desc_base + (q_lo + (((qt * g.KB + kb) * ST_QTILE_BYTES) >> 4));
                  const uint32_t d_addr = tmem_base + buf * 256 + qt * 128;  // NQ2 = 1: one 256-column accumulator
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    umma_bf16_2sm(d_addr, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
                }
              }
              if (!nofeed) umma_commit_2sm(&empty_bar[st]);
              if (step == nsteps - 1) {
                for (int sub = 0; sub < NSUB; ++sub) umma_commit_2sm(&acc_full[buf * 2 + sub]);
              }
            }
            __syncwarp();
          }
        }
        if (elect_one()) umma_commit_2sm(q_empty);
        __syncwarp();
      }
      if (dbg_st && lane == 0) {
        atomicAdd(&g_stream_stats[0], (unsigned long long)c_acc);
        atomicAdd(&g_stream_stats[1], (unsigned long long)c_full);
      }
    }
  } else if (warp >= ST_EPI_WARP0 && warp < ST_EPI_WARP0 + EPI_WARPS) {
    // ------------------------------------------------------------------ epilogue: lane x 128-column half
    const int ew = warp - ST_EPI_WARP0;
    const int quarter = warp & 3;
    const int half = ew >> 2;
    Epi epi;
    uint32_t tc = 0;
    long long c_wait = 0, c_proc = 0;
    const int sub = NSUB == 2 ? half : 0;
    const uint32_t acc_empty_leader[2] = {mapa_u32(smem_u32(&acc_empty[sub]), 0), mapa_u32(smem_u32(&acc_empty[2 + sub]), 0)};
    for (int seg = 0; geom2_segment(g, unit, seg, s, t0, ntile, tstep, last_seg); ++seg) {
      const int part = 0;
      // `half` selects the 128-column half of the one accumulator (NQ2 = 1) or the query tile (NQ2 = 2)
      int qslot[1] = {half * 128 + quarter * 32 + lane};
      const long long q = (long long)s * S2_QUERIES + (NQ2 == 2 ? half * 256 : 0) + rank * 128 + quarter * 32 + lane;
      long long qrow[1] = {q < g.Q ? q : -1};
      epi.template begin_segment<S2_SLOTS, 1>(ea, g, s, part, qrow, qslot, lane, scratch);
      for (long long ti = 0, t = t0; ti < ntile; ++ti, t += tstep, ++tc) {
        const uint32_t buf = tc & 1;
        epi.template pre_tile<COLS, 1>(ea, g, qslot, lane, scratch);
        const long long c0 = dbg_st ? clock64() : 0;
        mbar_wait(&acc_full[buf * 2 + sub], (tc >> 1) & 1);
        tc_fence_after();
        const long long c1 = dbg_st ? clock64() : 0;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + buf * 256 + half * 128;
        epi.template tile<COLS>(ea, g, 0, taddr, (unsigned long long)t * S2_ROWS + (NQ2 == 1 ? half * 128 : 0));
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(acc_empty_leader[buf]);
        if (dbg_st) {
          c_wait += c1 - c0;
          c_proc += clock64() - c1;
        }
      }
      epi.template end_segment<S2_SLOTS, 1>(ea, g, s, part, qslot, lane, scratch, last_seg);
    }
    if (dbg_st && lane == 0) {
      atomicAdd(&g_stream_stats[2], (unsigned long long)c_wait);
      atomicAdd(&g_stream_stats[3], (unsigned long long)c_proc);
      atomicAdd(&g_stream_stats[4], (unsigned long long)tc);
    }
    Epi::epilogue_exit(scratch, lane);
  } else if (warp >= ST_HELP_WARP0) {
    Epi::template helper<S2_SLOTS, EPI_WARPS>(ea, g, warp - ST_HELP_WARP0, lane, scratch);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA may exit (or free TMEM) while its peer can still signal it or read its operands
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

}  // namespace b200

// Exact inner-product top-K (K4): catalogue-scoring GEMM on tcgen05 fused with a threshold-filtered select.
// Replaces faiss.IndexFlatIP.search (reference src/serving/retrieval.py:171) and the np.dot+argsort eval twin
// (scripts/evaluate_model.py:217-232, src/evaluation/metrics.py:381-396).
//
// Per (CTA, query): a running threshold tau (= current k-th best of what this CTA has seen) lives in a register of
// the thread that owns the query's TMEM lane; a score enters the query's candidate buffer only if it beats tau
// (strictly: at equal score the earlier = lower row id wins, faiss' heap rule).  When a buffer is about to overflow a
// warp radix-selects its k best 64-bit keys (ordered score << 32 | ~row) and raises tau.  At segment end the k
// survivors are published; a second small kernel merges the <= max_parts partial lists of every query and sorts them
// (score desc, row asc).  Scores are never written to HBM.
#include <cfloat>
#include "stream_scores.cuh"
#include "../../include/b200rec.h"

namespace b200 {

// ---------------------------------------------------------------------------------------------------------------
// warp-level radix select helpers
// ---------------------------------------------------------------------------------------------------------------
// hist[256] holds digit counts; find digit d with  count(digits > d) < need <= count(digits >= d).
// Returns d; need is reduced by count(digits > d); bucket = hist[d].  Called by a full warp.
__device__ __forceinline__ int warp_pick_digit(const uint32_t* hist, int& need, int& bucket, int lane) {
  uint32_t c[8];
  uint32_t lane_total = 0;
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    c[b] = hist[lane * 8 + b];
    lane_total += c[b];
  }
  uint32_t incl = lane_total;  // sum over lanes >= lane
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_down_sync(FULL_MASK, incl, o);
    if (lane + o < 32) incl += t;
  }
  const uint32_t above = incl - lane_total;
  const bool mine = (above < (uint32_t)need) && ((uint32_t)need <= incl);
  int d = 0, nn = 0, bk = 0;
  if (mine) {
    uint32_t acc = above;
#pragma unroll
    for (int b = 7; b >= 0; --b) {
      if (acc + c[b] >= (uint32_t)need) {
        d = lane * 8 + b;
        nn = need - (int)acc;
        bk = (int)c[b];
        break;
      }
      acc += c[b];
    }
  }
  const uint32_t who = __ballot_sync(FULL_MASK, mine);
  const int src = __ffs(who) - 1;  // exactly one lane when need <= total
  d = __shfl_sync(FULL_MASK, d, src);
  need = __shfl_sync(FULL_MASK, nn, src);
  bucket = __shfl_sync(FULL_MASK, bk, src);
  return d;
}

// Keep the k largest of buf[0..n) (n > k, keys distinct), compacted to buf[0..k); returns the smallest kept key.
__device__ __noinline__ uint64_t warp_compact_topk(uint64_t* buf, int n, int k, uint32_t* hist, int lane) {
  uint64_t prefix = 0;
  int need = k, shift = 56;
  for (int pass = 0; pass < 8; ++pass, shift -= 8) {
    for (int i = lane; i < 256; i += 32) hist[i] = 0;
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
      const uint64_t key = __ldcg(buf + i);
      const bool match = (pass == 0) || ((key >> (shift + 8)) == prefix);
      if (match) atomicAdd(&hist[(uint32_t)(key >> shift) & 255u], 1u);
    }
    __syncwarp();
    int bucket;
    const int d = warp_pick_digit(hist, need, bucket, lane);
    prefix = (prefix << 8) | (uint64_t)d;
    __syncwarp();
    if (bucket == need) break;
  }
  if (shift < 0) shift = 0;
  int base = 0;
  uint64_t minkey = ~0ull;
  const uint32_t lt_mask = (1u << lane) - 1u;
  for (int i0 = 0; i0 < n; i0 += 32) {
    const int i = i0 + lane;
    const uint64_t key = (i < n) ? __ldcg(buf + i) : 0ull;
    const bool keep = (i < n) && ((key >> shift) >= prefix);
    const uint32_t b = __ballot_sync(FULL_MASK, keep);
    __syncwarp();
    if (keep) {
      __stcg(buf + base + __popc(b & lt_mask), key);
      minkey = key < minkey ? key : minkey;
    }
    base += __popc(b);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const uint64_t t = __shfl_xor_sync(FULL_MASK, minkey, o);
    minkey = t < minkey ? t : minkey;
  }
  __syncwarp();
  return minkey;
}

// ---------------------------------------------------------------------------------------------------------------
// top-K epilogue policy for stream_scores_kernel
// ---------------------------------------------------------------------------------------------------------------
struct TopkEpi {
  struct Args {
    uint64_t* cand;                 // [grid][128*NQ][cap]
    uint64_t* parts;                // [S][max_parts][128*NQ][k]
    const int64_t* excl_indptr;     // [Q+1] or null
    const int32_t* excl_rows;       // sorted per query
    int k;
    int cap;
  };
  float tau[2];
  int cnt[2];
  uint64_t* buf[2];
  const int32_t* ex_lo[2];
  int ex_n[2];

  template <int NQ, int QPT>
  __device__ __forceinline__ void begin_segment(const Args& ea, const StreamGeom& g, int s, const int (&qslot)[QPT],
                                                int lane) {
#pragma unroll
    for (int a = 0; a < QPT; ++a) {
      const long long q = (long long)s * 128 * NQ + qslot[a];
      tau[a] = (q < g.Q) ? -FLT_MAX : INFINITY;
      cnt[a] = 0;
      buf[a] = ea.cand + ((size_t)blockIdx.x * (128 * NQ) + qslot[a]) * ea.cap;
      ex_lo[a] = nullptr;
      ex_n[a] = 0;
      if (ea.excl_indptr != nullptr && q < g.Q) {
        const int64_t lo = ea.excl_indptr[q], hi = ea.excl_indptr[q + 1];
        ex_lo[a] = ea.excl_rows + lo;
        ex_n[a] = (int)(hi - lo);
      }
    }
  }

  // make room for a worst-case tile (BN appends) in every owned buffer
  template <int NQ, int BN, int QPT>
  __device__ __forceinline__ void pre_tile(const Args& ea, const StreamGeom& g, const int (&qslot)[QPT], int lane,
                                           uint32_t* hist) {
#pragma unroll
    for (int a = 0; a < QPT; ++a) {
      uint32_t todo = __ballot_sync(FULL_MASK, cnt[a] + BN > ea.cap);
      while (todo) {
        const int owner = __ffs(todo) - 1;
        todo &= todo - 1;
        uint64_t* b = reinterpret_cast<uint64_t*>(
            __shfl_sync(FULL_MASK, reinterpret_cast<unsigned long long>(buf[a]), owner));
        const int n = __shfl_sync(FULL_MASK, cnt[a], owner);
        __syncwarp();
        const uint64_t minkey = warp_compact_topk(b, n, ea.k, hist, lane);
        if (lane == owner) {
          cnt[a] = ea.k;
          tau[a] = ord_f32((uint32_t)(minkey >> 32));
        }
      }
    }
  }

  __device__ __forceinline__ bool excluded(int a, uint32_t row) const {
    int lo = 0, hi = ex_n[a];
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      const int32_t v = __ldg(ex_lo[a] + mid);
      if ((uint32_t)v < row)
        lo = mid + 1;
      else
        hi = mid;
    }
    return lo < ex_n[a] && (uint32_t)__ldg(ex_lo[a] + lo) == row;
  }

  __device__ __forceinline__ void consider(int a, float x, unsigned long long row, const StreamGeom& g) {
    if (x > tau[a] && row < (unsigned long long)g.N) {
      if (ex_n[a] == 0 || !excluded(a, (uint32_t)row)) {
        __stcg(buf[a] + cnt[a], make_key(x, (uint32_t)row));
        ++cnt[a];
      }
    }
  }

  template <int BN>
  __device__ __forceinline__ void tile(const Args& ea, const StreamGeom& g, int a, uint32_t taddr,
                                       unsigned long long row0) {
#pragma unroll 1
    for (int c = 0; c < BN; c += 64) {
      uint32_t v0[32], v1[32];
      tmem_ld_32x32(taddr + c, v0);
      tmem_ld_32x32(taddr + c + 32, v1);
      tmem_ld_wait();
      filter32(a, v0, row0 + c, g);
      filter32(a, v1, row0 + c + 32, g);
    }
  }

  __device__ __forceinline__ void filter32(int a, const uint32_t (&v)[32], unsigned long long row0,
                                           const StreamGeom& g) {
    float gm[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float m0 = fmax3(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1]), __uint_as_float(v[8 * j + 2]));
      const float m1 = fmax3(__uint_as_float(v[8 * j + 3]), __uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]));
      gm[j] = fmax3(m0, m1, fmaxf(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7])));
    }
    const float m = fmaxf(fmaxf(gm[0], gm[1]), fmaxf(gm[2], gm[3]));
    if (m > tau[a]) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (gm[j] > tau[a]) {
#pragma unroll
          for (int i = 0; i < 8; ++i) consider(a, __uint_as_float(v[8 * j + i]), row0 + 8 * j + i, g);
        }
      }
    }
  }

  template <int NQ, int QPT>
  __device__ __forceinline__ void end_segment(const Args& ea, const StreamGeom& g, int s, int part,
                                              const int (&qslot)[QPT], int lane, uint32_t* hist) {
#pragma unroll
    for (int a = 0; a < QPT; ++a) {
      for (int owner = 0; owner < 32; ++owner) {
        uint64_t* b = reinterpret_cast<uint64_t*>(
            __shfl_sync(FULL_MASK, reinterpret_cast<unsigned long long>(buf[a]), owner));
        int n = __shfl_sync(FULL_MASK, cnt[a], owner);
        const int slot = __shfl_sync(FULL_MASK, qslot[a], owner);
        const long long q = (long long)s * 128 * NQ + slot;
        if (q >= g.Q) continue;  // warp-uniform
        __syncwarp();
        if (n > ea.k) {
          warp_compact_topk(b, n, ea.k, hist, lane);
          n = ea.k;
        }
        uint64_t* dst = ea.parts + (((size_t)s * g.max_parts + part) * (128 * NQ) + slot) * ea.k;
        for (int i = lane; i < ea.k; i += 32) dst[i] = (i < n) ? __ldcg(b + i) : 0ull;
      }
    }
  }
};

// ---------------------------------------------------------------------------------------------------------------
// block-level exact select + sort of `n` keys produced by a loader; one CTA (256 threads) per query
// ---------------------------------------------------------------------------------------------------------------
constexpr int FIN_THREADS = 256;

template <class Loader>
__device__ void block_select_sort(const Loader& load, int n, int k_out, uint64_t* skeys /*[pow2 >= k_out]*/, int P,
                                  uint32_t* hist /*[256]*/, int* s_misc /*[8]*/) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // count valid keys
  if (tid == 0) s_misc[0] = 0;
  __syncthreads();
  int local = 0;
  for (int i = tid; i < n; i += FIN_THREADS) local += load(i) != 0ull;
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(FULL_MASK, local, o);
  if (lane == 0 && local) atomicAdd(&s_misc[0], local);
  __syncthreads();
  const int n_valid = s_misc[0];
  uint64_t prefix = 1;
  int shift = 0;
  if (n_valid > k_out) {
    prefix = 0;
    shift = 56;
    int need = k_out;
    for (int pass = 0; pass < 8; ++pass, shift -= 8) {
      hist[tid] = 0;
      __syncthreads();
      for (int i = tid; i < n; i += FIN_THREADS) {
        const uint64_t key = load(i);
        const bool match = (pass == 0) || ((key >> (shift + 8)) == prefix);
        if (match && key != 0ull) atomicAdd(&hist[(uint32_t)(key >> shift) & 255u], 1u);
      }
      __syncthreads();
      if (warp == 0) {
        int bucket;
        const int d = warp_pick_digit(hist, need, bucket, lane);
        if (lane == 0) {
          s_misc[1] = d;
          s_misc[2] = need;
          s_misc[3] = bucket;
        }
      }
      __syncthreads();
      prefix = (prefix << 8) | (uint64_t)s_misc[1];
      need = s_misc[2];
      const int bucket = s_misc[3];
      __syncthreads();
      if (bucket == need) break;
    }
    if (shift < 0) shift = 0;
  }
  // gather survivors, pad, bitonic sort descending
  if (tid == 0) s_misc[4] = 0;
  for (int i = tid; i < P; i += FIN_THREADS) skeys[i] = 0ull;
  __syncthreads();
  for (int i = tid; i < n; i += FIN_THREADS) {
    const uint64_t key = load(i);
    if (key != 0ull && (key >> shift) >= prefix) {
      const int pos = atomicAdd(&s_misc[4], 1);
      if (pos < P) skeys[pos] = key;
    }
  }
  __syncthreads();
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < P / 2; i += FIN_THREADS) {
        const int lo = 2 * i - (i & (stride - 1));  // index with bit `stride` cleared
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const uint64_t a = skeys[lo], b = skeys[hi];
        if ((a < b) == desc) {
          skeys[lo] = b;
          skeys[hi] = a;
        }
      }
      __syncthreads();
    }
  }
}

__device__ __forceinline__ void write_result(const uint64_t* skeys, int k_out, int64_t row_offset, float* out_s,
                                             int64_t* out_i) {
  for (int i = threadIdx.x; i < k_out; i += FIN_THREADS) {
    const uint64_t key = skeys[i];
    if (key == 0ull) {
      out_s[i] = -FLT_MAX;
      out_i[i] = -1;
    } else {
      out_s[i] = ord_f32((uint32_t)(key >> 32));
      out_i[i] = (int64_t)(~(uint32_t)key) + row_offset;
    }
  }
}

struct PartsLoader {
  const uint64_t* base;
  size_t part_stride;
  int k;
  __device__ __forceinline__ uint64_t operator()(int i) const {
    const int p = i / k, j = i - p * k;
    return __ldcg(base + (size_t)p * part_stride + j);
  }
};

template <int NQ>
__global__ void __launch_bounds__(FIN_THREADS)
topk_finalize_kernel(const StreamGeom g, const uint64_t* __restrict__ parts, int k, int P, int64_t row_offset,
                     float* __restrict__ out_scores, int64_t* __restrict__ out_ids) {
  extern __shared__ uint64_t fin_smem[];
  __shared__ uint32_t hist[256];
  __shared__ int s_misc[8];
  const int q = blockIdx.x;
  const int s = q / (128 * NQ), slot = q - s * 128 * NQ;
  const int nparts = geom_last_cta(g, s) - geom_first_cta(g, s) + 1;
  PartsLoader ld;
  ld.part_stride = (size_t)(128 * NQ) * k;
  ld.base = parts + ((size_t)s * g.max_parts * (128 * NQ) + slot) * k;
  ld.k = k;
  block_select_sort(ld, nparts * k, k, fin_smem, P, hist, s_misc);
  write_result(fin_smem, k, row_offset, out_scores + (size_t)q * k, out_ids + (size_t)q * k);
}

struct ListLoader {
  const float* scores;
  const int64_t* ids;
  size_t part_stride;  // Q * k_in
  int k_in;
  __device__ __forceinline__ uint64_t operator()(int i) const {
    const int p = i / k_in, j = i - p * k_in;
    const size_t off = (size_t)p * part_stride + j;
    const int64_t id = ids[off];
    return id < 0 ? 0ull : make_key(scores[off], (uint32_t)id);
  }
};

__global__ void __launch_bounds__(FIN_THREADS)
topk_merge_kernel(const float* __restrict__ scores, const int64_t* __restrict__ ids, int parts, long long Q, int k_in,
                  int k_out, int P, float* __restrict__ out_scores, int64_t* __restrict__ out_ids) {
  extern __shared__ uint64_t fin_smem[];
  __shared__ uint32_t hist[256];
  __shared__ int s_misc[8];
  const long long q = blockIdx.x;
  ListLoader ld;
  ld.scores = scores + (size_t)q * k_in;
  ld.ids = ids + (size_t)q * k_in;
  ld.part_stride = (size_t)Q * k_in;
  ld.k_in = k_in;
  block_select_sort(ld, parts * k_in, k_out, fin_smem, P, hist, s_misc);
  write_result(fin_smem, k_out, 0, out_scores + (size_t)q * k_out, out_ids + (size_t)q * k_out);
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

struct TopkPlan {
  int nq, bn;
  StreamGeom g;
  int cap;
  size_t cand_bytes, parts_bytes;
};

static int cap_for(int k, int bn) { return ((2 * k + bn + 31) / 32) * 32; }

static int plan_topk(TopkPlan& p, int64_t N, int64_t ld, int64_t Q, int k) {
  if (N <= 0 || Q <= 0) return fail("topk: empty catalogue or query set");
  if (N >= (1ll << 32) - 1) return fail("topk: a shard holds at most 2^32-2 rows");
  if (Q > INT32_MAX / 2) return fail("topk: too many queries");
  if (k < 1 || k > 2048) return fail("topk: k must be in [1, 2048] (got %d)", k);
  if (ld <= 0 || (ld % 64)) return fail("topk: leading dimension must be a positive multiple of 64 (got %lld)", (long long)ld);
  const int KB = (int)(ld / 64);
  const int sms = num_sms();
  const int q128 = (int)((Q + 127) / 128);
  bool ok = false;
  if (q128 >= 3 && stream_geom<4, 64>(p.g, N, (int)Q, KB, sms) && p.g.stages >= 4) {
    p.nq = 4, p.bn = 64, ok = true;
  } else if (q128 >= 2 && stream_geom<2, 128>(p.g, N, (int)Q, KB, sms) && p.g.stages >= 3) {
    p.nq = 2, p.bn = 128, ok = true;
  } else if (stream_geom<1, 256>(p.g, N, (int)Q, KB, sms) && p.g.stages >= 3) {
    p.nq = 1, p.bn = 256, ok = true;
  } else if (stream_geom<1, 64>(p.g, N, (int)Q, KB, sms)) {
    p.nq = 1, p.bn = 64, ok = true;
  }
  if (!ok) return fail("topk: dimension too large for the resident query tile (ld=%lld)", (long long)ld);
  p.cap = cap_for(k, p.bn);
  p.cand_bytes = (size_t)p.g.grid * 128 * p.nq * p.cap * sizeof(uint64_t);
  p.parts_bytes = (size_t)p.g.S * p.g.max_parts * 128 * p.nq * k * sizeof(uint64_t);
  return 0;
}

template <int NQ, int BN>
static int launch_topk(const TopkPlan& p, const void* catalogue, int64_t N, int64_t ld, const void* queries, int64_t Q,
                       int k, int64_t row_offset, const int64_t* excl_indptr, const int32_t* excl_rows,
                       float* out_scores, int64_t* out_ids, void* workspace, cudaStream_t st) {
  CUtensorMap tq, tx;
  if (make_tmap_bf16_2d(&tq, queries, (uint64_t)Q, (uint64_t)ld, (uint64_t)ld, 128)) return 1;
  if (make_tmap_bf16_2d(&tx, catalogue, (uint64_t)N, (uint64_t)ld, (uint64_t)ld, BN)) return 1;
  TopkEpi::Args ea;
  ea.cand = reinterpret_cast<uint64_t*>(workspace);
  ea.parts = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(workspace) + p.cand_bytes);
  ea.excl_indptr = excl_indptr;
  ea.excl_rows = excl_rows;
  ea.k = k;
  ea.cap = p.cap;
  auto kern = stream_scores_kernel<NQ, BN, TopkEpi>;
  static int smem_set = 0;
  if (smem_set < p.g.smem_bytes) {
    B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM_LIMIT));
    smem_set = ST_SMEM_LIMIT;
  }
  kern<<<p.g.grid, ST_THREADS, p.g.smem_bytes, st>>>(tq, tx, p.g, ea);
  B200_LAUNCH_OK("stream_scores_kernel<topk>");
  const int P = next_pow2(k);
  topk_finalize_kernel<NQ><<<(unsigned)Q, FIN_THREADS, P * sizeof(uint64_t), st>>>(p.g, ea.parts, k, P, row_offset,
                                                                                  out_scores, out_ids);
  B200_LAUNCH_OK("topk_finalize_kernel");
  return 0;
}

}  // namespace b200

extern "C" size_t b200rec_topk_workspace_bytes(int64_t N, int64_t ld, int64_t Q, int k) {
  b200::TopkPlan p;
  if (b200::plan_topk(p, N, ld, Q, k)) return 0;
  return p.cand_bytes + p.parts_bytes;
}

extern "C" int b200rec_flat_ip_topk(const void* catalogue, int64_t N, int64_t ld, const void* queries, int64_t Q, int k,
                                    int64_t row_offset, const int64_t* exclude_indptr, const int32_t* exclude_rows,
                                    float* out_scores, int64_t* out_ids, void* workspace, size_t workspace_bytes,
                                    void* stream) {
  using namespace b200;
  if (!catalogue || !queries || !out_scores || !out_ids || !workspace) return fail("topk: null pointer");
  TopkPlan p;
  if (plan_topk(p, N, ld, Q, k)) return 1;
  if (workspace_bytes < p.cand_bytes + p.parts_bytes)
    return fail("topk: workspace too small (%zu < %zu)", workspace_bytes, p.cand_bytes + p.parts_bytes);
  if ((exclude_indptr == nullptr) != (exclude_rows == nullptr)) return fail("topk: exclusion CSR needs both arrays");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (p.nq == 4) return launch_topk<4, 64>(p, catalogue, N, ld, queries, Q, k, row_offset, exclude_indptr, exclude_rows, out_scores, out_ids, workspace, st);
  if (p.nq == 2) return launch_topk<2, 128>(p, catalogue, N, ld, queries, Q, k, row_offset, exclude_indptr, exclude_rows, out_scores, out_ids, workspace, st);
  if (p.bn == 256) return launch_topk<1, 256>(p, catalogue, N, ld, queries, Q, k, row_offset, exclude_indptr, exclude_rows, out_scores, out_ids, workspace, st);
  return launch_topk<1, 64>(p, catalogue, N, ld, queries, Q, k, row_offset, exclude_indptr, exclude_rows, out_scores, out_ids, workspace, st);
}

extern "C" int b200rec_topk_merge(const float* scores, const int64_t* ids, int parts, int64_t Q, int k_in, int k_out,
                                  float* out_scores, int64_t* out_ids, void* stream) {
  using namespace b200;
  if (!scores || !ids || !out_scores || !out_ids) return fail("topk_merge: null pointer");
  if (parts < 1 || Q < 1 || k_in < 1) return fail("topk_merge: empty input");
  if (k_out < 1 || k_out > 2048) return fail("topk_merge: k_out must be in [1, 2048]");
  if ((int64_t)parts * k_in > INT32_MAX) return fail("topk_merge: too many candidates per query");
  const int P = next_pow2(k_out);
  topk_merge_kernel<<<(unsigned)Q, FIN_THREADS, P * sizeof(uint64_t), reinterpret_cast<cudaStream_t>(stream)>>>(
      scores, ids, parts, (long long)Q, k_in, k_out, P, out_scores, out_ids);
  B200_LAUNCH_OK("topk_merge_kernel");
  return 0;
}

// On-device ranking metrics of the offline evaluation: the per-user Python loops of
// Evaluator.evaluate (reference src/evaluation/metrics.py:240-319, helpers :74-231: recall / precision / NDCG / hit
// rate at several k, reciprocal rank, average precision, coverage) as ONE kernel over the [Q, K] id matrix the fused
// top-K kernel produced — the recommendation lists never visit the host.
//
// One warp per user.  32 positions per pass: each lane binary-searches its id in the user's sorted ground-truth list
// (CSR), a ballot gives the hit mask of the pass, and the warp walks the set bits in ASCENDING position order so that
// every fp64 accumulation (dcg += 1/log2(i+2), ap += hits/(i+1)) happens in the order of the reference's Python loop:
// per-user values are bit-identical to the reference's (the 1/log2 and ideal-DCG tables come from the host, built with
// the reference's own expressions).  Integer-byte work, HBM/latency-bound: 8·Q·K bytes of ids + the CSR.
#include "host_util.h"
#include "tc_common.cuh"
#include "../../include/b200rec.h"

namespace b200 {

constexpr int MAX_EVAL_K = 8;  // k values per call ([5,10,20,50,100] in the reference)

struct EvalKs {
  int32_t k[MAX_EVAL_K];
  int n;
};

__device__ __forceinline__ bool sorted_contains(const int64_t* __restrict__ v, int64_t n, int64_t x) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (__ldg(v + mid) < x) lo = mid + 1;
    else hi = mid;
  }
  return lo < n && __ldg(v + lo) == x;
}

__global__ void __launch_bounds__(256)
eval_metrics_kernel(const int64_t* __restrict__ pred, int64_t Q, int K, int64_t ld_pred,
                    const uint8_t* __restrict__ repeat, int64_t ld_rep, const int64_t* __restrict__ gt_indptr,
                    const int64_t* __restrict__ gt_rows, const int64_t* __restrict__ gt_count, const EvalKs ks,
                    const double* __restrict__ inv_log2, const double* __restrict__ idcg, double* __restrict__ per_user,
                    double* __restrict__ sums, uint32_t* __restrict__ coverage, int64_t n_items, int cov_k) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int ncol = 4 * ks.n + 2;
  for (int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); q < Q; q += warps) {
    const int64_t g0 = __ldg(gt_indptr + q), g1 = __ldg(gt_indptr + q + 1);
    const int64_t n_gt = gt_count ? __ldg(gt_count + q) : (g1 - g0);  // |ground truth| incl. items outside the catalogue
    int set_hits[MAX_EVAL_K];   // |set(pred[:k]) & gt|: repeated ids count once (recall / precision / hit rate)
    double dcg[MAX_EVAL_K];     // positional, as the reference's loop (NDCG)
#pragma unroll
    for (int j = 0; j < MAX_EVAL_K; ++j) { set_hits[j] = 0; dcg[j] = 0.0; }
    double rr = 0.0, ap = 0.0;
    int pos_hits = 0;
    for (int p0 = 0; p0 < K; p0 += 32) {
      const int p = p0 + lane;
      int64_t id = -1;
      if (p < K) id = __ldg(pred + q * ld_pred + p);
      const bool hit = id >= 0 && sorted_contains(gt_rows + g0, g1 - g0, id);
      const bool rep = hit && repeat && __ldg(repeat + q * ld_rep + p) != 0;
      if (coverage && id >= 0 && id < n_items && p < cov_k) atomicOr(coverage + (id >> 5), 1u << (id & 31));
      unsigned m = __ballot_sync(FULL_MASK, hit);
      const unsigned mrep = __ballot_sync(FULL_MASK, rep);
      while (m) {  // warp-uniform walk over the hits of this pass, ascending position
        const int b = __ffs(m) - 1;
        m &= m - 1;
        const int i = p0 + b;
        const bool first = ((mrep >> b) & 1u) == 0;
        ++pos_hits;
        if (rr == 0.0) rr = 1.0 / (double)(i + 1);
        ap += (double)pos_hits / (double)(i + 1);
        const double w = __ldg(inv_log2 + i);
#pragma unroll
        for (int j = 0; j < MAX_EVAL_K; ++j) {
          if (j < ks.n && i < ks.k[j]) {
            dcg[j] += w;
            if (first) ++set_hits[j];
          }
        }
      }
    }
    if (lane == 0) {
      double* o = per_user + q * ncol;
      const double ngt = (double)n_gt;
      for (int j = 0; j < ks.n; ++j) {
        const int k = ks.k[j];
        const int64_t ideal = n_gt < k ? n_gt : k;
        const double id_dcg = __ldg(idcg + ideal);
        o[4 * j + 0] = n_gt > 0 ? (double)set_hits[j] / ngt : 0.0;
        o[4 * j + 1] = (double)set_hits[j] / (double)k;
        o[4 * j + 2] = (n_gt > 0 && id_dcg != 0.0) ? dcg[j] / id_dcg : 0.0;
        o[4 * j + 3] = set_hits[j] > 0 ? 1.0 : 0.0;
      }
      o[4 * ks.n] = rr;
      o[4 * ks.n + 1] = n_gt > 0 ? ap / ngt : 0.0;
      if (sums)
        for (int c = 0; c < ncol; ++c) atomicAdd(sums + c, o[c]);
    }
  }
}

__global__ void popcount_kernel(const uint32_t* __restrict__ words, int64_t n, unsigned long long* __restrict__ out) {
  unsigned long long acc = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    acc += __popc(__ldg(words + i));
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(FULL_MASK, acc, o);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

}  // namespace b200

using namespace b200;

extern "C" int b200rec_eval_metrics(const int64_t* pred, int64_t Q, int K, int64_t ld_pred, const uint8_t* repeat,
                                    int64_t ld_repeat, const int64_t* gt_indptr, const int64_t* gt_rows,
                                    const int64_t* gt_count, const int32_t* k_values_host, int n_k,
                                    const double* inv_log2, const double* idcg, double* per_user, double* sums,
                                    uint32_t* coverage_bits, int64_t n_items, int coverage_k,
                                    unsigned long long* coverage_count, void* stream) {
  if (!pred || !gt_indptr || !gt_rows || !k_values_host || !inv_log2 || !idcg || !per_user)
    return fail("eval_metrics: null pointer");
  if (Q <= 0 || K <= 0) return fail("eval_metrics: empty prediction matrix");
  if (n_k <= 0 || n_k > MAX_EVAL_K) return fail("eval_metrics: 1..%d k values per call (got %d)", MAX_EVAL_K, n_k);
  EvalKs ks;
  ks.n = n_k;
  for (int j = 0; j < MAX_EVAL_K; ++j) ks.k[j] = j < n_k ? k_values_host[j] : 0;
  for (int j = 0; j < n_k; ++j)
    if (ks.k[j] <= 0) return fail("eval_metrics: k values must be positive");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int ncol = 4 * n_k + 2;
  if (sums) B200_CUDA_OK(cudaMemsetAsync(sums, 0, sizeof(double) * ncol, st));
  const int64_t words = (n_items + 31) / 32;
  if (coverage_bits) {
    if (n_items <= 0 || !coverage_count) return fail("eval_metrics: coverage needs n_items and a count output");
    B200_CUDA_OK(cudaMemsetAsync(coverage_bits, 0, sizeof(uint32_t) * words, st));
    B200_CUDA_OK(cudaMemsetAsync(coverage_count, 0, sizeof(unsigned long long), st));
  }
  const int64_t blocks = (Q + 7) / 8;
  const int grid = (int)(blocks < (int64_t)num_sms() * 8 ? blocks : (int64_t)num_sms() * 8);
  eval_metrics_kernel<<<grid, 256, 0, st>>>(pred, Q, K, ld_pred, repeat, ld_repeat, gt_indptr, gt_rows, gt_count, ks,
                                            inv_log2, idcg, per_user, sums, coverage_bits, n_items, coverage_k);
  B200_LAUNCH_OK("eval_metrics_kernel");
  if (coverage_bits) {
    const int64_t pb = (words + 255) / 256;
    popcount_kernel<<<(int)(pb < 1024 ? pb : 1024), 256, 0, st>>>(coverage_bits, words, coverage_count);
    B200_LAUNCH_OK("popcount_kernel");
  }
  return 0;
}

// Exact fp32 re-scoring of top-K candidates (the "fp32" storage of the drop-in flat index).
// The fused tensor-core pass scores fp32 catalogues with 3 split-bf16 piece products (h.h + h.m + m.h: ~2^-17 per
// product); an fp32 CPU search (faiss IndexFlatIP / np.dot, reference src/serving/retrieval.py:171,
// scripts/evaluate_model.py:222) is ~2^-23.  The index therefore asks the fused pass for k + margin candidates and this
// kernel recomputes their inner products from the fp32 rows kept in HBM: one warp per (query, candidate), 128-bit
// loads, fp32 FMAs, shuffle tree.  HBM-bound gather: Q * k_in * d * 4 bytes of random 4*d-byte rows.
#include <cfloat>
#include "host_util.h"
#include "../../include/b200rec.h"

namespace b200 {

__global__ void __launch_bounds__(256)
rescore_fp32_kernel(const float* __restrict__ q, int64_t ld_q, const float* __restrict__ rows, int64_t ld_rows,
                    int64_t n_rows, int64_t row_offset, int d, const int64_t* __restrict__ ids, int64_t total, int k_in,
                    float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const bool vec = ((d & 3) == 0) && ((ld_q & 3) == 0) && ((ld_rows & 3) == 0) &&
                   ((reinterpret_cast<uintptr_t>(q) & 15) == 0) && ((reinterpret_cast<uintptr_t>(rows) & 15) == 0);
  for (int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < total; w += warps) {
    const int64_t qi = w / k_in;
    const int64_t id = __ldg(ids + w);
    const int64_t r = id - row_offset;
    if (id < 0 || r < 0 || r >= n_rows) {
      if (lane == 0) out[w] = -FLT_MAX;
      continue;
    }
    const float* a = q + qi * ld_q;
    const float* b = rows + r * ld_rows;
    float acc = 0.f;
    if (vec) {
      for (int c = lane * 4; c < d; c += 128) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(a + c));
        const float4 y = __ldg(reinterpret_cast<const float4*>(b + c));
        acc = fmaf(x.x, y.x, acc);
        acc = fmaf(x.y, y.y, acc);
        acc = fmaf(x.z, y.z, acc);
        acc = fmaf(x.w, y.w, acc);
      }
    } else {
      for (int c = lane; c < d; c += 32) acc = fmaf(__ldg(a + c), __ldg(b + c), acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[w] = acc;
  }
}

}  // namespace b200

extern "C" int b200rec_rescore_fp32(const float* queries, int64_t ld_q, const float* rows, int64_t ld_rows,
                                    int64_t n_rows, int64_t row_offset, int d, const int64_t* ids, int64_t Q, int k_in,
                                    float* scores, void* stream) {
  using namespace b200;
  if (!queries || !rows || !ids || !scores) return fail("rescore_fp32: null pointer");
  if (Q <= 0 || k_in <= 0 || d <= 0 || n_rows <= 0) return fail("rescore_fp32: empty input");
  if (ld_q < d || ld_rows < d) return fail("rescore_fp32: leading dimension smaller than d");
  const int64_t total = Q * (int64_t)k_in;
  const int64_t blocks = (total + 7) / 8;
  const int grid = (int)(blocks < (int64_t)num_sms() * 16 ? blocks : (int64_t)num_sms() * 16);
  rescore_fp32_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(queries, ld_q, rows, ld_rows, n_rows,
                                                                               row_offset, d, ids, total, k_in, scores);
  B200_LAUNCH_OK("rescore_fp32_kernel");
  return 0;
}

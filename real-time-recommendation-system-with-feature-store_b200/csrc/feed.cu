// Device-side batch feed ("next" row f1): replaces the per-sample Python of
// MovieLensDataset.__getitem__ + collate_fn (reference src/training/datasets/movielens.py:86-162) and
// sample_negative_items (src/data/movielens.py:487-512) by two HBM-bound kernels per batch.
//   gather_rows      : out[b, :] = table[idx[b], :]                (user / positive-item / negative-item feature rows)
//   sample_negatives : R distinct items per row, uniform over the items the row's user has NOT interacted with
//                      (the reference draws np.random.choice(pool, R, replace=False) from all_items - positives;
//                      rejection sampling against the user's sorted positive list gives the same distribution)
#include "host_util.h"
#include "tc_common.cuh"
#include "../../include/b200rec.h"

namespace b200 {

// one warp per output row, 128-bit loads / stores when the row width and the pointers allow it
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ table, int64_t n_rows, int width, int64_t ld, const int64_t* __restrict__ idx,
                   int64_t B, float* __restrict__ out, int64_t ld_out, int* __restrict__ err) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  const bool vec = ((width & 3) == 0) && ((ld & 3) == 0) && ((ld_out & 3) == 0) &&
                   ((reinterpret_cast<uintptr_t>(table) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  for (int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); b < B; b += warps_total) {
    const int64_t r = __ldg(idx + b);
    const bool ok = r >= 0 && r < n_rows;
    if (!ok && lane == 0) atomicExch(err, 1);  // the reference's numpy fancy-indexing raises IndexError
    const float* src = table + (ok ? r : 0) * ld;
    float* dst = out + b * ld_out;
    if (vec) {
      for (int c = lane * 4; c < width; c += 128) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok) v = __ldg(reinterpret_cast<const float4*>(src + c));
        *reinterpret_cast<float4*>(dst + c) = v;
      }
    } else {
      for (int c = lane; c < width; c += 32) dst[c] = ok ? __ldg(src + c) : 0.f;
    }
  }
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__device__ __forceinline__ bool csr_contains(const int32_t* __restrict__ items, int n, int32_t v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(items + mid) < v) lo = mid + 1;
    else hi = mid;
  }
  return lo < n && __ldg(items + lo) == v;
}

// one thread per (row, negative slot) would need cross-thread duplicate checks; R is small (4-64), so one thread per
// row draws its R items in sequence: counter-based RNG keyed by (seed, global row, draw number), 64-bit multiply-high
// range reduction (bias < 2^-40), rejection against the user's sorted positives and the items already drawn.
__global__ void __launch_bounds__(256)
sample_negatives_kernel(const int64_t* __restrict__ user_of_row, int64_t B, const int64_t* __restrict__ pos_indptr,
                        const int32_t* __restrict__ pos_items, int64_t n_users, int64_t num_items, int R, uint64_t seed,
                        uint64_t row_base, int64_t* __restrict__ out, int* __restrict__ err) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int64_t u = __ldg(user_of_row + b);
  const int32_t* items = pos_items;
  int np = 0;
  if (u >= 0 && u < n_users) {  // users without interactions have an empty positive list (positive_items.get(u, []))
    const int64_t lo = __ldg(pos_indptr + u), hi = __ldg(pos_indptr + u + 1);
    items = pos_items + lo;
    np = (int)(hi - lo);
  }
  int64_t* o = out + b * R;
  if (num_items - np < R) {  // the reference returns the whole (too small) pool and collate_fn fails on the ragged batch
    atomicExch(err, 2);
    for (int j = 0; j < R; ++j) o[j] = -1;
    return;
  }
  const uint64_t key = splitmix64(seed ^ splitmix64(row_base + (uint64_t)b));
  uint64_t ctr = 0;
  for (int j = 0; j < R; ++j) {
    int64_t pick = -1;
    for (int tries = 0; tries < 4096 && pick < 0; ++tries) {
      const uint64_t r = splitmix64(key + (ctr++) * 0xD1B54A32D192ED03ull);
      const int64_t cand = (int64_t)__umul64hi(r, (uint64_t)num_items);
      bool bad = csr_contains(items, np, (int32_t)cand);
      for (int t = 0; t < j && !bad; ++t) bad = o[t] == cand;
      if (!bad) pick = cand;
    }
    if (pick < 0) {
      // pathologically dense user: take the first admissible item after a random start (still never a positive)
      const int64_t start = (int64_t)__umul64hi(splitmix64(key + (ctr++) * 0xD1B54A32D192ED03ull), (uint64_t)num_items);
      for (int64_t s = 0; s < num_items && pick < 0; ++s) {
        const int64_t cand = (start + s) % num_items;
        bool bad = csr_contains(items, np, (int32_t)cand);
        for (int t = 0; t < j && !bad; ++t) bad = o[t] == cand;
        if (!bad) pick = cand;
      }
    }
    o[j] = pick;
  }
}

}  // namespace b200

extern "C" int b200rec_gather_rows(const float* table, int64_t n_rows, int width, int64_t ld, const int64_t* idx,
                                   int64_t B, float* out, int64_t ld_out, int* err_flag, void* stream) {
  using namespace b200;
  if (!table || !idx || !out || !err_flag) return fail("gather_rows: null pointer");
  if (B <= 0 || width <= 0 || n_rows <= 0) return fail("gather_rows: empty input");
  if (ld < width || ld_out < width) return fail("gather_rows: leading dimension smaller than the row width");
  const int64_t blocks = (B + 7) / 8;
  const int grid = (int)(blocks < (int64_t)num_sms() * 16 ? blocks : (int64_t)num_sms() * 16);
  gather_rows_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(table, n_rows, width, ld, idx, B, out,
                                                                               ld_out, err_flag);
  B200_LAUNCH_OK("gather_rows_kernel");
  return 0;
}

extern "C" int b200rec_sample_negatives(const int64_t* user_of_row, int64_t B, const int64_t* pos_indptr,
                                        const int32_t* pos_items, int64_t n_users, int64_t num_items, int num_negatives,
                                        uint64_t seed, uint64_t row_base, int64_t* out_items, int* err_flag,
                                        void* stream) {
  using namespace b200;
  if (!user_of_row || !pos_indptr || !pos_items || !out_items || !err_flag) return fail("sample_negatives: null pointer");
  if (B <= 0 || num_negatives <= 0) return fail("sample_negatives: empty request");
  if (num_items <= 0 || num_items > INT32_MAX) return fail("sample_negatives: num_items must be in [1, 2^31)");
  if (num_negatives > 1024) return fail("sample_negatives: at most 1024 negatives per row");
  sample_negatives_kernel<<<(unsigned)((B + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      user_of_row, B, pos_indptr, pos_items, n_users, num_items, num_negatives, seed, row_base, out_items, err_flag);
  B200_LAUNCH_OK("sample_negatives_kernel");
  return 0;
}

// Embedding bags (K1): fused multi-field gather that writes the MLP's concatenated input directly, and the sparse
// row-gradient scatter-accumulate of nn.Embedding's backward (reference src/models/two_tower.py:113-126, :254-273;
// nn.Embedding(card+1, e, padding_idx=0) :46-50).  HBM-bound: one warp per sample row, 128-bit loads when the table
// row pitch allows it; the backward sorts (row, position) pairs (CUB radix sort = library plumbing), then one warp per
// distinct row sums its dY rows in registers and writes one coalesced gradient row.
#include <cub/cub.cuh>
#include "host_util.h"
#include "tc_common.cuh"
#include "../../include/b200rec.h"

namespace b200 {

constexpr int MAX_FIELDS = 16;
struct GatherArgs {
  const float* tables[MAX_FIELDS];
  const int64_t* idx[MAX_FIELDS];
  int64_t rows[MAX_FIELDS];
  int width[MAX_FIELDS];
  int tld[MAX_FIELDS];
  int coloff[MAX_FIELDS];
  int F;
};

__global__ void __launch_bounds__(256)
gather_concat_kernel(const float* __restrict__ numerical, int num_cols, int64_t ld_num, const GatherArgs ga, int64_t B,
                     float* __restrict__ out, int64_t ld_out, int* __restrict__ err) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); b < B; b += warps) {
    float* orow = out + b * ld_out;
    for (int c = lane; c < num_cols; c += 32) orow[c] = __ldg(numerical + b * ld_num + c);
    for (int f = 0; f < ga.F; ++f) {
      int64_t r = __ldg(ga.idx[f] + b);
      if (r < 0 || r >= ga.rows[f]) {  // nn.Embedding raises IndexError; flag it and read row 0
        if (lane == 0 && err) atomicExch(err, 1 + f);
        r = 0;
      }
      const float* trow = ga.tables[f] + r * ga.tld[f];
      float* dst = orow + ga.coloff[f];
      const int w = ga.width[f];
      const bool vec = ((ga.tld[f] & 3) == 0) && ((w & 3) == 0) && ((ga.coloff[f] & 3) == 0) && ((ld_out & 3) == 0) &&
                       ((reinterpret_cast<uintptr_t>(ga.tables[f]) & 15) == 0) &&
                       ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
      if (vec) {
        for (int c = lane; c < (w >> 2); c += 32)
          reinterpret_cast<float4*>(dst)[c] = __ldg(reinterpret_cast<const float4*>(trow) + c);
      } else {
        for (int c = lane; c < w; c += 32) dst[c] = __ldg(trow + c);
      }
    }
  }
}

// pos[i] = i and key[i] = idx[i], with every id outside [0, table_rows) replaced by the padding row: such a sample
// contributes no gradient (nn.Embedding raises IndexError in the forward — gather_concat_kernel flags it — so the
// backward must never hand an unchecked row to scatter_rows / sparse_adam, which write table[row]).
__global__ void keys_iota_kernel(const int64_t* __restrict__ idx, int64_t n, int64_t table_rows, int64_t padding_idx,
                                 int64_t* __restrict__ key, int32_t* __restrict__ pos) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int64_t r = idx[i];
    key[i] = (r < 0 || (table_rows > 0 && r >= table_rows)) ? padding_idx : r;
    pos[i] = (int32_t)i;
  }
}

// head[i] = 1 when sorted position i starts a new row that is not the padding row
__global__ void mark_heads_kernel(const int64_t* __restrict__ sidx, int64_t n, int64_t padding_idx,
                                  int32_t* __restrict__ head) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int64_t r = sidx[i];
    head[i] = (r != padding_idx && (i == 0 || sidx[i - 1] != r)) ? 1 : 0;
  }
}

// One warp per chunk of 32 consecutive SORTED positions (a hot row with thousands of duplicates — Zipf ids — is then
// spread over many warps instead of being walked serially by one): lanes own gradient columns, the 32 row ids /
// source positions are exchanged by shuffles so every dY load address is known up front, a run of equal rows is
// summed in registers and flushed once with atomicAdd (runs may continue in the neighbouring chunks; grad_rows is
// pre-zeroed).  slot = exclusive scan of the head flags, so position i belongs to output row slot[i] + head[i] - 1.
__global__ void __launch_bounds__(256)
segment_sum_kernel(const int64_t* __restrict__ sidx, const int32_t* __restrict__ spos, const int32_t* __restrict__ head,
                   const int32_t* __restrict__ slot, int64_t n, int64_t padding_idx, const float* __restrict__ dY,
                   int64_t ld_dy, int width, int64_t* __restrict__ unique_rows, float* __restrict__ grad_rows,
                   int32_t* __restrict__ n_unique) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t chunks = (n + 31) / 32;
  for (int64_t ch = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); ch < chunks; ch += warps) {
    const int64_t i = ch * 32 + lane;
    int32_t my_u = -1, my_pos = 0;
    if (i < n) {
      const int64_t r = sidx[i];
      const int32_t h = head[i], sl = slot[i];
      if (r != padding_idx) {
        my_u = sl + h - 1;
        my_pos = spos[i];
        if (h) unique_rows[my_u] = r;
      }
      if (i == n - 1) *n_unique = sl + h;
    }
    const int cnt = (int)((n - ch * 32) < 32 ? (n - ch * 32) : 32);
    for (int c0 = 0; c0 < width; c0 += 32) {
      const int c = c0 + lane;
      float acc = 0.f;
      int cur = -1;
      for (int j = 0; j < cnt; ++j) {
        const int32_t u = __shfl_sync(FULL_MASK, my_u, j);
        const int32_t p = __shfl_sync(FULL_MASK, my_pos, j);
        if (u != cur) {
          if (cur >= 0 && c < width) atomicAdd(grad_rows + (int64_t)cur * width + c, acc);
          acc = 0.f;
          cur = u;
        }
        if (u >= 0 && c < width) acc += __ldg(dY + (int64_t)p * ld_dy + c);
      }
      if (cur >= 0 && c < width) atomicAdd(grad_rows + (int64_t)cur * width + c, acc);
    }
  }
}

// dense[rows[u], :] = grad_rows[u, :]  (dense must be zeroed by the caller; rows are distinct)
__global__ void __launch_bounds__(256)
scatter_rows_kernel(const int64_t* __restrict__ rows, const float* __restrict__ grad_rows,
                    const int32_t* __restrict__ n_rows, int width, float* __restrict__ dense, int64_t ld, int accumulate) {
  const int lane = threadIdx.x & 31;
  const int n = __ldg(n_rows);
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t u = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); u < n; u += warps) {
    const int64_t r = __ldg(rows + u);
    for (int c = lane; c < width; c += 32) {
      const float g = __ldg(grad_rows + u * width + c);
      if (accumulate)
        dense[r * ld + c] += g;
      else
        dense[r * ld + c] = g;
    }
  }
}

struct SparseGradWs {
  int64_t* key;
  int64_t* sidx;
  int32_t* pos;
  int32_t* spos;
  int32_t* head;
  int32_t* slot;
  void* cub_tmp;
  size_t cub_bytes;
  size_t total;
};

static size_t align256(size_t v) { return (v + 255) / 256 * 256; }

static void plan_sparse_ws(SparseGradWs& w, int64_t B, void* base) {
  size_t sort_bytes = 0, scan_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const int64_t*)nullptr, (int64_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, (int)B);
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const int32_t*)nullptr, (int32_t*)nullptr, (int)B);
  w.cub_bytes = sort_bytes > scan_bytes ? sort_bytes : scan_bytes;
  uint8_t* p = reinterpret_cast<uint8_t*>(base);
  size_t off = 0;
  w.key = reinterpret_cast<int64_t*>(p + off); off += align256(sizeof(int64_t) * B);
  w.sidx = reinterpret_cast<int64_t*>(p + off); off += align256(sizeof(int64_t) * B);
  w.pos = reinterpret_cast<int32_t*>(p + off); off += align256(sizeof(int32_t) * B);
  w.spos = reinterpret_cast<int32_t*>(p + off); off += align256(sizeof(int32_t) * B);
  w.head = reinterpret_cast<int32_t*>(p + off); off += align256(sizeof(int32_t) * B);
  w.slot = reinterpret_cast<int32_t*>(p + off); off += align256(sizeof(int32_t) * B);
  w.cub_tmp = p + off; off += align256(w.cub_bytes);
  w.total = off;
}

}  // namespace b200

using namespace b200;

extern "C" int b200rec_gather_concat(const float* numerical, int64_t num_cols, int64_t ld_num,
                                     const float* const* tables_host, const int64_t* const* indices_host,
                                     const int64_t* table_rows_host, const int32_t* widths_host,
                                     const int32_t* table_ld_host, const int32_t* col_off_host, int F, int64_t B,
                                     float* out, int64_t ld_out, int32_t* err_flag, void* stream) {
  if (!out) return fail("gather_concat: null output");
  if (B <= 0) return fail("gather_concat: empty batch");
  if (F < 0 || F > MAX_FIELDS) return fail("gather_concat: at most %d fields per call (got %d)", MAX_FIELDS, F);
  if (num_cols > 0 && !numerical) return fail("gather_concat: numerical block missing");
  GatherArgs ga;
  ga.F = F;
  for (int f = 0; f < F; ++f) {
    if (!tables_host[f] || !indices_host[f]) return fail("gather_concat: null table or index pointer (field %d)", f);
    ga.tables[f] = tables_host[f];
    ga.idx[f] = indices_host[f];
    ga.rows[f] = table_rows_host[f];
    ga.width[f] = widths_host[f];
    ga.tld[f] = table_ld_host[f];
    ga.coloff[f] = col_off_host[f];
  }
  const int64_t blocks = (B + 7) / 8;
  const int grid = (int)(blocks < (int64_t)num_sms() * 16 ? blocks : (int64_t)num_sms() * 16);
  gather_concat_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(numerical, (int)num_cols, ld_num, ga, B,
                                                                                 out, ld_out, err_flag);
  B200_LAUNCH_OK("gather_concat_kernel");
  return 0;
}

extern "C" size_t b200rec_sparse_grad_workspace_bytes(int64_t B) {
  if (B <= 0 || B > INT32_MAX) return 0;
  SparseGradWs w;
  plan_sparse_ws(w, B, nullptr);
  return w.total;
}

extern "C" int b200rec_embedding_sparse_grad(const int64_t* idx, int64_t B, const float* dY, int64_t ld_dy, int width,
                                             int64_t padding_idx, int64_t table_rows, int64_t* unique_rows,
                                             float* grad_rows, int32_t* n_unique_out, void* workspace,
                                             size_t workspace_bytes, void* stream) {
  if (!idx || !dY || !unique_rows || !grad_rows || !n_unique_out || !workspace) return fail("sparse_grad: null pointer");
  if (B <= 0 || B > INT32_MAX || width <= 0) return fail("sparse_grad: bad sizes");
  SparseGradWs w;
  plan_sparse_ws(w, B, workspace);
  if (workspace_bytes < w.total) return fail("sparse_grad: workspace too small (%zu < %zu)", workspace_bytes, w.total);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const unsigned nb = (unsigned)((B + 255) / 256);
  // entries beyond *n_unique_out read as (row 0 = padding, zero gradient): lists can be concatenated and re-coalesced
  B200_CUDA_OK(cudaMemsetAsync(unique_rows, 0, sizeof(int64_t) * B, st));
  B200_CUDA_OK(cudaMemsetAsync(grad_rows, 0, sizeof(float) * B * width, st));
  if (padding_idx < 0 || (table_rows > 0 && padding_idx >= table_rows))
    return fail("sparse_grad: padding_idx %lld outside the table", (long long)padding_idx);
  keys_iota_kernel<<<nb, 256, 0, st>>>(idx, B, table_rows, padding_idx, w.key, w.pos);
  B200_LAUNCH_OK("keys_iota_kernel");
  int end_bit = 64;
  if (table_rows > 0) {
    end_bit = 1;
    while (end_bit < 63 && (1ll << end_bit) < table_rows) ++end_bit;
  }
  size_t tmp = w.cub_bytes;
  B200_CUDA_OK(cub::DeviceRadixSort::SortPairs(w.cub_tmp, tmp, w.key, w.sidx, w.pos, w.spos, (int)B, 0, end_bit, st));
  b200::g_launches.fetch_add(1);
  mark_heads_kernel<<<nb, 256, 0, st>>>(w.sidx, B, padding_idx, w.head);
  B200_LAUNCH_OK("mark_heads_kernel");
  tmp = w.cub_bytes;
  B200_CUDA_OK(cub::DeviceScan::ExclusiveSum(w.cub_tmp, tmp, w.head, w.slot, (int)B, st));
  b200::g_launches.fetch_add(1);
  const int64_t blocks = ((B + 31) / 32 + 7) / 8;
  const int grid = (int)(blocks < (int64_t)num_sms() * 16 ? blocks : (int64_t)num_sms() * 16);
  segment_sum_kernel<<<grid, 256, 0, st>>>(w.sidx, w.spos, w.head, w.slot, B, padding_idx, dY, ld_dy, width,
                                           unique_rows, grad_rows, n_unique_out);
  B200_LAUNCH_OK("segment_sum_kernel");
  return 0;
}

namespace b200 {
// dense[idx[b], :width] += dY[b, :width] for every sample whose id is a real row (padding row and out-of-range ids
// contribute nothing): the dense gradient nn.Embedding(sparse=False) produces, accumulated IN PLACE with fp32 atomics —
// one warp per sample, lanes along the row, so each atomic instruction covers one contiguous 128-byte segment.  Replaces
// sort + segment sum + scatter (14 launches) when the gradient lands in a pre-zeroed dense buffer anyway; the sorted
// path stays for row-sparse tables, whose optimizer needs each distinct row exactly once.
__global__ void __launch_bounds__(256)
scatter_add_rows_kernel(const int64_t* __restrict__ idx, int64_t B, const float* __restrict__ dY, int64_t ld_dy, int width,
                        int64_t padding_idx, int64_t table_rows, float* __restrict__ dense, int64_t ld,
                        int32_t* __restrict__ row_flags) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); b < B; b += warps) {
    const int64_t r = __ldg(idx + b);
    if (r == padding_idx || r < 0 || r >= table_rows) continue;
    const float* src = dY + b * ld_dy;
    float* dst = dense + r * ld;
    for (int c = lane; c < width; c += 32) atomicAdd(dst + c, __ldg(src + c));
    if (row_flags != nullptr && lane == 0) row_flags[r] = 1;   // this row's gradient is (possibly) non-zero this step
  }
}

// Row-sparse tables (config 4: 50 M rows): the optimiser needs every DISTINCT touched row once, with the gradients of its
// duplicates summed.  Sort + segment sum did that in ~10 launches and ~140 us per table (cub radix sort of 8192 keys,
// scan, a 32-block segment sum); here the first sample that claims a row (atomicCAS on a per-row slot) becomes its
// leader, every sample adds its gradient row to the LEADER's row of a compact [B, width] buffer, and rows_out[b] is the
// row id for leaders and -1 for everybody else (the sparse Adam skips -1).  One launch + a release launch that frees
// the slots again; duplicates are summed with fp32 atomics (order-dependent in the last bit, as in the dense scatter).
__global__ void __launch_bounds__(256)
sparse_claim_accumulate_kernel(const int64_t* __restrict__ idx, int64_t B, const float* __restrict__ dY, int64_t ld_dy,
                               int width, int64_t padding_idx, int64_t table_rows, int32_t* __restrict__ slot,
                               int64_t* __restrict__ rows_out, float* __restrict__ acc) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); b < B; b += warps) {
    const int64_t r = __ldg(idx + b);
    if (r == padding_idx || r < 0 || r >= table_rows) {
      if (lane == 0) rows_out[b] = -1;
      continue;
    }
    int old = 0;
    if (lane == 0) old = atomicCAS(slot + r, 0, (int)b + 1);   // 0 = free, b + 1 = claimed by sample b
    old = __shfl_sync(0xffffffffu, old, 0);
    const int64_t leader = old == 0 ? b : (int64_t)old - 1;
    if (lane == 0) rows_out[b] = old == 0 ? r : -1;
    const float* src = dY + b * ld_dy;
    float* dst = acc + leader * width;
    for (int c = lane; c < width; c += 32) atomicAdd(dst + c, __ldg(src + c));
  }
}

__global__ void sparse_release_kernel(const int64_t* __restrict__ rows, int64_t n, int32_t* __restrict__ slot) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int64_t r = rows[i];
    if (r >= 0) slot[r] = 0;
  }
}
}  // namespace b200

extern "C" int b200rec_scatter_add_rows(const int64_t* idx, int64_t B, const float* dY, int64_t ld_dy, int width,
                                        int64_t padding_idx, int64_t table_rows, float* dense, int64_t ld, void* stream) {
  using namespace b200;
  if (!idx || !dY || !dense) return fail("scatter_add_rows: null pointer");
  if (B <= 0 || width <= 0 || table_rows <= 0) return fail("scatter_add_rows: empty input");
  const int64_t blocks = (B + 7) / 8;
  const int grid = (int)(blocks < (int64_t)num_sms() * 16 ? blocks : (int64_t)num_sms() * 16);
  scatter_add_rows_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(idx, B, dY, ld_dy, width, padding_idx,
                                                                                    table_rows, dense, ld, nullptr);
  B200_LAUNCH_OK("scatter_add_rows_kernel");
  return 0;
}

extern "C" int b200rec_scatter_add_rows_flagged(const int64_t* idx, int64_t B, const float* dY, int64_t ld_dy, int width,
                                                int64_t padding_idx, int64_t table_rows, float* dense, int64_t ld,
                                                int32_t* row_flags, void* stream) {
  using namespace b200;
  if (!idx || !dY || !dense || !row_flags) return fail("scatter_add_rows_flagged: null pointer");
  if (B <= 0 || width <= 0 || table_rows <= 0) return fail("scatter_add_rows_flagged: empty input");
  const int64_t blocks = (B + 7) / 8;
  const int grid = (int)(blocks < (int64_t)num_sms() * 16 ? blocks : (int64_t)num_sms() * 16);
  scatter_add_rows_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(idx, B, dY, ld_dy, width, padding_idx,
                                                                                    table_rows, dense, ld, row_flags);
  B200_LAUNCH_OK("scatter_add_rows_kernel");
  return 0;
}


extern "C" int b200rec_sparse_claim_accumulate(const int64_t* idx, int64_t B, const float* dY, int64_t ld_dy, int width,
                                               int64_t padding_idx, int64_t table_rows, int32_t* slot, int64_t* rows_out,
                                               float* acc, void* stream) {
  using namespace b200;
  if (!idx || !dY || !slot || !rows_out || !acc) return fail("sparse_claim_accumulate: null pointer");
  if (B <= 0 || B >= INT32_MAX || width <= 0 || table_rows <= 0) return fail("sparse_claim_accumulate: bad sizes");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int64_t blocks = (B + 7) / 8;
  const int grid = (int)(blocks < (int64_t)num_sms() * 16 ? blocks : (int64_t)num_sms() * 16);
  sparse_claim_accumulate_kernel<<<grid, 256, 0, st>>>(idx, B, dY, ld_dy, width, padding_idx, table_rows, slot, rows_out, acc);
  B200_LAUNCH_OK("sparse_claim_accumulate_kernel");
  sparse_release_kernel<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(rows_out, B, slot);
  B200_LAUNCH_OK("sparse_release_kernel");
  return 0;
}

extern "C" int b200rec_scatter_rows(const int64_t* rows, const float* grad_rows, const int32_t* n_rows,
                                    int64_t max_rows, int width, float* dense, int64_t ld, int accumulate,
                                    void* stream) {
  if (!rows || !grad_rows || !n_rows || !dense) return fail("scatter_rows: null pointer");
  if (max_rows <= 0 || width <= 0) return fail("scatter_rows: empty input");
  const int64_t blocks = (max_rows + 7) / 8;
  const int grid = (int)(blocks < (int64_t)num_sms() * 16 ? blocks : (int64_t)num_sms() * 16);
  scatter_rows_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(rows, grad_rows, n_rows, width, dense,
                                                                                ld, accumulate);
  B200_LAUNCH_OK("scatter_rows_kernel");
  return 0;
}

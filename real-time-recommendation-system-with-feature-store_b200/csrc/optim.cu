// Optimiser step pieces (K5): global gradient-norm clipping and Adam with L2-coupled weight decay on flat fp32 buffers
// (reference src/training/trainers/two_tower.py:60-64 Adam(lr, weight_decay), :144 clip_grad_norm_(1.0), :146 step),
// plus the row-sparse Adam used for large embedding tables.  Pure HBM-bound streaming kernels, 128-bit accesses.
#include "host_util.h"
#include "tc_common.cuh"
#include "../../include/b200rec.h"

namespace b200 {

__global__ void __launch_bounds__(256)
sumsq_kernel(const float* __restrict__ x, int64_t n, double* __restrict__ out) {
  __shared__ double red[8];
  double s = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) ? n / 4 : 0;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldg(x4 + i);
    s += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
  }
  for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float v = __ldg(x + i);
    s += (double)v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL_MASK, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(out, t);
  }
}

// clip_grad_norm_: coef = min(1, max_norm / (total_norm + 1e-6)); also exports the norm
__global__ void clip_coef_kernel(const double* __restrict__ sumsq, float max_norm, float* __restrict__ coef,
                                 float* __restrict__ norm_out) {
  const float nrm = (float)sqrt(*sumsq);
  float c = max_norm / (nrm + 1e-6f);
  *coef = c < 1.f ? c : 1.f;
  if (norm_out) *norm_out = nrm;
}

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float cc, float wd, float beta1, float beta2,
                                         float eps, float step_size, float bias_c2_sqrt) {
  float gi = g * cc;
  gi = fmaf(wd, p, gi);
  m = beta1 * m + (1.f - beta1) * gi;
  v = beta2 * v + (1.f - beta2) * gi * gi;
  const float denom = sqrtf(v) / bias_c2_sqrt + eps;
  p = p - step_size * (m / denom);
}

// 28 B per parameter (p, m, v read + written, g read).  One element per thread and iteration: the compiler unrolls the
// grid-stride loop x4, i.e. 16 independent 4-byte loads in flight per thread — measured 5.2 TB/s (80 % of the copy peak);
// a float4 version of the same loop ran at 4.0 TB/s.  clear_grad: gradients that were non-zero are reset to zero on the
// way (the table gradients of a step are non-zero on a few thousand rows only), which replaces the separate fill of the
// whole gradient buffer before the next step.
__global__ void __launch_bounds__(256)
adam_dense_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                  int64_t n, float lr, float beta1, float beta2, float eps, float wd, float bias_c1, float bias_c2_sqrt,
                  const float* __restrict__ clip_coef, const float* __restrict__ hyper_dev, int clear_grad) {
  const float cc = clip_coef ? __ldg(clip_coef) : 1.f;
  if (hyper_dev) {  // CUDA-graph training: {lr, 1 - beta1^t, sqrt(1 - beta2^t)} written by b200rec_train_step_begin
    lr = __ldg(hyper_dev);
    bias_c1 = __ldg(hyper_dev + 1);
    bias_c2_sqrt = __ldg(hyper_dev + 2);
  }
  const float step_size = lr / bias_c1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float P = p[i], M = m[i], V = v[i];
    const float G = g[i];
    adam_one(P, G, M, V, cc, wd, beta1, beta2, eps, step_size, bias_c2_sqrt);
    p[i] = P;
    m[i] = M;
    v[i] = V;
    if (clear_grad && G != 0.f) g[i] = 0.f;
  }
}

// one warp per touched row
__global__ void __launch_bounds__(256)
sparse_adam_kernel(float* __restrict__ table, float* __restrict__ m, float* __restrict__ v, int64_t ld, int width,
                   const int64_t* __restrict__ rows, const float* __restrict__ grad_rows,
                   const int32_t* __restrict__ n_rows, float lr, float beta1, float beta2, float eps, float bias_c1,
                   float bias_c2_sqrt, const float* __restrict__ clip_coef, const float* __restrict__ hyper_dev) {
  const int lane = threadIdx.x & 31;
  const int n = __ldg(n_rows);
  const float cc = clip_coef ? __ldg(clip_coef) : 1.f;
  if (hyper_dev) {
    lr = __ldg(hyper_dev);
    bias_c1 = __ldg(hyper_dev + 1);
    bias_c2_sqrt = __ldg(hyper_dev + 2);
  }
  const float step_size = lr / bias_c1;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t u = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); u < n; u += warps) {
    const int64_t r = __ldg(rows + u);
    if (r < 0) continue;   // b200rec_sparse_claim_accumulate: a duplicate whose gradient went to its row's leader
    for (int c = lane; c < width; c += 32) {
      const int64_t o = r * ld + c;
      const float gi = __ldg(grad_rows + u * width + c) * cc;
      const float mi = beta1 * m[o] + (1.f - beta1) * gi;
      const float vi = beta2 * v[o] + (1.f - beta2) * gi * gi;
      m[o] = mi;
      v[o] = vi;
      table[o] -= step_size * (mi / (sqrtf(vi) / bias_c2_sqrt + eps));
    }
  }
}


// ---- dense embedding tables under the reference's dense Adam ------------------------------------------------------
// Adam with L2-coupled weight decay moves EVERY row each step (gi = wd * p), but the gradient of a step is non-zero
// only on the rows the batch touched (a few thousand of 10^6): the scatter kernel raises a per-row flag, and these
// kernels read (and reset) the gradient of flagged rows only — 24 B per parameter instead of 32 (28 + the reset), and
// the gradient-norm pass touches the flags and the flagged rows instead of the whole 282 MB buffer.
// One warp per row (lane = column mod 32), ROWS_PER_ITER rows in flight per warp.
constexpr int TBL_ROWS_PER_ITER = 4;
constexpr int TBL_MAX_EPL = 4;   // columns per lane: tables up to 128 wide

__global__ void __launch_bounds__(256)
table_sumsq_kernel(const float* __restrict__ g, int64_t ld, int64_t rows, int width, const int32_t* __restrict__ flags,
                   double* __restrict__ out) {
  __shared__ double red[8];
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  double s = 0.0;
  // lanes scan 32 flags at a time; flagged rows are then summed by the whole warp
  for (int64_t r0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32; r0 < rows; r0 += warps * 32) {
    const int64_t r = r0 + lane;
    unsigned hit = __ballot_sync(FULL_MASK, r < rows && __ldg(flags + r) != 0);
    while (hit) {
      const int j = __ffs(hit) - 1;
      hit &= hit - 1;
      const float* row = g + (r0 + j) * ld;
      for (int c = lane; c < width; c += 32) {
        const float v = __ldg(row + c);
        s += (double)v * v;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL_MASK, s, o);
  if (lane == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    if (t != 0.0) atomicAdd(out, t);
  }
}

template <int EPL>
__global__ void __launch_bounds__(256)
adam_table_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t ld,
                  int64_t rows, int width, int32_t* __restrict__ flags, float lr, float beta1, float beta2, float eps,
                  float wd, float bias_c1, float bias_c2_sqrt, const float* __restrict__ clip_coef,
                  const float* __restrict__ hyper_dev, int clear_grad) {
  const float cc = clip_coef ? __ldg(clip_coef) : 1.f;
  if (hyper_dev) {
    lr = __ldg(hyper_dev);
    bias_c1 = __ldg(hyper_dev + 1);
    bias_c2_sqrt = __ldg(hyper_dev + 2);
  }
  const float step_size = lr / bias_c1;
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  constexpr int R = TBL_ROWS_PER_ITER;
  for (int64_t r0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * R; r0 < rows; r0 += warps * R) {
    float P[R][EPL], M[R][EPL], V[R][EPL], G[R][EPL];
    int f[R];
#pragma unroll
    for (int j = 0; j < R; ++j) f[j] = (r0 + j < rows) ? flags[r0 + j] : 0;   // warp-uniform
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int64_t o = (r0 + j) * ld;
#pragma unroll
      for (int k = 0; k < EPL; ++k) {
        const int c = lane + 32 * k;
        const bool ok = r0 + j < rows && c < width;
        P[j][k] = ok ? p[o + c] : 0.f;
        M[j][k] = ok ? m[o + c] : 0.f;
        V[j][k] = ok ? v[o + c] : 0.f;
        G[j][k] = (ok && f[j] != 0) ? g[o + c] : 0.f;
      }
    }
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int64_t o = (r0 + j) * ld;
#pragma unroll
      for (int k = 0; k < EPL; ++k) {
        const int c = lane + 32 * k;
        if (r0 + j < rows && c < width) {
          adam_one(P[j][k], G[j][k], M[j][k], V[j][k], cc, wd, beta1, beta2, eps, step_size, bias_c2_sqrt);
          p[o + c] = P[j][k];
          m[o + c] = M[j][k];
          v[o + c] = V[j][k];
          if (clear_grad && f[j] != 0) g[o + c] = 0.f;
        }
      }
      if (clear_grad && f[j] != 0 && lane == 0) flags[r0 + j] = 0;   // this warp is the only reader of the row's flag
    }
  }
}

static int flat_grid(int64_t total) {
  int64_t g = (total + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 16;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

}  // namespace b200

using namespace b200;

extern "C" int b200rec_sumsq(const float* x, int64_t n, double* out, void* stream) {
  if (!x || !out) return fail("sumsq: null pointer");
  if (n <= 0) return fail("sumsq: empty input");
  sumsq_kernel<<<flat_grid((n + 3) / 4), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, n, out);
  B200_LAUNCH_OK("sumsq_kernel");
  return 0;
}

extern "C" int b200rec_clip_coef(const double* sumsq, float max_norm, float* coef, float* norm_out, void* stream) {
  if (!sumsq || !coef) return fail("clip_coef: null pointer");
  clip_coef_kernel<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream)>>>(sumsq, max_norm, coef, norm_out);
  B200_LAUNCH_OK("clip_coef_kernel");
  return 0;
}

extern "C" int b200rec_adam_dense(float* p, float* g, float* m, float* v, int64_t n, float lr, float beta1,
                                  float beta2, float eps, float weight_decay, float bias_c1, float bias_c2_sqrt,
                                  const float* clip_coef_dev, int clear_grad, void* stream) {
  if (!p || !g || !m || !v) return fail("adam_dense: null pointer");
  if (n <= 0) return fail("adam_dense: empty input");
  adam_dense_kernel<<<flat_grid(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bias_c1, bias_c2_sqrt, clip_coef_dev, nullptr, clear_grad);
  B200_LAUNCH_OK("adam_dense_kernel");
  return 0;
}

// the same update with {lr, 1 - beta1^t, sqrt(1 - beta2^t)} read from device memory (hyper_dev, see
// b200rec_train_step_begin): the form a CUDA-graph-captured training step uses
extern "C" int b200rec_adam_dense_dev(float* p, float* g, float* m, float* v, int64_t n, float beta1, float beta2,
                                      float eps, float weight_decay, const float* hyper_dev, const float* clip_coef_dev,
                                      int clear_grad, void* stream) {
  if (!p || !g || !m || !v || !hyper_dev) return fail("adam_dense_dev: null pointer");
  if (n <= 0) return fail("adam_dense_dev: empty input");
  adam_dense_kernel<<<flat_grid(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      p, g, m, v, n, 0.f, beta1, beta2, eps, weight_decay, 1.f, 1.f, clip_coef_dev, hyper_dev, clear_grad);
  B200_LAUNCH_OK("adam_dense_kernel");
  return 0;
}


extern "C" int b200rec_table_sumsq(const float* g, int64_t ld, int64_t rows, int width, const int32_t* row_flags,
                                   double* out, void* stream) {
  if (!g || !row_flags || !out) return fail("table_sumsq: null pointer");
  if (rows <= 0 || width <= 0 || ld < width) return fail("table_sumsq: bad shape");
  // one block per SM: the pass is a 4 B-per-row flag scan plus a few thousand rows, and every block ends in one fp64
  // atomicAdd on the same address (1184 blocks spent 25 us mostly there)
  const int64_t blocks = (rows + 255) / 256;
  const int grid = (int)(blocks < (int64_t)num_sms() ? blocks : (int64_t)num_sms());
  table_sumsq_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(g, ld, rows, width, row_flags, out);
  B200_LAUNCH_OK("table_sumsq_kernel");
  return 0;
}

extern "C" int b200rec_adam_table(float* p, float* g, float* m, float* v, int64_t ld, int64_t rows, int width,
                                  int32_t* row_flags, float lr, float beta1, float beta2, float eps, float weight_decay,
                                  float bias_c1, float bias_c2_sqrt, const float* hyper_dev, const float* clip_coef_dev,
                                  int clear_grad, void* stream) {
  if (!p || !g || !m || !v || !row_flags) return fail("adam_table: null pointer");
  if (rows <= 0 || width <= 0 || ld < width) return fail("adam_table: bad shape");
  if (width > 32 * TBL_MAX_EPL) return fail("adam_table: rows wider than %d columns use b200rec_adam_dense", 32 * TBL_MAX_EPL);
  const int64_t blocks = (rows + 8 * TBL_ROWS_PER_ITER - 1) / (8 * TBL_ROWS_PER_ITER);
  const int grid = (int)(blocks < (int64_t)num_sms() * 16 ? blocks : (int64_t)num_sms() * 16);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int epl = (width + 31) / 32;
#define ADAM_TABLE(E) adam_table_kernel<E><<<grid, 256, 0, st>>>(p, g, m, v, ld, rows, width, row_flags, lr, beta1, beta2, eps, \
                                                                 weight_decay, bias_c1, bias_c2_sqrt, clip_coef_dev, hyper_dev, clear_grad)
  if (epl == 1) ADAM_TABLE(1);
  else if (epl == 2) ADAM_TABLE(2);
  else if (epl == 3) ADAM_TABLE(3);
  else ADAM_TABLE(4);
#undef ADAM_TABLE
  B200_LAUNCH_OK("adam_table_kernel");
  return 0;
}

extern "C" int b200rec_sparse_adam(float* table, float* exp_avg, float* exp_avg_sq, int64_t ld, int width,
                                   const int64_t* rows, const float* grad_rows, const int32_t* n_rows, int64_t max_rows,
                                   float lr, float beta1, float beta2, float eps, float bias_c1, float bias_c2_sqrt,
                                   const float* clip_coef_dev, void* stream) {
  if (!table || !exp_avg || !exp_avg_sq || !rows || !grad_rows || !n_rows) return fail("sparse_adam: null pointer");
  if (max_rows <= 0 || width <= 0) return fail("sparse_adam: empty input");
  const int64_t blocks = (max_rows + 7) / 8;
  const int grid = (int)(blocks < (int64_t)num_sms() * 16 ? blocks : (int64_t)num_sms() * 16);
  sparse_adam_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      table, exp_avg, exp_avg_sq, ld, width, rows, grad_rows, n_rows, lr, beta1, beta2, eps, bias_c1, bias_c2_sqrt,
      clip_coef_dev, nullptr);
  B200_LAUNCH_OK("sparse_adam_kernel");
  return 0;
}

extern "C" int b200rec_sparse_adam_dev(float* table, float* exp_avg, float* exp_avg_sq, int64_t ld, int width,
                                       const int64_t* rows, const float* grad_rows, const int32_t* n_rows,
                                       int64_t max_rows, float beta1, float beta2, float eps, const float* hyper_dev,
                                       const float* clip_coef_dev, void* stream) {
  if (!table || !exp_avg || !exp_avg_sq || !rows || !grad_rows || !n_rows || !hyper_dev)
    return fail("sparse_adam_dev: null pointer");
  if (max_rows <= 0 || width <= 0) return fail("sparse_adam_dev: empty input");
  const int64_t blocks = (max_rows + 7) / 8;
  const int grid = (int)(blocks < (int64_t)num_sms() * 16 ? blocks : (int64_t)num_sms() * 16);
  sparse_adam_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      table, exp_avg, exp_avg_sq, ld, width, rows, grad_rows, n_rows, 0.f, beta1, beta2, eps, 1.f, 1.f, clip_coef_dev,
      hyper_dev);
  B200_LAUNCH_OK("sparse_adam_kernel");
  return 0;
}

// Tower-MLP pieces around the tensor-core GEMMs (K2): activation + BatchNorm1d (+ dropout) forward / backward,
// F.normalize backward, bias-gradient column sums.  All HBM-bound: coalesced along the feature dimension, one
// read of each activation per pass, statistics reduced in fp64 registers and merged with fp64 atomics.
// Reference order is Linear -> activation -> BatchNorm1d -> Dropout (src/models/two_tower.py:56-72,196-212).
#include "host_util.h"
#include "tc_common.cuh"
#include "tower_math.cuh"
#include "../../include/b200rec.h"

namespace b200 {

// ------------------------------------------------------------------ column statistics
// sums[0:H] += sum_b f1(b,h), sums[H:2H] += sum_b f2(b,h)  (fp64)
template <class F>
__device__ __forceinline__ void column_sums(int64_t B, int64_t H, double* sums, F f) {
  const int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H) return;
  double s1 = 0.0, s2 = 0.0;
  for (int64_t b = blockIdx.y; b < B; b += gridDim.y) {
    float v1, v2;
    f(b, h, v1, v2);
    s1 += (double)v1;
    s2 += (double)v2;
  }
  atomicAdd(sums + h, s1);
  atomicAdd(sums + H + h, s2);
}

__global__ void __launch_bounds__(128)
bn_stats_kernel(const float* __restrict__ z, int64_t B, int64_t H, int64_t ld, int act, double* __restrict__ sums) {
  column_sums(B, H, sums, [&](int64_t b, int64_t h, float& v1, float& v2) {
    const float a = act_fwd(act, __ldg(z + b * ld + h));
    v1 = a;
    v2 = a * a;
  });
}

// mean / invstd from the fp64 sums; running statistics with momentum (nn.BatchNorm1d defaults: eps 1e-5, momentum 0.1,
// unbiased variance in the running estimate)
__global__ void bn_finalize_kernel(const double* __restrict__ sums, int64_t B, int64_t H, float eps, float momentum,
                                   float* __restrict__ mean, float* __restrict__ invstd,
                                   float* __restrict__ running_mean, float* __restrict__ running_var) {
  const int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H) return;
  const double m = sums[h] / (double)B;
  double var = sums[H + h] / (double)B - m * m;
  if (var < 0.0) var = 0.0;
  mean[h] = (float)m;
  invstd[h] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) {
    const double unbiased = B > 1 ? var * (double)B / (double)(B - 1) : var;
    running_mean[h] = (float)((1.0 - momentum) * (double)running_mean[h] + (double)momentum * m);
    running_var[h] = (float)((1.0 - momentum) * (double)running_var[h] + (double)momentum * unbiased);
  }
}

// eval mode: mean / invstd from the running statistics
__global__ void bn_eval_stats_kernel(const float* __restrict__ running_mean, const float* __restrict__ running_var,
                                     int64_t H, float eps, float* __restrict__ mean, float* __restrict__ invstd) {
  const int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H) return;
  mean[h] = running_mean[h];
  invstd[h] = 1.0f / sqrtf(running_var[h] + eps);
}

__global__ void __launch_bounds__(256)
bn_apply_kernel(const float* __restrict__ z, int64_t B, int64_t H, int64_t ld, int act, const float* __restrict__ mean,
                const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                float drop_p, uint64_t seed, float* __restrict__ y, int64_t ld_y) {
  const int64_t total = B * H;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / H, h = i - b * H;
    const float a = act_fwd(act, __ldg(z + b * ld + h));
    float v = (a - __ldg(mean + h)) * __ldg(invstd + h);
    v = v * __ldg(gamma + h) + __ldg(beta + h);
    v *= drop_scale(drop_p, seed, b, h, H);
    y[b * ld_y + h] = v;
  }
}

__global__ void __launch_bounds__(128)
bn_bwd_stats_kernel(const float* __restrict__ dy, int64_t ld_dy, const float* __restrict__ z, int64_t ld_z, int64_t B,
                    int64_t H, int act, const float* __restrict__ mean, const float* __restrict__ invstd, float drop_p,
                    uint64_t seed, double* __restrict__ sums) {
  column_sums(B, H, sums, [&](int64_t b, int64_t h, float& v1, float& v2) {
    const float g = __ldg(dy + b * ld_dy + h) * drop_scale(drop_p, seed, b, h, H);
    const float xhat = (act_fwd(act, __ldg(z + b * ld_z + h)) - __ldg(mean + h)) * __ldg(invstd + h);
    v1 = g;
    v2 = g * xhat;
  });
}

// dz = act'(z) * gamma*invstd * (g - sum1/n - xhat*sum2/n)   (training, n > 0)   |   act'(z) * gamma*invstd * g  (eval)
// also dgamma[h] = sum2, dbeta[h] = sum1 (written by block 0 threads) and dbias[h] += sum_b dz via fp64 column sums.
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float* __restrict__ dy, int64_t ld_dy, const float* __restrict__ z, int64_t ld_z, int64_t B,
                    int64_t H, int act, const float* __restrict__ mean, const float* __restrict__ invstd,
                    const float* __restrict__ gamma, const double* __restrict__ sums, int training, float drop_p,
                    uint64_t seed, float* __restrict__ dz, int64_t ld_dz, int64_t B_stat) {
  const int64_t total = B * H;
  const double inv_n = 1.0 / (double)B_stat;  // rows the statistics were taken over (all replicas under data parallel)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / H, h = i - b * H;
    const float zz = __ldg(z + b * ld_z + h);
    const float g = __ldg(dy + b * ld_dy + h) * drop_scale(drop_p, seed, b, h, H);
    const float is = __ldg(invstd + h);
    float da;
    if (training) {
      const float xhat = (act_fwd(act, zz) - __ldg(mean + h)) * is;
      da = __ldg(gamma + h) * is * (g - (float)(sums[h] * inv_n) - xhat * (float)(sums[H + h] * inv_n));
    } else {
      da = __ldg(gamma + h) * is * g;
    }
    dz[b * ld_dz + h] = da * act_grad(act, zz);
  }
}

__global__ void __launch_bounds__(256)
act_dropout_kernel(const float* __restrict__ z, int64_t B, int64_t H, int64_t ld, int act, float drop_p, uint64_t seed,
                   float* __restrict__ y, int64_t ld_y) {
  const int64_t total = B * H;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / H, h = i - b * H;
    y[b * ld_y + h] = act_fwd(act, __ldg(z + b * ld + h)) * drop_scale(drop_p, seed, b, h, H);
  }
}
__global__ void __launch_bounds__(256)
act_dropout_bwd_kernel(const float* __restrict__ dy, int64_t ld_dy, const float* __restrict__ z, int64_t ld_z, int64_t B,
                       int64_t H, int act, float drop_p, uint64_t seed, float* __restrict__ dz, int64_t ld_dz) {
  const int64_t total = B * H;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / H, h = i - b * H;
    dz[b * ld_dz + h] = __ldg(dy + b * ld_dy + h) * drop_scale(drop_p, seed, b, h, H) * act_grad(act, __ldg(z + b * ld_z + h));
  }
}

__global__ void sums_to_float_kernel(const double* __restrict__ sums, int64_t n, float* __restrict__ out, int accumulate) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = accumulate ? out[i] + (float)sums[i] : (float)sums[i];
}

__global__ void __launch_bounds__(128)
colsum_kernel(const float* __restrict__ x, int64_t B, int64_t H, int64_t ld, double* __restrict__ sums) {
  const int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H) return;
  double s = 0.0;
  for (int64_t b = blockIdx.y; b < B; b += gridDim.y) s += (double)__ldg(x + b * ld + h);
  atomicAdd(sums + h, s);
}

// F.normalize backward, one warp per row: dO = (dE - E * <E, dE>) / norm
__global__ void __launch_bounds__(256)
normalize_bwd_kernel(const float* __restrict__ dE, const float* __restrict__ E, const float* __restrict__ norms,
                     int64_t B, int64_t D, float* __restrict__ dO) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < B; r += warps) {
    float dot = 0.f;
    for (int64_t c = lane; c < D; c += 32) dot = fmaf(__ldg(E + r * D + c), __ldg(dE + r * D + c), dot);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(FULL_MASK, dot, o);
    const float inv = 1.0f / __ldg(norms + r);
    for (int64_t c = lane; c < D; c += 32)
      dO[r * D + c] = (__ldg(dE + r * D + c) - __ldg(E + r * D + c) * dot) * inv;
  }
}

static dim3 col_grid(int64_t B, int64_t H, int threads) {
  const unsigned gx = (unsigned)((H + threads - 1) / threads);
  int64_t gy = (int64_t)num_sms() * 8 / gx;
  if (gy > (B + 15) / 16) gy = (B + 15) / 16;
  if (gy < 1) gy = 1;
  if (gy > 65535) gy = 65535;
  return dim3(gx, (unsigned)gy);
}
static int flat_grid(int64_t total) {
  int64_t g = (total + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 16;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

}  // namespace b200

using namespace b200;

extern "C" int b200rec_bn_forward(const float* z, int64_t B, int64_t H, int64_t ld, int act, int training, float eps,
                                  float momentum, const float* gamma, const float* beta, float* running_mean,
                                  float* running_var, float drop_p, uint64_t seed, float* mean, float* invstd,
                                  float* y, int64_t ld_y, double* scratch, void* stream) {
  if (!z || !gamma || !beta || !mean || !invstd || !y || !scratch) return fail("bn_forward: null pointer");
  if (B <= 0 || H <= 0) return fail("bn_forward: empty input");
  if (training && B < 2) return fail("bn_forward: Expected more than 1 value per channel when training (B=%lld)", (long long)B);
  if (!training && (!running_mean || !running_var)) return fail("bn_forward: eval mode needs running statistics");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const unsigned hb = (unsigned)((H + 127) / 128);
  if (training) {
    B200_CUDA_OK(cudaMemsetAsync(scratch, 0, sizeof(double) * 2 * H, st));
    bn_stats_kernel<<<col_grid(B, H, 128), 128, 0, st>>>(z, B, H, ld, act, scratch);
    B200_LAUNCH_OK("bn_stats_kernel");
    bn_finalize_kernel<<<hb, 128, 0, st>>>(scratch, B, H, eps, momentum, mean, invstd, running_mean, running_var);
    B200_LAUNCH_OK("bn_finalize_kernel");
  } else {
    bn_eval_stats_kernel<<<hb, 128, 0, st>>>(running_mean, running_var, H, eps, mean, invstd);
    B200_LAUNCH_OK("bn_eval_stats_kernel");
  }
  bn_apply_kernel<<<flat_grid(B * H), 256, 0, st>>>(z, B, H, ld, act, mean, invstd, gamma, beta,
                                                    training ? drop_p : 0.f, seed, y, ld_y);
  B200_LAUNCH_OK("bn_apply_kernel");
  return 0;
}

extern "C" int b200rec_bn_backward(const float* dy, int64_t ld_dy, const float* z, int64_t ld_z, int64_t B, int64_t H,
                                   int act, int training, const float* mean, const float* invstd, const float* gamma,
                                   float drop_p, uint64_t seed, float* dz, int64_t ld_dz, float* dgamma, float* dbeta,
                                   float* dbias, double* scratch, void* stream) {
  if (!dy || !z || !mean || !invstd || !gamma || !dz || !dgamma || !dbeta || !scratch)
    return fail("bn_backward: null pointer");
  if (B <= 0 || H <= 0) return fail("bn_backward: empty input");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const float p = training ? drop_p : 0.f;
  B200_CUDA_OK(cudaMemsetAsync(scratch, 0, sizeof(double) * 3 * H, st));
  bn_bwd_stats_kernel<<<col_grid(B, H, 128), 128, 0, st>>>(dy, ld_dy, z, ld_z, B, H, act, mean, invstd, p, seed, scratch);
  B200_LAUNCH_OK("bn_bwd_stats_kernel");
  bn_bwd_apply_kernel<<<flat_grid(B * H), 256, 0, st>>>(dy, ld_dy, z, ld_z, B, H, act, mean, invstd, gamma, scratch,
                                                        training, p, seed, dz, ld_dz, B);
  B200_LAUNCH_OK("bn_bwd_apply_kernel");
  const unsigned hb = (unsigned)((H + 127) / 128);
  sums_to_float_kernel<<<hb, 128, 0, st>>>(scratch, H, dbeta, 1);
  B200_LAUNCH_OK("sums_to_float_kernel");
  sums_to_float_kernel<<<hb, 128, 0, st>>>(scratch + H, H, dgamma, 1);
  B200_LAUNCH_OK("sums_to_float_kernel");
  if (dbias) {
    colsum_kernel<<<col_grid(B, H, 128), 128, 0, st>>>(dz, B, H, ld_dz, scratch + 2 * H);
    B200_LAUNCH_OK("colsum_kernel");
    sums_to_float_kernel<<<hb, 128, 0, st>>>(scratch + 2 * H, H, dbias, 1);
    B200_LAUNCH_OK("sums_to_float_kernel");
  }
  return 0;
}

// Data-parallel BatchNorm (batch statistics over ALL replicas, as a single process on the global batch computes
// them): the caller all-reduces (sum) the fp64 partial sums in `scratch[0, 2H)` between phase 0 and phase 1.
//   forward  phase 0: scratch = [sum act(z), sum act(z)^2] of the local rows
//            phase 1: mean / invstd / running stats from the reduced sums over B_total rows, then y for the local rows
//   backward phase 0: scratch = [sum dy', sum dy' xhat] of the local rows; dbeta / dgamma accumulate the LOCAL sums (the
//                     flat gradient all-reduce adds the replicas up once)
//            phase 1: dz for the local rows from the reduced sums over B_total rows (+ dbias = local colsum(dz))
extern "C" int b200rec_bn_forward_dp(const float* z, int64_t B, int64_t H, int64_t ld, int act, float eps, float momentum,
                                     const float* gamma, const float* beta, float* running_mean, float* running_var,
                                     float drop_p, uint64_t seed, float* mean, float* invstd, float* y, int64_t ld_y,
                                     double* scratch, int phase, int64_t B_total, void* stream) {
  if (!z || !scratch) return fail("bn_forward_dp: null pointer");
  if (B <= 0 || H <= 0) return fail("bn_forward_dp: empty input");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (phase == 0) {
    B200_CUDA_OK(cudaMemsetAsync(scratch, 0, sizeof(double) * 2 * H, st));
    bn_stats_kernel<<<col_grid(B, H, 128), 128, 0, st>>>(z, B, H, ld, act, scratch);
    B200_LAUNCH_OK("bn_stats_kernel");
    return 0;
  }
  if (!gamma || !beta || !mean || !invstd || !y) return fail("bn_forward_dp: null pointer");
  if (B_total < 2 || B_total < B) return fail("bn_forward_dp: B_total must be >= max(2, B)");
  const unsigned hb = (unsigned)((H + 127) / 128);
  bn_finalize_kernel<<<hb, 128, 0, st>>>(scratch, B_total, H, eps, momentum, mean, invstd, running_mean, running_var);
  B200_LAUNCH_OK("bn_finalize_kernel");
  bn_apply_kernel<<<flat_grid(B * H), 256, 0, st>>>(z, B, H, ld, act, mean, invstd, gamma, beta, drop_p, seed, y, ld_y);
  B200_LAUNCH_OK("bn_apply_kernel");
  return 0;
}

extern "C" int b200rec_bn_backward_dp(const float* dy, int64_t ld_dy, const float* z, int64_t ld_z, int64_t B, int64_t H,
                                      int act, const float* mean, const float* invstd, const float* gamma, float drop_p,
                                      uint64_t seed, float* dz, int64_t ld_dz, float* dgamma, float* dbeta, float* dbias,
                                      double* scratch, int phase, int64_t B_total, void* stream) {
  if (!dy || !z || !mean || !invstd || !gamma || !scratch) return fail("bn_backward_dp: null pointer");
  if (B <= 0 || H <= 0) return fail("bn_backward_dp: empty input");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const unsigned hb = (unsigned)((H + 127) / 128);
  if (phase == 0) {
    if (!dgamma || !dbeta) return fail("bn_backward_dp: null pointer");
    B200_CUDA_OK(cudaMemsetAsync(scratch, 0, sizeof(double) * 3 * H, st));
    bn_bwd_stats_kernel<<<col_grid(B, H, 128), 128, 0, st>>>(dy, ld_dy, z, ld_z, B, H, act, mean, invstd, drop_p, seed, scratch);
    B200_LAUNCH_OK("bn_bwd_stats_kernel");
    sums_to_float_kernel<<<hb, 128, 0, st>>>(scratch, H, dbeta, 1);
    B200_LAUNCH_OK("sums_to_float_kernel");
    sums_to_float_kernel<<<hb, 128, 0, st>>>(scratch + H, H, dgamma, 1);
    B200_LAUNCH_OK("sums_to_float_kernel");
    return 0;
  }
  if (!dz) return fail("bn_backward_dp: null pointer");
  if (B_total < B) return fail("bn_backward_dp: B_total must be >= B");
  bn_bwd_apply_kernel<<<flat_grid(B * H), 256, 0, st>>>(dy, ld_dy, z, ld_z, B, H, act, mean, invstd, gamma, scratch, 1,
                                                        drop_p, seed, dz, ld_dz, B_total);
  B200_LAUNCH_OK("bn_bwd_apply_kernel");
  if (dbias) {
    B200_CUDA_OK(cudaMemsetAsync(scratch + 2 * H, 0, sizeof(double) * H, st));
    colsum_kernel<<<col_grid(B, H, 128), 128, 0, st>>>(dz, B, H, ld_dz, scratch + 2 * H);
    B200_LAUNCH_OK("colsum_kernel");
    sums_to_float_kernel<<<hb, 128, 0, st>>>(scratch + 2 * H, H, dbias, 1);
    B200_LAUNCH_OK("sums_to_float_kernel");
  }
  return 0;
}

// step = ++*step_dev; hyper_dev = {lr, 1 - beta1^step, sqrt(1 - beta2^step)} for b200rec_adam_*_dev; the dropout seeds
// of this step are salted with a hash of (salt_key, step).  One thread.
__global__ void train_step_begin_kernel(long long* step_dev, const float* lr_dev, float beta1, float beta2,
                                        float* hyper_dev, unsigned long long salt_key) {
  const long long step = *step_dev + 1;
  *step_dev = step;
  hyper_dev[0] = *lr_dev;
  hyper_dev[1] = (float)(1.0 - pow((double)beta1, (double)step));
  hyper_dev[2] = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  unsigned long long x = salt_key + 0x9E3779B97F4A7C15ull * (unsigned long long)step;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  g_seed_salt = salt_key ? (x ^ (x >> 31)) : 0ull;
}

extern "C" int b200rec_train_step_begin(int64_t* step_dev, const float* lr_dev, float beta1, float beta2, float* hyper_dev,
                                        uint64_t salt_key, void* stream) {
  if (!step_dev || !lr_dev || !hyper_dev) return fail("train_step_begin: null pointer");
  train_step_begin_kernel<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<long long*>(step_dev), lr_dev, beta1, beta2, hyper_dev, (unsigned long long)salt_key);
  B200_LAUNCH_OK("train_step_begin_kernel");
  return 0;
}

extern "C" int b200rec_colsum(const float* x, int64_t B, int64_t H, int64_t ld, float* out, int accumulate,
                              double* scratch, void* stream) {
  if (!x || !out || !scratch) return fail("colsum: null pointer");
  if (B <= 0 || H <= 0) return fail("colsum: empty input");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  B200_CUDA_OK(cudaMemsetAsync(scratch, 0, sizeof(double) * H, st));
  colsum_kernel<<<col_grid(B, H, 128), 128, 0, st>>>(x, B, H, ld, scratch);
  B200_LAUNCH_OK("colsum_kernel");
  sums_to_float_kernel<<<(unsigned)((H + 127) / 128), 128, 0, st>>>(scratch, H, out, accumulate);
  B200_LAUNCH_OK("sums_to_float_kernel");
  return 0;
}

extern "C" int b200rec_normalize_bwd(const float* dE, const float* E, const float* norms, int64_t B, int64_t D,
                                     float* dO, void* stream) {
  if (!dE || !E || !norms || !dO) return fail("normalize_bwd: null pointer");
  if (B <= 0 || D <= 0) return fail("normalize_bwd: empty input");
  const int64_t blocks = (B + 7) / 8;
  const int grid = (int)(blocks < (int64_t)num_sms() * 16 ? blocks : (int64_t)num_sms() * 16);
  normalize_bwd_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dE, E, norms, B, D, dO);
  B200_LAUNCH_OK("normalize_bwd_kernel");
  return 0;
}

extern "C" int b200rec_act_dropout(const float* z, int64_t B, int64_t H, int64_t ld, int act, float drop_p,
                                   uint64_t seed, float* y, int64_t ld_y, void* stream) {
  if (!z || !y) return fail("act_dropout: null pointer");
  if (B <= 0 || H <= 0) return fail("act_dropout: empty input");
  act_dropout_kernel<<<flat_grid(B * H), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(z, B, H, ld, act, drop_p, seed,
                                                                                          y, ld_y);
  B200_LAUNCH_OK("act_dropout_kernel");
  return 0;
}

extern "C" int b200rec_act_dropout_bwd(const float* dy, int64_t ld_dy, const float* z, int64_t ld_z, int64_t B, int64_t H,
                                       int act, float drop_p, uint64_t seed, float* dz, int64_t ld_dz, void* stream) {
  if (!dy || !z || !dz) return fail("act_dropout_bwd: null pointer");
  if (B <= 0 || H <= 0) return fail("act_dropout_bwd: empty input");
  act_dropout_bwd_kernel<<<flat_grid(B * H), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      dy, ld_dy, z, ld_z, B, H, act, drop_p, seed, dz, ld_dz);
  B200_LAUNCH_OK("act_dropout_bwd_kernel");
  return 0;
}

// fused per-layer kernels (same translation unit: they share g_seed_salt with the kernels above)
#include "mlp_fused.cuh"

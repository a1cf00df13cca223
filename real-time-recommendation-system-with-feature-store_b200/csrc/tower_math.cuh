// Element-wise math shared by the tower kernels: activations (reference src/models/two_tower.py:77-86) and the
// counter-based dropout mask.  Included by ONE translation unit (tower_ops.cu): g_seed_salt is a __device__ variable.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

// ------------------------------------------------------------------ activations (two_tower.py:77-86)
__device__ __forceinline__ float act_fwd(int act, float z) {
  switch (act) {
    case 0: return z > 0.f ? z : 0.f;                                      // relu
    case 1: return 0.5f * z * (1.0f + erff(z * 0.70710678118654752440f));  // gelu (erf form, torch default)
    case 2: return z > 0.f ? z : 0.1f * z;                                 // leaky_relu(0.1)
    case 3: return tanhf(z);
    case 4: return 1.0f / (1.0f + expf(-z));
    default: return z;
  }
}
__device__ __forceinline__ float act_grad(int act, float z) {
  switch (act) {
    case 0: return z > 0.f ? 1.f : 0.f;
    case 1: {
      const float cdf = 0.5f * (1.0f + erff(z * 0.70710678118654752440f));
      const float pdf = 0.39894228040143267794f * expf(-0.5f * z * z);
      return cdf + z * pdf;
    }
    case 2: return z > 0.f ? 1.f : 0.1f;
    case 3: {
      const float t = tanhf(z);
      return 1.f - t * t;
    }
    case 4: {
      const float s = 1.0f / (1.0f + expf(-z));
      return s * (1.f - s);
    }
    default: return 1.f;
  }
}

// ------------------------------------------------------------------ counter-based dropout mask
// One 32-bit hash per element: the "lowbias32" integer finaliser (xorshift-multiply, full avalanche) over the element
// index mixed with the 64-bit seed; an element is kept when the top 24 bits reach p * 2^24.  The fused layer kernels
// re-derive a block's mask wherever they rebuild an operand from the saved pre-activations (forward consumer, weight
// gradient, data gradient: ~5 evaluations per element and step), so the mask has to cost a handful of integer
// instructions — the Philox-4x32-10 used before was ~100 per element and dominated the operand loaders.  Bit parity
// with torch's generator is impossible in a fused kernel either way.
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint32_t drop_threshold(float p) { return (uint32_t)(p * 16777216.0f); }
// 24 uniform bits of element idx under (s_lo, s_hi) = the two halves of the salted seed
__device__ __forceinline__ uint32_t drop_bits(uint32_t s_lo, uint32_t s_hi, unsigned long long idx) {
  const uint32_t x = mix32(((uint32_t)idx ^ s_lo) + (uint32_t)(idx >> 32) * 0x9E3779B1u);
  return (x ^ s_hi) >> 8;
}
// Per-step salt of every dropout seed, set on the device by b200rec_train_step_begin.  A training step captured in a
// CUDA graph replays with the kernel arguments of the capture, so what must change from step to step (the dropout
// streams here, Adam's bias corrections in optim.cu) lives in device memory.  0 (never set) in eager training.
__device__ unsigned long long g_seed_salt = 0ull;

// keep-scale of element (row, col): 0 when dropped, 1/(1-p) when kept; p == 0 -> 1
__device__ __forceinline__ float drop_scale(float p, uint64_t seed, int64_t row, int64_t col, int64_t H) {
  if (p <= 0.f) return 1.f;
  seed ^= g_seed_salt;
  const unsigned long long idx = (unsigned long long)row * (unsigned long long)H + (unsigned long long)col;
  return drop_bits((uint32_t)seed, (uint32_t)(seed >> 32), idx) >= drop_threshold(p) ? __frcp_rn(1.0f - p) : 0.f;
}

// ------------------------------------------------------------------ out-of-line activations for the fused layer kernels
// mlp_fused.cuh instantiates its loaders for ReLU (inline, branch-free) and for "any other activation": with the 5-way
// switch (erff / tanhf / expf) inlined in every loader the kernel was 41 K SASS instructions and spent 23 % of its issue
// slots waiting for instruction fetch.
__device__ __noinline__ float act_fwd_slow(int act, float z) { return act_fwd(act, z); }
__device__ __noinline__ float act_grad_slow(int act, float z) { return act_grad(act, z); }

}  // namespace b200

// Fused tower-MLP layer kernels (K2): Linear forward, data gradient and weight gradient on tcgen05 with everything
// around the GEMM folded into the operand loaders and the epilogue (reference src/models/two_tower.py:56-72,128-132:
// [Linear -> activation -> BatchNorm1d -> Dropout] x L -> Linear -> F.normalize, and what autograd derives from it).
//
// The unfused chain cost 6 launches per hidden layer forward (operand split x2, GEMM, BN statistics / finalize / apply)
// and ~12 backward; each of them moves the [B, H] activations through HBM again.  Here one launch per Linear and
// direction does all of it:
//
//   operand loaders (all 8 warps): read fp32 from global, apply the transform that PRODUCES the operand
//       forward input   y = Dropout(BN(act(z_prev)))          (z_prev = pre-activation of the block below)
//       backward input  dz = act'(z) * gamma*invstd * (g - sum(g)/n - xhat*sum(g*xhat)/n),  g = dy * dropout mask
//     split the value into bf16 pieces x = h + m (+ l) and store them as K-major SWIZZLE_128B tiles — the layout
//     tcgen05.mma reads — so no split-bf16 copy of any activation or weight ever exists in HBM;
//   tcgen05.mma: piece products h.h, h.m, m.h (+ h.l, l.h, m.m) = fp32-grade products, accumulated per magnitude class
//     in separate TMEM accumulators (small | mid | h.h even chunks | h.h odd chunks): tcgen05 adds with truncation, so
//     low-order products never meet a large partial sum and the h.h chains are halved;
//   epilogue (TMEM -> registers): + bias, write the pre-activation (forward) / dX (dgrad) / atomically add dW, db (wgrad);
//     the BatchNorm statistics of THIS block (forward: sum a, sum a^2) or the backward sums of the block below (dgrad:
//     sum g, sum g*xhat) are reduced over the tile's 128 rows through shared memory and merged with fp64 atomics; the
//     last Linear normalises its rows (F.normalize) in registers.
//
// A layer boundary is a kernel boundary because BatchNorm needs batch-wide sums (under data parallel the caller
// all-reduces exactly those sums between two launches).  mean / invstd are recomputed from the fp64 sums by every CTA
// that needs them (<= 512 features); the CTA (0,0) of the consuming forward launch applies the running-statistics
// update.  Operands are built with plain 128-bit global loads (the fp32 -> piece conversion has to pass through
// registers anyway); TMA feeds the kernels whose operands already exist in HBM in MMA layout (top-K, LSE, in-batch
// gradient, the generic GEMM).
#pragma once
#include "host_util.h"
#include "tc_common.cuh"
#include "tower_math.cuh"
#include "../../include/b200rec.h"

namespace b200 {

constexpr int MF_THREADS = 512;
constexpr int MF_EPI = MF_THREADS / 128;   // warps per TMEM lane quarter in the epilogue
constexpr int MF_MAXH = 512;          // widest BatchNorm block whose per-feature constants are staged in shared memory
constexpr int MF_SMEM_LIMIT = 232448;

// development builds only (tools/build_variant.py mf_timing -DMF_TIMING): phase time stamps of CTA (0,0,0)
#ifdef MF_TIMING
__device__ unsigned long long g_mf_stamps[32];
__device__ __forceinline__ void mf_stamp(int i) {
  if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_mf_stamps[i] = t;
  }
}
#define MF_STAMP(i) mf_stamp(i)
#else
#define MF_STAMP(i)
#endif

struct MfBlock {                      // device view of b200rec_bn_block
  const float* z;
  long long ldz;
  int H, act, training, update_running;
  const double* sums;
  const float* gamma;
  const float* beta;
  float* running_mean;
  float* running_var;
  long long* nbt;
  float eps, momentum, drop_p;
  unsigned long long seed;
  long long B_stat;
};

struct MfArgs {
  int mode;                           // 0 forward, 1 dgrad, 2 wgrad
  int np;                             // bf16 pieces per operand: 1 (bf16), 2 (3 products), 3 (6 products)
  int B, N, K;                        // batch rows, out features, in features of this Linear
  int NT;                             // accumulator columns of one CTA tile (multiple of 16, <= 128)
  int chunks;                         // 64-wide contraction chunks per CTA
  const float* x;                     // layer input when there is no block below: [B, K]
  long long ldx;
  MfBlock lower;                      // block whose output is this layer's input
  int has_lower;
  MfBlock own;                        // this layer's own block (backward of a hidden layer)
  int has_own;
  const float* w;                     // [N, K]
  long long ldw;
  const float* bias;
  float* out;                         // forward: z (or the normalised embedding); dgrad: dX
  long long ldo;
  double* out_sums;                   // forward: [2N] statistics of the own block; dgrad: [2K] backward sums of `lower`
  int out_act;
  int normalize;
  float* norms;
  const float* dy;                    // gradient wrt the block output (hidden layer) or wrt z (last layer): [B, N]
  long long lddy;
  const double* own_bsums;            // [2N] sum g, sum g*xhat over B_stat rows
  const double* own_bsums_local;      // wgrad: this replica's part (added to dbeta / dgamma)
  float* dw;
  long long lddw;
  float* db;
  float* dgamma;
  float* dbeta;
};

__device__ __forceinline__ void mf_store_unit(uint8_t* base, int piece_stride, int np, int r, int ch, const float (&v)[8]) {
  uint32_t hw[4], mw[4], lw[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float f0 = v[2 * j], f1 = v[2 * j + 1];
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(f0, f1);
    const float2 hf = __bfloat1622float2(h2);
    const float r0 = f0 - hf.x, r1 = f1 - hf.y;
    const __nv_bfloat162 m2 = __floats2bfloat162_rn(r0, r1);
    const float2 mf = __bfloat1622float2(m2);
    const __nv_bfloat162 l2 = __floats2bfloat162_rn(r0 - mf.x, r1 - mf.y);
    hw[j] = *reinterpret_cast<const uint32_t*>(&h2);
    mw[j] = *reinterpret_cast<const uint32_t*>(&m2);
    lw[j] = *reinterpret_cast<const uint32_t*>(&l2);
  }
  // K-major SW128: row r at r*128 B, its 16-byte chunk ch stored at position ch ^ (r & 7)
  uint8_t* p = base + r * 128 + ((ch ^ (r & 7)) << 4);
  *reinterpret_cast<uint4*>(p) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
  if (np >= 2) *reinterpret_cast<uint4*>(p + piece_stride) = make_uint4(mw[0], mw[1], mw[2], mw[3]);
  if (np >= 3) *reinterpret_cast<uint4*>(p + 2 * piece_stride) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}

// rows x 64 operand tile in two phases per pair of units, so that the global loads of both units are in flight before
// any of them is consumed: ld(r, c0, raw) issues UNCONDITIONAL loads (indices clamped into range), tf(r, c0, raw, v)
// turns them into the 8 contraction elements c0..c0+7 of tile row r (zero outside the matrix).  With loads inside
// bounds-checked branches the compiler serialised them: 8 dependent L2 round trips per unit.
// trans = 0: consecutive threads walk along the contraction (contiguous in the source); 1: along the tile rows.
template <int NRAW, class L, class T>
__device__ __forceinline__ void mf_load_tile(uint8_t* base, int piece_stride, int np, int rows, int trans, L ld, T tf) {
  const int units = rows * 8;
  for (int u0 = threadIdx.x; u0 < units; u0 += 2 * MF_THREADS) {
    const bool has1 = u0 + MF_THREADS < units;
    const int u1 = has1 ? u0 + MF_THREADS : u0;
    int r0, ch0, r1, ch1;
    if (trans) {
      ch0 = u0 / rows, r0 = u0 - ch0 * rows;
      ch1 = u1 / rows, r1 = u1 - ch1 * rows;
    } else {
      r0 = u0 >> 3, ch0 = u0 & 7;
      r1 = u1 >> 3, ch1 = u1 & 7;
    }
    float raw0[NRAW], raw1[NRAW];
    ld(r0, ch0 * 8, raw0);
    ld(r1, ch1 * 8, raw1);
    float v[8];
    tf(r0, ch0 * 8, raw0, v);
    mf_store_unit(base, piece_stride, np, r0, ch0, v);
    if (has1) {
      tf(r1, ch1 * 8, raw1, v);
      mf_store_unit(base, piece_stride, np, r1, ch1, v);
    }
  }
}

// 8 consecutive floats row[c..c+7], unconditional: vec (row base 16-byte aligned, limit % 8 == 0) -> two 128-bit loads of
// a unit that is entirely inside or entirely outside [0, limit) (outside: unit 0 is loaded and ignored); otherwise eight
// scalar loads with clamped indices.  The caller zeroes what lies outside.
__device__ __forceinline__ void mf_ld8(const float* row, int c, int limit, bool vec, float* v) {
  if (vec) {
    const int cc = c < limit ? c : 0;
    const float4 a = __ldg(reinterpret_cast<const float4*>(row + cc));
    const float4 b = __ldg(reinterpret_cast<const float4*>(row + cc + 4));
    v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __ldg(row + (c + j < limit ? c + j : limit - 1));
  }
}
__device__ __forceinline__ bool mf_vec_ok(const float* base, long long ld, int limit) {
  return ((reinterpret_cast<uintptr_t>(base) & 15) == 0) && ((ld & 3) == 0) && ((limit & 7) == 0);
}

struct MfStats {       // per-feature constants of a block in shared memory
  float* mean;
  float* invstd;
  float* p2;           // forward transform: scale = gamma*invstd         | backward transform: gamma*invstd
  float* p3;           // forward transform: shift = beta - mean*scale    | backward transform: sum(g)/n
  float* p4;           //                                                 | backward transform: sum(g*xhat)/n
};

// dropout stream of one block, hoisted out of the per-element path (same arithmetic as drop_scale in tower_math.cuh)
struct MfDrop {
  float keep;
  uint32_t thresh, s_lo, s_hi;
  long long H;
};
__device__ __forceinline__ MfDrop mf_drop(const MfBlock& k) {
  MfDrop d;
  d.keep = k.drop_p > 0.f ? __frcp_rn(1.0f - k.drop_p) : 1.f;
  d.thresh = k.drop_p > 0.f ? drop_threshold(k.drop_p) : 0u;
  const unsigned long long seed = k.seed ^ g_seed_salt;
  d.s_lo = (uint32_t)seed;
  d.s_hi = (uint32_t)(seed >> 32);
  d.H = k.H;
  return d;
}
template <bool DROP>
__device__ __forceinline__ float mf_keep(const MfDrop& d, long long row, int col) {
  if (!DROP) return 1.f;
  const unsigned long long idx = (unsigned long long)row * (unsigned long long)d.H + (unsigned long long)col;
  return drop_bits(d.s_lo, d.s_hi, idx) >= d.thresh ? d.keep : 0.f;   // thresh == 0 (p == 0): always kept, keep == 1
}
template <bool RELU>
__device__ __forceinline__ float mf_act(int act, float z) { return RELU ? fmaxf(z, 0.f) : act_fwd_slow(act, z); }
template <bool RELU>
__device__ __forceinline__ float mf_act_grad(int act, float z) { return RELU ? (z > 0.f ? 1.f : 0.f) : act_grad_slow(act, z); }

// y = Dropout(BN(act(z)))[b, h]:  act(z) * scale[h] + shift[h], times the keep-scale
template <bool RELU, bool DROP>
__device__ __forceinline__ float mf_fwd_val(const MfBlock& k, const MfStats& s, const MfDrop& d, long long b, int h, float zz) {
  const float y = fmaf(mf_act<RELU>(k.act, zz), s.p2[h], s.p3[h]);
  return y * mf_keep<DROP>(d, b, h);
}
// dz[b, n] = act'(z) * gamma*invstd * (g - sum(g)/n - xhat * sum(g*xhat)/n), g = dy * keep-scale  (training)
template <bool RELU, bool DROP>
__device__ __forceinline__ float mf_bwd1(const MfBlock& k, const MfStats& s, const MfDrop& d, float dy, float zz, long long b, int n) {
  const float g = dy * mf_keep<DROP>(d, b, n);
  const float xhat = (mf_act<RELU>(k.act, zz) - s.mean[n]) * s.invstd[n];
  const float da = s.p2[n] * (g - s.p3[n] - xhat * s.p4[n]);   // eval mode: p3 = p4 = 0
  return da * mf_act_grad<RELU>(k.act, zz);
}

template <int MODE, bool RELU, bool DROP>
__global__ void __launch_bounds__(MF_THREADS, 1) mlp_fused_kernel(const MfArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space (LDS / STS)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int np = a.np, NT = a.NT;
  const int a_piece = 128 * 128, b_piece = NT * 128;
  const int stage_bytes = np * (a_piece + b_piece);
  const int stage_region = 2 * stage_bytes;
  const int staging = 2 * 128 * (NT + 1) * 4;
  const int region = stage_region > staging ? stage_region : staging;
  float* s_par = reinterpret_cast<float*>(smem + region);          // 9 x MF_MAXH floats
  MfStats lo{s_par, s_par + MF_MAXH, s_par + 2 * MF_MAXH, s_par + 3 * MF_MAXH, nullptr};
  MfStats ow{s_par + 4 * MF_MAXH, s_par + 5 * MF_MAXH, s_par + 6 * MF_MAXH, s_par + 7 * MF_MAXH, s_par + 8 * MF_MAXH};
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_par + 9 * MF_MAXH);
  uint64_t* free_bar = bars;          // [2] the MMAs that read stage s have completed
  uint64_t* acc_bar = bars + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
  float* s_part = reinterpret_cast<float*>(bars + 4);              // [MF_EPI][128] row partials (normalise / db)
  float* s_bias = s_part + MF_EPI * 128;                           // [128] bias of this tile's out features (forward)

  MF_STAMP(0);
  if (warp == 1 && lane == 0) {
    mbar_init(&free_bar[0], 1);
    mbar_init(&free_bar[1], 1);
    mbar_init(acc_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);

  // ---------------------------------------------------------------- per-feature constants of the blocks involved
  // (one pass per block: all global loads of a feature are issued before the fp64 arithmetic that needs them)
  const bool cta0 = blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  MfDrop dlo{}, dow{};
  if (a.has_lower) {
    const MfBlock& k = a.lower;
    dlo = mf_drop(k);
    const bool upd = cta0 && MODE == 0 && k.training && k.update_running && k.running_mean;
    for (int h = threadIdx.x; h < k.H; h += MF_THREADS) {
      const float gam = __ldg(k.gamma + h), bet = __ldg(k.beta + h);
      float m, is;
      if (k.training) {
        const double s1 = k.sums[h], s2 = k.sums[k.H + h];
        const double n = (double)k.B_stat;
        const double mm = s1 / n;
        double var = s2 / n - mm * mm;
        if (var < 0.0) var = 0.0;
        m = (float)mm;
        is = (float)(1.0 / sqrt(var + (double)k.eps));
        if (upd) {   // nn.BatchNorm1d's momentum update with the unbiased variance, once per forward
          const double unbiased = k.B_stat > 1 ? var * n / (n - 1.0) : var;
          k.running_mean[h] = (float)((1.0 - k.momentum) * (double)k.running_mean[h] + (double)k.momentum * mm);
          k.running_var[h] = (float)((1.0 - k.momentum) * (double)k.running_var[h] + (double)k.momentum * unbiased);
        }
      } else {
        m = k.running_mean[h];
        is = 1.0f / sqrtf(k.running_var[h] + k.eps);
      }
      lo.mean[h] = m;
      lo.invstd[h] = is;
      const float sc = is * gam;
      lo.p2[h] = sc;
      lo.p3[h] = bet - m * sc;
    }
    if (upd && k.nbt && threadIdx.x == 0) *k.nbt += 1;
  }
  if (a.has_own) {
    const MfBlock& k = a.own;
    dow = mf_drop(k);
    const double inv_n = 1.0 / (double)k.B_stat;
    const bool affine = MODE == 2 && cta0 && a.dgamma && a.own_bsums_local;
    for (int n = threadIdx.x; n < k.H; n += MF_THREADS) {
      const float gam = __ldg(k.gamma + n);
      double b1 = 0.0, b2 = 0.0, l1 = 0.0, l2 = 0.0;
      if (k.training) b1 = a.own_bsums[n], b2 = a.own_bsums[k.H + n];
      if (affine) l1 = a.own_bsums_local[n], l2 = a.own_bsums_local[k.H + n];
      float m, is;
      if (k.training) {
        const double nn = (double)k.B_stat;
        const double mm = k.sums[n] / nn;
        double var = k.sums[k.H + n] / nn - mm * mm;
        if (var < 0.0) var = 0.0;
        m = (float)mm;
        is = (float)(1.0 / sqrt(var + (double)k.eps));
      } else {
        m = k.running_mean[n];
        is = 1.0f / sqrtf(k.running_var[n] + k.eps);
      }
      ow.mean[n] = m;
      ow.invstd[n] = is;
      ow.p2[n] = gam * is;
      ow.p3[n] = (float)(b1 * inv_n);      // eval mode: 0
      ow.p4[n] = (float)(b2 * inv_n);
      if (affine) {                        // BatchNorm affine gradients: dbeta = sum g, dgamma = sum g*xhat
        a.dbeta[n] += (float)l1;
        a.dgamma[n] += (float)l2;
      }
    }
  }
  if (MODE == 0 && threadIdx.x < 128) {
    const int n = blockIdx.y * NT + threadIdx.x;
    s_bias[threadIdx.x] = (a.bias != nullptr && n < a.N) ? __ldg(a.bias + n) : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  MF_STAMP(1);

  // ---------------------------------------------------------------- tile coordinates
  // forward: rows m0.. of the batch x out features n0..;  dgrad: rows m0.. x in features n0.. (contraction over N)
  // wgrad:   out features m0.. x in features n0.., contraction over this CTA's batch rows
  int m0, n0, c_begin;
  if (MODE == 2) {
    m0 = blockIdx.y * 128;
    n0 = blockIdx.z * NT;
    c_begin = blockIdx.x * a.chunks;
  } else {
    m0 = blockIdx.x * 128;
    n0 = blockIdx.y * NT;
    c_begin = 0;
  }
  int total_chunks;
  if (MODE == 0) total_chunks = (a.K + 63) / 64;
  else if (MODE == 1) total_chunks = (a.N + 63) / 64;
  else total_chunks = (a.B + 63) / 64;
  int nchunks = total_chunks - c_begin;
  if (nchunks > a.chunks) nchunks = a.chunks;

  const uint32_t idesc = umma_idesc_bf16(128, (uint32_t)NT);
  const uint64_t desc_base = umma_desc_k_sw128(0);
  const uint32_t smem_lo = (smem_u32(smem) & 0x3FFFFu) >> 4;
  uint32_t used = 0;   // accumulators that already hold a value (warp-uniform in the issuing warp)
  float dbacc = 0.f;   // wgrad: this thread's part of db[m0 + threadIdx.x % 128] (every unit of a thread has that row)

  for (int ci = 0; ci < nchunks; ++ci) {
    const int st = ci & 1;
    const int c0 = (c_begin + ci) * 64;   // first contraction index of this chunk
    if (ci >= 2) mbar_wait(&free_bar[st], ((ci >> 1) - 1) & 1);
    uint8_t* abase = smem + st * stage_bytes;
    uint8_t* bbase = abase + np * a_piece;
    if constexpr (MODE == 0) {
      // A[r, c] = input[m0 + r, c0 + c]
      const float* src = a.has_lower ? a.lower.z : a.x;
      const long long lds = a.has_lower ? a.lower.ldz : a.ldx;
      const bool va = mf_vec_ok(src, lds, a.K), vb = mf_vec_ok(a.w, a.ldw, a.K);
      mf_load_tile<8>(abase, a_piece, np, 128, 0,
          [&](int r, int c, float* raw) {
            const long long b = m0 + r < a.B ? m0 + r : a.B - 1;
            mf_ld8(src + b * lds, c0 + c, a.K, va, raw);
          },
          [&](int r, int c, const float* raw, float (&v)[8]) {
            const long long b = m0 + r;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int k = c0 + c + j;
              const bool ok = b < a.B && k < a.K;
              const float y = a.has_lower ? mf_fwd_val<RELU, DROP>(a.lower, lo, dlo, b, ok ? k : 0, raw[j]) : raw[j];
              v[j] = ok ? y : 0.f;
            }
          });
      // B[r, c] = W[n0 + r, c0 + c]
      mf_load_tile<8>(bbase, b_piece, np, NT, 0,
          [&](int r, int c, float* raw) {
            const int n = n0 + r < a.N ? n0 + r : a.N - 1;
            mf_ld8(a.w + (long long)n * a.ldw, c0 + c, a.K, vb, raw);
          },
          [&](int r, int c, const float* raw, float (&v)[8]) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = (n0 + r < a.N && c0 + c + j < a.K) ? raw[j] : 0.f;
          });
    } else if constexpr (MODE == 1) {
      // A[r, c] = dz[m0 + r, c0 + c]
      const bool vd = mf_vec_ok(a.dy, a.lddy, a.N), vz = a.has_own && mf_vec_ok(a.own.z, a.own.ldz, a.N);
      mf_load_tile<16>(abase, a_piece, np, 128, 0,
          [&](int r, int c, float* raw) {
            const long long b = m0 + r < a.B ? m0 + r : a.B - 1;
            mf_ld8(a.dy + b * a.lddy, c0 + c, a.N, vd, raw);
            if (a.has_own) mf_ld8(a.own.z + b * a.own.ldz, c0 + c, a.N, vz, raw + 8);
          },
          [&](int r, int c, const float* raw, float (&v)[8]) {
            const long long b = m0 + r;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int n = c0 + c + j;
              const bool ok = b < a.B && n < a.N;
              const float d = a.has_own ? mf_bwd1<RELU, DROP>(a.own, ow, dow, raw[j], raw[8 + j], b, ok ? n : 0) : raw[j];
              v[j] = ok ? d : 0.f;
            }
          });
      // B[r, c] = W[c0 + c, n0 + r]   (transposed read: consecutive threads -> consecutive in-features)
      mf_load_tile<8>(bbase, b_piece, np, NT, 1,
          [&](int r, int c, float* raw) {
            const int k = n0 + r < a.K ? n0 + r : a.K - 1;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int n = c0 + c + j < a.N ? c0 + c + j : a.N - 1;
              raw[j] = __ldg(a.w + (long long)n * a.ldw + k);
            }
          },
          [&](int r, int c, const float* raw, float (&v)[8]) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = (n0 + r < a.K && c0 + c + j < a.N) ? raw[j] : 0.f;
          });
    } else {
      // A[r, c] = dz[c0 + c, m0 + r]   (out features along the tile rows, batch along the contraction)
      mf_load_tile<16>(abase, a_piece, np, 128, 1,
          [&](int r, int c, float* raw) {
            const int n = m0 + r < a.N ? m0 + r : a.N - 1;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const long long b = c0 + c + j < a.B ? c0 + c + j : a.B - 1;
              raw[j] = __ldg(a.dy + b * a.lddy + n);
              if (a.has_own) raw[8 + j] = __ldg(a.own.z + b * a.own.ldz + n);
            }
          },
          [&](int r, int c, const float* raw, float (&v)[8]) {
            const int n = m0 + r;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const long long b = c0 + c + j;
              const bool ok = n < a.N && b < a.B;
              const float d = a.has_own ? mf_bwd1<RELU, DROP>(a.own, ow, dow, raw[j], raw[8 + j], b, ok ? n : 0) : raw[j];
              v[j] = ok ? d : 0.f;
              dbacc += v[j];
            }
          });
      // B[r, c] = input[c0 + c, n0 + r]
      const float* src = a.has_lower ? a.lower.z : a.x;
      const long long lds = a.has_lower ? a.lower.ldz : a.ldx;
      mf_load_tile<8>(bbase, b_piece, np, NT, 1,
          [&](int r, int c, float* raw) {
            const int k = n0 + r < a.K ? n0 + r : a.K - 1;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const long long b = c0 + c + j < a.B ? c0 + c + j : a.B - 1;
              raw[j] = __ldg(src + b * lds + k);
            }
          },
          [&](int r, int c, const float* raw, float (&v)[8]) {
            const int k = n0 + r;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const long long b = c0 + c + j;
              const bool ok = b < a.B && k < a.K;
              const float x = a.has_lower ? mf_fwd_val<RELU, DROP>(a.lower, lo, dlo, b, ok ? k : 0, raw[j]) : raw[j];
              v[j] = ok ? x : 0.f;
            }
          });
    }
    if (ci < 4) MF_STAMP(2 + 3 * ci);
    fence_proxy_async_smem();   // generic-proxy stores above -> visible to the tensor core's async-proxy reads
    __syncthreads();
    if (ci < 4) MF_STAMP(3 + 3 * ci);
    if (warp == 0) {
      tc_fence_after();
      uint32_t u = used;
      if (elect_one()) {
        const uint32_t a_lo = smem_lo + ((st * stage_bytes) >> 4);
        const uint32_t b_lo = a_lo + ((np * a_piece) >> 4);
        const uint32_t ap = a_piece >> 4, bp = b_piece >> 4;
        // piece products, smallest first: class 0 = small (m.m, l.h, h.l), 1 = mid (m.h, h.m), 2/3 = h.h (even / odd chunks)
        auto product = [&](int pa, int pb, int acc) {
          const uint64_t a_desc = desc_base + (a_lo + pa * ap);
          const uint64_t b_desc = desc_base + (b_lo + pb * bp);
          const uint32_t d_addr = tmem_base + acc * 128;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_bf16(d_addr, a_desc + 2 * k, b_desc + 2 * k, idesc, (u >> acc) & 1u);
            u |= 1u << acc;
          }
        };
        if (np == 3) {
          product(1, 1, 0);
          product(2, 0, 0);
          product(0, 2, 0);
        }
        if (np >= 2) {
          product(1, 0, 1);
          product(0, 1, 1);
        }
        product(0, 0, 2 + (ci & 1));
        umma_commit(&free_bar[st]);
        if (ci == nchunks - 1) umma_commit(acc_bar);
      }
      __syncwarp();
      if (ci < 4) MF_STAMP(4 + 3 * ci);
    }
    // every warp tracks the same accumulator bookkeeping (the epilogue needs it)
    if (np >= 3) used |= 1u;
    if (np >= 2) used |= 2u;
    used |= 4u << (ci & 1);
  }

  // ---------------------------------------------------------------- epilogue
  if (nchunks > 0) {
    if (MODE == 2 && a.db != nullptr && blockIdx.z == 0) {
      // db = colsum(dz): the loader already summed this thread's share of row m0 + (threadIdx.x & 127) in fp32
      if (threadIdx.x < 128) s_part[threadIdx.x] = 0.f;
      __syncthreads();
      atomicAdd(&s_part[threadIdx.x & 127], dbacc);
      __syncthreads();
      if (threadIdx.x < 128 && m0 + threadIdx.x < a.N) atomicAdd(a.db + m0 + threadIdx.x, s_part[threadIdx.x]);
    }
    MF_STAMP(14);
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    MF_STAMP(15);
    const int quarter = warp & 3, half = warp >> 2;   // half = which of the MF_EPI column-group sets this warp reads
    const int row_l = quarter * 32 + lane;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const int ngroups = NT / 16;
    auto load16 = [&](int g, float (&o)[16]) {   // sum of the accumulators that were used, smallest class first
      uint32_t w[4][16];
#pragma unroll
      for (int acc = 0; acc < 4; ++acc)
        if ((used >> acc) & 1u) tmem_ld_32x16(trow + acc * 128 + g * 16, w[acc]);
      tmem_ld_wait();
      bool first = true;
#pragma unroll
      for (int acc = 0; acc < 4; ++acc) {
        if (!((used >> acc) & 1u)) continue;
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] = first ? __uint_as_float(w[acc][j]) : o[j] + __uint_as_float(w[acc][j]);
        first = false;
      }
    };
    float* stage0 = reinterpret_cast<float*>(smem);
    float* stage1 = stage0 + 128 * (NT + 1);
    // column sums of a staged [128, NT] tile: thread = (column, which array, quarter of the rows); fp32 partials over 32
    // rows in four independent chains, merged in fp64 by the atomics (one per thread)
    auto column_sums = [&](const float* sa, const float* sb, bool square_b, double* dst, int limit) {
      const int col = threadIdx.x & 127, part = threadIdx.x >> 7, which = part & 1, r0 = (part >> 1) * (256 / MF_EPI);
      if (col < NT && n0 + col < limit) {
        const float* sg = which ? sb : sa;
        float p[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
        for (int r = 0; r < 256 / MF_EPI; r += 4) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float v = sg[(r0 + r + q) * (NT + 1) + col];
            p[q] += (which && square_b) ? v * v : v;
          }
        }
        atomicAdd(dst + which * limit + n0 + col, (double)p[0] + (double)p[1] + (double)p[2] + (double)p[3]);
      }
    };
    if constexpr (MODE == 0) {
      const long long b = m0 + row_l;
      const bool row_ok = b < a.B;
      float* orow = a.out + (row_ok ? b : 0) * a.ldo;
      const bool vec_ok = ((a.ldo & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.out) & 15) == 0);
      if (a.normalize) {
        // F.normalize(p=2, eps=1e-12) on rows that live in one tile (N <= NT)
        float ss = 0.f;
        for (int g = half; g < ngroups; g += MF_EPI) {
          float o[16];
          load16(g, o);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float v = (n0 + g * 16 + j < a.N) ? o[j] + s_bias[g * 16 + j] : 0.f;
            ss = fmaf(v, v, ss);
          }
        }
        s_part[half * 128 + row_l] = ss;
        __syncthreads();
        float tot = 0.f;
#pragma unroll
        for (int h2 = 0; h2 < MF_EPI; ++h2) tot += s_part[h2 * 128 + row_l];
        const float nrm = fmaxf(sqrtf(tot), 1e-12f);
        if (row_ok && half == 0 && a.norms) a.norms[b] = nrm;
        for (int g = half; g < ngroups; g += MF_EPI) {
          float o[16];
          load16(g, o);
          const int nb = n0 + g * 16;
#pragma unroll
          for (int j = 0; j < 16; ++j) o[j] = (o[j] + s_bias[g * 16 + j]) / nrm;
          if (row_ok) {
            if (vec_ok && nb + 16 <= a.N) {
#pragma unroll
              for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(orow + nb + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (nb + j < a.N) orow[nb + j] = o[j];
            }
          }
        }
      } else {
        const bool stats = a.out_sums != nullptr;
        for (int g = half; g < ngroups; g += MF_EPI) {
          float o[16];
          load16(g, o);
          if (g == half) MF_STAMP(18);
          const int nb = n0 + g * 16;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            o[j] += s_bias[g * 16 + j];
            if (stats) stage0[row_l * (NT + 1) + g * 16 + j] = (nb + j < a.N && row_ok) ? mf_act<RELU>(a.out_act, o[j]) : 0.f;
          }
          if (row_ok) {
            if (vec_ok && nb + 16 <= a.N) {
#pragma unroll
              for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(orow + nb + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (nb + j < a.N) orow[nb + j] = o[j];
            }
          }
        }
        MF_STAMP(19);
        if (stats) {
          tc_fence_before();
          __syncthreads();
          MF_STAMP(20);
          column_sums(stage0, stage0, true, a.out_sums, a.N);
        }
      }
    } else if constexpr (MODE == 1) {
      const long long b = m0 + row_l;
      const bool row_ok = b < a.B;
      const long long bc = row_ok ? b : a.B - 1;
      const bool sums = a.out_sums != nullptr;   // backward sums of the block below: sum g, sum g*xhat
      const bool vec_ok = ((a.ldo & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.out) & 15) == 0);
      const bool vz = sums && mf_vec_ok(a.lower.z, a.lower.ldz, a.K);
      float* orow = a.out + bc * a.ldo;
      for (int g = half; g < ngroups; g += MF_EPI) {
        const int kb = n0 + g * 16;
        float zz[16];
        if (sums) {   // issued before the TMEM read: the loads fly while the accumulators are fetched
          mf_ld8(a.lower.z + bc * a.lower.ldz, kb, a.K, vz, zz);
          mf_ld8(a.lower.z + bc * a.lower.ldz, kb + 8, a.K, vz, zz + 8);
        }
        float o[16];
        load16(g, o);
        if (g == half) MF_STAMP(18);
        if (row_ok) {
          if (vec_ok && kb + 16 <= a.K) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(orow + kb + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (kb + j < a.K) orow[kb + j] = o[j];
          }
        }
        if (sums) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const bool ok = kb + j < a.K && row_ok;
            const int k = ok ? kb + j : 0;
            const float gg = o[j] * mf_keep<DROP>(dlo, bc, k);
            const float gx = gg * ((mf_act<RELU>(a.lower.act, zz[j]) - lo.mean[k]) * lo.invstd[k]);
            stage0[row_l * (NT + 1) + g * 16 + j] = ok ? gg : 0.f;
            stage1[row_l * (NT + 1) + g * 16 + j] = ok ? gx : 0.f;
          }
        }
      }
      MF_STAMP(19);
      if (sums) {
        tc_fence_before();
        __syncthreads();
        MF_STAMP(20);
        column_sums(stage0, stage1, false, a.out_sums, a.K);
      }
    } else {
      // dW partial of this CTA: staged through shared memory so that one warp adds one row segment with coalesced
      // fp32 atomics (lanes = consecutive in-features) instead of 32 rows per instruction
      for (int g = half; g < ngroups; g += MF_EPI) {
        float o[16];
        load16(g, o);
#pragma unroll
        for (int j = 0; j < 16; ++j) stage0[row_l * (NT + 1) + g * 16 + j] = o[j];
      }
      tc_fence_before();
      __syncthreads();
      for (int r = warp; r < 128; r += MF_THREADS / 32) {
        const int n = m0 + r;
        if (n >= a.N) break;
        for (int c = lane; c < NT; c += 32) {
          const int k = n0 + c;
          if (k < a.K) atomicAdd(a.dw + (long long)n * a.lddw + k, stage0[r * (NT + 1) + c]);
        }
      }
    }
    tc_fence_before();
  }
  MF_STAMP(16);
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
  MF_STAMP(17);
}

static int mf_pad16(int v) { return (v + 15) / 16 * 16; }

static int mf_block(MfBlock& d, const b200rec_bn_block* s, const char* what) {
  if (!s->z || !s->gamma || !s->beta) return fail("%s: null pointer in a BatchNorm block", what);
  if (s->H < 1 || s->H > MF_MAXH) return fail("%s: block width %d outside [1, %d]", what, s->H, MF_MAXH);
  if (s->training && !s->sums) return fail("%s: training-mode block without batch statistics", what);
  if (!s->training && (!s->running_mean || !s->running_var)) return fail("%s: eval-mode block without running statistics", what);
  if (s->training && s->B_stat < 2) return fail("%s: Expected more than 1 value per channel when training", what);
  d.z = s->z, d.ldz = s->ldz, d.H = s->H, d.act = s->act, d.training = s->training, d.update_running = s->update_running;
  d.sums = s->sums, d.gamma = s->gamma, d.beta = s->beta, d.running_mean = s->running_mean, d.running_var = s->running_var;
  d.nbt = reinterpret_cast<long long*>(s->num_batches_tracked);
  d.eps = s->eps, d.momentum = s->momentum, d.drop_p = s->training ? s->drop_p : 0.f, d.seed = s->seed, d.B_stat = s->B_stat;
  return 0;
}

static int mf_launch(MfArgs& a, dim3 grid, cudaStream_t st) {
  const int stage_region = 2 * a.np * (128 * 128 + a.NT * 128);
  const int staging = 2 * 128 * (a.NT + 1) * 4;
  const int smem = 1024 + (stage_region > staging ? stage_region : staging) + 9 * MF_MAXH * 4 + 64 + MF_EPI * 128 * 4 + 128 * 4;
  if (smem > MF_SMEM_LIMIT) return fail("mlp_fused: %d bytes of shared memory needed", smem);
  // kernel variants: ReLU everywhere (inline, branch-free element loops) or any activation; dropout streams or none
  const bool relu = (!a.has_lower || a.lower.act == 0) && (!a.has_own || a.own.act == 0) && (a.mode != 0 || !a.out_sums || a.out_act == 0);
  const bool drop = (a.has_lower && a.lower.drop_p > 0.f) || (a.has_own && a.own.drop_p > 0.f);
  using KernelFn = void (*)(const MfArgs);
  static const KernelFn table[3][2][2] = {
      {{mlp_fused_kernel<0, false, false>, mlp_fused_kernel<0, false, true>}, {mlp_fused_kernel<0, true, false>, mlp_fused_kernel<0, true, true>}},
      {{mlp_fused_kernel<1, false, false>, mlp_fused_kernel<1, false, true>}, {mlp_fused_kernel<1, true, false>, mlp_fused_kernel<1, true, true>}},
      {{mlp_fused_kernel<2, false, false>, mlp_fused_kernel<2, false, true>}, {mlp_fused_kernel<2, true, false>, mlp_fused_kernel<2, true, true>}}};
  static bool attr_set = false;
  if (!attr_set) {
    for (int m = 0; m < 3; ++m)
      for (int r = 0; r < 2; ++r)
        for (int d = 0; d < 2; ++d)
          B200_CUDA_OK(cudaFuncSetAttribute(table[m][r][d], cudaFuncAttributeMaxDynamicSharedMemorySize, MF_SMEM_LIMIT));
    attr_set = true;
  }
  table[a.mode][relu ? 1 : 0][drop ? 1 : 0]<<<grid, MF_THREADS, smem, st>>>(a);
  B200_LAUNCH_OK("mlp_fused_kernel");
  return 0;
}

}  // namespace b200

extern "C" int b200rec_mlp_forward(const float* x, int64_t ldx, const b200rec_bn_block* lower, const float* w, int64_t ldw,
                                   const float* bias, int64_t B, int N, int K, int np, float* out, int64_t ldo,
                                   int out_act, double* out_sums, int normalize, float* norms, void* stream) {
  using namespace b200;
  if ((!x && !lower) || !w || !out) return fail("mlp_forward: null pointer");
  if (B <= 0 || N <= 0 || K <= 0 || B > INT32_MAX - 128) return fail("mlp_forward: bad sizes");
  if (np < 1 || np > 3) return fail("mlp_forward: np must be 1, 2 or 3");
  MfArgs a{};
  a.mode = 0, a.np = np, a.B = (int)B, a.N = N, a.K = K;
  a.NT = N >= 128 ? 128 : mf_pad16(N);
  if (normalize && N > a.NT) return fail("mlp_forward: the normalising epilogue needs N <= 128 (got %d)", N);
  a.chunks = (K + 63) / 64;
  a.x = x, a.ldx = ldx;
  if (lower) {
    if (mf_block(a.lower, lower, "mlp_forward")) return 1;
    if (a.lower.H != K) return fail("mlp_forward: lower block width %d != K %d", a.lower.H, K);
    a.has_lower = 1;
  }
  a.w = w, a.ldw = ldw, a.bias = bias, a.out = out, a.ldo = ldo, a.out_sums = out_sums, a.out_act = out_act;
  a.normalize = normalize, a.norms = norms;
  dim3 grid((unsigned)((B + 127) / 128), (unsigned)((N + a.NT - 1) / a.NT), 1);
  return mf_launch(a, grid, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int b200rec_mlp_dgrad(const float* dy, int64_t lddy, const b200rec_bn_block* own, const double* own_bsums,
                                 const float* w, int64_t ldw, int64_t B, int N, int K, int np, float* dx, int64_t lddx,
                                 const b200rec_bn_block* lower, double* lower_bsums, void* stream) {
  using namespace b200;
  if (!dy || !w || !dx) return fail("mlp_dgrad: null pointer");
  if (B <= 0 || N <= 0 || K <= 0 || B > INT32_MAX - 128) return fail("mlp_dgrad: bad sizes");
  if (np < 1 || np > 3) return fail("mlp_dgrad: np must be 1, 2 or 3");
  MfArgs a{};
  a.mode = 1, a.np = np, a.B = (int)B, a.N = N, a.K = K;
  a.NT = K >= 128 ? 128 : mf_pad16(K);
  a.chunks = (N + 63) / 64;
  if (own) {
    if (mf_block(a.own, own, "mlp_dgrad")) return 1;
    if (a.own.H != N) return fail("mlp_dgrad: own block width %d != N %d", a.own.H, N);
    if (a.own.training && !own_bsums) return fail("mlp_dgrad: training-mode block without backward sums");
    a.has_own = 1, a.own_bsums = own_bsums;
  }
  if (lower_bsums) {
    if (!lower) return fail("mlp_dgrad: backward sums requested without the lower block");
    if (mf_block(a.lower, lower, "mlp_dgrad")) return 1;
    if (a.lower.H != K) return fail("mlp_dgrad: lower block width %d != K %d", a.lower.H, K);
    a.has_lower = 1;
    a.out_sums = lower_bsums;
  }
  a.dy = dy, a.lddy = lddy, a.w = w, a.ldw = ldw, a.out = dx, a.ldo = lddx;
  dim3 grid((unsigned)((B + 127) / 128), (unsigned)((K + a.NT - 1) / a.NT), 1);
  return mf_launch(a, grid, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int b200rec_mlp_wgrad(const float* dy, int64_t lddy, const b200rec_bn_block* own, const double* own_bsums,
                                 const double* own_bsums_local, const float* x, int64_t ldx,
                                 const b200rec_bn_block* lower, int64_t B, int N, int K, int np, float* dw, int64_t lddw,
                                 float* db, float* dgamma, float* dbeta, void* stream) {
  using namespace b200;
  if (!dy || (!x && !lower) || !dw) return fail("mlp_wgrad: null pointer");
  if (B <= 0 || N <= 0 || K <= 0 || B > INT32_MAX - 128) return fail("mlp_wgrad: bad sizes");
  if (np < 1 || np > 3) return fail("mlp_wgrad: np must be 1, 2 or 3");
  MfArgs a{};
  a.mode = 2, a.np = np, a.B = (int)B, a.N = N, a.K = K;
  const int kcols = K;
  a.NT = kcols >= 128 ? 128 : mf_pad16(kcols);
  if (own) {
    if (mf_block(a.own, own, "mlp_wgrad")) return 1;
    if (a.own.H != N) return fail("mlp_wgrad: own block width %d != N %d", a.own.H, N);
    if (a.own.training && !own_bsums) return fail("mlp_wgrad: training-mode block without backward sums");
    a.has_own = 1, a.own_bsums = own_bsums, a.own_bsums_local = own_bsums_local ? own_bsums_local : own_bsums;
    a.dgamma = dgamma, a.dbeta = dbeta;
    if (dgamma && !a.own_bsums_local) return fail("mlp_wgrad: dgamma / dbeta requested without the backward sums");
    if ((dgamma == nullptr) != (dbeta == nullptr)) return fail("mlp_wgrad: dgamma and dbeta come together");
  }
  if (lower) {
    if (mf_block(a.lower, lower, "mlp_wgrad")) return 1;
    if (a.lower.H != K) return fail("mlp_wgrad: lower block width %d != K %d", a.lower.H, K);
    a.has_lower = 1;
  }
  a.dy = dy, a.lddy = lddy, a.x = x, a.ldx = ldx, a.dw = dw, a.lddw = lddw, a.db = db;
  // the batch is the contraction: split it over enough CTAs to fill the chip, each accumulating its 64-row chunks in
  // TMEM and adding one [N, K] partial with fp32 atomics
  const int tiles = ((N + 127) / 128) * ((kcols + a.NT - 1) / a.NT);
  const int total_chunks = (int)((B + 63) / 64);
  int splits = num_sms() / tiles;
  if (splits < 1) splits = 1;
  if (splits > total_chunks) splits = total_chunks;
  a.chunks = (total_chunks + splits - 1) / splits;
  splits = (total_chunks + a.chunks - 1) / a.chunks;
  dim3 grid((unsigned)splits, (unsigned)((N + 127) / 128), (unsigned)((kcols + a.NT - 1) / a.NT));
  return mf_launch(a, grid, reinterpret_cast<cudaStream_t>(stream));
}

#ifdef MF_TIMING
extern "C" int b200rec_debug_mf_stamps(unsigned long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, b200::g_mf_stamps, sizeof(unsigned long long) * 32) == cudaSuccess ? 0 : 1;
}
#endif

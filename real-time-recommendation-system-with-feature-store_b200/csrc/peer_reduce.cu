// One-shot all-reduce of a SMALL fp64 vector over NVLink peer memory: the synced-BatchNorm statistics of exact data
// parallel training (2 * H doubles per layer and direction, 8 exchanges per step) and the loss share.  An NCCL
// all-reduce of 1-4 KB costs 25-60 us per call inside the captured training step; this is one 1-block kernel:
//   1. every rank stores its vector into its own slot of EVERY rank's symmetric buffer (plain st.global to peer
//      addresses mapped by torch.distributed._symmetric_memory), fences, and raises its flag in every buffer to the
//      call's sequence number (st.release.sys);
//   2. it waits (ld.acquire.sys, bounded) until all flags of its own buffer carry that number, then sums the slots in
//      RANK ORDER — every replica computes bit-identical sums, which is what keeps replicas in lockstep.
// Two slot sets alternate by the parity of the sequence number: a rank can only be one call ahead of the slowest one
// (call s+1 waits for the flags s+1, which a peer raises after it finished reading set s), so set s is never
// overwritten while somebody reads it.  The sequence number lives in the buffer itself and is advanced by the kernel,
// so a captured CUDA graph replays correctly.  All ranks must issue the same sequence of calls on one stream.
// Each rank's kernel runs on its OWN GPU (one process per GPU): kernels never wait for another kernel of the same device.
#include "host_util.h"
#include "../../include/b200rec.h"

namespace b200 {

constexpr int PR_MAX_WORLD = 8;
constexpr int PR_HEADER_BYTES = 128 + 2 * PR_MAX_WORLD * 8;   // sequence number (padded) | flags[2][8]

struct PeerPtrs {
  unsigned char* p[PR_MAX_WORLD];
};

__device__ __forceinline__ void st_release_sys(unsigned long long* addr, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* addr) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(addr) : "memory");
  return v;
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* addr) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(addr) : "memory");
  return v;
}

__global__ void __launch_bounds__(256)
peer_allreduce_f64_kernel(double* __restrict__ data, int n, int rank, int world, PeerPtrs peers, int max_n,
                          int* __restrict__ status) {
  unsigned char* mine = peers.p[rank];
  unsigned long long* seqp = reinterpret_cast<unsigned long long*>(mine);
  const unsigned long long seq = *reinterpret_cast<volatile unsigned long long*>(seqp) + 1ull;
  const int par = (int)(seq & 1ull);
  auto flags = [&](unsigned char* base) { return reinterpret_cast<unsigned long long*>(base + 128) + par * PR_MAX_WORLD; };
  auto slot = [&](unsigned char* base, int r) {
    return reinterpret_cast<double*>(base + PR_HEADER_BYTES) + ((size_t)par * PR_MAX_WORLD + r) * max_n;
  };
  // 1. my vector -> my slot in every rank's buffer
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = data[i];
    for (int r = 0; r < world; ++r) slot(peers.p[r], rank)[i] = v;
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < world) st_release_sys(flags(peers.p[threadIdx.x]) + rank, seq);
  // 2. wait for everybody's vector, sum in rank order
  if ((int)threadIdx.x < world) {
    const unsigned long long* f = flags(mine) + threadIdx.x;
    unsigned int spins = 0;
    while (ld_acquire_sys(f) < seq) {
      if (++spins > (1u << 27)) {   // ~ seconds: a replica died or the call sequences diverged; do not hang the GPU
        if (status) atomicExch(status, 1);
        break;
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < world; ++r) s += ld_relaxed_sys_f64(slot(mine, r) + i);
    data[i] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) *reinterpret_cast<volatile unsigned long long*>(seqp) = seq;
}

}  // namespace b200

using namespace b200;

extern "C" size_t b200rec_peer_allreduce_bytes(int max_n) {
  if (max_n <= 0) return 0;
  return (size_t)PR_HEADER_BYTES + (size_t)2 * PR_MAX_WORLD * (size_t)max_n * sizeof(double);
}

extern "C" int b200rec_peer_allreduce_f64(double* data, int n, int rank, int world, const uint64_t* peer_buffers_host,
                                          int max_n, int* status_dev, void* stream) {
  if (!data || !peer_buffers_host) return fail("peer_allreduce_f64: null pointer");
  if (world < 1 || world > PR_MAX_WORLD || rank < 0 || rank >= world) return fail("peer_allreduce_f64: world must be 1..8");
  if (n <= 0 || n > max_n) return fail("peer_allreduce_f64: n = %d outside (0, %d]", n, max_n);
  PeerPtrs pp;
  for (int r = 0; r < PR_MAX_WORLD; ++r)
    pp.p[r] = r < world ? reinterpret_cast<unsigned char*>(static_cast<uintptr_t>(peer_buffers_host[r])) : nullptr;
  for (int r = 0; r < world; ++r)
    if (pp.p[r] == nullptr || (reinterpret_cast<uintptr_t>(pp.p[r]) & 127)) return fail("peer_allreduce_f64: bad peer buffer");
  peer_allreduce_f64_kernel<<<1, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(data, n, rank, world, pp, max_n, status_dev);
  B200_LAUNCH_OK("peer_allreduce_f64_kernel");
  return 0;
}

#include "host_util.h"
#include "../../include/b200rec.h"

namespace b200 {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    // resolved at run time from the driver: no link-time dependency on libcuda (the build box has no GPU)
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail("cuTensorMapEncodeTiled not available from the driver");
  // cuTensorMapEncodeTiled is a DRIVER call: on a thread that has made no runtime call yet (the autograd engine's
  // worker thread running a backward whose first library call builds tensor maps) no context is current and it fails
  // with CUDA_ERROR_INVALID_CONTEXT.  cudaFree(nullptr) binds the device's primary context to the calling thread.
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    cudaFree(nullptr);
    ctx_bound = true;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return fail("TMA base pointer must be 16-byte aligned");
  if ((ld * 2) % 16 != 0) return fail("TMA row pitch must be a multiple of 16 bytes (ld=%llu)", (unsigned long long)ld);
  if (box_rows == 0 || box_rows > 256) return fail("TMA box rows out of range: %u", box_rows);
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

}  // namespace b200

extern "C" {
const char* b200rec_last_error(void) { return b200::g_err; }
int b200rec_version(void) { return 1; }
int64_t b200rec_launch_count(void) { return b200::g_launches.load(); }
}

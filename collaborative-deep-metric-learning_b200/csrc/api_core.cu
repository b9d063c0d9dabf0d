// Context, error reporting and TMA tensor-map encoding for libcdml.
#include <stdarg.h>
#include <string.h>

#include "../../include/cdml.h"
#include "ctx.cuh"

namespace cdml {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int make_tmap_2d(cdml_ctx* ctx, CUtensorMap* map, const void* ptr, int dtype16, uint64_t inner, uint64_t outer,
                 uint64_t ld, uint32_t box_inner, uint32_t box_outer, int swizzle_bytes) {
  CDML_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA operand base %p is not 16-byte aligned", ptr);
  CDML_REQUIRE((ld * 2) % 16 == 0, "TMA operand pitch %llu elements is not a multiple of 16 bytes",
               (unsigned long long)ld);
  CDML_REQUIRE(box_inner * 2 <= 128 && box_outer <= 256, "bad TMA box %u x %u", box_inner, box_outer);
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = ctx->encode_tiled(map, dtype16 == CDML_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                                 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CDML_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (inner=%llu outer=%llu ld=%llu)", (int)r,
               (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld);
  return 0;
}
void* ctx_scratch(cdml_ctx* ctx, size_t bytes) {
  if (ctx->scratch_bytes >= bytes) return ctx->scratch;
  if (ctx->scratch != nullptr) {
    cudaDeviceSynchronize();   // growth is rare; nobody may still be using the old block
    cudaFree(ctx->scratch);
    ctx->scratch = nullptr, ctx->scratch_bytes = 0;
  }
  const size_t want = bytes + bytes / 2 + (1 << 20);
  if (cudaMalloc(&ctx->scratch, want) != cudaSuccess) {
    ctx->scratch = nullptr;
    set_error("ctx_scratch: cudaMalloc of %zu bytes failed", want);
    return nullptr;
  }
  ctx->scratch_bytes = want;
  return ctx->scratch;
}
}  // namespace cdml

extern "C" {

const char* cdml_last_error(void) { return cdml::g_err; }

int cdml_version(void) { return 100; }

int cdml_ctx_create(int device, cdml_ctx** out) {
  CDML_REQUIRE(out != nullptr, "cdml_ctx_create: out is NULL");
  *out = nullptr;
  int count = 0;
  CDML_CHECK_CUDA(cudaGetDeviceCount(&count));
  CDML_REQUIRE(device >= 0 && device < count, "cdml_ctx_create: device %d out of range (%d visible)", device, count);
  CDML_CHECK_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  CDML_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  CDML_REQUIRE(prop.major == 10, "libcdml is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
  cdml_ctx* c = new cdml_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  c->encode_tiled = nullptr;
  c->scratch = nullptr, c->scratch_bytes = 0;
  c->mine_stats = nullptr;
  {  // keep stream-ordered allocations cached across synchronisation points
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
    delete c;
    cdml::set_error("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: %s", cudaGetErrorString(e));
    return -2;
  }
  c->encode_tiled = reinterpret_cast<cdml_encode_tiled_fn>(fn);
  if (cudaMalloc(&c->dev_flags, sizeof(int32_t)) != cudaSuccess || cudaMemset(c->dev_flags, 0, sizeof(int32_t)) != cudaSuccess) {
    delete c;
    cdml::set_error("cdml_ctx_create: cudaMalloc of the error word failed");
    return -2;
  }
  *out = c;
  return 0;
}

int cdml_ctx_destroy(cdml_ctx* ctx) {
  if (ctx == nullptr) return 0;
  cudaFree(ctx->dev_flags);
  cudaFree(ctx->scratch);
  cudaFree(ctx->mine_stats);
  delete ctx;
  return 0;
}

int cdml_ctx_poll_errors(cdml_ctx* ctx, void* stream, int32_t* flags) {
  CDML_REQUIRE(ctx != nullptr && flags != nullptr, "cdml_ctx_poll_errors: NULL argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CDML_CHECK_CUDA(cudaMemcpyAsync(flags, ctx->dev_flags, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CDML_CHECK_CUDA(cudaMemsetAsync(ctx->dev_flags, 0, sizeof(int32_t), st));
  CDML_CHECK_CUDA(cudaStreamSynchronize(st));
  return 0;
}

}  // extern "C"

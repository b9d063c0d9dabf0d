// Host-side context shared by the translation units of libcdml.
#pragma once
#include "common.cuh"

typedef CUresult (*cdml_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                         const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                         CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct cdml_ctx {
  int device;
  int num_sms;
  int32_t* dev_flags;  // device error word
  cdml_encode_tiled_fn encode_tiled;
  void* scratch;         // grow-only device scratch for kernels that need a few MB of workspace (mining)
  size_t scratch_bytes;
  unsigned long long* mine_stats;  // device [2]: re-scans and mined anchors of the last cdml_mine_semihard (diagnostics)
};

namespace cdml {
// Returns a device buffer of at least `bytes` owned by the context (re-allocated only when it must grow).
void* ctx_scratch(cdml_ctx* ctx, size_t bytes);
}

namespace cdml {
// 2-D tensor map over a row-major 16-bit matrix: `inner` contiguous elements, `outer` rows of pitch ld elements.
// 128-byte swizzle (operand tiles) or 64-byte swizzle (the 32 x 64-byte store boxes of the epilogues); zero fill / clipping
// out of bounds.
int make_tmap_2d(cdml_ctx* ctx, CUtensorMap* map, const void* ptr, int dtype16, uint64_t inner, uint64_t outer,
                 uint64_t ld, uint32_t box_inner, uint32_t box_outer, int swizzle_bytes = 128);
}  // namespace cdml

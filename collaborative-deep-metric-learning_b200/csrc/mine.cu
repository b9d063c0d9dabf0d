// In-batch semi-hard negative mining (build-defined, the reference draws negatives at random).
//
// For triplet i with dp = |a_i - p_i|^2, over the candidate rows r = 3j+1, 3j+2 (positives and negatives of the
// batch) whose guid is neither a_i nor p_i:
//   (1) argmin d(a_i, r) with dp < d < dp + margin, else (2) argmin d with d >= dp + margin, else (3) keep row 3i+2.
// The B x 2B distance matrix is a tcgen05 GEMM (anchors x candidates, fp16 operands) whose epilogue keeps two running
// (distance,row) minima per anchor in registers and folds them into global 64-bit keys with atomicMin, so the matrix
// never reaches HBM.  d = 2 - 2 a.c (unit-norm embeddings).  Ties -> lowest row.
#include <algorithm>

#include "gemm_launch.cuh"

namespace cdml {

constexpr unsigned long long kNoCand = ~0ull;

template <int BN>
struct EpiMine {
  __device__ __forceinline__ void block_begin() const {}
  __device__ __forceinline__ void block_end() const {}
  const float* dp;        // [B] exact |a-p|^2
  const int32_t* guid;    // [B,3] int32 guids
  unsigned long long* semi;    // [B] best semi-hard (float bits << 32 | row)
  unsigned long long* beyond;  // [B]
  float margin;
  int cand;  // 1: columns are the positives' rows 3j+1, 2: the negatives' rows 3j+2
  __device__ __forceinline__ void run(uint32_t taddr, int row, int n0, int /*split*/, const GemmShape& s) const {
    const int lane = threadIdx.x & 31;
    const bool row_ok = row < s.M;
    const float dpi = row_ok ? __ldg(dp + row) : 0.f;
    const float lim = dpi + margin;
    const int ga = row_ok ? __ldg(guid + 3 * row) : 0;
    const int gp = row_ok ? __ldg(guid + 3 * row + 1) : 0;
    float sd = __int_as_float(0x7f800000), bd = __int_as_float(0x7f800000);
    int sr = -1, br = -1;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      const int nb = n0 + c * 32;
      if (nb >= s.N) break;
      uint32_t v[32];
      tmem_ld_32x32(taddr + c * 32, v);
      // one coalesced load of the chunk's 32 candidate guids, broadcast per column by shuffle
      const int g_lane = nb + lane < s.N ? __ldg(guid + 3 * (nb + lane) + cand) : ga;  // beyond N: never valid
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int col = nb + j;
        const int gj = __shfl_sync(0xffffffffu, g_lane, j);
        const float d = fmaxf(fmaf(-2.f, __uint_as_float(v[j]), 2.f), 0.f);
        const bool valid = gj != ga && gj != gp;
        const bool is_semi = valid && d > dpi && d < lim && d < sd;
        const bool is_beyond = valid && d >= lim && d < bd;
        sd = is_semi ? d : sd, sr = is_semi ? 3 * col + cand : sr;
        bd = is_beyond ? d : bd, br = is_beyond ? 3 * col + cand : br;
      }
    }
    if (row_ok) {
      if (sr >= 0)
        atomicMin(semi + row, (static_cast<unsigned long long>(__float_as_uint(sd)) << 32) | static_cast<uint32_t>(sr));
      if (br >= 0)
        atomicMin(beyond + row, (static_cast<unsigned long long>(__float_as_uint(bd)) << 32) | static_cast<uint32_t>(br));
    }
  }
};

__global__ void mine_prepare_kernel(const float* __restrict__ E, int64_t ld, int D, const int64_t* __restrict__ guid64,
                                    int64_t B, float* __restrict__ dp, int32_t* __restrict__ guid32,
                                    unsigned long long* __restrict__ semi, unsigned long long* __restrict__ beyond) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * 8 + warp; i < B; i += static_cast<int64_t>(gridDim.x) * 8) {
    const float* a = E + 3 * i * ld;
    const float* p = a + ld;
    float s = 0.f;
    for (int j = lane; j < D; j += 32) {
      const float t = a[j] - p[j];
      s = fmaf(t, t, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) dp[i] = s, semi[i] = kNoCand, beyond[i] = kNoCand;
    if (lane < 3) guid32[3 * i + lane] = static_cast<int32_t>(guid64[3 * i + lane]);
  }
}

__global__ void mine_finalize_kernel(const float* __restrict__ E, int64_t ld, int D,
                                     const unsigned long long* __restrict__ semi,
                                     const unsigned long long* __restrict__ beyond, int64_t B,
                                     int32_t* __restrict__ neg_row, float* __restrict__ d_an) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * 8 + warp; i < B; i += static_cast<int64_t>(gridDim.x) * 8) {
    const unsigned long long ks = semi[i], kb = beyond[i];
    const int r = ks != kNoCand ? static_cast<int>(static_cast<uint32_t>(ks))
                                : (kb != kNoCand ? static_cast<int>(static_cast<uint32_t>(kb)) : static_cast<int>(3 * i + 2));
    if (lane == 0) neg_row[i] = r;
    if (d_an != nullptr) {  // exact fp32 distance of the chosen negative
      const float* a = E + 3 * i * ld;
      const float* n = E + static_cast<int64_t>(r) * ld;
      float s = 0.f;
      for (int j = lane; j < D; j += 32) {
        const float t = a[j] - n[j];
        s = fmaf(t, t, s);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) d_an[i] = s;
    }
  }
}

}  // namespace cdml

using namespace cdml;

extern "C" int cdml_mine_semihard(cdml_ctx* ctx, const void* E16, int64_t ld16, int dtype16, const float* E32,
                                  int64_t ld32, const int64_t* guid, int64_t B, int D, float margin, int32_t* neg_row,
                                  float* d_an, void* stream) {
  CDML_REQUIRE(ctx && E16 && E32 && guid && neg_row, "cdml_mine_semihard: NULL argument");
  CDML_REQUIRE(B > 0 && 3 * B < (1ll << 31) && D > 0 && D % 8 == 0 && ld16 >= D && ld32 >= D,
               "cdml_mine_semihard: bad geometry B=%lld D=%d", (long long)B, D);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // scratch: dp [B] f32 | guid32 [3B] i32 | semi [B] u64 | beyond [B] u64
  uint8_t* ws = nullptr;
  const size_t bytes = static_cast<size_t>(B) * (4 + 12 + 8 + 8) + 64;
  CDML_CHECK_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&ws), bytes, st));
  unsigned long long* semi = reinterpret_cast<unsigned long long*>(ws);
  unsigned long long* beyond = semi + B;
  float* dp = reinterpret_cast<float*>(beyond + B);
  int32_t* guid32 = reinterpret_cast<int32_t*>(dp + B);
  const int grid = static_cast<int>(std::min<int64_t>((B + 7) / 8, ctx->num_sms * 8));
  mine_prepare_kernel<<<grid, 256, 0, st>>>(E32, ld32, D, guid, B, dp, guid32, semi, beyond);
  int rc = 0;
  const uint16_t* e16 = static_cast<const uint16_t*>(E16);
  for (int cand = 1; cand <= 2 && rc >= 0; ++cand) {
    EpiMine<kBN> epi{dp, guid32, semi, beyond, margin, cand};
    // A: anchors = rows 0,3,6,.. (pitch 3*ld16); B: candidates = rows cand, cand+3, .. ; both K-major, K = D
    if (resb_applicable(D) && B >= 8 * kBM)
      rc = launch_gemm_resb(ctx, e16, 3 * ld16, e16 + cand * ld16, 3 * ld16, B, B, D, dtype16, epi, st);
    else
      rc = launch_gemm<0, 0>(ctx, e16, 3 * ld16, e16 + cand * ld16, 3 * ld16, B, B, D, dtype16, 1, epi, st);
  }
  if (rc >= 0) {
    mine_finalize_kernel<<<grid, 256, 0, st>>>(E32, ld32, D, semi, beyond, B, neg_row, d_an);
    rc = cudaGetLastError() == cudaSuccess ? 0 : -2;
  }
  cudaFreeAsync(ws, st);
  return rc < 0 ? rc : 0;
}

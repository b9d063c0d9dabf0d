// In-batch semi-hard negative mining (build-defined, the reference draws negatives at random).
//
// For triplet i with dp = |a_i - p_i|^2, over the candidate rows r = 3j+1, 3j+2 (positives and negatives of the
// batch) whose guid is neither a_i nor p_i:
//   (1) argmin d(a_i, r) with dp < d < dp + margin, else (2) argmin d with d >= dp + margin, else (3) keep row 3i+2.
// The B x 2B distance matrix is a tcgen05 GEMM (anchors x candidates, fp16 operands) whose epilogue keeps a running
// (distance,row) minimum per anchor and folds it into a global 64-bit key with atomicMin, so the matrix
// never reaches HBM.  d = 2 - 2 a.c (unit-norm embeddings).  Ties -> lowest row.
#include <algorithm>

#include "gemm_launch.cuh"

namespace cdml {

constexpr unsigned long long kNoCand = ~0ull;

// (1)/(2) collapse into ONE criterion: the chosen row is argmin d over valid candidates with d > dp -- if that minimum
// is below dp+margin it is the semi-hard pick, otherwise it is exactly the closest beyond-margin pick.
// Common path per score: d = 2-2s, keep min over (d > dp).  Only when a 32-column chunk can beat the anchor's current
// best (a bound read from the global key at tile start; rare after the first tiles) are guids checked and the row
// recorded.  Result: 64-bit key (float bits of d << 32 | row) folded with atomicMin -> ties go to the lowest row.
template <int BN>
struct EpiMine {
  static constexpr bool kSplitColumns = true;
  struct State {
    float dpi, bound_d;
    int ga, gp;
    int g[4];   // candidate guid of column (chunk c0+i, lane) for the warp's four chunks
  };
  // the anchor's constants do not depend on the accumulator: fetched while the tile's MMAs are still running
  __device__ __forceinline__ void pre(State& st, int row, int n0, const GemmShape& s, int c0, int c1, uint32_t /*stg*/) const {
    static_assert(BN == 256, "a warp owns 4 chunks of the tile");
    const int lane = threadIdx.x & 31;
    const bool row_ok = row < s.M;
    const float inf = __int_as_float(0x7f800000);
    st.dpi = row_ok ? __ldg(dp + row) : inf;
    st.ga = row_ok ? __ldg(guid + 3 * row) : 0;
    st.gp = row_ok ? __ldg(guid + 3 * row + 1) : 0;
    st.bound_d = row_ok ? __uint_as_float(static_cast<uint32_t>(best[row] >> 32)) : inf;   // empty key reads as NaN
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int col = n0 + (c0 + i) * 32 + lane;
      st.g[i] = (c0 + i < c1 && col < s.N) ? __ldg(guid + 3 * col + cand) : -1;
    }
  }
  __device__ __forceinline__ void block_begin(uint32_t) const {}
  __device__ __forceinline__ void block_end(uint32_t) const {}
  const float* dp;        // [B] exact |a-p|^2
  const int32_t* guid;    // [B,3] int32 guids
  unsigned long long* best;  // [B] (float bits of d) << 32 | row
  int cand;  // 1: columns are the positives' rows 3j+1, 2: the negatives' rows 3j+2
  // Works in score space: d = 2 - 2s, so "d > dp" is "s < s_hi" and the closest candidate is the LARGEST such score.
  // The chunk's 32 candidate guids are parked in the warp's
  // staging buffer (one coalesced load per chunk) so the rare re-scan needs no global loads.
  __device__ __forceinline__ void run(uint32_t taddr, int row, int n0, int /*split*/, const GemmShape& s, int c0,
                                      int c1, uint32_t stg, State& st) const {
    const int lane = threadIdx.x & 31;
    const bool row_ok = row < s.M;
    const float inf = __int_as_float(0x7f800000);
    const float dpi = st.dpi;
    const float s_hi = row_ok ? 1.f - 0.5f * dpi : -inf;      // scores must stay BELOW this (rows beyond M: nothing does)
    const int ga = st.ga, gp = st.gp;
    float bound_d = st.bound_d;                                // best distance known for this anchor (any tile)
    if (!(bound_d == bound_d)) bound_d = inf;                  // empty key reads as NaN
    float s_lo = 1.f - 0.5f * bound_d;                         // a chunk matters only if it holds a score >= this
    float bd = inf;
    int br = -1;
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      const int c = c0 + cc;
      const int nb = n0 + c * 32;
      if (c >= c1 || nb >= s.N) break;
      uint32_t v[32];
      tmem_ld_32x32(taddr + c * 32, v);
      const int g_lane = nb + lane < s.N ? st.g[cc] : ga;   // beyond N: never valid
      __syncwarp();
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(stg + lane * 4), "r"(g_lane) : "memory");
      tmem_ld_wait();
      // Hot loop, one instruction per score: t = s_hi - score in packed fp32x2 subtracts, then a 3-input UNSIGNED
      // minimum over the raw bits.  Positive floats order like their bit patterns and every negative float is
      // >= 0x80000000, so the minimum is the smallest t > 0 (the closest candidate with d > dp) whenever one exists.
      uint32_t u0 = 0xffffffffu, u1 = 0xffffffffu;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 lo = sub2(s_hi, s_hi, __uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]));
        const float2 hi = sub2(s_hi, s_hi, __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
        u0 = umin3(u0, __float_as_uint(lo.x), __float_as_uint(lo.y));
        u1 = umin3(u1, __float_as_uint(hi.x), __float_as_uint(hi.y));
      }
      __syncwarp();                                            // guids visible to every lane of the warp
      // the chunk matters only if it holds a score in [s_lo, s_hi), i.e. 0 <= t <= s_hi - s_lo (monotone in fp32;
      // t == +0 is a harmless false alarm: the re-scan below applies the exact conditions)
      if (min(u0, u1) <= __float_as_uint(s_hi - s_lo)) {       // rare: this chunk may improve the anchor's best
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float sc = __uint_as_float(v[j]);
          const float d = fmaxf(fmaf(-2.f, sc, 2.f), 0.f);
          if (sc < s_hi && sc >= s_lo && d > dpi && d < bd) {
            int gj;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(gj) : "r"(stg + j * 4) : "memory");
            if (gj != ga && gj != gp) bd = d, br = 3 * (nb + j) + cand;
          }
        }
        if (br >= 0) s_lo = fmaxf(s_lo, 1.f - 0.5f * bd);
      }
    }
    if (br >= 0)
      atomicMin(best + row, (static_cast<unsigned long long>(__float_as_uint(bd)) << 32) | static_cast<uint32_t>(br));
  }
};

__global__ void mine_prepare_kernel(const float* __restrict__ E, int64_t ld, int D, const int64_t* __restrict__ guid64,
                                    int64_t B, float* __restrict__ dp, int32_t* __restrict__ guid32,
                                    unsigned long long* __restrict__ best) {
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
    if (lane == 0) dp[i] = s, best[i] = kNoCand;
    if (lane < 3) guid32[3 * i + lane] = static_cast<int32_t>(guid64[3 * i + lane]);
  }
}

__global__ void mine_finalize_kernel(const float* __restrict__ E, int64_t ld, int D,
                                     const unsigned long long* __restrict__ best, int64_t B,
                                     int32_t* __restrict__ neg_row, float* __restrict__ d_an) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * 8 + warp; i < B; i += static_cast<int64_t>(gridDim.x) * 8) {
    const unsigned long long kb = best[i];
    const int r = kb != kNoCand ? static_cast<int>(static_cast<uint32_t>(kb)) : static_cast<int>(3 * i + 2);
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
  // scratch: best [B] u64 | dp [B] f32 | guid32 [3B] i32
  const size_t bytes = static_cast<size_t>(B) * (8 + 4 + 12) + 64;
  uint8_t* ws = static_cast<uint8_t*>(ctx_scratch(ctx, bytes));
  if (ws == nullptr) return -2;
  unsigned long long* best = reinterpret_cast<unsigned long long*>(ws);
  float* dp = reinterpret_cast<float*>(best + B);
  int32_t* guid32 = reinterpret_cast<int32_t*>(dp + B);
  const int grid = static_cast<int>(std::min<int64_t>((B + 7) / 8, ctx->num_sms * 8));
  mine_prepare_kernel<<<grid, 256, 0, st>>>(E32, ld32, D, guid, B, dp, guid32, best);
  int rc = 0;
  const uint16_t* e16 = static_cast<const uint16_t*>(E16);
  for (int cand = 1; cand <= 2 && rc >= 0; ++cand) {
    EpiMine<kBN> epi{dp, guid32, best, cand};
    // A: anchors = rows 0,3,6,.. (pitch 3*ld16); B: candidates = rows cand, cand+3, .. ; both K-major, K = D
    if (resb_applicable(D) && B >= 8 * kBM)
      rc = launch_gemm_resb(ctx, e16, 3 * ld16, e16 + cand * ld16, 3 * ld16, B, B, D, dtype16, epi, st);
    else
      rc = launch_gemm<0, 0>(ctx, e16, 3 * ld16, e16 + cand * ld16, 3 * ld16, B, B, D, dtype16, 1, epi, st);
  }
  if (rc >= 0) {
    mine_finalize_kernel<<<grid, 256, 0, st>>>(E32, ld32, D, best, B, neg_row, d_an);
    rc = cudaGetLastError() == cudaSuccess ? 0 : -2;
  }
  return rc < 0 ? rc : 0;
}

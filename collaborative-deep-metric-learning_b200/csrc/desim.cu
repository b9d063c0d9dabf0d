// De-similarity post-filter of the KNN lists (faiss_knn.py:134-244: desim, fliter_fI, iter_desim_mp) -- integer / set work,
// HBM-bound.  The reference sweeps the columns of eI with 81 x Pool(22) numpy passes over a +1-shifted copy; rows are
// independent, so here ONE WARP owns a row of eI in registers and walks its columns:
//
//   prepare:  F[v, j] = fI[v, j] if fD[v, j] <= threshold and fI[v, j] != v and 0 <= fI[v, j] < 2^31 else -1   (j < f_end)
//             int32, row pitch fw_pad -- the table every pivot gathers from (fliter_fI, faiss_knn.py:146-155)
//   rows:     for c = 0 .. ke-1: v = eI[r, c]; if alive: every later alive entry that occurs in F[v, :] is dropped (-1);
//             finally the row's own id is dropped (faiss_knn.py:236-238).
//
// The only HBM traffic that matters is the gather of F rows (ke x fw_pad x 4 bytes per eI row, random 128-byte rows).
#include "../../include/cdml.h"
#include "ctx.cuh"

namespace cdml {

constexpr int kDesimWarps = 8;

__global__ void __launch_bounds__(256)
desim_prepare_kernel(const int64_t* __restrict__ fI, const float* __restrict__ fD, int64_t nf, int kf, int64_t ld_fi,
                     int64_t ld_fd, float threshold, int fw, int fw_pad, int32_t* __restrict__ F) {
  const int64_t total = nf * fw_pad;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t v = i / fw_pad;
    const int j = static_cast<int>(i - v * fw_pad);
    int32_t out = -1;
    if (j < fw) {
      const int64_t id = fI[v * ld_fi + j];
      const bool far = fD != nullptr && fD[v * ld_fd + j] > threshold;      // NaN distances are kept, like `fD > thr`
      if (!far && id != v && id >= 0 && id < (1ll << 31)) out = static_cast<int32_t>(id);
    }
    F[i] = out;
  }
}

// E = entries of the row per lane (ke <= 32 E), FC = 32-wide chunks of an F row (fw <= 32 FC).
//
// Two phases per row, the warp's state in shared memory:
//  (1) relation, fully parallel: a Bloom bitmap (8192 bits) and an open-addressing hash of the row's ids (id -> chain of
//      the columns holding it; duplicates chain up) are built once; then for EVERY column c -- alive or not, the answer
//      does not depend on it -- lane j tests the j-th feature neighbour of eI[r,c] against the bitmap (one LDS; ~0.1 % false
//      positives), and only the lanes that pass probe the hash and OR the later columns holding the id into kill[c]
//      (a ke-bit mask).  All ke gathers of F rows are independent, kUnroll of them in flight per warp.
//  (2) resolution, sequential but sparse: alive = valid columns; for each column c WITH a non-empty kill mask, ascending:
//      if alive[c]: alive &= ~kill[c].
// History on 4M x 81 rows (ncu, profiles/): column walk with fw x ke comparisons 184 ms -> hash probe per neighbour, walk
// in column order with 4 prefetched pivots 62.6 ms (14 k instructions per row, issue-bound) -> this form 35.1 ms -> one-wave
// grid, two-bit blocked Bloom words, 256-slot hash 32.0 ms (4.3 k instructions per row; 8 -> 4 gathers in flight per warp
// costs 35 %: the kernel lives on memory-level parallelism x resident warps, 70 registers and 5 KB of row state per warp).
// Tried and measured slower (39.9 ms, commit 90c5b28): a persistent CTA per SM whose warps stage all ke gathers of a row
// with cp.async.bulk behind an mbarrier (12 KB of staging per warp leaves 11 warps per SM; the kernel is bound by the
// dependent shared-memory chains of the probes, not by the number of gathers in flight).
template <int E>
struct DesimCfg {
  static constexpr int kSlots = E <= 1 ? 64 : E == 2 ? 128 : E <= 4 ? 256 : 512;     // load factor <= 0.5
  static constexpr int kWarps = E <= 2 ? 8 : E <= 4 ? 4 : 2;                        // static shared memory <= 30 KB
};
constexpr int kUnroll = 12;   // gathers in flight per warp: 4 -> 44.0 ms, 8 -> 32.0, 12 -> 30.8, 16 -> 33.3 (4M x 81 rows)
constexpr int kBloomWords = 256;

__device__ __forceinline__ uint32_t desim_mix(int32_t id) { return static_cast<uint32_t>(id) * 2654435761u; }
// Blocked Bloom filter: one 32-bit word per id (top 8 hash bits), two bits inside it -> one LDS per test, ~0.1 % false
// positives at 81 ids (a single bit per id gave ~1 %, i.e. a slow-path excursion on every fourth column).
__device__ __forceinline__ uint32_t bloom_word(uint32_t m) { return m >> 24; }
__device__ __forceinline__ uint32_t bloom_bits(uint32_t m) { return (1u << ((m >> 19) & 31)) | (1u << ((m >> 14) & 31)); }

template <int E, int FC>
__global__ void __launch_bounds__(DesimCfg<E>::kWarps * 32)
desim_rows_kernel(const int64_t* __restrict__ eI, int64_t n, int ke, int64_t ld_e, const int32_t* __restrict__ F,
                  int64_t nf, int fw, int fw_pad, int64_t* __restrict__ out, int64_t ld_o, int64_t row_offset, int32_t* flags) {
  constexpr int S = DesimCfg<E>::kSlots, W = DesimCfg<E>::kWarps;
  constexpr uint32_t kNone = 0xffffffffu;
  __shared__ int32_t s_key[W][S];
  __shared__ uint32_t s_head[W][S];           // first column of the chain of columns holding s_key
  __shared__ uint32_t s_bloom[W][kBloomWords];
  __shared__ uint32_t s_next[W][32 * E];      // next column holding the same id
  __shared__ int32_t s_val[W][32 * E];        // pivot id of a column, -1 = never a pivot
  __shared__ uint32_t s_kill[W][32 * E][E];   // columns a pivot removes
  __shared__ uint32_t s_any[W][E];            // columns with a non-empty kill mask
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int32_t* key = s_key[warp];
  uint32_t* head = s_head[warp];
  uint32_t* bloom = s_bloom[warp];
  uint32_t* next = s_next[warp];
  int32_t* val = s_val[warp];
  uint32_t(*kill)[E] = s_kill[warp];
  uint32_t* any = s_any[warp];
  for (int i = lane; i < S; i += 32) key[i] = -1, head[i] = kNone;      // rows undo their own insertions afterwards
  for (int i = lane; i < kBloomWords; i += 32) bloom[i] = 0;
  if (lane < E) any[lane] = 0;
  for (int i = lane; i < 32 * E * E; i += 32) (&kill[0][0])[i] = 0;     // rows re-zero the masks they used
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * W + warp; r < n; r += static_cast<int64_t>(gridDim.x) * W) {
    __syncwarp();
    int64_t raw[E];
    int32_t ent[E];       // valid pivot id, -1 = padding, -2 = id outside the feature table (left untouched)
    uint32_t slot[E];
    uint32_t alive[E];    // warp-uniform copy of the alive words
#pragma unroll
    for (int t = 0; t < E; ++t) {
      const int c = lane + 32 * t;
      raw[t] = c < ke ? eI[r * ld_e + c] : -1;
      if (raw[t] < 0) ent[t] = -1;
      else if (raw[t] >= nf) {           // the reference would raise IndexError in fI[col_eI] (faiss_knn.py:179)
        ent[t] = -2;
        atomicOr(flags, 2);
      } else ent[t] = static_cast<int32_t>(raw[t]);
      val[c] = ent[t] >= 0 ? ent[t] : -1;
      alive[t] = __ballot_sync(0xffffffffu, ent[t] >= 0);
      slot[t] = 0;
      if (ent[t] >= 0) {
        const uint32_t m = desim_mix(ent[t]);
        atomicOr(&bloom[bloom_word(m)], bloom_bits(m));
        uint32_t h = (m >> 8) & (S - 1);
        while (true) {
          const int32_t old = atomicCAS(&key[h], -1, ent[t]);
          if (old == -1 || old == ent[t]) break;
          h = (h + 1) & (S - 1);
        }
        slot[t] = h;
        next[c] = atomicExch(&head[h], static_cast<uint32_t>(c));
      }
    }
    __syncwarp();
    // (1) relation
    for (int c0 = 0; c0 < ke; c0 += kUnroll) {
      int32_t f[kUnroll][FC];
#pragma unroll
      for (int p = 0; p < kUnroll; ++p) {
        const int32_t v = c0 + p < ke ? val[c0 + p] : -1;
#pragma unroll
        for (int q = 0; q < FC; ++q) {
          const int j = lane + 32 * q;
          f[p][q] = (v >= 0 && j < fw) ? __ldg(F + static_cast<int64_t>(v) * fw_pad + j) : -1;
        }
      }
#pragma unroll
      for (int p = 0; p < kUnroll; ++p) {
#pragma unroll
        for (int q = 0; q < FC; ++q) {
          const int32_t id = f[p][q];
          const uint32_t m = desim_mix(id);
          const bool maybe = id >= 0 && (bloom[bloom_word(m)] & bloom_bits(m)) == bloom_bits(m);
          if (maybe) {                                   // rare: ~1 % of the lanes
            const int c = c0 + p;
            uint32_t h = (m >> 8) & (S - 1);
            int32_t k = key[h];
            while (k != id && k != -1) {
              h = (h + 1) & (S - 1);
              k = key[h];
            }
            if (k == id) {
              for (uint32_t col = head[h]; col != kNone; col = next[col])
                if (static_cast<int>(col) > c) {
                  atomicOr(&kill[c][col >> 5], 1u << (col & 31));
                  atomicOr(&any[c >> 5], 1u << (c & 31));
                }
            }
          }
        }
      }
    }
    __syncwarp();
    // (2) resolution over the columns that remove anything, ascending
#pragma unroll
    for (int t = 0; t < E; ++t) {
      for (uint32_t todo = any[t]; todo != 0; todo &= todo - 1) {
        const int c = 32 * t + __ffs(todo) - 1;
        if ((alive[t] >> (c & 31)) & 1u) {
#pragma unroll
          for (int u = 0; u < E; ++u) alive[u] &= ~kill[c][u];
        }
      }
    }
    __syncwarp();
#pragma unroll
    for (int t = 0; t < E; ++t) {
      if ((any[t] >> lane) & 1u) {                           // leave the row state clean: re-zero the masks that were used
#pragma unroll
        for (int u = 0; u < E; ++u) kill[lane + 32 * t][u] = 0;
      }
    }
    __syncwarp();
#pragma unroll
    for (int t = 0; t < E; ++t) {
      const int c = lane + 32 * t;
      if (ent[t] >= 0) {                                     // undo this row's insertions
        const uint32_t m = desim_mix(ent[t]);
        key[slot[t]] = -1, head[slot[t]] = kNone, bloom[bloom_word(m)] = 0;
      }
      if (lane == 0) any[t] = 0;
      if (c < ke) {
        int64_t w = raw[t] < 0 ? -1 : raw[t];
        if (ent[t] >= 0 && !((alive[t] >> lane) & 1u)) w = -1;
        if (w == r + row_offset) w = -1;                   // the row's own id (rows of a slice keep their global number)
        out[r * ld_o + c] = w;
      }
    }
  }
}

// faiss_knn.desim (faiss_knn.py:134-143): eI[i, j] = -1 wherever eI[i, j] occurs in row i of fI.
__global__ void __launch_bounds__(256)
desim_simple_kernel(const int64_t* __restrict__ eI, int64_t n, int ke, int64_t ld_e, const int64_t* __restrict__ fI, int kf,
                    int64_t ld_f, int64_t* __restrict__ out, int64_t ld_o) {
  const int64_t total = n * ke;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / ke;
    const int c = static_cast<int>(i - r * ke);
    const int64_t v = eI[r * ld_e + c];
    bool hit = false;
    for (int j = 0; j < kf; ++j) hit |= fI[r * ld_f + j] == v;
    out[r * ld_o + c] = hit ? -1 : v;
  }
}

template <int E, int FC>
static int launch_rows_fc(cdml_ctx* ctx, cudaStream_t st, const int64_t* eI, int64_t n, int ke, int64_t ld_e, const int32_t* F,
                          int64_t nf, int fw, int fw_pad, int64_t* out, int64_t ld_o, int64_t row_offset) {
  constexpr int W = DesimCfg<E>::kWarps;
  auto kern = desim_rows_kernel<E, FC>;
  static int resident = 0;               // CTAs per SM: the grid is exactly one wave (rows are strided over it)
  if (resident == 0) {
    CDML_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, W * 32, 0));
    if (resident < 1) resident = 1;
  }
  const int64_t blocks = (n + W - 1) / W, cap = static_cast<int64_t>(ctx->num_sms) * resident;
  kern<<<static_cast<int>(blocks < cap ? blocks : cap), W * 32, 0, st>>>(eI, n, ke, ld_e, F, nf, fw, fw_pad, out, ld_o,
                                                                        row_offset, ctx->dev_flags);
  CDML_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int E>
static int launch_rows(cdml_ctx* ctx, int fc, cudaStream_t st, const int64_t* eI, int64_t n, int ke, int64_t ld_e,
                       const int32_t* F, int64_t nf, int fw, int fw_pad, int64_t* out, int64_t ld_o, int64_t row_offset) {
  if (fc == 1) return launch_rows_fc<E, 1>(ctx, st, eI, n, ke, ld_e, F, nf, fw, fw_pad, out, ld_o, row_offset);
  return launch_rows_fc<E, 2>(ctx, st, eI, n, ke, ld_e, F, nf, fw, fw_pad, out, ld_o, row_offset);
}

static inline int desim_width(int kf, int f_end) { return f_end < kf ? f_end : kf; }
static inline int pad32(int w) { return (w + 31) / 32 * 32; }

}  // namespace cdml

extern "C" {

int64_t cdml_desim_workspace_bytes(int64_t nf, int kf, int f_end) {
  if (nf <= 0 || kf <= 0 || f_end <= 0) return 0;
  return nf * cdml::pad32(cdml::desim_width(kf, f_end)) * static_cast<int64_t>(sizeof(int32_t));
}

int cdml_desim(cdml_ctx* ctx, const int64_t* eI, int64_t n, int ke, int64_t ld_e, const int64_t* fI, const float* fD,
               int64_t nf, int kf, int64_t ld_fi, int64_t ld_fd, float fD_threshold, int f_end, void* workspace,
               int64_t* out, int64_t ld_out, int64_t row_offset, void* stream) {
  using namespace cdml;
  CDML_REQUIRE(ctx && eI && fI && out && workspace, "cdml_desim: NULL argument");
  CDML_REQUIRE(n >= 0 && ke > 0 && nf > 0 && kf > 0 && f_end > 0 && ld_e >= ke && ld_out >= ke && ld_fi >= kf &&
               (fD == nullptr || ld_fd >= kf), "cdml_desim: bad geometry");
  CDML_REQUIRE(ke <= 256, "cdml_desim: at most 256 neighbours per row (got %d; the reference uses 81)", ke);
  CDML_REQUIRE(nf < (1ll << 31), "cdml_desim: feature table of %lld rows needs 64-bit ids", (long long)nf);
  const int fw = desim_width(kf, f_end), fw_pad = pad32(fw);
  CDML_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "cdml_desim: workspace must be 16-byte aligned");
  CDML_REQUIRE(fw <= 64, "cdml_desim: at most 64 feature neighbours per pivot (got %d; the reference uses 31)", fw);
  if (n == 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int32_t* F = static_cast<int32_t*>(workspace);
  {
    const int64_t blocks = (nf * fw_pad + 255) / 256, cap = static_cast<int64_t>(ctx->num_sms) * 8;
    desim_prepare_kernel<<<static_cast<int>(blocks < cap ? blocks : cap), 256, 0, st>>>(fI, fD, nf, kf, ld_fi, ld_fd,
                                                                                      fD_threshold, fw, fw_pad, F);
    CDML_CHECK_CUDA(cudaGetLastError());
  }
  const int fc = fw_pad / 32;
  const int e = (ke + 31) / 32;
  if (e <= 1) return launch_rows<1>(ctx, fc, st, eI, n, ke, ld_e, F, nf, fw, fw_pad, out, ld_out, row_offset);
  if (e == 2) return launch_rows<2>(ctx, fc, st, eI, n, ke, ld_e, F, nf, fw, fw_pad, out, ld_out, row_offset);
  if (e == 3) return launch_rows<3>(ctx, fc, st, eI, n, ke, ld_e, F, nf, fw, fw_pad, out, ld_out, row_offset);
  if (e == 4) return launch_rows<4>(ctx, fc, st, eI, n, ke, ld_e, F, nf, fw, fw_pad, out, ld_out, row_offset);
  return launch_rows<8>(ctx, fc, st, eI, n, ke, ld_e, F, nf, fw, fw_pad, out, ld_out, row_offset);
}

int cdml_desim_simple(cdml_ctx* ctx, const int64_t* eI, int64_t n, int ke, int64_t ld_e, const int64_t* fI, int kf,
                      int64_t ld_f, int64_t* out, int64_t ld_out, void* stream) {
  using namespace cdml;
  CDML_REQUIRE(ctx && eI && fI && out, "cdml_desim_simple: NULL argument");
  CDML_REQUIRE(n >= 0 && ke > 0 && kf > 0 && ld_e >= ke && ld_out >= ke && ld_f >= kf, "cdml_desim_simple: bad geometry");
  if (n == 0) return 0;
  const int64_t blocks = (n * ke + 255) / 256, cap = static_cast<int64_t>(ctx->num_sms) * 8;
  desim_simple_kernel<<<static_cast<int>(blocks < cap ? blocks : cap), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      eI, n, ke, ld_e, fI, kf, ld_f, out, ld_out);
  CDML_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"

// Host-side text formats either side of the hot path -- plain C++ on HOST pointers, no CUDA:
//   * cdml_format_knn_rows   the knn_split* / strict_knn* / cross_knn* line format of faiss_knn.write_process
//                            (faiss_knn.py:267-283): per-row Python map/lambda/join over 10M x 80 neighbours;
//   * cdml_parse_features_txt the 'guid#f1,f2,...' feature text of online_data.read_features_txt (online_data.py:48-84;
//                            the reference's own log reports 345 s for it).
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/cdml.h"
#include "common.cuh"

namespace cdml {

// str(numpy.float32(v)): shortest digits that round-trip (numpy's Dragon4 in unique mode; std::to_chars gives the same
// digits), positional for 1e-4 <= |v| < 1e6 (compared as doubles: float32(0.0001) < 1e-4 prints as 1e-04) with at least one
// digit after the point, otherwise d[.ddd]e[+-]XX.
static char* format_f32_numpy(char* p, float v) {
  if (v != v) { memcpy(p, "nan", 3); return p + 3; }
  if (v == 0.0f) { const char* z = std::signbit(v) ? "-0.0" : "0.0"; const size_t n = strlen(z); memcpy(p, z, n); return p + n; }
  if (v < 0) { *p++ = '-'; v = -v; }
  if (v > 3.4028235e38f) { memcpy(p, "inf", 3); return p + 3; }
  char sci[48];
  const auto r = std::to_chars(sci, sci + sizeof(sci), v, std::chars_format::scientific);
  // sci = d[.ddd]e[+-]XX
  const char* e = sci;
  while (*e != 'e') ++e;
  char digits[16];
  int nd = 0;
  for (const char* q = sci; q < e; ++q)
    if (*q != '.') digits[nd++] = *q;
  int ex = 0;
  for (const char* q = e + 2; q < r.ptr; ++q) ex = ex * 10 + (*q - '0');
  if (e[1] == '-') ex = -ex;
  const double a = static_cast<double>(v);
  if (!(a >= 1e-4 && a < 1e6)) {             // scientific: the to_chars layout is numpy's
    const size_t n = static_cast<size_t>(r.ptr - sci);
    memcpy(p, sci, n);
    return p + n;
  }
  if (ex < 0) {                              // 0.000ddd
    *p++ = '0', *p++ = '.';
    for (int i = 0; i < -ex - 1; ++i) *p++ = '0';
    memcpy(p, digits, nd);
    return p + nd;
  }
  for (int i = 0; i <= ex; ++i) *p++ = i < nd ? digits[i] : '0';
  *p++ = '.';
  if (nd > ex + 1) {
    memcpy(p, digits + ex + 1, nd - ex - 1);
    return p + (nd - ex - 1);
  }
  *p++ = '0';
  return p;
}

// One feature value as Python's float() reads it (optional blanks and sign, decimal / exponent, inf / nan): text -> double.
static bool parse_py_float(const char* b, const char* e, double* out) {
  while (b < e && (*b == ' ' || *b == '\t' || *b == '\r' || *b == '\n' || *b == '\f' || *b == '\v')) ++b;
  while (e > b && (e[-1] == ' ' || e[-1] == '\t' || e[-1] == '\r' || e[-1] == '\n' || e[-1] == '\f' || e[-1] == '\v')) --e;
  if (b == e) return false;
  bool neg = false;
  if (*b == '+' || *b == '-') {
    neg = *b == '-';
    ++b;
    if (b == e || *b == '+' || *b == '-') return false;
  }
  double v;
  const auto r = std::from_chars(b, e, v, std::chars_format::general);
  if (r.ec == std::errc::result_out_of_range) {            // float('1e999') = inf, float('1e-999') = 0.0
    const char* q = b;
    while (q < e && *q != 'e' && *q != 'E') ++q;
    bool big = true;
    if (q < e && q + 1 < e && q[1] == '-') big = false;
    v = big ? __builtin_inf() : 0.0;
    if (r.ptr != e) return false;
  } else if (r.ec != std::errc() || r.ptr != e) {
    const size_t n = static_cast<size_t>(e - b);
    auto ieq = [&](const char* w) {
      if (strlen(w) != n) return false;
      for (size_t i = 0; i < n; ++i)
        if ((b[i] | 0x20) != w[i]) return false;
      return true;
    };
    if (ieq("inf") || ieq("infinity")) v = __builtin_inf();
    else if (ieq("nan")) v = __builtin_nan("");
    else return false;
  }
  *out = neg ? -v : v;
  return true;
}

}  // namespace cdml

extern "C" {

int64_t cdml_format_knn_rows(const float* D, const int64_t* I, int64_t nq, int k, int64_t ld, int64_t begin_index,
                             const char* guid_blob, const int64_t* guid_off, int64_t n_guids, char* out, int64_t cap) {
  using namespace cdml;
  CDML_REQUIRE(D && I && guid_blob && guid_off && out, "cdml_format_knn_rows: NULL argument");
  CDML_REQUIRE(nq >= 0 && k >= 1 && ld >= k && begin_index >= 0 && begin_index + nq <= n_guids,
               "cdml_format_knn_rows: rows [%lld, %lld) outside the decode map of %lld guids", (long long)begin_index,
               (long long)(begin_index + nq), (long long)n_guids);
  char* p = out;
  char* const end = out + cap;
  for (int64_t i = 0; i < nq; ++i) {
    const int64_t q = begin_index + i;
    const int64_t qlen = guid_off[q + 1] - guid_off[q];
    CDML_REQUIRE(p + qlen + 2 <= end, "cdml_format_knn_rows: output buffer of %lld bytes is too small", (long long)cap);
    memcpy(p, guid_blob + guid_off[q], qlen);
    p += qlen;
    *p++ = ',';
    for (int j = 1; j < k; ++j) {
      const int64_t idx = I[i * ld + j];
      const float dist = D[i * ld + j];
      if (!(idx > 0 && dist > 0.0f && dist < 1.4f)) continue;        // faiss_knn.py:277
      CDML_REQUIRE(idx < n_guids, "cdml_format_knn_rows: neighbour id %lld is not in the decode map", (long long)idx);
      const int64_t glen = guid_off[idx + 1] - guid_off[idx];
      CDML_REQUIRE(p + glen + 40 <= end, "cdml_format_knn_rows: output buffer of %lld bytes is too small", (long long)cap);
      memcpy(p, guid_blob + guid_off[idx], glen);
      p += glen;
      *p++ = '#';
      p = format_f32_numpy(p, dist);
      *p++ = '<';
    }
    CDML_REQUIRE(p + 1 <= end, "cdml_format_knn_rows: output buffer of %lld bytes is too small", (long long)cap);
    *p++ = '\n';
  }
  return p - out;
}

int64_t cdml_format_f32(const float* v, int64_t n, char* out, int64_t cap) {
  using namespace cdml;
  CDML_REQUIRE(v && out && cap >= 32 * n, "cdml_format_f32: need 32 bytes per value");
  char* p = out;
  for (int64_t i = 0; i < n; ++i) {
    p = format_f32_numpy(p, v[i]);
    *p++ = '\n';
  }
  return p - out;
}

int64_t cdml_parse_features_txt(const char* buf, int64_t len, int width, float* out, int64_t max_rows, int64_t* guid_begin,
                                int32_t* guid_len, int num_threads) {
  using namespace cdml;
  CDML_REQUIRE(buf && out && guid_begin && guid_len && width > 0 && len >= 0 && max_rows >= 0,
               "cdml_parse_features_txt: bad argument");
  std::vector<int64_t> starts;          // line starts; a final line without '\n' counts, an empty tail does not
  for (int64_t s = 0; s < len;) {
    starts.push_back(s);
    const void* nl = memchr(buf + s, '\n', static_cast<size_t>(len - s));
    s = nl ? static_cast<const char*>(nl) - buf + 1 : len;
  }
  const int64_t nl = static_cast<int64_t>(starts.size());
  starts.push_back(len);
  CDML_REQUIRE(nl <= max_rows, "cdml_parse_features_txt: %lld lines but room for %lld rows", (long long)nl, (long long)max_rows);
  std::vector<uint8_t> ok(static_cast<size_t>(nl), 0);
  auto work = [&](int64_t lo, int64_t hi) {
    for (int64_t i = lo; i < hi; ++i) {
      const char* b = buf + starts[i];
      const char* e = buf + starts[i + 1];
      if (e > b && e[-1] == '\n') --e;                       // line.strip('\n')
      while (e > b && e[-1] == '\n') --e;
      while (b < e && *b == '\n') ++b;
      const char* hash = static_cast<const char*>(memchr(b, '#', static_cast<size_t>(e - b)));
      if (hash == nullptr || memchr(hash + 1, '#', static_cast<size_t>(e - hash - 1)) != nullptr) continue;   // split('#') != 2 parts
      float* row = out + i * static_cast<int64_t>(width);
      int n = 0;
      bool good = true;
      const char* f = hash + 1;
      while (true) {
        const char* c = static_cast<const char*>(memchr(f, ',', static_cast<size_t>(e - f)));
        const char* fe = c ? c : e;
        double v;
        if (!parse_py_float(f, fe, &v)) { good = false; break; }
        if (n < width) row[n] = static_cast<float>(v);       // features[i] = feature: double -> float32
        ++n;
        if (!c) break;
        f = c + 1;
      }
      if (!good || n != width) continue;
      guid_begin[i] = b - buf;
      guid_len[i] = static_cast<int32_t>(hash - b);
      ok[i] = 1;
    }
  };
  int nt = num_threads > 0 ? num_threads : static_cast<int>(std::thread::hardware_concurrency());
  if (nt < 1) nt = 1;
  if (nt > nl) nt = static_cast<int>(nl > 0 ? nl : 1);
  std::vector<std::thread> pool;
  for (int t = 1; t < nt; ++t) pool.emplace_back(work, nl * t / nt, nl * (t + 1) / nt);
  work(0, nl / nt);
  for (auto& th : pool) th.join();
  int64_t kept = 0;                     // compact the kept rows to the front, in file order
  for (int64_t i = 0; i < nl; ++i) {
    if (!ok[i]) continue;
    if (kept != i) {
      memmove(out + kept * static_cast<int64_t>(width), out + i * static_cast<int64_t>(width), sizeof(float) * width);
      guid_begin[kept] = guid_begin[i], guid_len[kept] = guid_len[i];
    }
    ++kept;
  }
  return kept;
}

}  // extern "C"

// Host-side input pipeline behind include/b200rec_io.h: text -> packed batch arrays.
//
// The input is cut into T byte ranges; thread t finds the lines that start in its range, a prefix
// sum over the line counts gives every thread the first output row it owns, and each thread then
// parses its lines straight into the caller's arrays — so the output is independent of the thread
// count and the fixed-length formats need no merge step.  The variable-length formats parse into
// thread-local vectors that are stitched with one memcpy per thread.
//
// Reference behaviour followed (not its code): models/rank/deepfm/criteo_reader.py:61-103,
// tools/dataset/parser.cpp:37-77, models/rank/dnn/benchmark_reader.py:39-56.
#include "b200rec_io.h"

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <string_view>
#include <system_error>
#include <thread>
#include <vector>

namespace {

thread_local std::string g_error;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_error = buf;
  return code;
}

struct Line {
  const char* p;
  const char* e;
};

bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\v' || c == '\f'; }

// One line of the input with Python's l.strip() applied (strip) or only the newline / CR removed.
inline bool trim_line(const char* p, const char* e, bool strip, Line* out) {
  const char* a = p;
  const char* b = e;
  if (strip) {
    while (a < b && is_space(*a)) ++a;
    while (b > a && is_space(b[-1])) --b;
  } else if (b > a && b[-1] == '\r') {
    --b;
  }
  *out = {a, b};
  return a < b;  // stripped: blank == empty; raw: only truly empty lines are dropped
}

int thread_count(int requested, size_t len) {
  int t = requested > 0 ? requested : int(std::thread::hardware_concurrency());
  if (t < 1) t = 1;
  // below ~64 KiB per thread the spawn cost exceeds the parse
  size_t cap = std::max<size_t>(1, len >> 16);
  return int(std::min<size_t>(size_t(t), cap));
}

template <class Fn>
void run_threads(int T, Fn fn) {
  if (T == 1) { fn(0); return; }
  std::vector<std::thread> th;
  th.reserve((size_t)T);
  for (int t = 0; t < T; ++t) th.emplace_back([&fn, t] { fn(t); });
  for (auto& x : th) x.join();
}

// The non-empty lines of text[0,len) cut into T consecutive groups (thread t owns the lines that
// START in its byte range), found in parallel.  base[t] = index of group t's first line in the
// whole input, so thread t writes rows base[t].. and the output is independent of T.
struct Partition {
  int T = 1;
  std::vector<std::vector<Line>> lines;
  std::vector<size_t> base;
  std::vector<int64_t> skipped;
  size_t total = 0;
  int64_t total_skipped = 0;
};

template <class Keep>
Partition partition_lines(const char* text, size_t len, bool strip, int n_threads, Keep keep) {
  Partition pt;
  pt.T = thread_count(n_threads, len);
  const int T = pt.T;
  pt.lines.resize((size_t)T);
  pt.skipped.assign((size_t)T, 0);
  const char* const end = text + len;
  auto first_start = [&](int t) -> const char* {  // first line start at or after the t-th cut
    if (t <= 0) return text;
    if (t >= T) return end;
    const char* c = text + len * size_t(t) / size_t(T);
    if (c[-1] == '\n') return c;
    const char* nl = static_cast<const char*>(memchr(c, '\n', size_t(end - c)));
    return nl ? nl + 1 : end;
  };
  run_threads(T, [&](int t) {
    const char* p = first_start(t);
    const char* const stop = first_start(t + 1);
    auto& out = pt.lines[size_t(t)];
    out.reserve(size_t(stop - p) / 256 + 16);
    while (p < stop) {
      const char* nl = static_cast<const char*>(memchr(p, '\n', size_t(end - p)));
      const char* e = nl ? nl : end;
      Line ln;
      if (trim_line(p, e, strip, &ln)) {
        if (keep(ln)) out.push_back(ln); else ++pt.skipped[size_t(t)];
      }
      p = nl ? nl + 1 : end;
    }
  });
  pt.base.resize((size_t)T);
  for (int t = 0; t < T; ++t) {
    pt.base[size_t(t)] = pt.total;
    pt.total += pt.lines[size_t(t)].size();
    pt.total_skipped += pt.skipped[size_t(t)];
  }
  return pt;
}

inline bool keep_all(const Line&) { return true; }

struct Status {
  int code = B200REC_IO_OK;
  std::string msg;
  void set(int c, int64_t line, const char* what, std::string_view tok) {
    if (code != B200REC_IO_OK) return;
    code = c;
    char buf[384];
    snprintf(buf, sizeof buf, "line %lld: %s '%.*s'", (long long)(line + 1), what,
             int(std::min<size_t>(tok.size(), 120)), tok.data());
    msg = buf;
  }
};

// Runs body(t, lines of group t, index of its first line, status); returns the first error in
// LINE order so the message does not depend on scheduling.  ("line N" counts non-empty lines.)
template <class Body>
int parallel_lines(const Partition& pt, Body body) {
  std::vector<Status> st((size_t)pt.T);
  run_threads(pt.T, [&](int t) { body(t, pt.lines[size_t(t)], pt.base[size_t(t)], st[size_t(t)]); });
  for (auto& x : st)
    if (x.code != B200REC_IO_OK) return fail(x.code, "%s", x.msg.c_str());
  return B200REC_IO_OK;
}

bool parse_i64(std::string_view s, int64_t* out) {
  const char* a = s.data();
  const char* b = a + s.size();
  if (a < b && *a == '+') ++a;  // int("+5") is legal Python
  if (a == b) return false;
  auto r = std::from_chars(a, b, *out, 10);
  return r.ec == std::errc() && r.ptr == b;
}

bool parse_u64(std::string_view s, uint64_t* out) {
  const char* a = s.data();
  const char* b = a + s.size();
  if (a == b) return false;
  auto r = std::from_chars(a, b, *out, 10);
  return r.ec == std::errc() && r.ptr == b;
}

// float(token): correctly rounded double, like Python; the caller narrows to float32 like
// np.array(..).astype('float32').
bool parse_f64(std::string_view s, double* out) {
  const char* a = s.data();
  const char* b = a + s.size();
  if (a < b && *a == '+') ++a;
  if (a == b) return false;
  auto r = std::from_chars(a, b, *out, std::chars_format::general);
  if (r.ec == std::errc::result_out_of_range) {  // Python gives +-inf / +-0.0 here, not an error
    if (r.ptr != b) return false;
    std::string z(s);
    *out = strtod(z.c_str(), nullptr);
    return true;
  }
  return r.ec == std::errc() && r.ptr == b;
}

// ---- schema of the slot-text formats ------------------------------------------------------------
struct SlotSchema {
  std::vector<std::string> sparse;
  std::string label, dense;
  bool has_label = false, has_dense = false;
  int dense_dim = 0;

  // slot name -> 0..n_sparse-1 sparse, -2 label, -3 dense, -1 unknown.  A line has ~40 tokens, so
  // this lookup is the hot spot of the parser: names of <= 7 bytes (all of Criteo's sparse slots)
  // are packed into one integer and found in a small open-addressing table; longer names fall
  // back to a scan with a length check.
  std::vector<uint64_t> tab_key;
  std::vector<int> tab_val;
  std::vector<std::pair<std::string, int>> long_names;
  uint64_t tab_mask = 0;

  static uint64_t pack(std::string_view s) {
    uint64_t k = 0;
    memcpy(&k, s.data(), s.size());
    return k | (uint64_t(s.size() + 1) << 56);  // never 0: 0 marks an empty table cell
  }
  static uint64_t mix(uint64_t k) { return (k * 0x9E3779B97F4A7C15ULL) >> 40; }

  void add(const std::string& name, int kind) {
    if (name.size() > 7) { long_names.emplace_back(name, kind); return; }
    const uint64_t k = pack(name);
    for (uint64_t h = mix(k) & tab_mask;; h = (h + 1) & tab_mask) {
      if (tab_key[h] == k) return;  // a name listed twice keeps its first meaning
      if (tab_key[h] == 0) { tab_key[h] = k; tab_val[h] = kind; return; }
    }
  }

  void index() {
    size_t cells = 16;
    while (cells < 4 * (sparse.size() + 2)) cells *= 2;
    tab_key.assign(cells, 0);
    tab_val.assign(cells, -1);
    tab_mask = cells - 1;
    long_names.clear();
    if (has_dense) add(dense, -3);
    if (has_label) add(label, -2);
    for (size_t i = 0; i < sparse.size(); ++i) add(sparse[i], int(i));
  }

  int find(std::string_view name) const {
    if (name.size() <= 7) {
      const uint64_t k = pack(name);
      for (uint64_t h = mix(k) & tab_mask;; h = (h + 1) & tab_mask) {
        if (tab_key[h] == k) return tab_val[h];
        if (tab_key[h] == 0) return -1;
      }
    }
    for (const auto& ln : long_names)
      if (ln.first.size() == name.size() && memcmp(ln.first.data(), name.data(), name.size()) == 0) return ln.second;
    return -1;
  }
};

int make_schema(SlotSchema& sc, const char* label_slot, const char* const* sparse_slots,
                int n_sparse, const char* dense_slot, int dense_dim) {
  if (n_sparse < 0 || (n_sparse > 0 && !sparse_slots)) return fail(B200REC_IO_ERR_ARG, "sparse_slots is null");
  if (dense_slot && dense_dim <= 0) return fail(B200REC_IO_ERR_ARG, "dense_dim must be > 0 (got %d)", dense_dim);
  sc.has_label = label_slot != nullptr;
  if (sc.has_label) sc.label = label_slot;
  sc.has_dense = dense_slot != nullptr;
  if (sc.has_dense) sc.dense = dense_slot;
  sc.dense_dim = sc.has_dense ? dense_dim : 0;
  for (int i = 0; i < n_sparse; ++i) {
    if (!sparse_slots[i]) return fail(B200REC_IO_ERR_ARG, "sparse_slots[%d] is null", i);
    sc.sparse.emplace_back(sparse_slots[i]);
  }
  sc.index();
  return B200REC_IO_OK;
}

// Walks the `slot:value` tokens of one line.  on_token(kind, value_view) returns false to stop.
template <class F>
bool for_each_slot_token(const Line& ln, const SlotSchema& sc, int64_t line_no, Status& st, F on_token) {
  const char* p = ln.p;
  while (p <= ln.e) {
    const char* sp = static_cast<const char*>(memchr(p, ' ', size_t(ln.e - p)));
    const char* te = sp ? sp : ln.e;
    std::string_view tok(p, size_t(te - p));
    p = te + 1;
    if (tok.empty()) { if (!sp) break; continue; }  // "a  b".split(" ") yields '' — not a slot
    size_t colon = tok.find(':');
    std::string_view name = colon == std::string_view::npos ? tok : tok.substr(0, colon);
    int kind = sc.find(name);
    if (kind == -1) { if (!sp) break; continue; }
    if (colon == std::string_view::npos) {  // slot_feasign[1] -> IndexError in the reference
      st.set(B200REC_IO_ERR_PARSE, line_no, "slot token without ':value'", tok);
      return false;
    }
    // "a:b:c".split(":")[1] == "b": the value ends at a second colon
    std::string_view val = tok.substr(colon + 1);
    size_t c2 = val.find(':');
    if (c2 != std::string_view::npos) val = val.substr(0, c2);
    if (!on_token(kind, val, tok)) return false;
    if (!sp) break;
  }
  return true;
}

// ---- string hashes ------------------------------------------------------------------------------
// libstdc++'s 64-bit std::_Hash_bytes (MurmurHash64A, seed 0xc70f6907) = std::hash<std::string>.
uint64_t murmur64a(const char* data, size_t len) {
  const uint64_t mul = (uint64_t(0xc6a4a793UL) << 32) + uint64_t(0x5bd1e995UL);
  const uint64_t seed = 0xc70f6907UL;
  auto load8 = [](const char* p) { uint64_t r; memcpy(&r, p, 8); return r; };
  auto shift_mix = [](uint64_t v) { return v ^ (v >> 47); };
  const char* const end = data + (len & ~size_t(7));
  uint64_t hash = seed ^ (uint64_t(len) * mul);
  for (const char* p = data; p != end; p += 8) {
    const uint64_t d = shift_mix(load8(p) * mul) * mul;
    hash ^= d;
    hash *= mul;
  }
  if ((len & 7) != 0) {
    uint64_t d = 0;
    for (int n = int(len & 7) - 1; n >= 0; --n) d = (d << 8) + uint8_t(end[n]);
    hash ^= d;
    hash *= mul;
  }
  hash = shift_mix(hash) * mul;
  hash = shift_mix(hash);
  return hash;
}

// xxHash32 (Y. Collet, XXH32 as published in the xxHash specification).
uint32_t xxh32(const char* data, size_t len, uint32_t seed) {
  const uint32_t P1 = 2654435761U, P2 = 2246822519U, P3 = 3266489917U, P4 = 668265263U, P5 = 374761393U;
  auto rotl = [](uint32_t x, int r) { return (x << r) | (x >> (32 - r)); };
  auto rd32 = [](const uint8_t* p) { return uint32_t(p[0]) | uint32_t(p[1]) << 8 | uint32_t(p[2]) << 16 | uint32_t(p[3]) << 24; };
  const uint8_t* p = reinterpret_cast<const uint8_t*>(data);
  const uint8_t* const end = p + len;
  uint32_t h;
  if (len >= 16) {
    uint32_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
    const uint8_t* const limit = end - 16;
    do {
      v1 = rotl(v1 + rd32(p) * P2, 13) * P1; p += 4;
      v2 = rotl(v2 + rd32(p) * P2, 13) * P1; p += 4;
      v3 = rotl(v3 + rd32(p) * P2, 13) * P1; p += 4;
      v4 = rotl(v4 + rd32(p) * P2, 13) * P1; p += 4;
    } while (p <= limit);
    h = rotl(v1, 1) + rotl(v2, 7) + rotl(v3, 12) + rotl(v4, 18);
  } else {
    h = seed + P5;
  }
  h += uint32_t(len);
  while (p + 4 <= end) { h = rotl(h + rd32(p) * P3, 17) * P4; p += 4; }
  while (p < end) { h = rotl(h + uint32_t(*p) * P5, 11) * P1; ++p; }
  h ^= h >> 15; h *= P2;
  h ^= h >> 13; h *= P3;
  h ^= h >> 16;
  return h;
}

const double kContMin[13] = {0, -3, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
const double kContDiff[13] = {20, 603, 100, 50, 64000, 500, 100, 50, 500, 10, 10, 10, 50};

}  // namespace

extern "C" {

int b200rec_io_abi_version(void) { return B200REC_IO_ABI_VERSION; }
const char* b200rec_io_last_error(void) { return g_error.c_str(); }

uint64_t b200rec_io_hash_std_string(const char* s, size_t len) { return murmur64a(s, len); }
uint32_t b200rec_io_xxh32(const char* s, size_t len, uint32_t seed) { return xxh32(s, len, seed); }

int b200rec_io_count_lines(const char* text, size_t len, int64_t* n_lines) {
  if ((!text && len) || !n_lines) return fail(B200REC_IO_ERR_ARG, "null argument");
  *n_lines = int64_t(partition_lines(text, len, false, 0, keep_all).total);
  return B200REC_IO_OK;
}

int b200rec_io_parse_slot_text(const char* text, size_t len, const char* label_slot,
                               const char* const* sparse_slots, int n_sparse,
                               const char* dense_slot, int dense_dim, int64_t* label,
                               int64_t* ids, float* dense, int64_t cap, int64_t* n_out,
                               int n_threads) {
  return b200rec_io_parse_slot_text_ex(text, len, label_slot, sparse_slots, n_sparse, dense_slot,
                                       dense_dim, 0, label, ids, dense, cap, n_out, n_threads);
}

int b200rec_io_parse_slot_text_ex(const char* text, size_t len, const char* label_slot,
                                  const char* const* sparse_slots, int n_sparse,
                                  const char* dense_slot, int dense_dim, int flags, int64_t* label,
                                  int64_t* ids, float* dense, int64_t cap, int64_t* n_out,
                                  int n_threads) {
  if ((!text && len) || !n_out) return fail(B200REC_IO_ERR_ARG, "null argument");
  if (flags & ~(B200REC_IO_DENSE_LOG1P | B200REC_IO_SKIP_EMPTY_SPARSE))
    return fail(B200REC_IO_ERR_ARG, "unknown flags 0x%x", flags);
  const bool log1p_dense = (flags & B200REC_IO_DENSE_LOG1P) != 0;
  const bool skip_empty = (flags & B200REC_IO_SKIP_EMPTY_SPARSE) != 0;
  SlotSchema sc;
  if (int rc = make_schema(sc, label_slot, sparse_slots, n_sparse, dense_slot, dense_dim)) return rc;
  if ((sc.has_label && !label) || (n_sparse > 0 && !ids) || (sc.has_dense && !dense))
    return fail(B200REC_IO_ERR_ARG, "output buffer for a declared slot is null");
  const Partition pt = partition_lines(text, len, true, n_threads, keep_all);
  *n_out = 0;
  if (int64_t(pt.total) > cap)
    return fail(B200REC_IO_ERR_CAPACITY, "%zu samples but cap = %lld", pt.total, (long long)cap);
  const int F = n_sparse, Dn = sc.dense_dim;
  int rc = parallel_lines(pt, [&](int, const std::vector<Line>& lines, size_t base, Status& st) {
    std::vector<uint8_t> seen((size_t)F + 1);
    for (size_t i = 0; i < lines.size() && st.code == B200REC_IO_OK; ++i) {
      const int64_t n = int64_t(base + i);
      int64_t* row = ids ? ids + n * F : nullptr;
      float* drow = dense ? dense + n * Dn : nullptr;
      std::fill(seen.begin(), seen.end(), 0);
      for (int f = 0; f < F; ++f) row[f] = 0;
      for (int j = 0; j < Dn; ++j) drow[j] = 0.f;
      if (sc.has_label) label[n] = 0;
      int n_dense = 0;
      for_each_slot_token(lines[i], sc, n, st, [&](int kind, std::string_view val, std::string_view tok) {
        if (kind == -3) {
          double v;
          if (!parse_f64(val, &v)) { st.set(B200REC_IO_ERR_PARSE, n, "bad float in", tok); return false; }
          if (n_dense >= Dn) { st.set(B200REC_IO_ERR_RAGGED, n, "too many dense values at", tok); return false; }
          drow[n_dense++] = float(log1p_dense ? std::log(v + 1.0) : v);
          return true;
        }
        if (skip_empty && val.empty()) return true;
        int64_t v;
        if (!parse_i64(val, &v)) { st.set(B200REC_IO_ERR_PARSE, n, "bad integer in", tok); return false; }
        uint8_t& s = seen[kind == -2 ? size_t(F) : size_t(kind)];
        if (s) { st.set(B200REC_IO_ERR_RAGGED, n, "second value for a fixed-length slot at", tok); return false; }
        s = 1;
        if (kind == -2) label[n] = v; else row[kind] = v;
        return true;
      });
      if (st.code == B200REC_IO_OK && n_dense != 0 && n_dense != Dn)
        st.set(B200REC_IO_ERR_RAGGED, n, "dense slot is shorter than dense_dim:", sc.dense);
    }
  });
  if (rc) return rc;
  *n_out = int64_t(pt.total);
  return B200REC_IO_OK;
}

int b200rec_io_parse_slot_text_lod(const char* text, size_t len, const char* label_slot,
                                   const char* const* sparse_slots, int n_sparse,
                                   const char* dense_slot, int dense_dim, int64_t* label,
                                   int64_t* keys, int64_t* offsets, float* dense, int64_t cap,
                                   int64_t keys_cap, int64_t* n_out, int64_t* n_keys_out,
                                   int n_threads) {
  if ((!text && len) || !n_out || !n_keys_out || !offsets) return fail(B200REC_IO_ERR_ARG, "null argument");
  SlotSchema sc;
  if (int rc = make_schema(sc, label_slot, sparse_slots, n_sparse, dense_slot, dense_dim)) return rc;
  if ((sc.has_label && !label) || (keys_cap > 0 && !keys) || (sc.has_dense && !dense))
    return fail(B200REC_IO_ERR_ARG, "output buffer for a declared slot is null");
  const Partition pt = partition_lines(text, len, true, n_threads, keep_all);
  *n_out = 0;
  *n_keys_out = 0;
  if (int64_t(pt.total) > cap)
    return fail(B200REC_IO_ERR_CAPACITY, "%zu samples but cap = %lld", pt.total, (long long)cap);
  const int F = n_sparse, Dn = sc.dense_dim;
  struct Local { std::vector<int64_t> keys; std::vector<int32_t> bag_len; };
  std::vector<Local> local((size_t)pt.T);
  int rc = parallel_lines(pt, [&](int t, const std::vector<Line>& lines, size_t base, Status& st) {
    Local& L = local[size_t(t)];
    L.bag_len.reserve(lines.size() * size_t(F));
    std::vector<std::vector<int64_t>> bags((size_t)F);
    for (size_t i = 0; i < lines.size() && st.code == B200REC_IO_OK; ++i) {
      const int64_t n = int64_t(base + i);
      float* drow = dense ? dense + n * Dn : nullptr;
      for (auto& b : bags) b.clear();
      for (int j = 0; j < Dn; ++j) drow[j] = 0.f;
      if (sc.has_label) label[n] = 0;
      int n_dense = 0;
      bool label_seen = false;
      for_each_slot_token(lines[i], sc, n, st, [&](int kind, std::string_view val, std::string_view tok) {
        if (kind == -3) {
          double v;
          if (!parse_f64(val, &v)) { st.set(B200REC_IO_ERR_PARSE, n, "bad float in", tok); return false; }
          if (n_dense >= Dn) { st.set(B200REC_IO_ERR_RAGGED, n, "too many dense values at", tok); return false; }
          drow[n_dense++] = float(v);
          return true;
        }
        int64_t v;
        if (!parse_i64(val, &v)) { st.set(B200REC_IO_ERR_PARSE, n, "bad integer in", tok); return false; }
        if (kind == -2) {
          if (label_seen) { st.set(B200REC_IO_ERR_RAGGED, n, "second label at", tok); return false; }
          label_seen = true;
          label[n] = v;
        } else {
          bags[size_t(kind)].push_back(v);
        }
        return true;
      });
      if (st.code == B200REC_IO_OK && n_dense != 0 && n_dense != Dn)
        st.set(B200REC_IO_ERR_RAGGED, n, "dense slot is shorter than dense_dim:", sc.dense);
      for (int f = 0; f < F; ++f) {
        auto& b = bags[size_t(f)];
        if (b.empty()) b.push_back(0);  // padded like criteo_reader.py:88-89
        L.bag_len.push_back(int32_t(b.size()));
        L.keys.insert(L.keys.end(), b.begin(), b.end());
      }
    }
  });
  if (rc) return rc;
  int64_t total = 0;
  for (auto& L : local) total += int64_t(L.keys.size());
  if (total > keys_cap)
    return fail(B200REC_IO_ERR_CAPACITY, "%lld keys but keys_cap = %lld", (long long)total, (long long)keys_cap);
  int64_t at = 0;
  for (int t = 0; t < pt.T; ++t) {
    Local& L = local[size_t(t)];
    int64_t* off = offsets + int64_t(pt.base[size_t(t)]) * F;
    for (size_t i = 0; i < L.bag_len.size(); ++i) { off[i] = at; at += L.bag_len[i]; }
    if (!L.keys.empty()) memcpy(keys + (at - int64_t(L.keys.size())), L.keys.data(), L.keys.size() * sizeof(int64_t));
  }
  offsets[int64_t(pt.total) * F] = at;
  *n_out = int64_t(pt.total);
  *n_keys_out = total;
  return B200REC_IO_OK;
}

int b200rec_io_parse_multislot(const char* text, size_t len, const int* slot_is_float, int n_slots,
                               uint64_t* keys, int64_t* koffsets, int64_t keys_cap, float* fvals,
                               int64_t* foffsets, int64_t fvals_cap, int64_t cap, int64_t* n_out,
                               int64_t* n_keys_out, int64_t* n_fvals_out, int n_threads) {
  if ((!text && len) || !n_out || !n_keys_out || !n_fvals_out || !slot_is_float || n_slots <= 0)
    return fail(B200REC_IO_ERR_ARG, "null argument / n_slots <= 0");
  int n_int = 0, n_float = 0;
  for (int s = 0; s < n_slots; ++s) (slot_is_float[s] ? n_float : n_int)++;
  if ((n_int && (!koffsets || (keys_cap > 0 && !keys))) || (n_float && (!foffsets || (fvals_cap > 0 && !fvals))))
    return fail(B200REC_IO_ERR_ARG, "output buffer for a declared slot type is null");
  const Partition pt = partition_lines(text, len, true, n_threads, keep_all);
  *n_out = *n_keys_out = *n_fvals_out = 0;
  if (int64_t(pt.total) > cap)
    return fail(B200REC_IO_ERR_CAPACITY, "%zu samples but cap = %lld", pt.total, (long long)cap);
  struct Local {
    std::vector<uint64_t> keys; std::vector<float> fvals;
    std::vector<int32_t> klen, flen;
  };
  std::vector<Local> local((size_t)pt.T);
  int rc = parallel_lines(pt, [&](int t, const std::vector<Line>& lines, size_t base, Status& st) {
    Local& L = local[size_t(t)];
    for (size_t i = 0; i < lines.size() && st.code == B200REC_IO_OK; ++i) {
      const int64_t n = int64_t(base + i);
      const char* p = lines[i].p;
      const char* const e = lines[i].e;
      auto next = [&](std::string_view* tok) {  // tokens separated by runs of blanks (str.split())
        while (p < e && is_space(*p)) ++p;
        if (p >= e) return false;
        const char* a = p;
        while (p < e && !is_space(*p)) ++p;
        *tok = std::string_view(a, size_t(p - a));
        return true;
      };
      for (int s = 0; s < n_slots && st.code == B200REC_IO_OK; ++s) {
        std::string_view tok;
        int64_t cnt;
        if (!next(&tok)) { st.set(B200REC_IO_ERR_PARSE, n, "line ends before slot", std::to_string(s)); break; }
        if (!parse_i64(tok, &cnt) || cnt <= 0) { st.set(B200REC_IO_ERR_PARSE, n, "slot length must be a positive integer, got", tok); break; }
        (slot_is_float[s] ? L.flen : L.klen).push_back(int32_t(cnt));
        for (int64_t k = 0; k < cnt; ++k) {
          if (!next(&tok)) { st.set(B200REC_IO_ERR_PARSE, n, "line ends inside slot", std::to_string(s)); break; }
          if (slot_is_float[s]) {
            double v;
            if (!parse_f64(tok, &v)) { st.set(B200REC_IO_ERR_PARSE, n, "bad float", tok); break; }
            L.fvals.push_back(float(v));
          } else {
            uint64_t v;
            if (!parse_u64(tok, &v)) { st.set(B200REC_IO_ERR_PARSE, n, "bad feasign", tok); break; }
            L.keys.push_back(v);
          }
        }
      }
      std::string_view extra;
      if (st.code == B200REC_IO_OK && next(&extra))
        st.set(B200REC_IO_ERR_PARSE, n, "tokens left after the last slot:", extra);
    }
  });
  if (rc) return rc;
  int64_t nk = 0, nf = 0;
  for (auto& L : local) { nk += int64_t(L.keys.size()); nf += int64_t(L.fvals.size()); }
  if (nk > keys_cap) return fail(B200REC_IO_ERR_CAPACITY, "%lld keys but keys_cap = %lld", (long long)nk, (long long)keys_cap);
  if (nf > fvals_cap) return fail(B200REC_IO_ERR_CAPACITY, "%lld floats but fvals_cap = %lld", (long long)nf, (long long)fvals_cap);
  int64_t ka = 0, fa = 0;
  for (int t = 0; t < pt.T; ++t) {
    Local& L = local[size_t(t)];
    if (n_int) {
      int64_t* off = koffsets + int64_t(pt.base[size_t(t)]) * n_int;
      for (size_t i = 0; i < L.klen.size(); ++i) { off[i] = ka; ka += L.klen[i]; }
      if (!L.keys.empty()) memcpy(keys + (ka - int64_t(L.keys.size())), L.keys.data(), L.keys.size() * sizeof(uint64_t));
    }
    if (n_float) {
      int64_t* off = foffsets + int64_t(pt.base[size_t(t)]) * n_float;
      for (size_t i = 0; i < L.flen.size(); ++i) { off[i] = fa; fa += L.flen[i]; }
      if (!L.fvals.empty()) memcpy(fvals + (fa - int64_t(L.fvals.size())), L.fvals.data(), L.fvals.size() * sizeof(float));
    }
  }
  if (n_int) koffsets[int64_t(pt.total) * n_int] = ka;
  if (n_float) foffsets[int64_t(pt.total) * n_float] = fa;
  *n_out = int64_t(pt.total);
  *n_keys_out = nk;
  *n_fvals_out = nf;
  return B200REC_IO_OK;
}

int b200rec_io_parse_criteo_tsv(const char* text, size_t len, int hash_kind, int64_t hash_dim,
                                const double* cont_min, const double* cont_diff, int64_t* label,
                                int64_t* ids, float* dense, int64_t cap, int64_t* n_out,
                                int64_t* n_skipped_out, int n_threads) {
  if ((!text && len) || !n_out || !label || !ids || !dense) return fail(B200REC_IO_ERR_ARG, "null argument");
  if (hash_kind != B200REC_IO_HASH_STD && hash_kind != B200REC_IO_HASH_XXH32)
    return fail(B200REC_IO_ERR_ARG, "unknown hash_kind %d", hash_kind);
  if (hash_dim <= 0) return fail(B200REC_IO_ERR_ARG, "hash_dim must be > 0");
  if (!cont_min) cont_min = kContMin;
  if (!cont_diff) cont_diff = kContDiff;
  constexpr int kCols = 40, kDense = 13, kSparse = 26;
  // getline / rstrip('\n') keep leading blanks and tabs: only the newline is removed.  parser.cpp
  // drops lines that do not have exactly 40 columns (:50-52); the Python reader has no such check.
  const bool filter = hash_kind == B200REC_IO_HASH_STD;
  const Partition pt = partition_lines(text, len, false, n_threads, [&](const Line& ln) {
    if (!filter) return true;
    int tabs = 0;
    for (const char* q = ln.p; q < ln.e; ++q) tabs += (*q == '\t');
    return tabs + 1 == kCols;
  });
  *n_out = 0;
  if (n_skipped_out) *n_skipped_out = pt.total_skipped;
  if (int64_t(pt.total) > cap)
    return fail(B200REC_IO_ERR_CAPACITY, "%zu samples but cap = %lld", pt.total, (long long)cap);
  int rc = parallel_lines(pt, [&](int, const std::vector<Line>& lines, size_t base, Status& st) {
    std::string salted;
    for (size_t i = 0; i < lines.size() && st.code == B200REC_IO_OK; ++i) {
      const int64_t n = int64_t(base + i);
      const char* p = lines[i].p;
      const char* const e = lines[i].e;
      for (int col = 0; col < kCols; ++col) {
        if (p > e) { st.set(B200REC_IO_ERR_PARSE, n, "fewer than 40 columns in", std::string_view(lines[i].p, size_t(e - lines[i].p))); break; }
        const char* tab = static_cast<const char*>(memchr(p, '\t', size_t(e - p)));
        const char* te = tab ? tab : e;
        std::string_view tok(p, size_t(te - p));
        p = te + 1;
        if (col == 0) {
          if (!parse_i64(tok, &label[n])) { st.set(B200REC_IO_ERR_PARSE, n, "bad label", tok); break; }
        } else if (col <= kDense) {
          double v = 0.0;
          if (!tok.empty()) {
            if (!parse_f64(tok, &v)) { st.set(B200REC_IO_ERR_PARSE, n, "bad number", tok); break; }
            v = (v - cont_min[col - 1]) / cont_diff[col - 1];
          }
          dense[n * kDense + (col - 1)] = float(v);
        } else {
          uint64_t h;
          if (hash_kind == B200REC_IO_HASH_STD) {
            h = murmur64a(tok.data(), tok.size());
          } else {
            salted.assign(std::to_string(col));
            salted.append(tok.data(), tok.size());
            h = xxh32(salted.data(), salted.size(), 0);
          }
          ids[n * kSparse + (col - 1 - kDense)] = int64_t(h % uint64_t(hash_dim));
        }
      }
    }
  });
  if (rc) return rc;
  *n_out = int64_t(pt.total);
  return B200REC_IO_OK;
}

int b200rec_io_parse_din(const char* text, size_t len, int64_t* hist_items, int64_t* hist_cats,
                         int64_t* offsets, int64_t* target_item, int64_t* target_cat, float* label,
                         int64_t cap, int64_t keys_cap, int64_t* n_out, int64_t* n_keys_out,
                         int64_t* n_skipped_out, int n_threads) {
  if ((!text && len) || !n_out || !n_keys_out || !offsets || !target_item || !target_cat || !label ||
      (keys_cap > 0 && (!hist_items || !hist_cats)))
    return fail(B200REC_IO_ERR_ARG, "null argument");
  const Partition pt = partition_lines(text, len, true, n_threads, [](const Line& ln) {
    int semis = 0;
    for (const char* q = ln.p; q < ln.e; ++q) semis += (*q == ';');
    return semis >= 4;
  });
  *n_out = *n_keys_out = 0;
  if (n_skipped_out) *n_skipped_out = pt.total_skipped;
  if (int64_t(pt.total) > cap)
    return fail(B200REC_IO_ERR_CAPACITY, "%zu samples but cap = %lld", pt.total, (long long)cap);
  struct Local { std::vector<int64_t> items, cats; std::vector<int32_t> lens; };
  std::vector<Local> local((size_t)pt.T);
  int rc = parallel_lines(pt, [&](int t, const std::vector<Line>& lines, size_t base, Status& st) {
    Local& L = local[size_t(t)];
    L.lens.reserve(lines.size());
    for (size_t i = 0; i < lines.size() && st.code == B200REC_IO_OK; ++i) {
      const int64_t n = int64_t(base + i);
      const char* p = lines[i].p;
      const char* const e = lines[i].e;
      std::string_view field[5];
      for (int f = 0; f < 5; ++f) {      // the first five `;`-separated fields; the rest is ignored
        const char* semi = static_cast<const char*>(memchr(p, ';', size_t(e - p)));
        const char* fe = (semi && f < 4) ? semi : (semi ? semi : e);
        field[f] = std::string_view(p, size_t(fe - p));
        p = semi ? semi + 1 : e;
      }
      auto id_list = [&](std::string_view s, std::vector<int64_t>& out) -> int {  // str.split()
        int cnt = 0;
        const char* q = s.data();
        const char* const qe = q + s.size();
        while (q < qe) {
          while (q < qe && is_space(*q)) ++q;
          if (q >= qe) break;
          const char* a = q;
          while (q < qe && !is_space(*q)) ++q;
          int64_t v;
          if (!parse_i64(std::string_view(a, size_t(q - a)), &v)) {
            st.set(B200REC_IO_ERR_PARSE, n, "bad id", std::string_view(a, size_t(q - a)));
            return -1;
          }
          out.push_back(v);
          ++cnt;
        }
        return cnt;
      };
      const int ni = id_list(field[0], L.items);
      if (ni < 0) break;
      const int nc = id_list(field[1], L.cats);
      if (nc < 0) break;
      if (ni != nc) { st.set(B200REC_IO_ERR_RAGGED, n, "item and category histories differ in length:", field[1]); break; }
      auto strip = [](std::string_view s) {
        while (!s.empty() && is_space(s.front())) s.remove_prefix(1);
        while (!s.empty() && is_space(s.back())) s.remove_suffix(1);
        return s;
      };
      double lab;
      if (!parse_i64(strip(field[2]), &target_item[n])) { st.set(B200REC_IO_ERR_PARSE, n, "bad target item", field[2]); break; }
      if (!parse_i64(strip(field[3]), &target_cat[n])) { st.set(B200REC_IO_ERR_PARSE, n, "bad target category", field[3]); break; }
      if (!parse_f64(strip(field[4]), &lab)) { st.set(B200REC_IO_ERR_PARSE, n, "bad label", field[4]); break; }
      label[n] = float(lab);
      L.lens.push_back(int32_t(ni));
    }
  });
  if (rc) return rc;
  int64_t total = 0;
  for (auto& L : local) total += int64_t(L.items.size());
  if (total > keys_cap)
    return fail(B200REC_IO_ERR_CAPACITY, "%lld ids but keys_cap = %lld", (long long)total, (long long)keys_cap);
  int64_t at = 0;
  for (int t = 0; t < pt.T; ++t) {
    Local& L = local[size_t(t)];
    int64_t* off = offsets + int64_t(pt.base[size_t(t)]);
    for (size_t i = 0; i < L.lens.size(); ++i) { off[i] = at; at += L.lens[i]; }
    if (!L.items.empty()) {
      memcpy(hist_items + (at - int64_t(L.items.size())), L.items.data(), L.items.size() * sizeof(int64_t));
      memcpy(hist_cats + (at - int64_t(L.cats.size())), L.cats.data(), L.cats.size() * sizeof(int64_t));
    }
  }
  offsets[int64_t(pt.total)] = at;
  *n_out = int64_t(pt.total);
  *n_keys_out = total;
  return B200REC_IO_OK;
}

}  // extern "C"

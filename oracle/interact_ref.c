/* Plain-C restatements of two more pieces of the path (TEST INFRASTRUCTURE — see oracle/__init__.py):
 * the DLRM dot interaction and the uint64 feasign fold.  Independent of oracle/nets.py /
 * oracle/readers.py (no torch, no numpy): a second CPU restatement the first ones are checked
 * against (tests/test_oracle_c.py).  Build: make -C oracle -> oracle/_build/libinteract_ref.so
 *
 * dot_interact_ref follows /root/reference/models/rank/dlrm/net.py:97-115 statement by statement:
 *   T[b]    = [e_1 .. e_F, x]                 (:97-100; x, the bottom-MLP output, is the last row)
 *   Z       = T T^T                            (:103)
 *   Zflat   = triu(Z, 1) + tril(MIN_FLOAT, -1 if self_interaction else 0), then masked_select of
 *             everything > MIN_FLOAT, row-major (:105-113): the strict upper triangle — and with
 *             self_interaction the diagonal positions too, which hold 0 because triu(Z,1) zeroed them
 *   R       = concat([x, Zflat])               (:115)
 * and its analytic backward: dT[i] = sum_{j != i} dZ(i,j) T[j]  (+ dR[:d] into the x row).
 */
#include <stdint.h>
#include <stddef.h>

static int pair_index(int N, int self, int i, int j) { /* i < j, or i <= j with self */
  int start = 0;
  for (int k = 0; k < i; ++k) start += self ? N - k : N - 1 - k;
  return start + (self ? j - i : j - i - 1);
}

void dot_interact_ref_fwd(const float* T, double* R, int64_t B, int N, int d, int self) {
  const int P = self ? N * (N + 1) / 2 : N * (N - 1) / 2;
  for (int64_t b = 0; b < B; ++b) {
    const float* t = T + (size_t)b * N * d;
    double* r = R + (size_t)b * (d + P);
    for (int c = 0; c < d; ++c) r[c] = t[(N - 1) * d + c];
    int p = 0;
    for (int i = 0; i < N; ++i) {
      for (int j = self ? i : i + 1; j < N; ++j) {
        double acc = 0.0;
        if (j != i)
          for (int c = 0; c < d; ++c) acc += (double)t[i * d + c] * (double)t[j * d + c];
        r[d + p++] = acc;
      }
    }
  }
}

void dot_interact_ref_bwd(const float* T, const double* dR, double* dT, int64_t B, int N, int d,
                          int self) {
  const int P = self ? N * (N + 1) / 2 : N * (N - 1) / 2;
  for (int64_t b = 0; b < B; ++b) {
    const float* t = T + (size_t)b * N * d;
    const double* g = dR + (size_t)b * (d + P);
    double* o = dT + (size_t)b * N * d;
    for (int i = 0; i < N; ++i) {
      for (int c = 0; c < d; ++c) {
        double acc = (i == N - 1) ? g[c] : 0.0;
        for (int j = 0; j < N; ++j) {
          if (j == i) continue;
          const int p = i < j ? pair_index(N, self, i, j) : pair_index(N, self, j, i);
          acc += g[d + p] * (double)t[j * d + c];
        }
        o[i * d + c] = acc;
      }
    }
  }
}

/* include/b200rec.h: b200rec_hash_keys — splitmix64 finaliser of key ^ (slot+1)*golden ratio. */
void hash_keys_ref(const uint64_t* keys, const int32_t* slot_of_key, int64_t n, uint64_t V,
                   int reserve_zero, int64_t* rows) {
  for (int64_t i = 0; i < n; ++i) {
    uint64_t z = keys[i];
    if (slot_of_key) z ^= (uint64_t)(slot_of_key[i] + 1) * 0x9E3779B97F4A7C15ULL;
    z ^= z >> 30;
    z *= 0xBF58476D1CE4E5B9ULL;
    z ^= z >> 27;
    z *= 0x94D049BB133111EBULL;
    z ^= z >> 31;
    rows[i] = reserve_zero ? (keys[i] == 0 ? 0 : (int64_t)(1 + z % (V - 1))) : (int64_t)(z % V);
  }
}

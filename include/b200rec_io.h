/* b200rec_io.h — C ABI of the HOST-side input pipeline.
 *
 * libb200rec_io.so is plain C++17 (no CUDA, no torch): it turns the reference's text formats into
 * the packed arrays the device path consumes — ONE `[n,F]` int64 id matrix, ONE `[n,Dn]` float
 * matrix and ONE `[n]` int64 label vector per batch (the reference feeds 28 separate arrays per
 * sample, models/rank/deepfm/dygraph_model.py:41-50).  Output buffers are the caller's (pinned
 * host memory when the next hop is an H2D copy); nothing is allocated behind the caller's back
 * except per-call scratch that is freed before returning.
 *
 * Formats and the reference code each entry point replaces:
 *   slot text   `slot:value slot:value ...`    doc/custom_reader.md:5-24,
 *                                               models/rank/deepfm/criteo_reader.py:61-103
 *   multislot   `<count> v.. <count> v.. ...`   the QueueDataset/InMemoryDataset wire format,
 *                                               tools/dataset/README.MD:24-33, output of
 *                                               tools/dataset/parser.cpp:54-75
 *   criteo tsv  label \t 13 ints \t 26 tokens   tools/dataset/parser.cpp:37-77 (hash kind 0),
 *                                               models/rank/dnn/benchmark_reader.py:39-56 (kind 1)
 *
 * Conventions: every function returns B200REC_IO_OK or a negative code, the message is in
 * b200rec_io_last_error() (thread-local); `n_threads <= 0` means "all hardware threads"; results do
 * not depend on the thread count.  Blank lines are skipped.  A malformed token is an ERROR that
 * names the line — the Python readers raise at the same places.
 */
#ifndef B200REC_IO_H_
#define B200REC_IO_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200REC_IO_ABI_VERSION 1

#define B200REC_IO_OK 0
#define B200REC_IO_ERR_ARG (-1)      /* null pointer, bad size */
#define B200REC_IO_ERR_PARSE (-2)    /* malformed token / line */
#define B200REC_IO_ERR_CAPACITY (-3) /* output buffer too small */
#define B200REC_IO_ERR_RAGGED (-4)   /* a slot's length differs from the fixed-length schema */

int b200rec_io_abi_version(void);
const char* b200rec_io_last_error(void);

/* Non-empty lines in text[0,len): an upper bound of the samples any parser below emits, i.e. the
 * `cap` a caller needs (lines holding only blanks are counted here but skipped by the parsers). */
int b200rec_io_count_lines(const char* text, size_t len, int64_t* n_lines);

/* ---- slot text, fixed length (criteo_reader.py:61-103) -------------------------------------------
 * Tokens are separated by single spaces; a token is `slot:value`; tokens whose slot is not in the
 * schema are ignored (criteo_reader.py:72-73).  `label_slot` (may be NULL) is an integer slot written
 * to label[n]; sparse_slots[f] goes to ids[n, f]; dense_slot's values fill dense[n, 0..dense_dim).
 * A slot missing on a line is filled with 0 / 0.0 (criteo_reader.py:80-89).  More than one value in
 * a sparse/label slot, or a dense slot whose length is neither 0 nor dense_dim, is
 * B200REC_IO_ERR_RAGGED (use the _lod entry point for multi-hot slots). */
int b200rec_io_parse_slot_text(const char* text, size_t len, const char* label_slot,
                               const char* const* sparse_slots, int n_sparse,
                               const char* dense_slot, int dense_dim, int64_t* label,
                               int64_t* ids, float* dense, int64_t cap, int64_t* n_out,
                               int n_threads);

/* Same, with the two per-model variations of the reference's readers selected by `flags`:
 *   B200REC_IO_DENSE_LOG1P        dense value = log(v + 1), evaluated in double before narrowing to
 *                                 float32 (models/rank/dcn_v2/reader.py:63-64)
 *   B200REC_IO_SKIP_EMPTY_SPARSE  a sparse token with an empty value (`7:`) is ignored instead of
 *                                 being an error (dcn_v2/reader.py:55-57) */
#define B200REC_IO_DENSE_LOG1P 1
#define B200REC_IO_SKIP_EMPTY_SPARSE 2
int b200rec_io_parse_slot_text_ex(const char* text, size_t len, const char* label_slot,
                                  const char* const* sparse_slots, int n_sparse,
                                  const char* dense_slot, int dense_dim, int flags, int64_t* label,
                                  int64_t* ids, float* dense, int64_t cap, int64_t* n_out,
                                  int n_threads);

/* ---- slot text, variable length (doc/custom_reader.md:15-24) -------------------------------------
 * Same tokens; every (sample, sparse slot) pair is a bag of >= 1 keys: bag b = n*n_sparse + f holds
 * keys[offsets[b] .. offsets[b+1]); a missing slot is the bag {0} like the padded reader.
 * offsets has cap*n_sparse + 1 entries; keys has keys_cap.  This is the (keys, offsets) pair
 * b200rec_gather_pool_sum consumes. */
int b200rec_io_parse_slot_text_lod(const char* text, size_t len, const char* label_slot,
                                   const char* const* sparse_slots, int n_sparse,
                                   const char* dense_slot, int dense_dim, int64_t* label,
                                   int64_t* keys, int64_t* offsets, float* dense, int64_t cap,
                                   int64_t keys_cap, int64_t* n_out, int64_t* n_keys_out,
                                   int n_threads);

/* ---- multislot wire format ------------------------------------------------------------------------
 * Each line holds n_slots groups `<count> v1 .. v_count`, tokens separated by blanks.
 * slot_is_float[s] != 0: values are floats and go to fvals / foffsets (bags n*n_float + j);
 * otherwise they are uint64 feasigns and go to keys / koffsets (bags n*n_int + i).  Counts of 0 are
 * rejected, as Paddle's MultiSlot feed does. */
int b200rec_io_parse_multislot(const char* text, size_t len, const int* slot_is_float, int n_slots,
                               uint64_t* keys, int64_t* koffsets, int64_t keys_cap, float* fvals,
                               int64_t* foffsets, int64_t fvals_cap, int64_t cap, int64_t* n_out,
                               int64_t* n_keys_out, int64_t* n_fvals_out, int n_threads);

/* ---- raw Criteo TSV -------------------------------------------------------------------------------
 * Column 0 label, 1..13 integer features, 14..39 categorical tokens, tab separated.
 * dense[n, j] = column empty ? 0 : (value - cont_min[j]) / cont_diff[j], computed in double
 * (parser.cpp:55-63, benchmark_reader.py:43-49); NULL cont_min/cont_diff = the constants both
 * reference readers hard-code.  ids[n, f] = hash(column 14+f) mod hash_dim with
 *   hash_kind 0: 64-bit libstdc++ std::hash<std::string> of the token   (parser.cpp:68)
 *   hash_kind 1: xxHash32(seed 0) of str(column index) + token           (benchmark_reader.py:50-52)
 * Lines that do not have exactly 40 columns are skipped and counted (parser.cpp:49-51); with
 * hash_kind 1 a short line is an error (the Python reader raises) and extra columns are ignored. */
#define B200REC_IO_HASH_STD 0
#define B200REC_IO_HASH_XXH32 1
int b200rec_io_parse_criteo_tsv(const char* text, size_t len, int hash_kind, int64_t hash_dim,
                                const double* cont_min, const double* cont_diff, int64_t* label,
                                int64_t* ids, float* dense, int64_t cap, int64_t* n_out,
                                int64_t* n_skipped_out, int n_threads);

/* ---- DIN behaviour logs (models/rank/din/dinReader.py:45-57) ---------------------------------------
 * Line: `hist_items;hist_cats;target_item;target_cat;label` — the two histories are blank-separated
 * id lists of equal length, label is a float.  Lines with fewer than 5 `;` fields are skipped
 * (dinReader.py:52-53) and counted.  Output is LoD: sample n owns hist_items / hist_cats
 * [offsets[n], offsets[n+1]).  Grouping by length, padding and the mask (dinReader.py:61-144) are
 * batch-level and live in the caller (paddlerec_b200/dataio.py: DinBatchReader). */
int b200rec_io_parse_din(const char* text, size_t len, int64_t* hist_items, int64_t* hist_cats,
                         int64_t* offsets, int64_t* target_item, int64_t* target_cat, float* label,
                         int64_t cap, int64_t keys_cap, int64_t* n_out, int64_t* n_keys_out,
                         int64_t* n_skipped_out, int n_threads);

/* The two string hashes, exposed for tests and for callers that hash elsewhere. */
uint64_t b200rec_io_hash_std_string(const char* s, size_t len);
uint32_t b200rec_io_xxh32(const char* s, size_t len, uint32_t seed);

#ifdef __cplusplus
}
#endif
#endif /* B200REC_IO_H_ */

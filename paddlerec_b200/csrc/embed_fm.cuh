// K1 / K2: DeepFM's fused multi-slot gather + FM interaction, forward and backward.
//
// Replaces FM.forward of the reference (models/rank/deepfm/net.py:105-139) and its autograd
// (SURVEY.md §8a rows A1-A3, A7).  HBM-bound: per sample the forward moves
//   F*8 (ids) + Dn*4 (dense) + F*4D (rows) + F*4 (first-order) + (F+Dn)*4D (feat) + 8 (y1,y2)
// algorithmic bytes (4532 B at F=26, Dn=13, D=16) and ~3 flops per byte, so the design goal is
// to keep as many independent 128-bit row loads in flight as the LSU allows and to touch
// every byte once:
//   * ids/dense of a tile of samples are staged to shared memory with coalesced loads;
//   * TPR lanes share one sample, each lane owns VEC consecutive floats of every row, so one
//     warp instruction fetches 32/TPR random rows with full 32 B-sector use (D=16: 8 rows);
//   * the running sum S and sum of squares Q stay in registers over all F+Dn fields; the
//     first-order lookups are spread over the TPR lanes; a xor-shuffle finishes y1/y2;
//   * feat is written once with evict-first stores (it is larger than L2 at B=65536).
#pragma once

#include <stdlib.h>

#include "common.cuh"
#include "segreduce.cuh"

namespace b200rec {

// Tunables (env B200REC_K1_UNROLL = 8|13|26, B200REC_K1_CACHE = 0|1, B200REC_K1_CTAS = persistent
// CTAs per SM), read once per process:
// U = independent row loads in flight per lane; CACHE = let row loads allocate in L1 (useful for
// the fused slot layout, where the first-order weight sits in the same 128-byte line as the row).
struct K1Config {
  int unroll;
  int cache_rows;
  int ctas_per_sm;
};
static K1Config k1_config() {
  static K1Config cfg = [] {
    K1Config c{13, 0, 8};
    if (const char* s = getenv("B200REC_K1_UNROLL")) c.unroll = atoi(s);
    if (const char* s = getenv("B200REC_K1_CACHE")) c.cache_rows = atoi(s);
    if (const char* s = getenv("B200REC_K1_CTAS")) c.ctas_per_sm = atoi(s);
    if (c.ctas_per_sm < 1 || c.ctas_per_sm > 16) c.ctas_per_sm = 8;
    if (c.unroll != 8 && c.unroll != 13 && c.unroll != 26) c.unroll = 13;
    return c;
  }();
  return cfg;
}

template <int TPR>
struct FwdGeom {
  static constexpr int kThreads = 128;
  static constexpr int kSamples = kThreads / TPR;  // samples per tile (32 at D=16)
};

// Persistent CTAs: each loops over tiles of kSamples samples (static stride).  While tile t is
// processed, the ids/dense of tile t+gridDim.x stream into the other shared-memory buffer with
// cp.async, so the id round trip is off the critical path.  Per tile and lane: all row loads and
// first-order lookups of a batch of kFieldUnroll fields are issued (predicated PTX, no branches)
// before the first one is consumed.
template <int VEC, int TPR, int kFieldUnroll, bool CACHE>
__global__ void __launch_bounds__(FwdGeom<TPR>::kThreads)
embed_fm_fwd_kernel(const float* __restrict__ W, const float* __restrict__ W1,
                    const int64_t* __restrict__ ids, const float* __restrict__ dense,
                    const float* __restrict__ dense_w, const float* __restrict__ dense_w1,
                    float* __restrict__ feat, float* __restrict__ y1, float* __restrict__ y2,
                    float* __restrict__ S, int64_t B, int F, int Dn, int D, int64_t V,
                    int64_t pad, int64_t ldw, int64_t ldw1) {
  constexpr int kThreads = FwdGeom<TPR>::kThreads;
  constexpr int SPB = FwdGeom<TPR>::kSamples;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const size_t ids_elems = (size_t)SPB * F;
  const size_t dense_elems = (size_t)SPB * Dn;
  const size_t buf_bytes = (ids_elems * 8 + dense_elems * 4 + 15) / 16 * 16;

  const int64_t ntiles = (B + SPB - 1) / SPB;
  const int s = threadIdx.x / TPR;
  const int r = threadIdx.x % TPR;
  const bool lane_ok = r * VEC < D;
  const int N = F + Dn;

  auto stage = [&](int64_t tile, int buf) {
    int64_t* s_ids = reinterpret_cast<int64_t*>(smem_raw + (size_t)buf * buf_bytes);
    float* s_dense = reinterpret_cast<float*>(s_ids + ids_elems);
    const int64_t b0 = tile * SPB;
    const int nb = (int)min((int64_t)SPB, B - b0);
    for (int i = threadIdx.x; i < nb * F; i += kThreads) cp_async_8(s_ids + i, ids + b0 * F + i);
    for (int i = threadIdx.x; i < nb * Dn; i += kThreads)
      cp_async_4(s_dense + i, dense + b0 * Dn + i);
  };

  int64_t tile = blockIdx.x;
  if (tile < ntiles) stage(tile, 0);
  cp_async_commit();
  for (int it = 0; tile < ntiles; tile += gridDim.x, ++it) {
    const int64_t next = tile + gridDim.x;
    if (next < ntiles) stage(next, (it + 1) & 1);
    cp_async_commit();
    cp_async_wait<1>();  // everything but the newest group (the next tile) has landed
    __syncthreads();

    const int64_t* s_ids = reinterpret_cast<const int64_t*>(smem_raw + (size_t)(it & 1) * buf_bytes);
    const float* s_dense = reinterpret_cast<const float*>(s_ids + ids_elems);
    const int64_t b0 = tile * SPB;
    const int nb = (int)min((int64_t)SPB, B - b0);
    const bool sample_ok = s < nb;
    const int64_t b = b0 + s;

    Vec<VEC> Ssum = vzero<VEC>();
    Vec<VEC> Q = vzero<VEC>();
    float first = 0.f;

    if (sample_ok) {
      const int64_t* my_ids = s_ids + (size_t)s * F;
      float* feat_row = feat + (size_t)b * N * D + r * VEC;
      for (int f0 = 0; f0 < F; f0 += kFieldUnroll) {
        Vec<VEC> e[kFieldUnroll];
        float w1v[kFieldUnroll];
        // phase 1: issue every load of the batch before anything consumes one
#pragma unroll
        for (int j = 0; j < kFieldUnroll; ++j) {
          const int f = f0 + j;
          const int64_t id = (f < F) ? my_ids[f] : pad;
          const bool in_range = (uint64_t)id < (uint64_t)V;
          const bool live = (f < F) && in_range && id != pad;
          const size_t row = live ? (size_t)id : 0;
          e[j] = ld_row_pred<VEC, CACHE>(W + row * ldw + r * VEC, live && lane_ok);
          w1v[j] = ld_row_pred<1, true>(W1 + row * ldw1, live && ((f & (TPR - 1)) == r)).v[0];
          if ((f < F) && !in_range && r == 0) atomicAdd(&g_oob_count, 1ull);
        }
        // phase 2: consume
#pragma unroll
        for (int j = 0; j < kFieldUnroll; ++j) {
          const int f = f0 + j;
          if (f < F) {
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
              Ssum.v[k] += e[j].v[k];
              Q.v[k] = fmaf(e[j].v[k], e[j].v[k], Q.v[k]);
            }
            if (lane_ok) st_stream<VEC>(feat_row + (size_t)f * D, e[j]);
          }
        }
#pragma unroll
        for (int j = 0; j < kFieldUnroll; ++j) first += w1v[j];
      }
      const float* my_dense = s_dense + (size_t)s * Dn;
      for (int j = 0; j < Dn; ++j) {
        const float x = my_dense[j];
        Vec<VEC> e = vzero<VEC>();
        if (lane_ok) {
          const Vec<VEC> w = ld_cached<VEC>(dense_w + (size_t)j * D + r * VEC);
#pragma unroll
          for (int k = 0; k < VEC; ++k) e.v[k] = x * w.v[k];
          st_stream<VEC>(feat_row + (size_t)(F + j) * D, e);
        }
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          Ssum.v[k] += e.v[k];
          Q.v[k] = fmaf(e.v[k], e.v[k], Q.v[k]);
        }
        if ((j & (TPR - 1)) == r) first = fmaf(x, __ldg(dense_w1 + j), first);
      }
    }

    float t = 0.f;
#pragma unroll
    for (int k = 0; k < VEC; ++k) t += Ssum.v[k] * Ssum.v[k] - Q.v[k];
    t = group_sum<TPR>(t);
    first = group_sum<TPR>(first);
    if (sample_ok) {
      if (r == 0) {
        y1[b] = first;
        y2[b] = 0.5f * t;
      }
      if (S != nullptr && lane_ok) st_plain<VEC>(S + (size_t)b * D + r * VEC, Ssum);
    }
    __syncthreads();  // buffer (it&1) is free for the prefetch issued two iterations later
  }
  cp_async_wait<0>();
}

struct K1Args {
  const float* W; const float* W1; const int64_t* ids; const float* dense; const float* dense_w;
  const float* dense_w1; float* feat; float* y1; float* y2; float* S;
  int64_t B; int F, Dn, D; int64_t V, pad, ldw, ldw1;
};

template <int VEC, int TPR, int U, bool C>
static int k1_launch(const K1Args& a, int64_t grid, size_t smem, cudaStream_t st) {
  auto kern = embed_fm_fwd_kernel<VEC, TPR, U, C>;
  if (smem > 48 * 1024)
    B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(unsigned)grid, FwdGeom<TPR>::kThreads, smem, st>>>(
      a.W, a.W1, a.ids, a.dense, a.dense_w, a.dense_w1, a.feat, a.y1, a.y2, a.S, a.B, a.F, a.Dn, a.D,
      a.V, a.pad, a.ldw, a.ldw1);
  return B200REC_OK;
}

static int launch_embed_fm_fwd(const float* W, const float* W1, const int64_t* ids,
                               const float* dense, const float* dense_w, const float* dense_w1,
                               float* feat, float* y1, float* y2, float* S, int64_t B, int F,
                               int Dn, int D, int64_t V, int64_t pad, int64_t ldw, int64_t ldw1,
                               cudaStream_t st) {
  RowShape rs;
  B200_REQUIRE(pick_row_shape(D, &rs), "embed_fm_fwd: unsupported D=%d", D);
  B200_REQUIRE(ldw >= D && ldw1 >= 1 && ldw % rs.vec == 0,
               "embed_fm_fwd: bad row strides ldw=%lld ldw1=%lld", (long long)ldw, (long long)ldw1);
  if (rs.vec == 4)
    B200_REQUIRE(aligned16(W) && aligned16(feat) && aligned16(dense_w) && (!S || aligned16(S)),
                 "embed_fm_fwd: W/feat/dense_w/S must be 16-byte aligned for D%%4==0");
  if (rs.vec == 2)
    B200_REQUIRE(aligned8(W) && aligned8(feat) && aligned8(dense_w) && (!S || aligned8(S)),
                 "embed_fm_fwd: W/feat/dense_w/S must be 8-byte aligned for D%%2==0");
  if (B == 0) return B200REC_OK;
  B200_DISPATCH_ROW_SHAPE(rs, {
    constexpr int SPB = FwdGeom<TPR>::kSamples;
    const size_t buf = ((size_t)SPB * F * sizeof(int64_t) + (size_t)SPB * Dn * sizeof(float) + 15) /
                       16 * 16;
    const size_t smem = 2 * buf;
    B200_REQUIRE(smem <= 200 * 1024, "embed_fm_fwd: F=%d Dn=%d tile does not fit shared memory", F,
                 Dn);
    const K1Config cfg = k1_config();
    // same 128-byte line as the row (fused slot layout)?  then let the row load allocate in L1
    const bool same_line = (W1 >= W && W1 < W + ldw && ldw1 == ldw);
    const bool cache = cfg.cache_rows < 0 ? same_line : cfg.cache_rows != 0;
    const int64_t ntiles = (B + SPB - 1) / SPB;
    const int64_t grid = min(ntiles, (int64_t)sm_count() * cfg.ctas_per_sm);
    K1Args a{W, W1, ids, dense, dense_w, dense_w1, feat, y1, y2, S, B, F, Dn, D, V, pad, ldw, ldw1};
    int rc;
    if (cfg.unroll == 8)
      rc = cache ? k1_launch<VEC, TPR, 8, true>(a, grid, smem, st)
                 : k1_launch<VEC, TPR, 8, false>(a, grid, smem, st);
    else if (cfg.unroll == 26)
      rc = cache ? k1_launch<VEC, TPR, 26, true>(a, grid, smem, st)
                 : k1_launch<VEC, TPR, 26, false>(a, grid, smem, st);
    else
      rc = cache ? k1_launch<VEC, TPR, 13, true>(a, grid, smem, st)
                 : k1_launch<VEC, TPR, 13, false>(a, grid, smem, st);
    if (rc != B200REC_OK) return rc;
  });
  B200_LAUNCH_CHECK();
  return B200REC_OK;
}

// ------------------------------------------------------------------------------------------
// K2a: sparse part of the backward.  One group of TPR lanes per DISTINCT id (segment of the
// sorted positions).  dfeat is never materialised for the sparse fields:
//   dfeat[p] = gy2[b]*(S[b]-feat[p]) + dfeat_dnn[p]      (net.py:123-137 differentiated)
// and summed in the fixed (stable-sort) order of the segment => deterministic.
constexpr int kBwdThreads = 256;

struct FmRowContrib {
  const float* feat;
  const float* S;
  const float* dfeat_dnn;  // may be null
  const float* gy1;
  const float* gy2;
  int F, N, D;
  template <int VEC>
  __device__ __forceinline__ void add(int p, int r, bool lane_ok, Vec<VEC>& acc,
                                      float& acc1) const {
    const int b = p / F;
    const int f = p - b * F;
    const float g2 = __ldg(gy2 + b);
    acc1 += __ldg(gy1 + b);
    if (lane_ok) {
      const size_t off = ((size_t)b * N + f) * D + r * VEC;
      const Vec<VEC> fe = ld_row<VEC>(feat + off);
      const Vec<VEC> sv = ld_cached<VEC>(S + (size_t)b * D + r * VEC);
      Vec<VEC> dd = vzero<VEC>();
      if (dfeat_dnn != nullptr) dd = ld_row<VEC>(dfeat_dnn + off);
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc.v[k] += fmaf(g2, sv.v[k] - fe.v[k], dd.v[k]);
    }
  }
};

// K2b: dense-feature part: ddense_w[j,:] = sum_b dense[b,j]*dfeat[b,F+j,:],
// ddense_w1[j] = sum_b gy1[b]*dense[b,j].  Persistent CTAs accumulate in registers over a
// strided set of samples, reduce once, and write one partial per CTA; a second tiny kernel adds
// the partials in a fixed order (deterministic, no float atomics).
constexpr int kDenseChunk = 8;  // dense features held in registers at once (Dn=13 -> 2 passes)

template <int VEC, int TPR>
__global__ void __launch_bounds__(kBwdThreads, 2)
embed_fm_bwd_dense_kernel(const float* __restrict__ feat, const float* __restrict__ S,
                          const float* __restrict__ dfeat_dnn, const float* __restrict__ gy1,
                          const float* __restrict__ gy2, const float* __restrict__ dense,
                          float* __restrict__ partials /*[grid, Dn*D + Dn]*/, int64_t B, int F,
                          int Dn, int D) {
  constexpr int SPB = kBwdThreads / TPR;
  constexpr int kWarps = kBwdThreads / 32;
  __shared__ float s_red[kWarps][32 * VEC];
  __shared__ float s_red1[kWarps];
  const int s = threadIdx.x / TPR;
  const int r = threadIdx.x % TPR;
  const int warp = threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  const bool lane_ok = r * VEC < D;
  const int N = F + Dn;
  float* my_partial = partials + (size_t)blockIdx.x * ((size_t)Dn * D + Dn);

  for (int j0 = 0; j0 < Dn; j0 += kDenseChunk) {
    Vec<VEC> acc[kDenseChunk];
    float acc1[kDenseChunk];
#pragma unroll
    for (int j = 0; j < kDenseChunk; ++j) {
      acc[j] = vzero<VEC>();
      acc1[j] = 0.f;
    }
    for (int64_t b = (int64_t)blockIdx.x * SPB + s; b < B; b += (int64_t)gridDim.x * SPB) {
      const float g2 = __ldg(gy2 + b);
      const float g1 = __ldg(gy1 + b);
      Vec<VEC> sv = vzero<VEC>();
      if (lane_ok) sv = ld_cached<VEC>(S + (size_t)b * D + r * VEC);
#pragma unroll
      for (int j = 0; j < kDenseChunk; ++j) {
        if (j0 + j < Dn) {
          const float x = __ldg(dense + (size_t)b * Dn + j0 + j);
          if (lane_ok) {
            const size_t off = ((size_t)b * N + F + j0 + j) * D + r * VEC;
            const Vec<VEC> fe = ld_row<VEC>(feat + off);
            Vec<VEC> dd = vzero<VEC>();
            if (dfeat_dnn != nullptr) dd = ld_row<VEC>(dfeat_dnn + off);
#pragma unroll
            for (int k = 0; k < VEC; ++k)
              acc[j].v[k] = fmaf(x, fmaf(g2, sv.v[k] - fe.v[k], dd.v[k]), acc[j].v[k]);
          }
          if (r == 0) acc1[j] = fmaf(g1, x, acc1[j]);
        }
      }
    }
    // reduce over the samples of the CTA, one dense feature at a time
#pragma unroll
    for (int j = 0; j < kDenseChunk; ++j) {
      if (j0 + j < Dn) {  // uniform across the CTA
        Vec<VEC> a = acc[j];
        float a1 = acc1[j];
#pragma unroll
        for (int o = TPR; o < 32; o <<= 1) {
#pragma unroll
          for (int k = 0; k < VEC; ++k) a.v[k] += __shfl_xor_sync(0xffffffffu, a.v[k], o);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        if (lane < TPR) {
#pragma unroll
          for (int k = 0; k < VEC; ++k) s_red[warp][lane * VEC + k] = a.v[k];
        }
        if (lane == 0) s_red1[warp] = a1;
        __syncthreads();
        if (threadIdx.x < D) {
          float t = 0.f;
#pragma unroll
          for (int w = 0; w < kWarps; ++w) t += s_red[w][threadIdx.x];
          my_partial[(size_t)(j0 + j) * D + threadIdx.x] = t;
        }
        if (threadIdx.x == 0) {
          float t = 0.f;
#pragma unroll
          for (int w = 0; w < kWarps; ++w) t += s_red1[w];
          my_partial[(size_t)Dn * D + j0 + j] = t;
        }
        __syncthreads();
      }
    }
  }
}

// out[k] = sum_g partials[g][k] (deterministic: fixed strided order + fixed shared-memory tree).
// `len` is small (<= a few thousand) but G can be hundreds: 32 columns x 8 g-lanes per block so
// the G-long sums are not one serial dependent chain per column.
constexpr int kRedCols = 32;
constexpr int kRedLanes = 8;
__global__ void __launch_bounds__(kRedCols * kRedLanes)
reduce_partials_kernel(const float* __restrict__ partials, int G, int len,
                       float* __restrict__ out0, int len0, float* __restrict__ out1) {
  __shared__ float s[kRedLanes][kRedCols + 1];
  const int cx = threadIdx.x % kRedCols, gy = threadIdx.x / kRedCols;
  const int k = blockIdx.x * kRedCols + cx;
  float t = 0.f;
  if (k < len)
    for (int g = gy; g < G; g += kRedLanes) t += partials[(size_t)g * len + k];
  s[gy][cx] = t;
  __syncthreads();
  if (gy == 0 && k < len) {
    float r = 0.f;
#pragma unroll
    for (int y = 0; y < kRedLanes; ++y) r += s[y][cx];
    if (k < len0)
      out0[k] = r;
    else
      out1[k - len0] = r;
  }
}
static inline unsigned reduce_partials_grid(int len) { return (unsigned)((len + kRedCols - 1) / kRedCols); }
constexpr int kRedThreads = kRedCols * kRedLanes;

static int bwd_dense_grid() { return sm_count() * 2; }

static int launch_embed_fm_bwd(const float* feat, const float* S, const float* dfeat_dnn,
                               const float* gy1, const float* gy2, const float* dense,
                               const int32_t* seg_offsets, const int32_t* sorted_pos,
                               const int32_t* num_unique, float* dW_rows, float* dW1_rows,
                               SegOut so, float* ddense_w, float* ddense_w1, int64_t B, int F,
                               int Dn, int D, void* ws, size_t ws_bytes, cudaStream_t st) {
  RowShape rs;
  B200_REQUIRE(pick_row_shape(D, &rs), "embed_fm_bwd: unsupported D=%d", D);
  const int align = rs.vec * 4;
  auto ok = [&](const void* p) { return (reinterpret_cast<uintptr_t>(p) % align) == 0; };
  B200_REQUIRE(ok(feat) && ok(S) && ok(dfeat_dnn) && ok(dW_rows),
               "embed_fm_bwd: feat/S/dfeat_dnn/dW_rows must be %d-byte aligned", align);
  B200_REQUIRE(so.ld_rows >= D && so.ld_rows % rs.vec == 0 && so.ld_rows1 >= 1 && so.zero_pad >= 0,
               "embed_fm_bwd: bad output strides");
  B200_REQUIRE(B * F < (int64_t)INT32_MAX, "embed_fm_bwd: B*F must fit int32");
  const int G = bwd_dense_grid();
  const size_t dense_bytes = align_up((size_t)G * ((size_t)Dn * D + Dn) * sizeof(float), 256);
  const size_t need = dense_bytes + seg_workspace_bytes(B * F, D);
  if (ws_bytes < need) {
    set_error("embed_fm_bwd: workspace %zu < %zu bytes", ws_bytes, need);
    return B200REC_ERR_WORKSPACE;
  }
  if (B == 0) {
    if (Dn > 0) {
      B200_CUDA(cudaMemsetAsync(ddense_w, 0, (size_t)Dn * D * sizeof(float), st));
      B200_CUDA(cudaMemsetAsync(ddense_w1, 0, (size_t)Dn * sizeof(float), st));
    }
    return B200REC_OK;
  }
  const int N = F + Dn;
  int rc = B200REC_OK;
  B200_DISPATCH_ROW_SHAPE(rs, {
    const int64_t n = B * F;
    if (n > 0) {
      FmRowContrib contrib{feat, S, dfeat_dnn, gy1, gy2, F, N, D};
      rc = launch_seg_reduce<VEC, TPR, FmRowContrib>(seg_offsets, sorted_pos, num_unique, contrib,
                                                     dW_rows, dW1_rows, so, n, D,
                                                     static_cast<unsigned char*>(ws) + dense_bytes,
                                                     st);
    }
    if (rc == B200REC_OK && Dn > 0) {
      embed_fm_bwd_dense_kernel<VEC, TPR><<<G, kBwdThreads, 0, st>>>(
          feat, S, dfeat_dnn, gy1, gy2, dense, static_cast<float*>(ws), B, F, Dn, D);
    }
  });
  if (rc != B200REC_OK) return rc;
  B200_LAUNCH_CHECK();
  if (Dn > 0) {
    const int len = Dn * D + Dn;
    reduce_partials_kernel<<<reduce_partials_grid(len), kRedThreads, 0, st>>>(
        static_cast<const float*>(ws), G, len, ddense_w, Dn * D, ddense_w1);
    B200_LAUNCH_CHECK();
  }
  return B200REC_OK;
}

}  // namespace b200rec

// Deterministic segmented reduction over positions grouped by id (the scatter-add half of every
// embedding backward).  Three kernels share one "contribution" functor:
//   light:   one TPR-lane group per distinct id walks its (short) segment sequentially;
//   hot:     ids that occur more than kHotThreshold times (skewed / Zipf traffic, DIN targets) are
//            deferred: the light kernel reserves ceil(len/kHotChunk) work items per such id, and
//            one CTA per item reduces its chunk (strided partial sums + a fixed shared-memory
//            tree) into a partial row;
//   combine: adds the partial rows of each hot id in chunk order.
// All orders are fixed by the stable sort => bit-reproducible results, no float atomics, and a
// single popular id no longer serialises on one lane group (or one CTA).
#pragma once

#include "common.cuh"

namespace b200rec {

constexpr int kSegThreads = 256;

// Output geometry: rows[u*ld_rows + d]; rows1[u*ld_rows1] (+ zero_pad zeroed floats after it, used
// when rows1 is the tail of a fused [emb | w1 | pad] gradient row).
struct SegOut {
  int64_t ld_rows;
  int64_t ld_rows1;
  int zero_pad;
};
constexpr int kHotThreshold = 64;   // segments longer than this go to the hot path
constexpr int kHotChunk = 2048;     // positions per hot work item (one CTA)

// workspace layout: int32 counter | int32 items[2*max_items] (segment, chunk) | float partials
static inline size_t seg_max_items(int64_t n) { return (size_t)(n / kHotThreshold) + 1; }
static inline size_t seg_workspace_bytes(int64_t n, int D) {
  const size_t items = seg_max_items(n);
  return align_up(16 + items * 2 * sizeof(int32_t), 256) + items * (size_t)(D + 1) * sizeof(float) +
         256;
}

template <int VEC, int TPR, typename Contrib>
__global__ void __launch_bounds__(kSegThreads)
seg_light_kernel(const int32_t* __restrict__ seg_offsets, const int32_t* __restrict__ sorted_pos,
                 const int32_t* __restrict__ num_unique, Contrib contrib, float* __restrict__ rows,
                 float* __restrict__ rows1, int D, SegOut so, int32_t* __restrict__ hot_count,
                 int32_t* __restrict__ hot_items) {
  constexpr int GPB = kSegThreads / TPR;
  const int U = num_unique[0];
  const int r = threadIdx.x % TPR;
  const bool lane_ok = r * VEC < D;
  for (int64_t u = (int64_t)blockIdx.x * GPB + threadIdx.x / TPR; u < U;
       u += (int64_t)gridDim.x * GPB) {
    const int beg = seg_offsets[u];
    const int end = seg_offsets[u + 1];
    if (end - beg > kHotThreshold) {
      if (r == 0) {
        const int chunks = (end - beg + kHotChunk - 1) / kHotChunk;
        const int first = atomicAdd(hot_count, chunks);
        for (int c = 0; c < chunks; ++c) {
          hot_items[2 * (first + c)] = (int32_t)u;
          hot_items[2 * (first + c) + 1] = c | (c == 0 ? (chunks << 16) : 0);  // head knows #chunks
        }
      }
      continue;
    }
    Vec<VEC> acc = vzero<VEC>();
    float acc1 = 0.f;
    for (int i = beg; i < end; ++i) contrib.template add<VEC>(sorted_pos[i], r, lane_ok, acc, acc1);
    if (lane_ok) st_plain<VEC>(rows + (size_t)u * so.ld_rows + r * VEC, acc);
    if (rows1 != nullptr && r == 0) {
      float* p1 = rows1 + (size_t)u * so.ld_rows1;
      p1[0] = acc1;
      for (int j = 1; j <= so.zero_pad; ++j) p1[j] = 0.f;
    }
  }
}

template <int VEC, int TPR, typename Contrib>
__global__ void __launch_bounds__(kSegThreads)
seg_hot_kernel(const int32_t* __restrict__ seg_offsets, const int32_t* __restrict__ sorted_pos,
               Contrib contrib, int D, const int32_t* __restrict__ hot_count,
               const int32_t* __restrict__ hot_items, float* __restrict__ partials) {
  constexpr int GPB = kSegThreads / TPR;
  __shared__ float s_acc[GPB][TPR * VEC + 1];
  __shared__ float s_acc1[GPB];
  const int n_items = hot_count[0];
  const int g = threadIdx.x / TPR;
  const int r = threadIdx.x % TPR;
  const bool lane_ok = r * VEC < D;
  for (int h = blockIdx.x; h < n_items; h += gridDim.x) {
    const int u = hot_items[2 * h];
    const int chunk = hot_items[2 * h + 1] & 0xffff;
    const int beg = seg_offsets[u] + chunk * kHotChunk;
    const int end = min(seg_offsets[u + 1], beg + kHotChunk);
    Vec<VEC> acc = vzero<VEC>();
    float acc1 = 0.f;
    for (int i = beg + g; i < end; i += GPB)
      contrib.template add<VEC>(sorted_pos[i], r, lane_ok, acc, acc1);
#pragma unroll
    for (int k = 0; k < VEC; ++k) s_acc[g][r * VEC + k] = acc.v[k];
    if (r == 0) s_acc1[g] = acc1;
    __syncthreads();
    for (int stride = GPB / 2; stride > 0; stride >>= 1) {  // fixed tree over the groups
      if (g < stride) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) s_acc[g][r * VEC + k] += s_acc[g + stride][r * VEC + k];
        if (r == 0) s_acc1[g] += s_acc1[g + stride];
      }
      __syncthreads();
    }
    if (g == 0) {
      float* out = partials + (size_t)h * (D + 1);
      if (lane_ok) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) out[r * VEC + k] = s_acc[0][r * VEC + k];
      }
      if (r == 0) out[D] = s_acc1[0];
    }
    __syncthreads();
  }
}

// one thread per (hot id, column): rows[u] = sum_c partials[first+c] in chunk order
__global__ void seg_combine_kernel(const int32_t* __restrict__ hot_count,
                                   const int32_t* __restrict__ hot_items,
                                   const float* __restrict__ partials, float* __restrict__ rows,
                                   float* __restrict__ rows1, int D, SegOut so) {
  const int n_items = hot_count[0];
  for (int h = blockIdx.x; h < n_items; h += gridDim.x) {
    const int meta = hot_items[2 * h + 1];
    if ((meta & 0xffff) != 0) continue;  // not a head item
    const int chunks = meta >> 16;
    const int u = hot_items[2 * h];
    for (int d = threadIdx.x; d <= D; d += blockDim.x) {
      float t = 0.f;
      for (int c = 0; c < chunks; ++c) t += partials[(size_t)(h + c) * (D + 1) + d];
      if (d < D) {
        rows[(size_t)u * so.ld_rows + d] = t;
      } else if (rows1 != nullptr) {
        float* p1 = rows1 + (size_t)u * so.ld_rows1;
        p1[0] = t;
        for (int j = 1; j <= so.zero_pad; ++j) p1[j] = 0.f;
      }
    }
  }
}

// Launches light + hot + combine.  `ws` must hold seg_workspace_bytes(n, D).
// A hot id is limited to 32767 chunks (6.7e7 positions) by the item encoding.
template <int VEC, int TPR, typename Contrib>
static int launch_seg_reduce(const int32_t* seg_offsets, const int32_t* sorted_pos,
                             const int32_t* num_unique, const Contrib& contrib, float* rows,
                             float* rows1, SegOut so, int64_t n, int D, void* ws,
                             cudaStream_t st) {
  B200_REQUIRE(n < (int64_t)32767 * kHotChunk, "segment reduce: n=%lld too large", (long long)n);
  unsigned char* base = static_cast<unsigned char*>(ws);
  int32_t* hot_count = reinterpret_cast<int32_t*>(base);
  int32_t* hot_items = reinterpret_cast<int32_t*>(base + 16);
  const size_t items = seg_max_items(n);
  float* partials = reinterpret_cast<float*>(
      base + align_up(16 + items * 2 * sizeof(int32_t), 256));
  B200_CUDA(cudaMemsetAsync(hot_count, 0, sizeof(int32_t), st));
  constexpr int GPB = kSegThreads / TPR;
  const int64_t want = (n + GPB - 1) / GPB;
  const unsigned grid = (unsigned)min(want, (int64_t)sm_count() * 64);
  seg_light_kernel<VEC, TPR, Contrib><<<grid, kSegThreads, 0, st>>>(
      seg_offsets, sorted_pos, num_unique, contrib, rows, rows1, D, so, hot_count, hot_items);
  B200_LAUNCH_CHECK();
  if (n > kHotThreshold) {
    const unsigned hgrid = (unsigned)min((int64_t)items, (int64_t)sm_count() * 8);
    seg_hot_kernel<VEC, TPR, Contrib><<<hgrid, kSegThreads, 0, st>>>(
        seg_offsets, sorted_pos, contrib, D, hot_count, hot_items, partials);
    B200_LAUNCH_CHECK();
    seg_combine_kernel<<<hgrid, 128, 0, st>>>(hot_count, hot_items, partials, rows, rows1, D, so);
    B200_LAUNCH_CHECK();
  }
  return B200REC_OK;
}

}  // namespace b200rec

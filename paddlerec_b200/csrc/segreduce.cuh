// Deterministic segmented reduction over positions grouped by id (the scatter-add half of every
// embedding backward).  Two kernels share one "contribution" functor:
//   light: one TPR-lane group per distinct id walks its (short) segment sequentially;
//   hot:   ids that occur more than kHotThreshold times (skewed / Zipf traffic, DIN targets) are
//          deferred to a list and reduced by a whole CTA each — strided partial sums, then a fixed
//          shared-memory tree — so a popular id no longer serialises on one lane group.
// Both orders are fixed by the stable sort => bit-reproducible results, no float atomics.
#pragma once

#include "common.cuh"

namespace b200rec {

constexpr int kSegThreads = 256;
constexpr int kHotThreshold = 64;

// workspace: [0] hot counter (int32), [1..] hot list (segment indices)
static inline size_t seg_workspace_bytes(int64_t n) {
  return ((size_t)(n / kHotThreshold) + 8) * sizeof(int32_t);
}

template <int VEC, int TPR, typename Contrib>
__global__ void __launch_bounds__(kSegThreads)
seg_light_kernel(const int32_t* __restrict__ seg_offsets, const int32_t* __restrict__ sorted_pos,
                 const int32_t* __restrict__ num_unique, Contrib contrib, float* __restrict__ rows,
                 float* __restrict__ rows1, int D, int32_t* __restrict__ hot) {
  constexpr int GPB = kSegThreads / TPR;
  const int U = num_unique[0];
  const int r = threadIdx.x % TPR;
  const bool lane_ok = r * VEC < D;
  for (int64_t u = (int64_t)blockIdx.x * GPB + threadIdx.x / TPR; u < U;
       u += (int64_t)gridDim.x * GPB) {
    const int beg = seg_offsets[u];
    const int end = seg_offsets[u + 1];
    if (end - beg > kHotThreshold) {
      if (r == 0) hot[1 + atomicAdd(hot, 1)] = (int32_t)u;
      continue;
    }
    Vec<VEC> acc = vzero<VEC>();
    float acc1 = 0.f;
    for (int i = beg; i < end; ++i) contrib.template add<VEC>(sorted_pos[i], r, lane_ok, acc, acc1);
    if (lane_ok) st_plain<VEC>(rows + (size_t)u * D + r * VEC, acc);
    if (rows1 != nullptr && r == 0) rows1[u] = acc1;
  }
}

template <int VEC, int TPR, typename Contrib>
__global__ void __launch_bounds__(kSegThreads)
seg_hot_kernel(const int32_t* __restrict__ seg_offsets, const int32_t* __restrict__ sorted_pos,
               Contrib contrib, float* __restrict__ rows, float* __restrict__ rows1, int D,
               const int32_t* __restrict__ hot) {
  constexpr int GPB = kSegThreads / TPR;
  __shared__ float s_acc[GPB][TPR * VEC + 1];
  __shared__ float s_acc1[GPB];
  const int n_hot = hot[0];
  const int g = threadIdx.x / TPR;
  const int r = threadIdx.x % TPR;
  const bool lane_ok = r * VEC < D;
  for (int h = blockIdx.x; h < n_hot; h += gridDim.x) {
    const int u = hot[1 + h];
    const int beg = seg_offsets[u];
    const int end = seg_offsets[u + 1];
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
      if (lane_ok) {
        Vec<VEC> o;
#pragma unroll
        for (int k = 0; k < VEC; ++k) o.v[k] = s_acc[0][r * VEC + k];
        st_plain<VEC>(rows + (size_t)u * D + r * VEC, o);
      }
      if (rows1 != nullptr && r == 0) rows1[u] = s_acc1[0];
    }
    __syncthreads();
  }
}

// Launches light + hot.  `ws` must hold seg_workspace_bytes(n).
template <int VEC, int TPR, typename Contrib>
static int launch_seg_reduce(const int32_t* seg_offsets, const int32_t* sorted_pos,
                             const int32_t* num_unique, const Contrib& contrib, float* rows,
                             float* rows1, int64_t n, int D, void* ws, cudaStream_t st) {
  int32_t* hot = static_cast<int32_t*>(ws);
  B200_CUDA(cudaMemsetAsync(hot, 0, sizeof(int32_t), st));
  constexpr int GPB = kSegThreads / TPR;
  const int64_t want = (n + GPB - 1) / GPB;
  const unsigned grid = (unsigned)min(want, (int64_t)sm_count() * 64);
  seg_light_kernel<VEC, TPR, Contrib><<<grid, kSegThreads, 0, st>>>(
      seg_offsets, sorted_pos, num_unique, contrib, rows, rows1, D, hot);
  B200_LAUNCH_CHECK();
  if (n > kHotThreshold) {
    const unsigned hgrid = (unsigned)min((int64_t)(n / kHotThreshold), (int64_t)sm_count() * 8);
    seg_hot_kernel<VEC, TPR, Contrib><<<hgrid, kSegThreads, 0, st>>>(seg_offsets, sorted_pos, contrib,
                                                                    rows, rows1, D, hot);
    B200_LAUNCH_CHECK();
  }
  return B200REC_OK;
}

}  // namespace b200rec

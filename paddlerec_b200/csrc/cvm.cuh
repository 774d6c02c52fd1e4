// continuous_value_model (CVM) forward/backward — the op the GPUBox branch applies to every
// sparse_embedding output (models/rank/wide_deep/net.py:81-88, models/rank/dnn/net.py:72-79;
// show_click built at wide_deep/static_model.py:88-94).
//   input x [N, D+2]: columns 0,1 = show, click statistics stored with the embedding
//   use_cvm = 1: y [N, D+2], y0 = log(x0+1), y1 = log(x1+1) - y0, y[2:] = x[2:]
//   use_cvm = 0: y [N, D]   = x[:, 2:]
//   backward   : dx[:, 2:] = dy[:, (2|0):],  dx[:, 0:2] = show_click[:, 0:2]
//                (the table thereby accumulates show/click counts — Paddle's cvm_grad semantics)
// Streaming, HBM-bound: 2*4*(D+2) bytes per row.
#pragma once

#include "common.cuh"

namespace b200rec {

__global__ void cvm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t N, int D,
                               int use_cvm) {
  const int in_w = D + 2, out_w = use_cvm ? D + 2 : D;
  const int64_t total = N * out_w;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / out_w;
    const int c = (int)(i - n * out_w);
    const float* row = x + n * in_w;
    float v;
    if (!use_cvm) {
      v = row[c + 2];
    } else if (c == 0) {
      v = logf(row[0] + 1.f);
    } else if (c == 1) {
      v = logf(row[1] + 1.f) - logf(row[0] + 1.f);
    } else {
      v = row[c];
    }
    y[i] = v;
  }
}

__global__ void cvm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ show_click,
                               float* __restrict__ dx, int64_t N, int D, int use_cvm) {
  const int in_w = D + 2, out_w = use_cvm ? D + 2 : D;
  const int64_t total = N * in_w;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / in_w;
    const int c = (int)(i - n * in_w);
    dx[i] = (c < 2) ? show_click[n * 2 + c] : dy[n * out_w + (use_cvm ? c : c - 2)];
  }
}

}  // namespace b200rec

// K4: DIN attention pooling — attention-unit MLP over [h, t, h-t, h*t], additive mask, E^-0.5
// scale, softmax over the history and weighted-sum pooling, fused (forward).
//
// Reference: DINLayer.forward, models/rank/din/net.py:155-173 (≈12 Paddle ops that materialise the
// [B,L,4E] concat = 205 KB/sample at L=100, E=128).  Here the concat never exists:
//     [h, t, h-t, h*t] W1 = h (Wa+Wc) + (h*t) Wd + t (Wb-Wc)
// so per history position only two E x 80 products remain (half the FLOPs of the reference's
// formulation) and the t-term  tb = t (Wb-Wc) + b1  is one [B,E]x[E,80] GEMM done once per sample
// by the caller.  FLOP-bound (2*L*(2E*80 + 80*40 + 40) flop per sample, 4.8 MFLOP at L=100) on the
// fp32 SIMT pipe: this round's kernel is a register-tiled shared-memory GEMM, not tensor-core code
// (the 1e-4 parity bar rules out single-pass TF32/BF16; a split-precision tcgen05 version is listed
// in DESIGN.md as next).
//
//   kernel A (din_scores_kernel): persistent CTAs over tiles of 64 flattened (b,l) positions;
//     both E x 80 weight blocks live transposed+padded in shared memory for the whole kernel;
//     thread (og,pg) owns a 4-position x 5-output register tile; layer 2 (80->40) and layer 3
//     (40->1) run from shared memory / warp shuffles; writes the masked, scaled score per position.
//   kernel B (din_softmax_pool_kernel): one CTA per sample: softmax over L, then
//     out[b,:] = sum_l w_l h[b,l,:]  (h is re-read; it was just touched and mostly sits in L2).
#pragma once

#include "common.cuh"

namespace b200rec {

constexpr int kDinH1 = 80;
constexpr int kDinH2 = 40;
constexpr int kDinTile = 64;      // positions per tile
constexpr int kDinThreads = 256;  // 16 output groups x 16 position groups
constexpr int kDinMaxE = 128;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

struct DinSmem {
  // all sizes in floats; EP = E + 4 (row padding keeps the 128-bit LDS conflict-free)
  static __host__ __device__ size_t floats(int E) {
    const size_t EP = E + 4;
    return 2 * (size_t)kDinH1 * EP            // WacT, WdT  [80][EP]
           + (size_t)kDinH1 * kDinH2          // W2s [80][40]
           + (size_t)kDinTile * EP            // hs [64][EP]
           + hts_floats(E)                    // hts [64][EP], reused for z1s [64][81]
           + 2 * kDinH2 + 8;                  // b2, W3
  }
  static __host__ __device__ size_t hts_floats(int E) {
    const size_t EP = E + 4;
    return (size_t)kDinTile * (EP > (size_t)(kDinH1 + 1) ? EP : (size_t)(kDinH1 + 1));
  }
};

__global__ void __launch_bounds__(kDinThreads, 1)
din_scores_kernel(const float* __restrict__ hist, const float* __restrict__ tseq,
                  const float* __restrict__ tb, const float* __restrict__ Wac,
                  const float* __restrict__ Wd, const float* __restrict__ W2,
                  const float* __restrict__ b2, const float* __restrict__ W3,
                  const float* __restrict__ b3, const int64_t* __restrict__ mask,
                  float* __restrict__ scores, int64_t P, int L, int E, float scale) {
  extern __shared__ __align__(16) float sm[];
  const int EP = E + 4;
  float* WacT = sm;                               // [80][EP]
  float* WdT = WacT + (size_t)kDinH1 * EP;        // [80][EP]
  float* W2s = WdT + (size_t)kDinH1 * EP;         // [80][40]
  float* hs = W2s + kDinH1 * kDinH2;              // [64][EP]
  float* hts = hs + (size_t)kDinTile * EP;        // [64][EP]
  float* z1s = hts;                               // [64][81]  (after GEMM1)
  float* b2s = hts + DinSmem::hts_floats(E);      // [40]
  float* W3s = b2s + kDinH2;                      // [40]

  const int tid = threadIdx.x;
  // one-time: weights -> shared (transposed [o][k])
  for (int i = tid; i < E * kDinH1; i += kDinThreads) {
    const int k = i / kDinH1, o = i - k * kDinH1;
    WacT[(size_t)o * EP + k] = Wac[i];
    WdT[(size_t)o * EP + k] = Wd[i];
  }
  for (int i = tid; i < kDinH1 * kDinH2; i += kDinThreads) W2s[i] = W2[i];
  if (tid < kDinH2) {
    b2s[tid] = b2[tid];
    W3s[tid] = W3[tid];
  }
  const float b3v = b3[0];
  __syncthreads();

  const int og = tid & 15;        // outputs og + 16*j, j<5
  const int pg = tid >> 4;        // positions pg*4 + i, i<4
  const int p2 = tid >> 2;        // layer 2: position
  const int qg = tid & 3;         // layer 2: outputs qg*10 + j, j<10
  const int chunks = E / 4;
  const int64_t ntiles = (P + kDinTile - 1) / kDinTile;

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t p0 = tile * kDinTile;
    // ---- load h and h*t rows of the tile ------------------------------------------------------
    for (int i = tid; i < kDinTile * chunks; i += kDinThreads) {
      const int row = i / chunks, c = i - row * chunks;
      const int64_t p = p0 + row;
      float4 h = make_float4(0.f, 0.f, 0.f, 0.f), tt = h;
      if (p < P) {
        const int64_t b = p / L;
        h = __ldg(reinterpret_cast<const float4*>(hist + (size_t)p * E) + c);
        tt = __ldg(reinterpret_cast<const float4*>(tseq + (size_t)b * E) + c);
      }
      *reinterpret_cast<float4*>(hs + (size_t)row * EP + c * 4) = h;
      *reinterpret_cast<float4*>(hts + (size_t)row * EP + c * 4) =
          make_float4(h.x * tt.x, h.y * tt.y, h.z * tt.z, h.w * tt.w);
    }
    __syncthreads();

    // ---- GEMM1: [64 x 2E] @ [2E x 80], 4x5 register tile --------------------------------------
    float acc[4][5];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 5; ++j) acc[i][j] = 0.f;
    for (int c = 0; c < chunks; ++c) {
      float4 hv[4], htv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        hv[i] = *reinterpret_cast<const float4*>(hs + (size_t)(pg * 4 + i) * EP + c * 4);
        htv[i] = *reinterpret_cast<const float4*>(hts + (size_t)(pg * 4 + i) * EP + c * 4);
      }
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const float4 wa = *reinterpret_cast<const float4*>(WacT + (size_t)(og + 16 * j) * EP + c * 4);
        const float4 wd = *reinterpret_cast<const float4*>(WdT + (size_t)(og + 16 * j) * EP + c * 4);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float a = acc[i][j];
          a = fmaf(hv[i].x, wa.x, a); a = fmaf(hv[i].y, wa.y, a);
          a = fmaf(hv[i].z, wa.z, a); a = fmaf(hv[i].w, wa.w, a);
          a = fmaf(htv[i].x, wd.x, a); a = fmaf(htv[i].y, wd.y, a);
          a = fmaf(htv[i].z, wd.z, a); a = fmaf(htv[i].w, wd.w, a);
          acc[i][j] = a;
        }
      }
    }
    __syncthreads();  // everyone is done reading hts before it becomes z1s
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = pg * 4 + i;
      const int64_t p = p0 + row;
      const int64_t b = (p < P) ? p / L : 0;
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const int o = og + 16 * j;
        z1s[row * (kDinH1 + 1) + o] = sigmoidf_(acc[i][j] + __ldg(tb + (size_t)b * kDinH1 + o));
      }
    }
    __syncthreads();

    // ---- layer 2 (80 -> 40, sigmoid) and layer 3 (40 -> 1) -------------------------------------
    float a2[10];
#pragma unroll
    for (int j = 0; j < 10; ++j) a2[j] = 0.f;
    for (int o = 0; o < kDinH1; ++o) {
      const float z = z1s[p2 * (kDinH1 + 1) + o];
      const float* w = W2s + o * kDinH2 + qg * 10;
#pragma unroll
      for (int j = 0; j < 10; ++j) a2[j] = fmaf(z, w[j], a2[j]);
    }
    float part = 0.f;
#pragma unroll
    for (int j = 0; j < 10; ++j)
      part = fmaf(sigmoidf_(a2[j] + b2s[qg * 10 + j]), W3s[qg * 10 + j], part);
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    const int64_t p = p0 + p2;
    if (qg == 0 && p < P) {
      const float m = (mask != nullptr) ? (float)mask[p] : 0.f;
      scores[p] = ((part + b3v) + m) * scale;   // mask is added BEFORE the scale (net.py:166-168)
    }
    __syncthreads();  // smem tile is reused by the next iteration
  }
}

// One CTA per sample: w = softmax(scores[b,:]); out[b,:] = sum_l w_l * hist[b,l,:].
__global__ void __launch_bounds__(128)
din_softmax_pool_kernel(const float* __restrict__ hist, const float* __restrict__ scores,
                        float* __restrict__ weights, float* __restrict__ out, int L, int E) {
  extern __shared__ float sw[];  // [L]
  __shared__ float red[4];
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* sc = scores + (size_t)b * L;
  float mx = -INFINITY;
  for (int l = tid; l < L; l += 128) {
    const float v = sc[l];
    sw[l] = v;
    mx = fmaxf(mx, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((tid & 31) == 0) red[tid >> 5] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  float sum = 0.f;
  for (int l = tid; l < L; l += 128) {
    const float e = __expf(sw[l] - mx);
    sw[l] = e;
    sum += e;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((tid & 31) == 0) red[tid >> 5] = sum;
  __syncthreads();
  const float inv = 1.f / (((red[0] + red[1]) + red[2]) + red[3]);
  for (int l = tid; l < L; l += 128) {
    const float w = sw[l] * inv;
    sw[l] = w;
    weights[(size_t)b * L + l] = w;
  }
  __syncthreads();
  const float* hb = hist + (size_t)b * L * E;
  for (int e = tid; e < E; e += 128) {
    float acc = 0.f;
    for (int l = 0; l < L; ++l) acc = fmaf(sw[l], hb[(size_t)l * E + e], acc);
    out[(size_t)b * E + e] = acc;
  }
}

static int launch_din_attn_fwd(const float* hist, const float* tseq, const float* tb,
                               const float* Wac, const float* Wd, const float* W2, const float* b2,
                               const float* W3, const float* b3, const int64_t* mask, float* scores,
                               float* weights, float* out, int64_t B, int L, int E, float scale,
                               cudaStream_t st) {
  B200_REQUIRE(E > 0 && E % 4 == 0 && E <= kDinMaxE, "din_attn_fwd: E=%d must be a multiple of 4, <=%d",
               E, kDinMaxE);
  B200_REQUIRE(L > 0 && L <= 8192, "din_attn_fwd: L=%d out of range", L);
  B200_REQUIRE(aligned16(hist) && aligned16(tseq), "din_attn_fwd: hist/tseq must be 16-byte aligned");
  if (B == 0) return B200REC_OK;
  const int64_t P = B * L;
  const size_t smem = DinSmem::floats(E) * sizeof(float);
  B200_CUDA(cudaFuncSetAttribute(din_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
  const int64_t ntiles = (P + kDinTile - 1) / kDinTile;
  const unsigned grid = (unsigned)min(ntiles, (int64_t)sm_count());
  din_scores_kernel<<<grid, kDinThreads, smem, st>>>(hist, tseq, tb, Wac, Wd, W2, b2, W3, b3, mask,
                                                     scores, P, L, E, scale);
  B200_LAUNCH_CHECK();
  din_softmax_pool_kernel<<<(unsigned)B, 128, (size_t)L * sizeof(float), st>>>(hist, scores, weights,
                                                                               out, L, E);
  B200_LAUNCH_CHECK();
  return B200REC_OK;
}

}  // namespace b200rec

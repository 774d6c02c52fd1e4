// Hand-written tcgen05 / TMEM / TMA GEMMs for the dense tower in split precision ("bf16x3").
//
// Reference path: DNN.forward models/rank/deepfm/net.py:169-174 (Linear -> ReLU chain), its autograd
// (tools/trainer.py:151 loss.backward()), and the same chain in dcn_v2/net.py:178-184,
// wide_deep/net.py:95-97, din/net.py:175-184.  An fp32 product a@w is evaluated on the 5th-gen
// tensor cores as a_hi@w_hi + a_lo@w_hi + a_hi@w_lo (bf16 operands, exact products, fp32
// accumulation in TMEM; ~2^-16 relative), which meets the 1e-4 logit bar that plain TF32 misses.
//
// Operand format ("planes"): a matrix X[R, C] lives in HBM as bf16 [R, 2*ld] with the rounded value
// hi(X) in columns [0, C) and the residual lo(X) = bf16(X - hi) in columns [ld, ld + C); ld is C
// rounded up to 8 so that both planes are 16-byte aligned TMA tensors.  Same bytes as fp32.
//
// Two kernels, both warp-specialised (warp 0 = TMA producer, warp 1 = tcgen05.mma issuer and TMEM
// owner, warps 2-5 = epilogue) with mbarrier pipelines smem full/empty and TMEM full/empty:
//
//   tc_gemm_kmajor_kernel   D[M,N] = A[M,K] . B[N,K]^T     both operands K-major (reduction dim
//       contiguous): forward (A = activations, B = W^T planes) and dX (A = dZ planes, B = W
//       planes).  Persistent over 128 x BN output tiles, two TMEM accumulator stages so the
//       epilogue of tile t overlaps the main loop of tile t+1.  Fused epilogues: +bias, ReLU,
//       hi/lo split (emits the NEXT GEMM's operand directly), ReLU mask + bias column-sum in
//       backward, or plain fp32.
//   tc_gemm_dw_kernel       dW[K,N] = sum_m A[m,K] . G[m,N]   both operands MN-major (the batch dim
//       is the reduction), split over the batch across CTAs, fp32 partial tiles + fixed-order
//       reduce (deterministic, no atomics).
//
// Shared-memory tiles are 128-byte-swizzled exactly as TMA writes them (CU_TENSOR_MAP_SWIZZLE_128B)
// and described to the tensor core by 64-bit UMMA descriptors (layout type SWIZZLE_128B, version 1).
#pragma once

#include <cuda.h>  // CUtensorMap and enums only; cuTensorMapEncodeTiled is resolved at run time
#include <cuda_bf16.h>

#include "common.cuh"

namespace b200rec {
namespace tc {

constexpr int kBM = 128;         // output rows per tile = UMMA M (cta_group::1)
constexpr int kBK = 64;          // bf16 per k-block = one 128-byte swizzle row
constexpr int kUK = 16;          // UMMA K for 16-bit operands
constexpr int kThreads = 192;    // dW kernel: warp 0 TMA, warp 1 MMA, warps 2..5 epilogue
constexpr int kEpiWarps = 8;     // K-major kernel: two epilogue warpgroups (alternate 32-col chunks)
constexpr int kThreadsK = 64 + 32 * kEpiWarps;
constexpr uint32_t kStagingPerWarp = 4096;   // one 32x32 chunk: hi + lo bf16 tiles, or one fp32 tile
constexpr int kTmemCols = 512;   // whole TMEM of the SM (one CTA per SM by shared-memory size)
constexpr int kAccStride = 256;  // column offset of the second accumulator stage
constexpr int kMaxBN = 256;
constexpr int kDwBR = 32;        // batch rows per k-block of the dW kernel
constexpr uint32_t kSmemBudget = 227u * 1024u - 2048u;

// device-side error word: a pipeline wait that exceeds its budget records where and traps, so a
// protocol bug surfaces as a CUDA error instead of a hung GPU.
__device__ unsigned int g_tc_timeout = 0u;

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, unsigned where) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) {  // ~2 s at 2 GHz
      atomicExch(&g_tc_timeout, where | 0x80000000u);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
// 2-D tiled TMA load: box lands in shared memory 128B-swizzled, completion on an mbarrier.
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar,
                                            int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(m), "r"(bar), "r"(x), "r"(y)
      : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// all tcgen05.mma issued so far by this thread -> one arrival on `bar` when they complete
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   bar)
               : "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc], bf16 x bf16 -> fp32
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                       uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// ---- cta_group::2 (CTA pair) variants ----------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* m, uint32_t bar,
                                                int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(m), "r"(bar), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void tc_mma2(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                        uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrival on the barrier at the same offset in BOTH CTAs of the pair
__device__ __forceinline__ void tc_commit2(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
      "[%0], %1;" ::"r"(bar),
      "h"((uint16_t)3)
      : "memory");
}
// arrive on the barrier at local offset `bar` in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(bar), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote)
               : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets row (lane base + t)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ------------------------------------------------------------------------------------------------
// descriptors
// ------------------------------------------------------------------------------------------------
// UMMA shared-memory descriptor, SWIZZLE_128B (bits: [0,14) addr>>4, [16,30) LBO>>4, [32,46)
// SBO>>4, [46,48) version = 1 on sm_100, [61,64) layout type = 2).
// layout: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes,
                                              uint32_t sbo_bytes, uint32_t layout = 2u) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)layout << 61);
}
// instruction descriptor for kind::f16: fp32 accumulate, bf16 x bf16, M = 128, N = n
__host__ __device__ inline uint32_t umma_idesc(int n, int a_mn_major, int b_mn_major,
                                               int m = 128) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a_mn_major & 1) << 15) |
         ((uint32_t)(b_mn_major & 1) << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------
// epilogue description (plain pointers: this struct crosses into the kernel by value)
// ------------------------------------------------------------------------------------------------
struct Epilogue {
  const float* bias;                // [N] added before ReLU, or null
  int relu;                         // max(., 0)
  float* out_f32;                   // [M, ld_f32] or null
  int64_t ld_f32;
  __nv_bfloat16* out_planes;        // [M, 2*ldp] hi | lo, or null
  int64_t ldp;
  const __nv_bfloat16* mask_src;    // keep v where mask_src[m, n] > 0 (hi plane of the saved
  int64_t ld_mask;                  //   activation; row pitch in elements), or null
  float* colsum;                    // [tiles_m, N] per-tile column sums of the emitted v, or null
  // CrossNetV2 (dcn_v2/net.py:222-226): v = x0[m,n] * (acc + bias[n]) + xl[m,n]
  const float* cross_x0;
  const float* cross_xl;
  int64_t ld_cross;
  // also write out_planes[m, N] = 1.0 (hi) / 0 (lo): the column of ones that makes the dW GEMM of
  // the next layer produce its bias gradient as row N (needs ldp > N)
  int ones_col;
  // outputs that leave the SM as TMA bulk stores of swizzled shared-memory tiles (host decides:
  // needs 16-byte aligned base and pitch; at most one of the two uses the staging buffer)
  int tma_planes, tma_f32;
  const float* addend;              // [M, ld_add] added to acc (+bias) before everything else, or null
  int64_t ld_add;
  float* aux_f32;                   // [M, ld_aux] receives acc + bias + addend, i.e. the value BEFORE
  int64_t ld_aux;                   //   the CrossNet / ReLU / mask transforms, or null
};

__device__ __forceinline__ void split_bf16(float x, float& hi_f, __nv_bfloat16& hi,
                                           __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  hi_f = __bfloat162float(hi);
  lo = __float2bfloat16_rn(x - hi_f);
}
__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

// sum of v[j] over the 32 lanes for every j: 31 shuffles; lane j ends with column j's total
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
  for (int h = 16; h >= 1; h >>= 1) {
    const bool up = (lane & h) != 0;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const float send = up ? v[i] : v[i + h];
      const float keep = up ? v[i + h] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
    }
  }
  return v[0];
}

// hi/lo split of two values, packed for a 32-bit store (element a at the lower address)
__device__ __forceinline__ void split_pack2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
  hi = *reinterpret_cast<const uint32_t*>(&h2);
  const float ha = __uint_as_float(hi << 16), hb = __uint_as_float(hi & 0xFFFF0000u);
  const __nv_bfloat162 l2 = __floats2bfloat162_rn(a - ha, b - hb);
  lo = *reinterpret_cast<const uint32_t*>(&l2);
}

// ---- staged outputs: the warp's 32x32 chunk is written to shared memory in the swizzled layout
// the TMA engine expects and leaves as ONE bulk tensor store per plane (coalesced, asynchronous,
// no L1 involvement; rows/columns outside the matrix are clipped by the tensor map).
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(m),
               "r"(src), "r"(x), "r"(y)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c,
                                             uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
// 32 rows x 32 bf16 (64-byte rows, SWIZZLE_64B): hi tile at sbuf, lo tile at sbuf + 2048
__device__ __forceinline__ void stage_planes(uint32_t sbuf, int lane, const float (&v)[32]) {
  const uint32_t rowoff = (uint32_t)lane * 64u;
  const uint32_t sw = ((uint32_t)lane >> 1) & 3u;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split_pack2(v[8 * g + 2 * j], v[8 * g + 2 * j + 1], h[j], l[j]);
    const uint32_t off = rowoff + (((uint32_t)g ^ sw) << 4);
    st_shared_v4(sbuf + off, h[0], h[1], h[2], h[3]);
    st_shared_v4(sbuf + 2048u + off, l[0], l[1], l[2], l[3]);
  }
}
// 32 rows x 32 fp32 (128-byte rows, SWIZZLE_128B) at sbuf
__device__ __forceinline__ void stage_f32(uint32_t sbuf, int lane, const float (&v)[32]) {
  const uint32_t rowoff = (uint32_t)lane * 128u;
  const uint32_t sw = (uint32_t)lane & 7u;
#pragma unroll
  for (int g = 0; g < 8; ++g)
    st_shared_v4(sbuf + rowoff + (((uint32_t)g ^ sw) << 4), __float_as_uint(v[4 * g]),
                 __float_as_uint(v[4 * g + 1]), __float_as_uint(v[4 * g + 2]),
                 __float_as_uint(v[4 * g + 3]));
}

// Full 32-column chunk, everything 16-byte aligned: vector loads/stores only.  `mk` holds the 32
// mask values (bf16 pairs) fetched before the TMEM wait.
__device__ __forceinline__ void epilogue_fast(const Epilogue& ep, float (&v)[32], int64_t row,
                                              int col0, const uint4 (&mk)[4], bool store_f32,
                                              bool store_planes) {
  if (ep.bias != nullptr) {
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + col0) + g);
      v[4 * g] += b.x; v[4 * g + 1] += b.y; v[4 * g + 2] += b.z; v[4 * g + 3] += b.w;
    }
  }
  if (ep.addend != nullptr) {
    const float4* ad = reinterpret_cast<const float4*>(ep.addend + row * ep.ld_add + col0);
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const float4 a = __ldg(ad + g);
      v[4 * g] += a.x; v[4 * g + 1] += a.y; v[4 * g + 2] += a.z; v[4 * g + 3] += a.w;
    }
  }
  if (ep.aux_f32 != nullptr) {
    float4* o = reinterpret_cast<float4*>(ep.aux_f32 + row * ep.ld_aux + col0);
#pragma unroll
    for (int g = 0; g < 8; ++g) o[g] = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
  }
  if (ep.cross_x0 != nullptr) {
    const float4* x0 = reinterpret_cast<const float4*>(ep.cross_x0 + row * ep.ld_cross + col0);
    const float4* xl = reinterpret_cast<const float4*>(ep.cross_xl + row * ep.ld_cross + col0);
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const float4 a = __ldg(x0 + g), b = __ldg(xl + g);
      v[4 * g] = fmaf(a.x, v[4 * g], b.x);
      v[4 * g + 1] = fmaf(a.y, v[4 * g + 1], b.y);
      v[4 * g + 2] = fmaf(a.z, v[4 * g + 2], b.z);
      v[4 * g + 3] = fmaf(a.w, v[4 * g + 3], b.w);
    }
  }
  if (ep.relu) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  if (ep.mask_src != nullptr) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const uint32_t w[4] = {mk[g].x, mk[g].y, mk[g].z, mk[g].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        // bf16 > 0  <=>  its bit pattern, as a signed integer in the high half, is > 0
        if (!((int)(w[j] << 16) > 0)) v[8 * g + 2 * j] = 0.f;
        if (!((int)(w[j] & 0xFFFF0000u) > 0)) v[8 * g + 2 * j + 1] = 0.f;
      }
    }
  }
  if (ep.out_f32 != nullptr && store_f32) {
    float4* o = reinterpret_cast<float4*>(ep.out_f32 + row * ep.ld_f32 + col0);
#pragma unroll
    for (int g = 0; g < 8; ++g) o[g] = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
  }
  if (ep.out_planes != nullptr && store_planes) {
    uint4* oh = reinterpret_cast<uint4*>(ep.out_planes + row * 2 * ep.ldp + col0);
    uint4* ol = reinterpret_cast<uint4*>(ep.out_planes + row * 2 * ep.ldp + ep.ldp + col0);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      uint32_t h[4], l[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) split_pack2(v[8 * g + 2 * j], v[8 * g + 2 * j + 1], h[j], l[j]);
      oh[g] = make_uint4(h[0], h[1], h[2], h[3]);
      ol[g] = make_uint4(l[0], l[1], l[2], l[3]);
    }
  }
}

// One row (this thread) x 32 columns of the accumulator -> outputs.  n_valid = columns of this chunk
// that exist in the matrix; row_ok = the row exists.
__device__ __forceinline__ void epilogue_chunk(const Epilogue& ep, float (&v)[32], int64_t row,
                                               bool row_ok, int col0, int n_valid, int lane,
                                               float* s_colsum /* [32] for this warp or null */,
                                               bool store_f32 = true, bool store_planes = true) {
  if (ep.bias != nullptr || ep.cross_x0 != nullptr || ep.relu || ep.mask_src != nullptr ||
      ep.addend != nullptr || ep.aux_f32 != nullptr) {
#pragma unroll
    for (int g = 0; g < 8; ++g) {   // groups of 4 columns
      const int c = g * 4;
      if (c < n_valid) {
        if (ep.bias != nullptr) {
          if (c + 4 <= n_valid && (reinterpret_cast<uintptr_t>(ep.bias + col0 + c) & 15u) == 0) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + c));
            v[c] += b.x; v[c + 1] += b.y; v[c + 2] += b.z; v[c + 3] += b.w;
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (c + j < n_valid) v[c + j] += __ldg(ep.bias + col0 + c + j);
          }
        }
        if (ep.addend != nullptr && row_ok) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (c + j < n_valid) v[c + j] += __ldg(ep.addend + row * ep.ld_add + col0 + c + j);
        }
        if (ep.aux_f32 != nullptr && row_ok) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (c + j < n_valid) ep.aux_f32[row * ep.ld_aux + col0 + c + j] = v[c + j];
        }
        if (ep.cross_x0 != nullptr && row_ok) {
          const float* x0 = ep.cross_x0 + row * ep.ld_cross + col0 + c;
          const float* xl = ep.cross_xl + row * ep.ld_cross + col0 + c;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (c + j < n_valid) v[c + j] = fmaf(__ldg(x0 + j), v[c + j], __ldg(xl + j));
        }
        if (ep.relu) {
#pragma unroll
          for (int j = 0; j < 4; ++j) v[c + j] = fmaxf(v[c + j], 0.f);
        }
      }
    }
    if (ep.mask_src != nullptr) {
      const __nv_bfloat16* mrow = ep.mask_src + row * ep.ld_mask + col0;
#pragma unroll
      for (int g = 0; g < 4; ++g) {   // groups of 8 columns = 16 bytes of bf16
        const int c = g * 8;
        if (c + 8 <= n_valid && row_ok && (reinterpret_cast<uintptr_t>(mrow + c) & 15u) == 0) {
          const uint4 q = __ldg(reinterpret_cast<const uint4*>(mrow + c));
          const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float a0 = __uint_as_float(w[j] << 16);
            const float a1 = __uint_as_float(w[j] & 0xFFFF0000u);
            if (!(a0 > 0.f)) v[c + 2 * j] = 0.f;
            if (!(a1 > 0.f)) v[c + 2 * j + 1] = 0.f;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (c + j < n_valid && row_ok) {
              if (!(__bfloat162float(mrow[c + j]) > 0.f)) v[c + j] = 0.f;
            }
        }
      }
    }
  }
  if (!row_ok) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = 0.f;
  }
  if (row_ok) {
    if (ep.out_f32 != nullptr && store_f32) {
      float* o = ep.out_f32 + row * ep.ld_f32 + col0;
      const bool vec_ok = (reinterpret_cast<uintptr_t>(o) & 15u) == 0;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const int c = g * 4;
        if (c + 4 <= n_valid && vec_ok) {
          *reinterpret_cast<float4*>(o + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (c + j < n_valid) o[c + j] = v[c + j];
        }
      }
    }
    if (ep.out_planes != nullptr && store_planes) {
      __nv_bfloat16* oh = ep.out_planes + row * 2 * ep.ldp + col0;
      __nv_bfloat16* ol = oh + ep.ldp;
      const bool vec_ok = ((reinterpret_cast<uintptr_t>(oh) | reinterpret_cast<uintptr_t>(ol)) & 15u) == 0;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int c = g * 8;
        if (c >= n_valid) break;
        __nv_bfloat16 h[8], l[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float hf;
          split_bf16(v[c + j], hf, h[j], l[j]);
        }
        if (c + 8 <= n_valid && vec_ok) {
          *reinterpret_cast<uint4*>(oh + c) = make_uint4(pack_bf16(h[0], h[1]), pack_bf16(h[2], h[3]),
                                                         pack_bf16(h[4], h[5]), pack_bf16(h[6], h[7]));
          *reinterpret_cast<uint4*>(ol + c) = make_uint4(pack_bf16(l[0], l[1]), pack_bf16(l[2], l[3]),
                                                         pack_bf16(l[4], l[5]), pack_bf16(l[6], l[7]));
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (c + j < n_valid) {
              oh[c + j] = h[j];
              ol[c + j] = l[j];
            }
        }
      }
    }
  }
  if (s_colsum != nullptr) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j >= n_valid) v[j] = 0.f;
    const float t = warp_transpose_sum(v, lane);
    s_colsum[lane] = t;
  }
}

// ------------------------------------------------------------------------------------------------
// K-major GEMM: D[M,N] = A[M,K] . B[N,K]^T  (three tcgen05.mma per 16-wide k step)
//   BK = 64: 128-byte swizzle rows (2 pipeline stages at BN = 208)
//   BK = 32: 64-byte swizzle rows, half-size stages -> 5 stages at BN = 208: same bytes in shared
//            memory, but a freed slot is refilled twice as early, which hides the TMA latency
// ------------------------------------------------------------------------------------------------
template <int BK>
__global__ void __launch_bounds__(kThreadsK, 1)
tc_gemm_kmajor_kernel(const __grid_constant__ CUtensorMap tmA_hi,
                      const __grid_constant__ CUtensorMap tmA_lo,
                      const __grid_constant__ CUtensorMap tmB_hi,
                      const __grid_constant__ CUtensorMap tmB_lo,
                      const __grid_constant__ CUtensorMap tmO_hi,
                      const __grid_constant__ CUtensorMap tmO_lo,
                      const __grid_constant__ CUtensorMap tmO_f32, int M, int N, int K, int BN,
                      int stages, Epilogue ep) {
  constexpr uint32_t kRow = BK * 2;                       // bytes per operand row in a stage
  constexpr uint32_t kLayout = BK == 64 ? 2u : 4u;        // SWIZZLE_128B / SWIZZLE_64B
  constexpr uint32_t kSbo = 8u * kRow;                    // 8-row swizzle atom
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;          // swizzled tiles: atom-aligned
  uint8_t* sm = smem_raw + (base - raw);
  const uint32_t a_bytes = kBM * kRow;
  const uint32_t b_bytes = (uint32_t)BN * kRow;
  const uint32_t stage_bytes = 2u * a_bytes + 2u * b_bytes;
  // [stages][output staging: 4 KB per epilogue warp][barriers][tmem ptr][colsum scratch]
  const uint32_t stage_end = (uint32_t)stages * stage_bytes;
  const uint32_t staging = base + stage_end;
  const uint32_t bar_off = stage_end + kEpiWarps * kStagingPerWarp;
  const uint32_t bar0 = base + bar_off;
  const uint32_t bar_full = bar0, bar_empty = bar0 + 8u * stages;
  const uint32_t bar_tfull = bar0 + 16u * stages, bar_tempty = bar_tfull + 16u;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + (size_t)bar_off + 16 * stages + 32);
  float* s_colsum = reinterpret_cast<float*>(sm + (size_t)bar_off + 16 * stages + 64);
  // s_colsum: [4 lane quarters][kMaxBN], only with ep.colsum

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (M + kBM - 1) / kBM, tiles_n = (N + BN - 1) / BN;
  const int n_tiles = tiles_m * tiles_n;
  const int nkb = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA_hi); prefetch_tmap(&tmA_lo); prefetch_tmap(&tmB_hi); prefetch_tmap(&tmB_lo);
    if (ep.tma_planes) { prefetch_tmap(&tmO_hi); prefetch_tmap(&tmO_lo); }
    if (ep.tma_f32) prefetch_tmap(&tmO_f32);
    for (int s = 0; s < stages; ++s) {
      mbar_init(bar_full + 8u * s, 1);
      mbar_init(bar_empty + 8u * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_tfull + 8u * s, 1);
      mbar_init(bar_tempty + 8u * s, kEpiWarps);   // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int m0 = (tile / tiles_n) * kBM, n0 = (tile % tiles_n) * BN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(bar_empty + 8u * stage, phase ^ 1u, 0x100u + stage);
          const uint32_t full = bar_full + 8u * stage;
          const uint32_t sA = base + (uint32_t)stage * stage_bytes;
          mbar_expect_tx(full, stage_bytes);
          tma_load_2d(sA, &tmA_hi, full, kb * BK, m0);
          tma_load_2d(sA + a_bytes, &tmA_lo, full, kb * BK, m0);
          tma_load_2d(sA + 2u * a_bytes, &tmB_hi, full, kb * BK, n0);
          tma_load_2d(sA + 2u * a_bytes + b_bytes, &tmB_lo, full, kb * BK, n0);
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int n0 = (tile % tiles_n) * BN;
        int n_cur = N - n0;
        n_cur = n_cur >= BN ? BN : ((n_cur + 15) & ~15);
        const uint32_t idesc = umma_idesc(n_cur, 0, 0);
        mbar_wait(bar_tempty + 8u * acc, acc_phase ^ 1u, 0x200u + acc);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * kAccStride;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(bar_full + 8u * stage, phase, 0x300u + stage);
          tc_fence_after();
          const uint32_t sA = base + (uint32_t)stage * stage_bytes;
          const uint32_t sB = sA + 2u * a_bytes;
          int ksteps = (K - kb * BK + kUK - 1) / kUK;
          ksteps = ksteps > BK / kUK ? BK / kUK : ksteps;
          for (int j = 0; j < ksteps; ++j) {
            const uint32_t ko = (uint32_t)j * (kUK * 2);   // 32 bytes further inside the swizzle row
            const uint64_t a_hi = umma_desc(sA + ko, 16, kSbo, kLayout);
            const uint64_t a_lo = umma_desc(sA + a_bytes + ko, 16, kSbo, kLayout);
            const uint64_t b_hi = umma_desc(sB + ko, 16, kSbo, kLayout);
            const uint64_t b_lo = umma_desc(sB + b_bytes + ko, 16, kSbo, kLayout);
            tc_mma(d_tmem, a_hi, b_hi, idesc, (kb | j) != 0 ? 1u : 0u);
            tc_mma(d_tmem, a_lo, b_hi, idesc, 1u);
            tc_mma(d_tmem, a_hi, b_lo, idesc, 1u);
          }
          tc_commit(bar_empty + 8u * stage);                 // smem slot free when these complete
          if (kb == nkb - 1) tc_commit(bar_tfull + 8u * acc);  // accumulator ready
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // epilogue: a warp may touch TMEM lanes 32*(warp%4) .. +31 only; the two warps that share a
    // lane quarter take alternate 32-column chunks
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool want_colsum = ep.colsum != nullptr;
    const bool tma_p = ep.tma_planes != 0 && !want_colsum, tma_f = ep.tma_f32 != 0 && !want_colsum;
    const uint32_t sbuf = staging + (uint32_t)(warp - 2) * kStagingPerWarp;
    const bool base_aligned =
        (ep.out_f32 == nullptr || ((reinterpret_cast<uintptr_t>(ep.out_f32) | (ep.ld_f32 * 4)) & 15u) == 0) &&
        (ep.out_planes == nullptr || ((reinterpret_cast<uintptr_t>(ep.out_planes) | (ep.ldp * 2)) & 15u) == 0) &&
        (ep.mask_src == nullptr || ((reinterpret_cast<uintptr_t>(ep.mask_src) | (ep.ld_mask * 2)) & 15u) == 0) &&
        (ep.bias == nullptr || (reinterpret_cast<uintptr_t>(ep.bias) & 15u) == 0) &&
        (ep.addend == nullptr || ((reinterpret_cast<uintptr_t>(ep.addend) | (ep.ld_add * 4)) & 15u) == 0) &&
        (ep.aux_f32 == nullptr || ((reinterpret_cast<uintptr_t>(ep.aux_f32) | (ep.ld_aux * 4)) & 15u) == 0) &&
        (ep.cross_x0 == nullptr ||
         ((reinterpret_cast<uintptr_t>(ep.cross_x0) | reinterpret_cast<uintptr_t>(ep.cross_xl) |
           (ep.ld_cross * 4)) & 15u) == 0);
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int m_blk = tile / tiles_n;
      const int m0 = m_blk * kBM, n0 = (tile % tiles_n) * BN;
      const int n_tile = (N - n0) < BN ? (N - n0) : BN;
      const int64_t row = (int64_t)m0 + q * 32 + lane;
      const bool row_ok = row < M;
      // warp-uniform on purpose: tcgen05.ld is .sync.aligned, so the 32 lanes must not split into a
      // fast and a slow path inside the chunk loop (a partially valid last row block goes slow)
      const bool fast_ok =
          base_aligned && ((int64_t)m0 + q * 32 + 31 < M) && (n0 & 7) == 0 && !want_colsum;
      mbar_wait(bar_tfull + 8u * acc, acc_phase, 0x400u + acc);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * kAccStride;
      for (int c0 = half * 32; c0 < n_tile; c0 += 64) {
        uint32_t r[32];
        __syncwarp();
        tmem_ld32(t_row + (uint32_t)c0, r);
        const int nv = (n_tile - c0) < 32 ? (n_tile - c0) : 32;
        const bool fast = fast_ok && nv == 32;
        uint4 mk[4];
        if (fast && ep.mask_src != nullptr) {    // in flight while the TMEM load completes
          const uint4* mp = reinterpret_cast<const uint4*>(ep.mask_src + row * ep.ld_mask + n0 + c0);
#pragma unroll
          for (int g = 0; g < 4; ++g) mk[g] = __ldg(mp + g);
        }
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        // Full 32-column chunks leave through the TMA staging buffer; a partial chunk (tile width
        // not a multiple of 32) is stored directly: its 32-wide box would spill into the
        // neighbouring tile's columns (the tensor map clips only at the matrix edge).
        const bool via_tma = (tma_p || tma_f) && nv == 32;
        const bool st_f = !(tma_f && via_tma), st_p = !(tma_p && via_tma);
        if (fast)
          epilogue_fast(ep, v, row, n0 + c0, mk, st_f, st_p);
        else
          epilogue_chunk(ep, v, row, row_ok, n0 + c0, nv, lane,
                         want_colsum ? s_colsum + q * kMaxBN + c0 : nullptr, st_f, st_p);
        if (via_tma) {
          if (lane == 0) bulk_wait_read0();      // the previous store has finished reading sbuf
          __syncwarp();
          if (tma_p)
            stage_planes(sbuf, lane, v);
          else
            stage_f32(sbuf, lane, v);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (tma_p) {
              tma_store_2d(&tmO_hi, sbuf, n0 + c0, m0 + q * 32);
              tma_store_2d(&tmO_lo, sbuf + 2048u, n0 + c0, m0 + q * 32);
            } else {
              tma_store_2d(&tmO_f32, sbuf, n0 + c0, m0 + q * 32);
            }
            bulk_commit();
          }
        }
      }
      if (ep.ones_col && ep.out_planes != nullptr && row_ok && half == 0 && n0 + n_tile == N) {
        __nv_bfloat16* oh = ep.out_planes + row * 2 * ep.ldp + N;
        oh[0] = __float2bfloat16_rn(1.f);
        oh[ep.ldp] = __float2bfloat16_rn(0.f);
      }
      // accumulator drained -> the MMA warp may overwrite this stage
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8u * acc);
      if (want_colsum) {
        named_bar_sync(1, 32 * kEpiWarps);
        const int t = threadIdx.x - 64;
        for (int c = t; c < n_tile; c += 32 * kEpiWarps) {
          const float s = ((s_colsum[c] + s_colsum[kMaxBN + c]) + s_colsum[2 * kMaxBN + c]) +
                          s_colsum[3 * kMaxBN + c];
          ep.colsum[(int64_t)m_blk * N + n0 + c] = s;
        }
        named_bar_sync(1, 32 * kEpiWarps);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    if ((tma_p || tma_f) && lane == 0) bulk_wait0();   // shared memory must outlive the stores
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// K-major GEMM on CTA PAIRS (cta_group::2): the two SMs of a cluster compute one 256 x BN tile.
// Each CTA loads its own 128 rows of A but only HALF of the B tile; tcgen05.mma.cta_group::2 (issued
// by the leader CTA alone) reads B from both shared memories, so the L2 -> SM operand traffic per
// flop drops by ~30 % (A 32 KB + B 26 KB per k-block per SM instead of 32 + 52 at BN = 208) — the
// 1-CTA kernel above is bound by exactly that traffic — and the freed shared memory holds a third
// pipeline stage.  Protocol (as CUTLASS's 2-SM pipelines): both producers signal the LEADER's full
// barrier (peer bit of the barrier address cleared), the leader arms it with the bytes of both;
// tcgen05.commit multicasts the slot release and the accumulator-ready signal to both CTAs; the
// peer's epilogue warps arrive remotely on the leader's TMEM-empty barrier.
// ------------------------------------------------------------------------------------------------
template <int BK>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreadsK, 1)
tc_gemm_kmajor2_kernel(const __grid_constant__ CUtensorMap tmA_hi,
                      const __grid_constant__ CUtensorMap tmA_lo,
                      const __grid_constant__ CUtensorMap tmB_hi,
                      const __grid_constant__ CUtensorMap tmB_lo,
                      const __grid_constant__ CUtensorMap tmO_hi,
                      const __grid_constant__ CUtensorMap tmO_lo,
                      const __grid_constant__ CUtensorMap tmO_f32, int M, int N, int K, int BN,
                      int stages, Epilogue ep) {
  constexpr uint32_t kRow = BK * 2;                       // bytes per operand row in a stage
  constexpr uint32_t kLayout = BK == 64 ? 2u : 4u;        // SWIZZLE_128B / SWIZZLE_64B
  constexpr uint32_t kSbo = 8u * kRow;                    // 8-row swizzle atom
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;          // swizzled tiles: atom-aligned
  uint8_t* sm = smem_raw + (base - raw);
  const uint32_t cta_rank = cluster_ctarank();          // 0 = leader: issues every tcgen05.mma
  const bool leader = cta_rank == 0;
  const int half_bn = BN / 2;                            // B rows held by EACH CTA of the pair
  const uint32_t a_bytes = kBM * kRow;
  const uint32_t b_bytes = (uint32_t)half_bn * kRow;
  const uint32_t stage_bytes = 2u * a_bytes + 2u * b_bytes;
  // [stages][output staging: 4 KB per epilogue warp][barriers][tmem ptr][colsum scratch]
  const uint32_t stage_end = (uint32_t)stages * stage_bytes;
  const uint32_t staging = base + stage_end;
  const uint32_t bar_off = stage_end + kEpiWarps * kStagingPerWarp;
  const uint32_t bar0 = base + bar_off;
  const uint32_t bar_full = bar0, bar_empty = bar0 + 8u * stages;
  const uint32_t bar_tfull = bar0 + 16u * stages, bar_tempty = bar_tfull + 16u;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + (size_t)bar_off + 16 * stages + 32);
  float* s_colsum = reinterpret_cast<float*>(sm + (size_t)bar_off + 16 * stages + 64);
  // s_colsum: [4 lane quarters][kMaxBN], only with ep.colsum

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (M + 2 * kBM - 1) / (2 * kBM), tiles_n = (N + BN - 1) / BN;   // 256-row pairs
  const int n_tiles = tiles_m * tiles_n;
  const int nkb = (K + BK - 1) / BK;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA_hi); prefetch_tmap(&tmA_lo); prefetch_tmap(&tmB_hi); prefetch_tmap(&tmB_lo);
    if (ep.tma_planes) { prefetch_tmap(&tmO_hi); prefetch_tmap(&tmO_lo); }
    if (ep.tma_f32) prefetch_tmap(&tmO_f32);
    for (int s = 0; s < stages; ++s) {
      mbar_init(bar_full + 8u * s, 1);
      mbar_init(bar_empty + 8u * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_tfull + 8u * s, 1);
      mbar_init(bar_tempty + 8u * s, 2 * kEpiWarps);   // the epilogue warps of BOTH CTAs (leader's copy is used)
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc2(smem_u32(tmem_slot), kTmemCols);   // collective over the CTA pair (same warp id in both)
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync();      // barriers of both CTAs are initialised before any remote arrive / TMA signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < n_tiles; tile += n_clusters) {
        const int m0 = (tile / tiles_n) * 2 * kBM + (int)cta_rank * kBM;
        const int n0 = (tile % tiles_n) * BN;
        int n_cur = N - n0;
        n_cur = n_cur >= BN ? BN : ((n_cur + 15) & ~15);
        const int nb0 = n0 + (int)cta_rank * (n_cur / 2);     // this CTA's half of the B rows
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(bar_empty + 8u * stage, phase ^ 1u, 0x100u + stage);
          // transaction bytes of BOTH CTAs land on the leader's barrier (peer bit cleared)
          const uint32_t full = (bar_full + 8u * stage) & 0xFEFFFFFFu;
          const uint32_t sA = base + (uint32_t)stage * stage_bytes;
          if (leader) mbar_expect_tx(bar_full + 8u * stage, 2u * stage_bytes);
          tma_load_2d_2sm(sA, &tmA_hi, full, kb * BK, m0);
          tma_load_2d_2sm(sA + a_bytes, &tmA_lo, full, kb * BK, m0);
          tma_load_2d_2sm(sA + 2u * a_bytes, &tmB_hi, full, kb * BK, nb0);
          tma_load_2d_2sm(sA + 2u * a_bytes + b_bytes, &tmB_lo, full, kb * BK, nb0);
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = cluster_id; tile < n_tiles; tile += n_clusters) {
        const int n0 = (tile % tiles_n) * BN;
        int n_cur = N - n0;
        n_cur = n_cur >= BN ? BN : ((n_cur + 15) & ~15);
        const uint32_t idesc = umma_idesc(n_cur, 0, 0, 2 * kBM);   // M = 256 over the two SMs
        mbar_wait(bar_tempty + 8u * acc, acc_phase ^ 1u, 0x200u + acc);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * kAccStride;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(bar_full + 8u * stage, phase, 0x300u + stage);
          tc_fence_after();
          const uint32_t sA = base + (uint32_t)stage * stage_bytes;
          const uint32_t sB = sA + 2u * a_bytes;
          int ksteps = (K - kb * BK + kUK - 1) / kUK;
          ksteps = ksteps > BK / kUK ? BK / kUK : ksteps;
          for (int j = 0; j < ksteps; ++j) {
            const uint32_t ko = (uint32_t)j * (kUK * 2);   // 32 bytes further inside the swizzle row
            const uint64_t a_hi = umma_desc(sA + ko, 16, kSbo, kLayout);
            const uint64_t a_lo = umma_desc(sA + a_bytes + ko, 16, kSbo, kLayout);
            const uint64_t b_hi = umma_desc(sB + ko, 16, kSbo, kLayout);
            const uint64_t b_lo = umma_desc(sB + b_bytes + ko, 16, kSbo, kLayout);
            tc_mma2(d_tmem, a_hi, b_hi, idesc, (kb | j) != 0 ? 1u : 0u);
            tc_mma2(d_tmem, a_lo, b_hi, idesc, 1u);
            tc_mma2(d_tmem, a_hi, b_lo, idesc, 1u);
          }
          tc_commit2(bar_empty + 8u * stage);                 // frees the slot in BOTH CTAs
          if (kb == nkb - 1) tc_commit2(bar_tfull + 8u * acc);  // both epilogues may start
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // epilogue: a warp may touch TMEM lanes 32*(warp%4) .. +31 only; the two warps that share a
    // lane quarter take alternate 32-column chunks
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool want_colsum = ep.colsum != nullptr;
    const bool tma_p = ep.tma_planes != 0 && !want_colsum, tma_f = ep.tma_f32 != 0 && !want_colsum;
    const uint32_t sbuf = staging + (uint32_t)(warp - 2) * kStagingPerWarp;
    const bool base_aligned =
        (ep.out_f32 == nullptr || ((reinterpret_cast<uintptr_t>(ep.out_f32) | (ep.ld_f32 * 4)) & 15u) == 0) &&
        (ep.out_planes == nullptr || ((reinterpret_cast<uintptr_t>(ep.out_planes) | (ep.ldp * 2)) & 15u) == 0) &&
        (ep.mask_src == nullptr || ((reinterpret_cast<uintptr_t>(ep.mask_src) | (ep.ld_mask * 2)) & 15u) == 0) &&
        (ep.bias == nullptr || (reinterpret_cast<uintptr_t>(ep.bias) & 15u) == 0) &&
        (ep.addend == nullptr || ((reinterpret_cast<uintptr_t>(ep.addend) | (ep.ld_add * 4)) & 15u) == 0) &&
        (ep.aux_f32 == nullptr || ((reinterpret_cast<uintptr_t>(ep.aux_f32) | (ep.ld_aux * 4)) & 15u) == 0) &&
        (ep.cross_x0 == nullptr ||
         ((reinterpret_cast<uintptr_t>(ep.cross_x0) | reinterpret_cast<uintptr_t>(ep.cross_xl) |
           (ep.ld_cross * 4)) & 15u) == 0);
    for (int tile = cluster_id; tile < n_tiles; tile += n_clusters) {
      const int m_blk = (tile / tiles_n) * 2 + (int)cta_rank;
      const int m0 = m_blk * kBM, n0 = (tile % tiles_n) * BN;
      const int n_tile = (N - n0) < BN ? (N - n0) : BN;
      const int64_t row = (int64_t)m0 + q * 32 + lane;
      const bool row_ok = row < M;
      // warp-uniform on purpose: tcgen05.ld is .sync.aligned, so the 32 lanes must not split into a
      // fast and a slow path inside the chunk loop (a partially valid last row block goes slow)
      const bool fast_ok =
          base_aligned && ((int64_t)m0 + q * 32 + 31 < M) && (n0 & 7) == 0 && !want_colsum;
      mbar_wait(bar_tfull + 8u * acc, acc_phase, 0x400u + acc);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * kAccStride;
      for (int c0 = half * 32; c0 < n_tile; c0 += 64) {
        uint32_t r[32];
        __syncwarp();
        tmem_ld32(t_row + (uint32_t)c0, r);
        const int nv = (n_tile - c0) < 32 ? (n_tile - c0) : 32;
        const bool fast = fast_ok && nv == 32;
        uint4 mk[4];
        if (fast && ep.mask_src != nullptr) {    // in flight while the TMEM load completes
          const uint4* mp = reinterpret_cast<const uint4*>(ep.mask_src + row * ep.ld_mask + n0 + c0);
#pragma unroll
          for (int g = 0; g < 4; ++g) mk[g] = __ldg(mp + g);
        }
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        // Full 32-column chunks leave through the TMA staging buffer; a partial chunk (tile width
        // not a multiple of 32) is stored directly: its 32-wide box would spill into the
        // neighbouring tile's columns (the tensor map clips only at the matrix edge).
        const bool via_tma = (tma_p || tma_f) && nv == 32;
        const bool st_f = !(tma_f && via_tma), st_p = !(tma_p && via_tma);
        if (fast)
          epilogue_fast(ep, v, row, n0 + c0, mk, st_f, st_p);
        else
          epilogue_chunk(ep, v, row, row_ok, n0 + c0, nv, lane,
                         want_colsum ? s_colsum + q * kMaxBN + c0 : nullptr, st_f, st_p);
        if (via_tma) {
          if (lane == 0) bulk_wait_read0();      // the previous store has finished reading sbuf
          __syncwarp();
          if (tma_p)
            stage_planes(sbuf, lane, v);
          else
            stage_f32(sbuf, lane, v);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (tma_p) {
              tma_store_2d(&tmO_hi, sbuf, n0 + c0, m0 + q * 32);
              tma_store_2d(&tmO_lo, sbuf + 2048u, n0 + c0, m0 + q * 32);
            } else {
              tma_store_2d(&tmO_f32, sbuf, n0 + c0, m0 + q * 32);
            }
            bulk_commit();
          }
        }
      }
      if (ep.ones_col && ep.out_planes != nullptr && row_ok && half == 0 && n0 + n_tile == N) {
        __nv_bfloat16* oh = ep.out_planes + row * 2 * ep.ldp + N;
        oh[0] = __float2bfloat16_rn(1.f);
        oh[ep.ldp] = __float2bfloat16_rn(0.f);
      }
      // accumulator drained -> the MMA warp may overwrite this stage
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(bar_tempty + 8u * acc, 0u);   // the leader's MMA warp waits on it
      if (want_colsum) {
        named_bar_sync(1, 32 * kEpiWarps);
        const int t = threadIdx.x - 64;
        for (int c = t; c < n_tile; c += 32 * kEpiWarps) {
          const float s = ((s_colsum[c] + s_colsum[kMaxBN + c]) + s_colsum[2 * kMaxBN + c]) +
                          s_colsum[3 * kMaxBN + c];
          ep.colsum[(int64_t)m_blk * N + n0 + c] = s;
        }
        named_bar_sync(1, 32 * kEpiWarps);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    if ((tma_p || tma_f) && lane == 0) bulk_wait0();   // shared memory must outlive the stores
  }
  tc_fence_before();
  cluster_sync();      // nobody leaves while the peer may still signal / read its shared memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// dW GEMM: P[split][K,N] = sum_{m in split} A[m,K]^T G[m,N]   (both operands MN-major)
// ------------------------------------------------------------------------------------------------
struct DwDebug {
  uint32_t lbo_a, lbo_b, sbo;   // 0 = default
};

__global__ void __launch_bounds__(kThreads, 1)
tc_gemm_dw_kernel(const __grid_constant__ CUtensorMap tmA_hi,
                  const __grid_constant__ CUtensorMap tmA_lo,
                  const __grid_constant__ CUtensorMap tmG_hi,
                  const __grid_constant__ CUtensorMap tmG_lo, int M, int Kin, int N, int BN,
                  int stages, int rows_per_split, float* __restrict__ partials, DwDebug dbg) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const int nb = (BN + 63) / 64;                         // 64-column chunks of the G tile
  const uint32_t chunk = kDwBR * 128u;                   // one TMA box: BR rows x 128 B
  const uint32_t a_bytes = 2u * chunk, b_bytes = (uint32_t)nb * chunk;
  const uint32_t stage_bytes = 2u * a_bytes + 2u * b_bytes;
  const uint32_t bar0 = base + (uint32_t)stages * stage_bytes;
  const uint32_t bar_full = bar0, bar_empty = bar0 + 8u * stages, bar_tfull = bar0 + 16u * stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + (size_t)stages * stage_bytes + 16 * stages + 32);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = (N + BN - 1) / BN;
  const int k0 = (blockIdx.x / tiles_n) * kBM, n0 = (blockIdx.x % tiles_n) * BN;
  const int split = blockIdx.y;
  const int r0 = split * rows_per_split;
  int r1 = r0 + rows_per_split;
  r1 = r1 > M ? M : r1;
  const int nkb = r1 > r0 ? (r1 - r0 + kDwBR - 1) / kDwBR : 0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA_hi); prefetch_tmap(&tmA_lo); prefetch_tmap(&tmG_hi); prefetch_tmap(&tmG_lo);
    for (int s = 0; s < stages; ++s) {
      mbar_init(bar_full + 8u * s, 1);
      mbar_init(bar_empty + 8u * s, 1);
    }
    mbar_init(bar_tfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(bar_empty + 8u * stage, phase ^ 1u, 0x500u + stage);
        const uint32_t full = bar_full + 8u * stage;
        const uint32_t sA = base + (uint32_t)stage * stage_bytes;
        const uint32_t sB = sA + 2u * a_bytes;
        const int r = r0 + kb * kDwBR;
        mbar_expect_tx(full, stage_bytes);
        for (int c = 0; c < 2; ++c) {
          tma_load_2d(sA + c * chunk, &tmA_hi, full, k0 + c * 64, r);
          tma_load_2d(sA + a_bytes + c * chunk, &tmA_lo, full, k0 + c * 64, r);
        }
        for (int c = 0; c < nb; ++c) {
          tma_load_2d(sB + c * chunk, &tmG_hi, full, n0 + c * 64, r);
          tma_load_2d(sB + b_bytes + c * chunk, &tmG_lo, full, n0 + c * 64, r);
        }
        if (++stage == stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && nkb > 0) {
      int stage = 0;
      uint32_t phase = 0;
      int n_cur = N - n0;
      n_cur = n_cur >= BN ? BN : ((n_cur + 15) & ~15);
      const uint32_t idesc = umma_idesc(n_cur, 1, 1);
      const uint32_t lbo_a = dbg.lbo_a ? dbg.lbo_a : chunk;
      const uint32_t lbo_b = dbg.lbo_b ? dbg.lbo_b : chunk;
      const uint32_t sbo = dbg.sbo ? dbg.sbo : 1024u;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(bar_full + 8u * stage, phase, 0x600u + stage);
        tc_fence_after();
        const uint32_t sA = base + (uint32_t)stage * stage_bytes;
        const uint32_t sB = sA + 2u * a_bytes;
#pragma unroll
        for (int j = 0; j < kDwBR / kUK; ++j) {
          const uint32_t ko = (uint32_t)j * (kUK * 128u);  // 16 batch rows further down the box
          const uint64_t a_hi = umma_desc(sA + ko, lbo_a, sbo);
          const uint64_t a_lo = umma_desc(sA + a_bytes + ko, lbo_a, sbo);
          const uint64_t g_hi = umma_desc(sB + ko, lbo_b, sbo);
          const uint64_t g_lo = umma_desc(sB + b_bytes + ko, lbo_b, sbo);
          tc_mma(tmem_base, a_hi, g_hi, idesc, (kb | j) != 0 ? 1u : 0u);
          tc_mma(tmem_base, a_lo, g_hi, idesc, 1u);
          tc_mma(tmem_base, a_hi, g_lo, idesc, 1u);
        }
        tc_commit(bar_empty + 8u * stage);
        if (kb == nkb - 1) tc_commit(bar_tfull);
        if (++stage == stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    const int q = warp & 3;
    const int krow = k0 + q * 32 + lane;
    const int n_tile = (N - n0) < BN ? (N - n0) : BN;
    float* out = partials + ((size_t)split * Kin + (size_t)krow) * N + n0;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(out) & 15u) == 0;
    if (nkb > 0) {
      mbar_wait(bar_tfull, 0, 0x700u);
      tc_fence_after();
    }
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int c0 = 0; c0 < n_tile; c0 += 32) {
      uint32_t r[32];
      if (nkb > 0) {
        tmem_ld32(t_row + (uint32_t)c0, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
      if (krow < Kin) {
        const int nv = (n_tile - c0) < 32 ? (n_tile - c0) : 32;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const int c = g * 4;
          if (c + 4 <= nv && vec_ok) {
            *reinterpret_cast<uint4*>(out + c0 + c) = make_uint4(r[c], r[c + 1], r[c + 2], r[c + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (c + j < nv) out[c0 + c + j] = __uint_as_float(r[c + j]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// dW[k,n] = sum_s P[s][k,n] in fixed order
__global__ void tc_dw_reduce_kernel(const float* __restrict__ P, int splits, int64_t KN,
                                    float* __restrict__ dW) {
  const int64_t i4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 >= KN) return;
  if (i4 + 4 <= KN && (KN & 3) == 0) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < splits; ++s) {
      const float4 p = __ldg(reinterpret_cast<const float4*>(P + (size_t)s * KN + i4));
      acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
    }
    *reinterpret_cast<float4*>(dW + i4) = acc;
  } else {
    for (int64_t i = i4; i < KN && i < i4 + 4; ++i) {
      float acc = 0.f;
      for (int s = 0; s < splits; ++s) acc += P[(size_t)s * KN + i];
      dW[i] = acc;
    }
  }
}

// out[n] = sum_t P[t][n] (per-tile column sums -> bias gradient), fixed order, coalesced over n
__global__ void tc_colsum_reduce_kernel(const float* __restrict__ P, int tiles, int N,
                                        float* __restrict__ out) {
  // blockDim = (32 columns, 8 row-slices): slice y sums tiles y, y+8, ...; then a fixed-order
  // combine over the 8 slices in shared memory.
  __shared__ float s[8][33];
  const int n = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (n < N)
    for (int t = threadIdx.y; t < tiles; t += 8) acc += P[(size_t)t * N + n];
  s[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) t += s[y][threadIdx.x];
    out[n] = t;
  }
}

// W fp32 [K,N] -> Wp planes [K, 2*ldn] (straight) and WTp planes [N, 2*ldk] (transposed)
__global__ void tc_prep_weight_kernel(const float* __restrict__ W, int K, int N,
                                      __nv_bfloat16* __restrict__ Wp, int64_t ldn,
                                      __nv_bfloat16* __restrict__ WTp, int64_t ldk) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  for (int y = threadIdx.y; y < 32; y += blockDim.y) {
    const int k = k0 + y, n = n0 + threadIdx.x;
    float w = 0.f;
    if (k < K && n < N) {
      w = W[(size_t)k * N + n];
      float hf;
      __nv_bfloat16 h, l;
      split_bf16(w, hf, h, l);
      if (Wp != nullptr) {
        Wp[(size_t)k * 2 * ldn + n] = h;
        Wp[(size_t)k * 2 * ldn + ldn + n] = l;
      }
    }
    tile[y][threadIdx.x] = w;
  }
  __syncthreads();
  if (WTp == nullptr) return;
  for (int y = threadIdx.y; y < 32; y += blockDim.y) {
    const int n = n0 + y, k = k0 + threadIdx.x;
    if (k < K && n < N) {
      float hf;
      __nv_bfloat16 h, l;
      split_bf16(tile[threadIdx.x][y], hf, h, l);
      WTp[(size_t)n * 2 * ldk + k] = h;
      WTp[(size_t)n * 2 * ldk + ldk + k] = l;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// bf16 matrix [rows, width] with row pitch `pitch_elems`; box = box_w x box_rows, 128B swizzle,
// out-of-bounds elements read as zero.
static int make_map(CUtensorMap* m, const void* base, int64_t width, int64_t rows,
                    int64_t pitch_elems, int box_w, int box_rows, int elem_bytes = 2) {
  const CUtensorMapSwizzle swz = box_w * elem_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                                           : CU_TENSOR_MAP_SWIZZLE_64B;
  EncodeTiledFn enc = encode_tiled_fn();
  if (enc == nullptr) {
    set_error("tc_gemm: cuTensorMapEncodeTiled is not available from this driver");
    return B200REC_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0 || (pitch_elems * elem_bytes) % 16 != 0) {
    set_error("tc_gemm: operand planes must be 16-byte aligned with a pitch multiple of 8 "
              "(base %p, pitch %lld)", base, (long long)pitch_elems);
    return B200REC_ERR_INVALID;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)width, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)pitch_elems * elem_bytes};
  const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(m, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                            : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                         2, const_cast<void*>(base), dims,
                         strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("tc_gemm: cuTensorMapEncodeTiled failed (%d) width %lld rows %lld pitch %lld box %dx%d",
              (int)r, (long long)width, (long long)rows, (long long)pitch_elems, box_w, box_rows);
    return B200REC_ERR_CUDA;
  }
  return B200REC_OK;
}

static int pick_bn(int N) {
  // <= 224 columns: two 128-byte-row stages + the output staging must fit 227 KB of shared memory
  constexpr int kCap = 224;
  const int tiles = (N + kCap - 1) / kCap;
  int bn = (N + tiles - 1) / tiles;
  bn = (bn + 15) & ~15;
  return bn < 16 ? 16 : bn;
}

static DwDebug g_dw_debug = {0u, 0u, 0u};
static int g_bn_override = 0;
static int g_two_cta = 1;     // K-major GEMM on CTA pairs (cta_group::2) when M >= 4096 (0: never)
static int g_tma_store = 1;   // epilogue outputs through TMA bulk stores (0: direct stores)
static int g_bk = 64;   // k-block of the K-major kernel: 64 (SWIZZLE_128B) or 32 (SWIZZLE_64B)

// D = A . B^T with A planes [M, 2*lda] (logical [M,K]) and B planes [N, 2*ldb] (logical [N,K]).
static int launch_gemm_kmajor(const void* A, int64_t lda, const void* B, int64_t ldb, int64_t M,
                              int N, int K, const Epilogue& ep, cudaStream_t st) {
  B200_REQUIRE(M >= 0 && N > 0 && K > 0, "tc_gemm: bad sizes M=%lld N=%d K=%d", (long long)M, N, K);
  B200_REQUIRE(M < (1ll << 31), "tc_gemm: M too large");
  if (M == 0) return B200REC_OK;
  const int BN = g_bn_override > 0 ? g_bn_override : pick_bn(N);
  const int BK = g_bk == 64 ? 64 : 32;
  // CTA pairs pay off when there are enough 256-row tiles to fill the 74 clusters
  const bool pair = g_two_cta != 0 && M >= 4096 && BN % 16 == 0 && ep.colsum == nullptr;
  const uint32_t stage_bytes =
      (2u * kBM + 2u * (uint32_t)(pair ? BN / 2 : BN)) * (uint32_t)BK * 2u;
  const uint32_t staging_bytes = kEpiWarps * kStagingPerWarp;
  int stages = (int)((kSmemBudget - 4096u - 16u * kMaxBN - staging_bytes) / stage_bytes);
  stages = stages > 8 ? 8 : stages;
  B200_REQUIRE(stages >= 2, "tc_gemm: tile does not fit shared memory");
  const size_t smem =
      (size_t)stages * stage_bytes + staging_bytes + 16 * stages + 64 + 16 * kMaxBN + 1024;
  const __nv_bfloat16* a = static_cast<const __nv_bfloat16*>(A);
  const __nv_bfloat16* b = static_cast<const __nv_bfloat16*>(B);
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  int rc;
  if ((rc = make_map(&ma_hi, a, K, M, 2 * lda, BK, kBM)) != B200REC_OK) return rc;
  if ((rc = make_map(&ma_lo, a + lda, K, M, 2 * lda, BK, kBM)) != B200REC_OK) return rc;
  const int b_rows = pair ? BN / 2 : BN;      // each CTA of a pair loads half of the B tile
  if ((rc = make_map(&mb_hi, b, K, N, 2 * ldb, BK, b_rows)) != B200REC_OK) return rc;
  if ((rc = make_map(&mb_lo, b + ldb, K, N, 2 * ldb, BK, b_rows)) != B200REC_OK) return rc;
  // outputs through TMA stores where base and pitch allow it (else: direct stores from registers)
  Epilogue e2 = ep;
  CUtensorMap mo_hi = ma_hi, mo_lo = ma_hi, mo_f32 = ma_hi;   // placeholders when unused
  e2.tma_planes = e2.tma_f32 = 0;
  if (g_tma_store && ep.colsum == nullptr) {
    if (ep.out_planes != nullptr && (reinterpret_cast<uintptr_t>(ep.out_planes) & 15u) == 0 &&
        ep.ldp % 8 == 0) {
      if ((rc = make_map(&mo_hi, ep.out_planes, N, M, 2 * ep.ldp, 32, 32)) != B200REC_OK) return rc;
      if ((rc = make_map(&mo_lo, ep.out_planes + ep.ldp, N, M, 2 * ep.ldp, 32, 32)) != B200REC_OK)
        return rc;
      e2.tma_planes = 1;
    } else if (ep.out_planes == nullptr && ep.out_f32 != nullptr &&
               (reinterpret_cast<uintptr_t>(ep.out_f32) & 15u) == 0 && ep.ld_f32 % 4 == 0) {
      if ((rc = make_map(&mo_f32, ep.out_f32, N, M, ep.ld_f32, 32, 32, 4)) != B200REC_OK) return rc;
      e2.tma_f32 = 1;
    }
  }
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA(cudaFuncSetAttribute(tc_gemm_kmajor_kernel<64>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    B200_CUDA(cudaFuncSetAttribute(tc_gemm_kmajor_kernel<32>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    B200_CUDA(cudaFuncSetAttribute(tc_gemm_kmajor2_kernel<64>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    B200_CUDA(cudaFuncSetAttribute(tc_gemm_kmajor2_kernel<32>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  if (pair) {
    const int pair_tiles = (int)((M + 2 * kBM - 1) / (2 * kBM)) * ((N + BN - 1) / BN);
    const int clusters = pair_tiles < sm_count() / 2 ? pair_tiles : sm_count() / 2;
    if (BK == 64)
      tc_gemm_kmajor2_kernel<64><<<2 * clusters, kThreadsK, smem, st>>>(
          ma_hi, ma_lo, mb_hi, mb_lo, mo_hi, mo_lo, mo_f32, (int)M, N, K, BN, stages, e2);
    else
      tc_gemm_kmajor2_kernel<32><<<2 * clusters, kThreadsK, smem, st>>>(
          ma_hi, ma_lo, mb_hi, mb_lo, mo_hi, mo_lo, mo_f32, (int)M, N, K, BN, stages, e2);
    B200_LAUNCH_CHECK();
    return B200REC_OK;
  }
  const int tiles = (int)((M + kBM - 1) / kBM) * ((N + BN - 1) / BN);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  if (BK == 64)
    tc_gemm_kmajor_kernel<64><<<grid, kThreadsK, smem, st>>>(ma_hi, ma_lo, mb_hi, mb_lo, mo_hi, mo_lo,
                                                             mo_f32, (int)M, N, K, BN, stages, e2);
  else
    tc_gemm_kmajor_kernel<32><<<grid, kThreadsK, smem, st>>>(ma_hi, ma_lo, mb_hi, mb_lo, mo_hi, mo_lo,
                                                             mo_f32, (int)M, N, K, BN, stages, e2);
  B200_LAUNCH_CHECK();
  return B200REC_OK;
}

struct DwPlan {
  int BN, tiles, splits, rows_per_split, stages;
  size_t smem, ws_bytes;
};

static DwPlan plan_dw(int64_t M, int K, int N) {
  DwPlan p;
  p.BN = g_bn_override > 0 ? g_bn_override : pick_bn(N);
  p.tiles = ((K + kBM - 1) / kBM) * ((N + p.BN - 1) / p.BN);
  int splits = sm_count() / p.tiles;
  splits = splits < 1 ? 1 : splits;
  int64_t rps = (M + splits - 1) / splits;
  rps = (rps + kDwBR - 1) / kDwBR * kDwBR;
  rps = rps < kDwBR ? kDwBR : rps;
  p.rows_per_split = (int)rps;
  p.splits = (int)((M + rps - 1) / rps);
  p.splits = p.splits < 1 ? 1 : p.splits;
  const int nb = (p.BN + 63) / 64;
  const uint32_t stage_bytes = (4u + 2u * nb) * kDwBR * 128u;
  p.stages = (int)((kSmemBudget - 4096u) / stage_bytes);
  p.stages = p.stages > 8 ? 8 : p.stages;
  p.smem = (size_t)p.stages * stage_bytes + 16 * p.stages + 64 + 1024;
  p.ws_bytes = (size_t)p.splits * K * N * sizeof(float);
  return p;
}

// dW[K,N] = A^T G with A planes [M, 2*lda] (logical [M,K]) and G planes [M, 2*ldg] (logical [M,N])
static int launch_gemm_dw(const void* A, int64_t lda, const void* G, int64_t ldg, int64_t M, int K,
                          int N, float* dW, void* ws, size_t ws_bytes, cudaStream_t st) {
  B200_REQUIRE(M >= 0 && N > 0 && K > 0, "tc_gemm_dw: bad sizes");
  B200_REQUIRE(M < (1ll << 31), "tc_gemm_dw: M too large");
  const int64_t KN = (int64_t)K * N;
  if (M == 0) {
    B200_CUDA(cudaMemsetAsync(dW, 0, (size_t)KN * sizeof(float), st));
    return B200REC_OK;
  }
  const DwPlan p = plan_dw(M, K, N);
  if (ws_bytes < p.ws_bytes) {
    set_error("tc_gemm_dw: workspace %zu < %zu bytes", ws_bytes, p.ws_bytes);
    return B200REC_ERR_WORKSPACE;
  }
  const __nv_bfloat16* a = static_cast<const __nv_bfloat16*>(A);
  const __nv_bfloat16* g = static_cast<const __nv_bfloat16*>(G);
  CUtensorMap ma_hi, ma_lo, mg_hi, mg_lo;
  int rc;
  if ((rc = make_map(&ma_hi, a, K, M, 2 * lda, 64, kDwBR)) != B200REC_OK) return rc;
  if ((rc = make_map(&ma_lo, a + lda, K, M, 2 * lda, 64, kDwBR)) != B200REC_OK) return rc;
  if ((rc = make_map(&mg_hi, g, N, M, 2 * ldg, 64, kDwBR)) != B200REC_OK) return rc;
  if ((rc = make_map(&mg_lo, g + ldg, N, M, 2 * ldg, 64, kDwBR)) != B200REC_OK) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA(cudaFuncSetAttribute(tc_gemm_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   227 * 1024));
    attr_set = true;
  }
  float* P = p.splits == 1 ? dW : static_cast<float*>(ws);
  dim3 grid(p.tiles, p.splits);
  tc_gemm_dw_kernel<<<grid, kThreads, p.smem, st>>>(ma_hi, ma_lo, mg_hi, mg_lo, (int)M, K, N, p.BN,
                                                    p.stages, p.rows_per_split, P, g_dw_debug);
  B200_LAUNCH_CHECK();
  if (p.splits > 1) {
    const int64_t n4 = (KN + 3) / 4;
    tc_dw_reduce_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(P, p.splits, KN, dW);
    B200_LAUNCH_CHECK();
  }
  return B200REC_OK;
}

}  // namespace tc
}  // namespace b200rec

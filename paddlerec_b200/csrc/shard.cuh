// K5 helpers: row-cyclic table sharding (owner = id mod world, local row = id div world).
//
// Behavioural spec: the PSGPU/HeterPS pull/push of the reference's GPUBox trainer
// (tools/static_gpubox_trainer.py:152-159,244-259): keys are sharded over the GPUs of one box,
// each batch exchanges keys -> owners and rows -> requesters.  The exchange itself is an NCCL
// all-to-all issued by the host (torch.distributed); this file produces the bucket order.
#pragma once

#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace b200rec {

__global__ void shard_keys_kernel(const int64_t* __restrict__ ids, int64_t n, int world, int64_t V,
                                  uint32_t* __restrict__ owner, int32_t* __restrict__ pos,
                                  unsigned long long* __restrict__ counts) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t id = ids[i];
  const bool in_range = (uint64_t)id < (uint64_t)V;
  const uint32_t o = in_range ? (uint32_t)(id % world) : 0u;
  owner[i] = o;
  pos[i] = (int32_t)i;
  // warp-aggregated histogram (integer atomics: order-independent result)
  const unsigned active = __activemask();
  const unsigned peers = __match_any_sync(active, o);
  if ((int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31))
    atomicAdd(counts + o, (unsigned long long)__popc(peers));
}

__global__ void shard_emit_kernel(const int64_t* __restrict__ ids,
                                  const int32_t* __restrict__ sorted_pos, int64_t n, int world,
                                  int64_t V, int64_t* __restrict__ send_ids,
                                  int64_t* __restrict__ perm, int32_t* __restrict__ inv_perm) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int32_t p = sorted_pos[k];
  const int64_t id = ids[p];
  const bool in_range = (uint64_t)id < (uint64_t)V;
  send_ids[k] = in_range ? id / world : (int64_t)-1;
  perm[p] = k;
  inv_perm[k] = p;
}

struct ShardPlan {
  size_t off_owner_in, off_owner_out, off_pos_in, off_pos_out, off_cub, cub_bytes, total;
  int bits;
};

static int make_shard_plan(int64_t n, int world, ShardPlan* p) {
  B200_REQUIRE(n >= 0 && n < (int64_t)INT32_MAX, "shard_bucketize: n must fit int32");
  B200_REQUIRE(world >= 1 && world <= 65536, "shard_bucketize: bad world=%d", world);
  int bits = 1;
  while ((1 << bits) < world) ++bits;
  p->bits = bits;
  size_t cub_bytes = 0;
  B200_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const uint32_t*)nullptr,
                                            (uint32_t*)nullptr, (const int32_t*)nullptr,
                                            (int32_t*)nullptr, (int)n, 0, bits));
  p->cub_bytes = cub_bytes;
  size_t off = 0;
  p->off_owner_in = off;  off += align_up((size_t)n * 4, 256);
  p->off_owner_out = off; off += align_up((size_t)n * 4, 256);
  p->off_pos_in = off;    off += align_up((size_t)n * 4, 256);
  p->off_pos_out = off;   off += align_up((size_t)n * 4, 256);
  p->off_cub = off;       off += align_up(cub_bytes, 256);
  p->total = off + 256;
  return B200REC_OK;
}

static int launch_shard_bucketize(const int64_t* ids, int64_t n, int world, int64_t V,
                                  int64_t* send_ids, int64_t* perm, int32_t* inv_perm,
                                  int64_t* counts, void* ws, size_t ws_bytes, cudaStream_t st) {
  ShardPlan p;
  int rc = make_shard_plan(n, world, &p);
  if (rc != B200REC_OK) return rc;
  B200_CUDA(cudaMemsetAsync(counts, 0, (size_t)world * sizeof(int64_t), st));
  if (n == 0) return B200REC_OK;
  if (ws_bytes < p.total) {
    set_error("shard_bucketize: workspace %zu < %zu bytes", ws_bytes, p.total);
    return B200REC_ERR_WORKSPACE;
  }
  unsigned char* base =
      reinterpret_cast<unsigned char*>(align_up((size_t)(uintptr_t)ws, 256));
  uint32_t* owner_in = reinterpret_cast<uint32_t*>(base + p.off_owner_in);
  uint32_t* owner_out = reinterpret_cast<uint32_t*>(base + p.off_owner_out);
  int32_t* pos_in = reinterpret_cast<int32_t*>(base + p.off_pos_in);
  int32_t* pos_out = reinterpret_cast<int32_t*>(base + p.off_pos_out);
  const unsigned grid = (unsigned)((n + 255) / 256);
  shard_keys_kernel<<<grid, 256, 0, st>>>(ids, n, world, V, owner_in, pos_in,
                                          reinterpret_cast<unsigned long long*>(counts));
  B200_LAUNCH_CHECK();
  size_t cub_bytes = p.cub_bytes;
  B200_CUDA(cub::DeviceRadixSort::SortPairs(base + p.off_cub, cub_bytes, owner_in, owner_out,
                                            pos_in, pos_out, (int)n, 0, p.bits, st));
  shard_emit_kernel<<<grid, 256, 0, st>>>(ids, pos_out, n, world, V, send_ids, perm, inv_perm);
  B200_LAUNCH_CHECK();
  return B200REC_OK;
}

}  // namespace b200rec

// uint64 feasign -> table row fold (include/b200rec.h: b200rec_hash_keys).  One key per thread,
// grid-stride; 8-byte loads/stores are fully coalesced (256 B per warp instruction), so the
// kernel streams at HBM rate: 16 bytes per key (+4 with a slot array).
#pragma once

#include "common.cuh"

namespace b200rec {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z ^= z >> 30;
  z *= 0xBF58476D1CE4E5B9ULL;
  z ^= z >> 27;
  z *= 0x94D049BB133111EBULL;
  z ^= z >> 31;
  return z;
}

__global__ void __launch_bounds__(256)
hash_keys_kernel(const uint64_t* __restrict__ keys, const int32_t* __restrict__ slot_of_key,
                 int64_t n, uint64_t V, int reserve_zero, int64_t* __restrict__ rows) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t k = keys[i];
    const uint64_t salt = slot_of_key ? (uint64_t)(slot_of_key[i] + 1) * 0x9E3779B97F4A7C15ULL : 0ULL;
    const uint64_t z = mix64(k ^ salt);
    int64_t r;
    if (reserve_zero) {
      r = (k == 0) ? 0 : (int64_t)(1 + z % (V - 1));
    } else {
      r = (int64_t)(z % V);
    }
    rows[i] = r;
  }
}

}  // namespace b200rec

// Warp-level transpose-reduce used by the fused InstanceNorm statistics.
#pragma once
#include <cuda_runtime.h>

namespace rehr {

// Sum v[0..15] of each of the 32 lanes column-wise: afterwards lane L holds the total of column
// ((L>>4)&1)*8 + ((L>>3)&1)*4 + ((L>>2)&1)*2 + ((L>>1)&1) in v[0] (both lanes of a pair hold it).
__device__ __forceinline__ void warp_colsum16(float (&v)[16], int lane) {
#pragma unroll
  for (int k = 8, mask = 16; k >= 1; k >>= 1, mask >>= 1) {
    const bool upper = (lane & mask) != 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < k) {
        const float send = upper ? v[i] : v[i + k];
        const float keep = upper ? v[i + k] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
      }
    }
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// column owned by an (even) lane after warp_colsum16
__device__ __forceinline__ int warp_colsum16_col(int lane) { return lane >> 1; }

}  // namespace rehr

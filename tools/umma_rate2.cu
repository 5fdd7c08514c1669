// Probe 2: minimal per-MMA issue cost.  Descriptors precomputed into registers; 16 MMAs fully unrolled inside ONE
// elect_one region per outer iteration.  Variants: (0) asm with "memory" clobber, (1) asm without clobber,
// (2) descriptors advanced by adding a constant to the low word (no precompute array).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include "../rehrseg_b200/csrc/ptx.cuh"
using namespace rehr;
struct P { int N, iters, mode; long long* out; };
__device__ __forceinline__ bool elect_one() { uint32_t p; asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(p)); return p != 0; }
__device__ __forceinline__ void umma_nc(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc));
}
__global__ void __launch_bounds__(128, 1) k(const P p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar; __shared__ uint64_t bar2[2]; __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2[0], 1); mbar_init(&bar2[1], 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&tslot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tbase = tslot;
  if (warp == 1) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const uint32_t idesc = make_idesc_bf16(128, p.N, 0, 0);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 48 * 1024;
    uint64_t ad[16], bd[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      ad[i] = make_smem_desc(a0 + ((i / 2) / 3 * 10 + (i / 2) % 3) * 64 + (i & 1) * 32, 0, 640, 4);
      bd[i] = make_smem_desc(b0 + (i / 2) * 6144 + (i & 1) * 32, 0, 512, 4);
    }
    long long t0 = clock64();
    for (int it = 0; it < p.iters; ++it) {
      if (elect_one()) {
        if (p.mode == 0) {
#pragma unroll
          for (int i = 0; i < 16; ++i) umma_bf16(tbase, ad[i], bd[i], idesc, 1u);
        } else if (p.mode == 1) {
#pragma unroll
          for (int i = 0; i < 16; ++i) umma_nc(tbase, ad[i], bd[i], idesc);
        } else if (p.mode >= 3) {
          // mode 3: 4 MMAs + one commit;  mode 4: 4 MMAs + two commits;  mode 5: 4 MMAs, no commit (baseline)
#pragma unroll
          for (int i = 0; i < 4; ++i) umma_bf16(tbase, ad[i], bd[i], idesc, 1u);
          if (p.mode == 3 || p.mode == 4) umma_commit(&bar2[0]);
          if (p.mode == 4) umma_commit(&bar2[1]);
        } else {
          uint64_t a = ad[0], b = bd[0];
#pragma unroll
          for (int i = 0; i < 16; ++i) { umma_nc(tbase, a, b, idesc); a += 2; b += 2; }
        }
      }
      __syncwarp();
    }
    long long t1 = clock64();
    if (elect_one()) umma_commit(&bar);
    mbar_wait(&bar, 0, nullptr, 1);
    long long t2 = clock64();
    if ((threadIdx.x & 31) == 0) { p.out[0] = t1 - t0; p.out[1] = t2 - t0; }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}
int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int mode : {0, 2, 3, 4, 5}) for (int N : {32, 96}) {
    P p; p.N = N; p.iters = 200; p.mode = mode; p.out = d;
    k<<<1, 128, 100 * 1024>>>(p);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("err\n"); return 1; }
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("mode=%d N=%3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (N/2=%d)\n", mode, N, h[0] / (mode >= 3 ? 800.0 : 3200.0), h[1] / (mode >= 3 ? 800.0 : 3200.0), N / 2);
  }
}

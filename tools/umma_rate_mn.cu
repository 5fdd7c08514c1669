// Probe 3: completion rate of tcgen05.mma with BOTH operands MN-major (the weight-gradient shape of wgrad_march.cu: K = voxels)
// when consecutive MMAs accumulate into (a) ONE TMEM accumulator, (b) `acc` accumulators in rotation.  If a dependent accumulate
// chain is slower than the issue rate, spreading the K steps of a step over several accumulators is the fix.
//   A: 32 channels (64-byte rows, 64B swizzle), 4 in-plane offsets stacked along M through LBO = one row, SBO = 10 rows (halo tile)
//   B: 32 channels (64-byte rows), N = 32 * planes, LBO = one plane slot (8 KB), SBO = 8 rows
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include "../rehrseg_b200/csrc/ptx.cuh"
using namespace rehr;
struct P { int N, iters, acc, kmajor; long long* out; };
__device__ __forceinline__ bool elect_one() { uint32_t p; asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(p)); return p != 0; }
__global__ void __launch_bounds__(128, 1) k(const P p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar; __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&tslot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tbase = tslot;
  if (warp == 1) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const uint32_t idesc = make_idesc_bf16(128, p.N, p.kmajor ? 0 : 1, p.kmajor ? 0 : 1);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 48 * 1024;
    uint64_t ad[8], bd[8];
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      if (p.kmajor) {
        ad[ks] = make_smem_desc(a0 + ks * 32, 0, 640, 4);
        bd[ks] = make_smem_desc(b0 + ks * 32, 0, 512, 4);
      } else {
        ad[ks] = make_smem_desc(a0 + 2 * ks * 10 * 64, 64, 640, 4);      // 16 voxels per K step = 2 groups of 8 halo rows
        bd[ks] = make_smem_desc(b0 + ks * 16 * 64, 8192, 512, 4);
      }
    }
    long long t0 = clock64();
    for (int it = 0; it < p.iters; ++it) {
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) umma_bf16(tbase + (uint32_t)((ks % p.acc) * 128), ad[ks], bd[ks], idesc, 1u);
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
  for (int kmajor : {1, 0}) for (int N : {32, 64, 96, 128}) for (int acc : {1, 2, 3, 4}) {
    P p; p.N = N; p.iters = 400; p.acc = acc; p.kmajor = kmajor; p.out = d;
    k<<<1, 128, 100 * 1024>>>(p);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("err\n"); return 1; }
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("%s N=%3d accumulators=%d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (tensor floor N/2=%d)\n", kmajor ? "K-major " : "MN-major", N, acc,
           h[0] / 3200.0, h[1] / 3200.0, N / 2);
  }
}

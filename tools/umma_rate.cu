// Hardware probe: tcgen05.mma issue/execute rate (cycles per MMA) for M=128, K=16, various N, with the A descriptor
// (a) atom-aligned, (b) row-shifted, (c) SBO = 10 rows, for 64 B and 128 B swizzled rows.  One CTA, one issuing thread,
// ITER back-to-back MMAs into the same accumulator, timed with clock64 around issue..commit completion.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstring>
#include "../rehrseg_b200/csrc/ptx.cuh"
using namespace rehr;

struct P { int rowb, shift, grp, N, iters, kadv, mode; long long* out; };
__device__ __forceinline__ bool elect_one() { uint32_t p; asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(p)); return p != 0; }

__global__ void __launch_bounds__(128, 1) rate_kernel(const P p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&tslot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tbase = tslot;
  if (p.mode == 0 && threadIdx.x == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const uint32_t layout = swizzle_layout_for_bytes(p.rowb);
    const uint32_t idesc = make_idesc_bf16(128, p.N, 0, 0);
    const uint32_t a0 = smem_u32(smem) + p.shift * p.rowb;
    const uint32_t b0 = smem_u32(smem) + 48 * 1024;
    const int ksteps = p.rowb / 32;
    long long t0 = clock64();
    for (int it = 0; it < p.iters; ++it) {
      const int k = p.kadv ? (it % ksteps) : 0;
      const uint32_t tap = p.kadv ? ((it / ksteps) % 9) : 0;
      const uint32_t aoff = ((tap / 3) * p.grp + tap % 3) * p.rowb;
      const uint64_t ad = make_smem_desc(a0 + aoff + k * 32, 0, p.grp * p.rowb, layout);
      const uint64_t bd = make_smem_desc(b0 + k * 32, 0, 8 * p.rowb, layout);
      umma_bf16(tbase, ad, bd, idesc, 1u);
    }
    long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0, nullptr, 1);
    long long t2 = clock64();
    p.out[0] = t1 - t0;
    p.out[1] = t2 - t0;
  }
  if (p.mode == 1 && warp == 1) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const uint32_t layout = swizzle_layout_for_bytes(p.rowb);
    const uint32_t idesc = make_idesc_bf16(128, p.N, 0, 0);
    const uint32_t a0 = smem_u32(smem) + p.shift * p.rowb;
    const uint32_t b0 = smem_u32(smem) + 48 * 1024;
    const int ksteps = p.rowb / 32;
    long long t0 = clock64();
    for (int tap = 0, it = 0; it < p.iters; ++tap) {
      if (tap == 9) tap = 0;
      const uint32_t aoff = p.kadv ? ((tap / 3) * p.grp + tap % 3) * p.rowb : 0;
      for (int k = 0; k < ksteps && it < p.iters; ++k, ++it) {
        const uint64_t ad = make_smem_desc(a0 + aoff + (p.kadv ? k * 32 : 0), 0, p.grp * p.rowb, layout);
        const uint64_t bd = make_smem_desc(b0 + (p.kadv ? k * 32 : 0), 0, 8 * p.rowb, layout);
        if (elect_one()) umma_bf16(tbase, ad, bd, idesc, 1u);
      }
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
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 2000;
  for (int mode : {0, 1}) for (int rowb : {64}) for (int N : {32, 96, 192, 256}) for (int cfg : {0, 4}) {
    P p; p.mode = mode; p.rowb = rowb; p.N = N; p.iters = iters; p.out = d;
    p.shift = (cfg == 1 || cfg >= 3) ? 1 : 0; p.grp = (cfg >= 2) ? 10 : 8; p.kadv = cfg == 4;
    rate_kernel<<<1, 128, 100 * 1024>>>(p);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("mode=%d rowb=%3d N=%3d shift=%d grp=%2d kadv=%d : issue %.1f cyc/MMA, complete %.1f cyc/MMA (ideal N/2=%d, A-read 32)\n", mode, rowb, N, p.shift, p.grp, p.kadv,
           (double)h[0] / iters, (double)h[1] / iters, N / 2);
  }
  return 0;
}

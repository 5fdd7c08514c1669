// Hardware probe: can a K-major swizzled UMMA A-descriptor start at an arbitrary ROW of a TMA-written tile
// (row shift not a multiple of 8) and use an SBO that is not a multiple of the swizzle atom?  This is what the
// halo-resident conv kernel needs (tap shifts become descriptor offsets).
//   A tile: G[R=200 rows][C ch] bf16 loaded by ONE TMA box with swizzle S (C=64 -> 128B, C=32 -> 64B, C=16 -> 32B)
//   B tile: identity [N=C][K=C]  => D[m][n] = A_view[m][n]
//   A_view[m] := G[(m/8)*grp + m%8 + shift]   (grp = 10 rows => SBO = 10*rowbytes)
// Prints, per (C, shift, base_offset mode), the number of mismatching elements.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../rehrseg_b200/csrc/ptx.cuh"
using namespace rehr;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct P {
  CUtensorMap a_map, b_map;
  int C, rows, shift, grp, bo_mode, N;
  float* out;  // [128][N]
};

__device__ __forceinline__ uint64_t make_desc_bo(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout, uint32_t bo) {
  uint64_t d = make_smem_desc(saddr, lbo, sbo, layout);
  d |= (uint64_t)(bo & 7) << 49;
  return d;
}

__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ P p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rowb = p.C * 2;
  uint8_t* sa = smem;
  uint8_t* sb = smem + 64 * 1024;
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&tslot, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tslot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar[0], p.rows * rowb + p.N * rowb);
    tma_load_2d(&p.a_map, &bar[0], sa, 0, 0);
    tma_load_2d(&p.b_map, &bar[0], sb, 0, 0);
    mbar_wait(&bar[0], 0, nullptr, 1);
    tc_fence_after();
    const uint32_t layout = swizzle_layout_for_bytes((int)rowb);
    const uint32_t idesc = make_idesc_bf16(128, p.N, 0, 0);
    const uint32_t a0 = smem_u32(sa) + p.shift * rowb;
    const uint32_t pat = rowb * 8;  // swizzle repeat (1024 / 512 / 256)
    for (int k = 0; k < p.C / 16; ++k) {
      uint32_t bo = 0;
      if (p.bo_mode == 1) bo = ((a0 % pat) / rowb) & 7;          // row phase of the start inside the pattern
      if (p.bo_mode == 2) bo = (a0 >> 7) & 7;                   // PTX text formula
      const uint64_t ad = make_desc_bo(a0 + k * 32, 0, p.grp * rowb, layout, bo);
      const uint64_t bd = make_smem_desc(smem_u32(sb) + k * 32, 0, 8 * rowb, layout);
      umma_bf16(tbase, ad, bd, idesc, k > 0);
    }
    umma_commit(&bar[1]);
  }
  mbar_wait(&bar[1], 0, nullptr, 2);
  tc_fence_after();
  for (int c0 = 0; c0 < p.N; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tbase + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) p.out[(warp * 32 + lane) * p.N + c0 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 64);
}

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  const int R = 200;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int C : {64, 32, 16}) {
    std::vector<__nv_bfloat16> hG(R * C), hI(C * C);
    // value = row + chan/64 is exact in bf16 only for small ranges; use row*... keep <= 8 bits: encode row in value, chan via identity position
    for (int r = 0; r < R; ++r)
      for (int c = 0; c < C; ++c) hG[r * C + c] = __float2bfloat16((float)((r * 7 + c * 3) % 251));
    for (int n = 0; n < C; ++n)
      for (int k = 0; k < C; ++k) hI[n * C + k] = __float2bfloat16(n == k ? 1.f : 0.f);
    __nv_bfloat16 *dG, *dI;
    float* dO;
    cudaMalloc(&dG, hG.size() * 2);
    cudaMalloc(&dI, hI.size() * 2);
    cudaMalloc(&dO, 128 * C * 4);
    cudaMemcpy(dG, hG.data(), hG.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dI, hI.data(), hI.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMapSwizzle sw = C == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (C == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    P p;
    memset(&p, 0, sizeof(p));
    {
      cuuint64_t gd[2] = {(cuuint64_t)C, (cuuint64_t)R};
      cuuint64_t gs[1] = {(cuuint64_t)C * 2};
      cuuint32_t bd[2] = {(cuuint32_t)C, (cuuint32_t)R};
      cuuint32_t es[2] = {1, 1};
      CUresult r = enc(&p.a_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dG, gd, gs, bd, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode A failed %d\n", (int)r); return 1; }
      cuuint64_t gd2[2] = {(cuuint64_t)C, (cuuint64_t)C};
      cuuint32_t bd2[2] = {(cuuint32_t)C, (cuuint32_t)C};
      r = enc(&p.b_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dI, gd2, gs, bd2, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode B failed %d\n", (int)r); return 1; }
    }
    p.C = C; p.rows = R; p.N = C; p.out = dO;
    for (int grp : {8, 10}) {
      for (int shift : {0, 1, 2, 3, 5, 8, 11, 21}) {
        for (int bo_mode : {0, 1, 2}) {
          p.shift = shift; p.grp = grp; p.bo_mode = bo_mode;
          cudaMemset(dO, 0, 128 * C * 4);
          probe_kernel<<<1, 128, 100 * 1024>>>(p);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("C=%d grp=%d shift=%d bo=%d: CUDA error %s\n", C, grp, shift, bo_mode, cudaGetErrorString(e)); return 2; }
          std::vector<float> hO(128 * C);
          cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost);
          int bad = 0;
          for (int m = 0; m < 128; ++m) {
            const int r = (m / 8) * grp + m % 8 + shift;
            for (int c = 0; c < C; ++c) {
              const float want = r < R ? (float)((r * 7 + c * 3) % 251) : 0.f;
              if (hO[m * C + c] != want) ++bad;
            }
          }
          printf("C=%2d grp=%2d shift=%2d bo_mode=%d mismatches=%d/%d\n", C, grp, shift, bo_mode, bad, 128 * C);
        }
      }
    }
    cudaFree(dG); cudaFree(dI); cudaFree(dO);
  }
  return 0;
}

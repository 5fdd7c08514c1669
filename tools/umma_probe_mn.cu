// Hardware probe 3: MN-major swizzled UMMA descriptors for the weight-gradient kernel.
//   A (M side)  = X halo tile G[rows = voxels][C ch], TMA-written with swizzle S(C): M = 128 = (128/C) "atoms" of C channels,
//                 atom j starts LBO bytes after atom j-1 (LBO = lbo_rows rows: overlapping, row-shifted views = the kw taps);
//                 K = 16 voxels = 2 groups of 8 rows, groups SBO = grp rows apart.
//   B (N side)  = identity I[k][n] (16 x 16), MN-major SW32  =>  D[m][n] = A_view[k = n][m]
//   expected: A_view[k][m] = G[(k/8)*grp + k%8 + shift + (m/C)*lbo_rows][m % C]
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstring>
#include <vector>
#include "../rehrseg_b200/csrc/ptx.cuh"
using namespace rehr;
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
struct P { CUtensorMap a_map, b_map; int C, rows, shift, grp, lbo_rows; float* out; };

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ P p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rowb = p.C * 2;
  uint8_t* sa = smem;
  uint8_t* sb = smem + 64 * 1024;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&tslot, 32); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tbase = tslot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar[0], p.rows * rowb + 16 * 32);
    tma_load_2d(&p.a_map, &bar[0], sa, 0, 0);
    tma_load_2d(&p.b_map, &bar[0], sb, 0, 0);
    mbar_wait(&bar[0], 0, nullptr, 1);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, 16, 1, 1);
    const uint64_t ad = make_smem_desc(smem_u32(sa) + p.shift * rowb, p.lbo_rows * rowb, p.grp * rowb, swizzle_layout_for_bytes((int)rowb));
    const uint64_t bd = make_smem_desc(smem_u32(sb), 0, 8 * 32, 6);
    umma_bf16(tbase, ad, bd, idesc, 0);
    umma_commit(&bar[1]);
  }
  mbar_wait(&bar[1], 0, nullptr, 2);
  tc_fence_after();
  uint32_t v[16];
  tmem_ld16(tbase + ((uint32_t)(warp * 32) << 16), v);
  tmem_ld_wait();
  for (int i = 0; i < 16; ++i) p.out[(warp * 32 + lane) * 16 + i] = __uint_as_float(v[i]);
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 32);
}

int main() {
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  const int R = 200;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int C : {32, 64, 16}) {
    std::vector<__nv_bfloat16> hG(R * C), hI(16 * 16);
    for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) hG[r * C + c] = __float2bfloat16((float)((r * 7 + c * 3) % 251));
    for (int k = 0; k < 16; ++k) for (int n = 0; n < 16; ++n) hI[k * 16 + n] = __float2bfloat16(k == n ? 1.f : 0.f);
    __nv_bfloat16 *dG, *dI; float* dO;
    cudaMalloc(&dG, hG.size() * 2); cudaMalloc(&dI, hI.size() * 2); cudaMalloc(&dO, 128 * 16 * 4);
    cudaMemcpy(dG, hG.data(), hG.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dI, hI.data(), hI.size() * 2, cudaMemcpyHostToDevice);
    P p; memset(&p, 0, sizeof(p));
    CUtensorMapSwizzle sw = C == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (C == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    cuuint64_t gd[2] = {(cuuint64_t)C, (cuuint64_t)R}; cuuint64_t gs[1] = {(cuuint64_t)C * 2};
    cuuint32_t bd[2] = {(cuuint32_t)C, (cuuint32_t)R}; cuuint32_t es[2] = {1, 1};
    if (enc(&p.a_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dG, gd, gs, bd, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("enc A fail\n"); return 1; }
    cuuint64_t gd2[2] = {16, 16}; cuuint64_t gs2[1] = {32}; cuuint32_t bd2[2] = {16, 16};
    if (enc(&p.b_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dI, gd2, gs2, bd2, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("enc B fail\n"); return 1; }
    p.C = C; p.rows = R; p.out = dO;
    for (int grp : {8, 10}) for (int shift : {0, 1, 3, 10, 21}) for (int lbo_rows : {1, 2, 8, 17, 64}) {
      p.shift = shift; p.grp = grp; p.lbo_rows = lbo_rows;
      cudaMemset(dO, 0, 128 * 16 * 4);
      probe<<<1, 128, 100 * 1024>>>(p);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("C=%d: CUDA error %s\n", C, cudaGetErrorString(e)); return 2; }
      std::vector<float> hO(128 * 16);
      cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0, total = 0;
      for (int m = 0; m < 128; ++m) for (int k = 0; k < 16; ++k) {
        const int r = (k / 8) * grp + k % 8 + shift + (m / C) * lbo_rows;
        if (r >= R) continue;
        ++total;
        if (hO[m * 16 + k] != (float)((r * 7 + (m % C) * 3) % 251)) ++bad;
      }
      printf("C=%2d grp=%2d shift=%2d lbo_rows=%2d mismatches=%d/%d\n", C, grp, shift, lbo_rows, bad, total);
    }
    cudaFree(dG); cudaFree(dI); cudaFree(dO);
  }
  return 0;
}

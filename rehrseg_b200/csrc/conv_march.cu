// "Marching" convolution kernel for sm_100a: k = 3x3x3, stride 1, pad 1, few channels (the layers that hold
// 3/4 of the nnU-Net FLOPs: 32->32, 64->32, 64->64, 128->64 at 128^3 / 64^3 -- SURVEY.md section 7.3).
//
// Why a second kernel: with Cout = 32..64 the generic tapped GEMM (conv_engine.cu) re-fetches every input voxel
// 27 times from L2 (one TMA box per tap) and is L2->SMEM bound at ~10 % of tensor peak.  Here
//   * a CTA owns a column of output voxels: a 16(h) x 8(w) in-plane tile, marched along d for a segment of planes;
//   * each input plane (18 x 10 halo'd voxels x Cin, ONE TMA box per 64-channel chunk) is loaded once into a shared
//     memory ring; the 9 in-plane taps (kh,kw) are NOT re-loaded: they are UMMA descriptor offsets into that tile
//     (row shift kh*10+kw, 8-row groups 10 rows apart).  Verified on B200: swizzled K-major descriptors use absolute
//     smem address bits, so arbitrary row shifts are legal (profiles/r01_umma_shifted_descriptor_probe.log);
//   * the 3 depth taps are fused into ONE tcgen05.mma: B = [W(kd=2) | W(kd=1) | W(kd=0)] (N = 3*Ct) and the
//     accumulators of output planes p-1, p, p+1 sit in adjacent TMEM column slots, so input plane p is read from
//     shared memory once per (kh,kw,k-step) instead of three times.  This lifts the A-operand SMEM read bound
//     (128x16 bf16 = 4 KB per MMA at 128 B/clk = 32 clk) above the MMA time (N/2 = 48 clk for Ct = 32);
//   * all 27*Cin*Ct weights stay resident in shared memory (Ct = output-channel tile chosen to fit);
//   * epilogue warps drain one finished plane at a time (bias, InstanceNorm partial sums, activation, bf16 store),
//     re-zero the TMEM slot (every MMA accumulates; there is no per-column "overwrite" flag) and hand it back.
// The same kernel is the input-gradient of those layers (weights packed transposed + tap-flipped).
//
// Reference call sites replaced: ConvDropoutNormReLU.conv of the full-resolution nnU-Net stages (built at
// models/seg_model.py:174-191), sr_head.0 (models/seg_model.py:197), FLAVR Conv3DSimple k3 (resnet_3D.py:19-33).
#include "engine.h"
#include "ptx.cuh"
#include "reduce.cuh"

#include <algorithm>
#include <cstring>
#include <mutex>

namespace rehr {

int encode_tiled_bf16(CUtensorMap* m, const void* base, int rank, const unsigned long long* gdim,
                      const unsigned long long* gstride_bytes, const unsigned* box, int swizzle_bytes);

static constexpr int kMarchThreads = 192;  // warp0 TMA, warp1 MMA (+TMEM alloc), warps 2..5 epilogue
static constexpr int kTileH = 16, kTileW = 8;
static constexpr int kHaloH = kTileH + 2, kHaloW = kTileW + 2, kHaloRows = kHaloH * kHaloW;  // 18 x 10 = 180
static constexpr int kMaxRing = 8;
static constexpr int kMaxSlots = 16;

struct alignas(64) MarchParams {
  CUtensorMap x_map;  // 5-D NDHWC, box (BK, 10, 18, 1, 1)
  CUtensorMap w_map;  // 2-D [n_ct*9*chunks*3Ct rows][BK], box (BK, 3Ct)
  int N, D, H, W, Cin, Cout;
  int Ct, n_ct, BK, chunks;
  int tiles_h, tiles_w, Ds, n_seg;
  int ring, slots;
  int total_items, items_per_ct;
  uint32_t w_bytes, plane_bytes, chunk_stride, slot_stride, wtile_bytes;
  void* out;
  int out_f32;
  long long out_ld;
  const float* bias;
  int act;
  float slope;
  float* stats;  // [N][tiles_per_sample][Cout][2] or null
  int* err;
};

__device__ __forceinline__ float march_act(float v, int act, float slope) {
  if (act == REHR_ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == REHR_ACT_LRELU) return v > 0.f ? v : v * slope;
  return v;
}

struct ItemCoord {
  int ct, n, seg, th, tw;
};
__device__ __forceinline__ ItemCoord decode_item(const MarchParams& p, int item) {
  ItemCoord c;
  c.ct = item / p.items_per_ct;
  int r = item - c.ct * p.items_per_ct;
  c.tw = r % p.tiles_w;
  r /= p.tiles_w;
  c.th = r % p.tiles_h;
  r /= p.tiles_h;
  c.seg = r % p.n_seg;
  c.n = r / p.n_seg;
  return c;
}

__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}

// BKT = channels per K chunk (16 / 32 / 64 -> 32 / 64 / 128 B swizzled rows), CHUNKS = Cin / BKT.  Compile-time so that
// the MMA issue sequence of one input plane (9 * CHUNKS * BKT/16 instructions) is fully unrolled with constant
// descriptor increments: a single thread must issue one tcgen05.mma every ~50 clk (tools/umma_rate2.cu measures
// 40 clk/MMA for this code shape vs 120-280 clk/MMA with run-time descriptor arithmetic).
template <int BKT, int CHUNKS>
__global__ void __launch_bounds__(kMarchThreads, 1) conv_march_kernel(const __grid_constant__ MarchParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_w = smem;                                   // resident weights
  uint8_t* s_ring = smem + ((p.w_bytes + 1023u) & ~1023u);  // input plane ring
  uint8_t* tail = s_ring + (size_t)p.ring * p.slot_stride;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);  // [ring]
  uint64_t* empty_bar = full_bar + kMaxRing;               // [ring]
  uint64_t* tfull_bar = empty_bar + kMaxRing;              // [slots]
  uint64_t* tempty_bar = tfull_bar + kMaxSlots;            // [slots]
  uint64_t* wfull_bar = tempty_bar + kMaxSlots;            // [1]
  uint64_t* wfree_bar = wfull_bar + 1;                     // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfree_bar + 1);
  float* part = reinterpret_cast<float*>(tmem_slot + 4);   // [4 warps][2][64]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.x_map);
    tma_prefetch_desc(&p.w_map);
    for (int i = 0; i < p.ring; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < p.slots; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    mbar_init(wfull_bar, 1);
    mbar_init(wfree_bar, 1);
    fence_mbar_init();
  } else if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t rowb = (uint32_t)p.BK * 2u;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int cur_ct = -1;
      uint32_t wfree_phase = 0;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        const ItemCoord c = decode_item(p, item);
        if (c.ct != cur_ct) {
          if (cur_ct >= 0) {  // wait until every MMA that reads the old weights has retired
            mbar_wait(wfree_bar, wfree_phase, p.err, 21);
            wfree_phase ^= 1u;
          }
          cur_ct = c.ct;
          mbar_arrive_expect_tx(wfull_bar, p.w_bytes);
          const int tiles = 9 * p.chunks;
          for (int t = 0; t < tiles; ++t)
            tma_load_2d(&p.w_map, wfull_bar, s_w + (size_t)t * p.wtile_bytes, 0, (c.ct * tiles + t) * 3 * p.Ct);
        }
        const int d0 = c.seg * p.Ds, d1 = min(p.D, d0 + p.Ds);
        const int pa = max(d0 - 1, 0), pb = min(d1, p.D - 1);
        const int h0 = c.th * kTileH - 1, w0 = c.tw * kTileW - 1;
        for (int pl = pa; pl <= pb; ++pl) {
          mbar_wait(&empty_bar[stage], phase ^ 1u, p.err, 22);
          mbar_arrive_expect_tx(&full_bar[stage], p.plane_bytes);
          uint8_t* dst = s_ring + (size_t)stage * p.slot_stride;
          for (int ch = 0; ch < p.chunks; ++ch)
            tma_load_5d(&p.x_map, &full_bar[stage], dst + (size_t)ch * p.chunk_stride, ch * p.BK, w0, h0, pl, c.n);
          if (++stage == p.ring) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    // The whole warp runs the (warp-uniform) control flow and barrier waits; one elected lane issues the MMAs.
    constexpr uint32_t kRowB = BKT * 2;
    constexpr int kSteps = BKT / 16;
    constexpr uint32_t kLayout = kRowB == 128 ? 2u : (kRowB == 64 ? 4u : 6u);
    // high words of the K-major descriptors: SBO (bits 32..45), version 1 (bit 46), swizzle mode (bits 61..63)
    constexpr uint32_t kAHi = ((kHaloW * kRowB) >> 4) | (1u << 14) | (kLayout << 29);
    constexpr uint32_t kBHi = ((8u * kRowB) >> 4) | (1u << 14) | (kLayout << 29);
    const uint32_t sw_lo = smem_u32(s_w) >> 4, sring_lo = smem_u32(s_ring) >> 4;
    const uint32_t slot_lo = p.slot_stride >> 4, chunk_lo = p.chunk_stride >> 4, wtile_lo = p.wtile_bytes >> 4;
    int stage = 0;
    uint32_t phase = 0;
    int cur_ct = -1;
    uint32_t wfull_phase = 0;
    uint32_t tempty_par = 0;  // bit s = parity to wait for on tempty_bar[s]
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      const ItemCoord c = decode_item(p, item);
      if (c.ct != cur_ct) {
        cur_ct = c.ct;
        mbar_wait(wfull_bar, wfull_phase, p.err, 31);
        wfull_phase ^= 1u;
        tc_fence_after();
      }
      const int d0 = c.seg * p.Ds, d1 = min(p.D, d0 + p.Ds);
      const int pa = max(d0 - 1, 0), pb = min(d1, p.D - 1);
      int next_open = d0;
      for (int pl = pa; pl <= pb; ++pl) {
        const int qa = max(pl - 1, d0), qb = min(pl + 1, d1 - 1);
        while (next_open <= qb) {  // first touch of an output plane's TMEM slot: wait until it was drained + zeroed
          const int s = (next_open - d0) % p.slots;
          mbar_wait(&tempty_bar[s], (tempty_par >> s) & 1u, p.err, 32);
          tempty_par ^= 1u << s;
          ++next_open;
        }
        mbar_wait(&full_bar[stage], phase, p.err, 33);
        tc_fence_after();
        const uint32_t a_lo = sring_lo + (uint32_t)stage * slot_lo;
        // contiguous TMEM slot runs covering output planes qa..qb (one run unless the slot ring wraps)
        int ra = qa;
        while (ra <= qb) {
          int rb = ra;
          while (rb < qb && ((rb + 1 - d0) % p.slots) != 0) ++rb;
          const uint32_t idesc = make_idesc_bf16(128, (rb - ra + 1) * p.Ct, 0, 0);
          const uint32_t b_lo = sw_lo + (((uint32_t)((ra - pl + 1) * p.Ct) * kRowB) >> 4);
          const uint32_t d_tmem = tmem_base + (uint32_t)(((ra - d0) % p.slots) * p.Ct);
          if (elect_one_sync()) {
#pragma unroll
            for (int khw = 0; khw < 9; ++khw) {
#pragma unroll
              for (int ch = 0; ch < CHUNKS; ++ch) {
                const uint32_t a_t = a_lo + (uint32_t)((((khw / 3) * kHaloW + (khw % 3)) * kRowB) >> 4) + (uint32_t)ch * chunk_lo;
                const uint32_t b_t = b_lo + (uint32_t)(khw * CHUNKS + ch) * wtile_lo;
#pragma unroll
                for (int k = 0; k < kSteps; ++k) {
                  const uint64_t ad = ((uint64_t)kAHi << 32) | (uint64_t)(a_t + 2u * k);
                  const uint64_t bd = ((uint64_t)kBHi << 32) | (uint64_t)(b_t + 2u * k);
                  umma_bf16(d_tmem, ad, bd, idesc, 1u);
                }
              }
            }
          }
          __syncwarp();
          ra = rb + 1;
        }
        if (elect_one_sync()) {
          umma_commit(&empty_bar[stage]);
          // finished output planes
          if (pl - 1 >= d0) umma_commit(&tfull_bar[(pl - 1 - d0) % p.slots]);
          if (pl == pb) {
            for (int q = max(pl, d0); q <= d1 - 1; ++q) umma_commit(&tfull_bar[(q - d0) % p.slots]);
          }
        }
        __syncwarp();
        if (++stage == p.ring) {
          stage = 0;
          phase ^= 1u;
        }
      }
      // tell the producer the weights may be replaced (only signalled when the next item needs other weights)
      const int nitem = item + gridDim.x;
      if (nitem < p.total_items && nitem / p.items_per_ct != c.ct) {
        if (elect_one_sync()) umma_commit(wfree_bar);
        __syncwarp();
      }
    }
  } else {
    // ============================== epilogue (warps 2..5) ==============================
    const int q4 = warp & 3;
    const int row = q4 * 32 + lane;
    const int hl = row >> 3, wl = row & 7;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q4 * 32) << 16);
    // zero the whole accumulator space once; afterwards every drained slot is re-zeroed
    {
      uint32_t z[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) z[i] = 0u;
      for (int c0 = 0; c0 < 512; c0 += 16) tmem_st16(lane_addr + (uint32_t)c0, z);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0)
        for (int s = 0; s < p.slots; ++s) mbar_arrive(&tempty_bar[s]);
    }
    uint32_t tfull_par = 0;
    const int et = threadIdx.x - 64;
    const int nchunk = p.Ct / 16;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      const ItemCoord c = decode_item(p, item);
      const int d0 = c.seg * p.Ds, d1 = min(p.D, d0 + p.Ds);
      const int oh = c.th * kTileH + hl, ow = c.tw * kTileW + wl;
      const bool valid = oh < p.H && ow < p.W;
      const int cbase = c.ct * p.Ct;
      float ts1[4], ts2[4];  // running column sums owned by this lane pair (Ct <= 64 -> 4 chunks of 16)
#pragma unroll
      for (int i = 0; i < 4; ++i) ts1[i] = ts2[i] = 0.f;
      for (int q = d0; q < d1; ++q) {
        const int s = (q - d0) % p.slots;
        mbar_wait(&tfull_bar[s], (tfull_par >> s) & 1u, p.err, 41);
        tfull_par ^= 1u << s;
        tc_fence_after();
        const uint32_t taddr = lane_addr + (uint32_t)(s * p.Ct);
        const long long vox = (((long long)c.n * p.D + q) * p.H + oh) * p.W + ow;
#pragma unroll 1
        for (int ci = 0; ci < nchunk; ++ci) {
          uint32_t v[16];
          tmem_ld16(taddr + (uint32_t)(ci * 16), v);
          tmem_ld_wait();
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float x = __uint_as_float(v[i]);
            if (p.bias != nullptr) x += __ldg(p.bias + cbase + ci * 16 + i);
            f[i] = x;
          }
          if (p.stats != nullptr) {
            float s1[16], s2[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float x = valid ? f[i] : 0.f;
              s1[i] = x;
              s2[i] = x * x;
            }
            warp_colsum16(s1, lane);
            warp_colsum16(s2, lane);
            // accumulate in a fixed register (compile-time index) per chunk
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (k == ci) {
                ts1[k] += s1[0];
                ts2[k] += s2[0];
              }
          }
          if (valid) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = march_act(f[i], p.act, p.slope);
            const int cc = cbase + ci * 16;
            if (p.out_f32) {
              float* o = reinterpret_cast<float*>(p.out) + vox * p.out_ld + cc;
              if (cc + 16 <= p.Cout && (p.out_ld & 3) == 0) {
#pragma unroll
                for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
              } else {
                for (int i = 0; i < 16; ++i)
                  if (cc + i < p.Cout) o[i] = f[i];
              }
            } else {
              __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + vox * p.out_ld + cc;
              if (cc + 16 <= p.Cout && (p.out_ld & 7) == 0) {
                uint4 lo, hi;
                lo.x = pack_bf16x2(f[0], f[1]);
                lo.y = pack_bf16x2(f[2], f[3]);
                lo.z = pack_bf16x2(f[4], f[5]);
                lo.w = pack_bf16x2(f[6], f[7]);
                hi.x = pack_bf16x2(f[8], f[9]);
                hi.y = pack_bf16x2(f[10], f[11]);
                hi.z = pack_bf16x2(f[12], f[13]);
                hi.w = pack_bf16x2(f[14], f[15]);
                reinterpret_cast<uint4*>(o)[0] = lo;
                reinterpret_cast<uint4*>(o)[1] = hi;
              } else {
                for (int i = 0; i < 16; ++i)
                  if (cc + i < p.Cout) o[i] = __float2bfloat16(f[i]);
              }
            }
          }
          // re-zero the chunk just read
          uint32_t z[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) z[i] = 0u;
          tmem_st16(taddr + (uint32_t)(ci * 16), z);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[s]);
      }
      if (p.stats != nullptr) {
        // per-warp column totals -> shared -> one partial per (item, channel)
        if ((lane & 1) == 0) {
          const int col = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k < nchunk) {
              part[(q4 * 2 + 0) * 64 + k * 16 + col] = ts1[k];
              part[(q4 * 2 + 1) * 64 + k * 16 + col] = ts2[k];
            }
        }
        named_bar_sync(1, 128);
        if (et < p.Ct) {
          float a = 0.f, b = 0.f;
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            a += part[(w * 2 + 0) * 64 + et];
            b += part[(w * 2 + 1) * 64 + et];
          }
          const int tiles_per_sample = p.n_seg * p.tiles_h * p.tiles_w;
          const int tile = (c.seg * p.tiles_h + c.th) * p.tiles_w + c.tw;
          if (cbase + et < p.Cout) {
            float* dst = p.stats + (((long long)c.n * tiles_per_sample + tile) * p.Cout + cbase + et) * 2;
            dst[0] = a;
            dst[1] = b;
          }
        }
        named_bar_sync(1, 128);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// Host planning
// ------------------------------------------------------------------------------------------------
static const size_t kMarchWeightBudget = 112 * 1024;

int march_ct(int cin, int cout) {
  if (cin % 16 != 0 || cout % 16 != 0) return 0;
  if (cin != 16 && cin != 32 && cin != 64 && cin != 128) return 0;  // instantiated (BK, chunks) variants
  for (int ct : {64, 48, 32, 16}) {
    if (cout % ct != 0) continue;
    if ((size_t)27 * cin * ct * 2 <= kMarchWeightBudget) return ct;
  }
  return 0;
}

struct MarchPlan {
  MarchParams p;
  size_t smem;
  int grid;
};

static size_t march_tail_bytes() { return (2 * kMaxRing + 2 * kMaxSlots + 2) * 8 + 16 + 4 * 2 * 64 * 4; }

static int plan_march(const rehr_tensor& x, const rehr_tensor& y, MarchPlan* out, int ds_override) {
  MarchParams& p = out->p;
  memset(&p, 0, sizeof(p));
  const int ct = march_ct(x.c, y.c);
  if (ct == 0) return REHR_UNSUPPORTED;
  if (x.n != y.n || x.d != y.d || x.h != y.h || x.w != y.w) return REHR_BAD_SHAPE;
  p.N = x.n; p.D = x.d; p.H = x.h; p.W = x.w; p.Cin = x.c; p.Cout = y.c;
  p.Ct = ct;
  p.n_ct = y.c / ct;
  p.BK = std::min(x.c, 64);
  p.chunks = x.c / p.BK;
  p.tiles_h = (p.H + kTileH - 1) / kTileH;
  p.tiles_w = (p.W + kTileW - 1) / kTileW;
  p.slots = std::min(kMaxSlots, 512 / ct);
  const uint32_t rowb = p.BK * 2;
  p.wtile_bytes = 3 * ct * rowb;
  p.w_bytes = 9 * p.chunks * p.wtile_bytes;
  p.chunk_stride = (kHaloRows * rowb + 1023u) & ~1023u;
  p.slot_stride = p.chunks * p.chunk_stride;
  p.plane_bytes = p.chunks * kHaloRows * rowb;
  const size_t fixed = 1024 + ((p.w_bytes + 1023u) & ~1023u) + march_tail_bytes();
  const size_t budget = 227 * 1024;
  if (fixed + 2 * (size_t)p.slot_stride > budget) return REHR_UNSUPPORTED;
  p.ring = (int)std::min<size_t>(kMaxRing, (budget - fixed) / p.slot_stride);
  out->smem = fixed + (size_t)p.ring * p.slot_stride;
  // depth segment: as long as possible while keeping >= ~6 work items per SM
  const int sms = sm_count();
  const long long cols = (long long)p.n_ct * p.N * p.tiles_h * p.tiles_w;
  int ds = p.D;
  while (ds > 8 && cols * ((p.D + ds - 1) / ds) < 6LL * sms) ds = (ds + 1) / 2;
  if (ds_override > 0) ds = ds_override;
  p.Ds = ds;
  p.n_seg = (p.D + ds - 1) / ds;
  p.items_per_ct = p.N * p.n_seg * p.tiles_h * p.tiles_w;
  p.total_items = p.items_per_ct * p.n_ct;
  out->grid = std::min(p.total_items, sms);
  return REHR_OK;
}

int march_stats_tiles(const rehr_tensor& x, const rehr_tensor& y) {
  MarchPlan pl;
  if (plan_march(x, y, &pl, 0) != REHR_OK) return 0;
  return pl.p.n_seg * pl.p.tiles_h * pl.p.tiles_w;
}

int launch_march(const rehr_tensor& x, const void* w_march, const float* bias, const rehr_tensor& y, int y_is_f32, int act,
                 float slope, float* stats, cudaStream_t stream) {
  MarchPlan pl;
  int rc = plan_march(x, y, &pl, 0);
  if (rc != REHR_OK) return rc;
  MarchParams& p = pl.p;
  if (x.ld % 8 != 0) return REHR_BAD_ALIGNMENT;
  p.out = y.ptr;
  p.out_f32 = y_is_f32;
  p.out_ld = y.ld;
  p.bias = bias;
  p.act = act;
  p.slope = slope;
  p.stats = stats;
  p.err = nullptr;
  {
    const unsigned long long gdim[5] = {(unsigned long long)x.c, (unsigned long long)x.w, (unsigned long long)x.h,
                                        (unsigned long long)x.d, (unsigned long long)x.n};
    const unsigned long long pitch = (unsigned long long)x.ld * 2;
    const unsigned long long gstr[4] = {pitch, pitch * x.w, pitch * x.w * x.h, pitch * x.w * x.h * x.d};
    const unsigned box[5] = {(unsigned)p.BK, (unsigned)kHaloW, (unsigned)kHaloH, 1u, 1u};
    rc = encode_tiled_bf16(&p.x_map, x.ptr, 5, gdim, gstr, box, p.BK * 2);
    if (rc != REHR_OK) return rc;
  }
  {
    const unsigned long long rows = (unsigned long long)p.n_ct * 9 * p.chunks * 3 * p.Ct;
    const unsigned long long gdim[2] = {(unsigned long long)p.BK, rows};
    const unsigned long long gstr[1] = {(unsigned long long)p.BK * 2};
    const unsigned box[2] = {(unsigned)p.BK, (unsigned)(3 * p.Ct)};
    rc = encode_tiled_bf16(&p.w_map, w_march, 2, gdim, gstr, box, p.BK * 2);
    if (rc != REHR_OK) return rc;
  }
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    cudaError_t e;
    e = cudaFuncSetAttribute(conv_march_kernel<16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) attr_err = e;
    e = cudaFuncSetAttribute(conv_march_kernel<32, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) attr_err = e;
    e = cudaFuncSetAttribute(conv_march_kernel<64, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) attr_err = e;
    e = cudaFuncSetAttribute(conv_march_kernel<64, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) attr_err = e;
  });
  if (attr_err != cudaSuccess) {
    g_last_cuda_error = (int)attr_err;
    return REHR_CUDA_ERROR;
  }
  if (p.BK == 16 && p.chunks == 1)
    conv_march_kernel<16, 1><<<pl.grid, kMarchThreads, pl.smem, stream>>>(p);
  else if (p.BK == 32 && p.chunks == 1)
    conv_march_kernel<32, 1><<<pl.grid, kMarchThreads, pl.smem, stream>>>(p);
  else if (p.BK == 64 && p.chunks == 1)
    conv_march_kernel<64, 1><<<pl.grid, kMarchThreads, pl.smem, stream>>>(p);
  else if (p.BK == 64 && p.chunks == 2)
    conv_march_kernel<64, 2><<<pl.grid, kMarchThreads, pl.smem, stream>>>(p);
  else
    return REHR_UNSUPPORTED;
  REHR_CHECK_LAUNCH();
  return REHR_OK;
}

// ------------------------------------------------------------------------------------------------
// Weight packing for the marching kernel:
//   dst[ct][khw][chunk][j*Ct + col][BK]   with  j = 2 - kd,  co = ct*Ct + col,  ci = chunk*BK + k
//   src element = w[co*s_co + ci*s_ci + (flip ? 26 - t : t)],  t = (kd*3 + kh)*3 + kw
// forward of W[Cout][Cin][27]: s_co = Cin*27, s_ci = 27, flip = 0;
// input-gradient (dx[B] from dy[A]) of W[A][B][27]: cout := B, cin := A, s_co = 27, s_ci = B*27, flip = 1.
// ------------------------------------------------------------------------------------------------
__global__ void pack_march_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int cout, int cin, int Ct, int BK,
                                  long long s_co, long long s_ci, int flip) {
  const int chunks = cin / BK;
  const long long total = (long long)cout * cin * 27;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % BK);
    long long r = i / BK;
    const int rowi = (int)(r % (3 * Ct));
    r /= 3 * Ct;
    const int chunk = (int)(r % chunks);
    r /= chunks;
    const int khw = (int)(r % 9);
    const int ct = (int)(r / 9);
    const int j = rowi / Ct, col = rowi % Ct;
    const int kd = 2 - j, kh = khw / 3, kw = khw % 3;
    const int t = (kd * 3 + kh) * 3 + kw;
    const int co = ct * Ct + col, ci = chunk * BK + k;
    dst[i] = __float2bfloat16(src[co * s_co + ci * s_ci + (flip ? 26 - t : t)]);
  }
}

}  // namespace rehr

using namespace rehr;

extern "C" {

int rehr_conv3d_march_supported(const rehr_conv_desc* d, int cin, int cout) {
  if (!d) return 0;
  if (d->kd != 3 || d->kh != 3 || d->kw != 3 || d->sd != 1 || d->sh != 1 || d->sw != 1 || d->pd != 1 || d->ph != 1 || d->pw != 1)
    return 0;
  return march_ct(cin, cout) > 0 ? 1 : 0;
}

size_t rehr_conv3d_march_weight_bytes(int cin, int cout) {
  return march_ct(cin, cout) > 0 ? (size_t)27 * cin * cout * 2 : 0;
}

int rehr_pack_weight_march(const float* src, void* dst_bf16, int cout, int cin, long long s_co, long long s_ci, int flip,
                           rehr_stream stream) {
  if (!src || !dst_bf16) return REHR_BAD_SHAPE;
  const int ct = march_ct(cin, cout);
  if (ct == 0) return REHR_UNSUPPORTED;
  const long long total = (long long)cout * cin * 27;
  const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
  pack_march_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, reinterpret_cast<__nv_bfloat16*>(dst_bf16), cout, cin, ct,
                                                             std::min(cin, 64), s_co, s_ci, flip);
  REHR_CHECK_LAUNCH();
  return REHR_OK;
}

int rehr_conv3d_march_stats_tiles(const rehr_tensor* x, const rehr_tensor* y) {
  if (!x || !y) return 0;
  return march_stats_tiles(*x, *y);
}

int rehr_conv3d_march_fwd(const rehr_tensor* x, const void* w_march, const float* bias, const rehr_tensor* y, int y_is_f32,
                          int act, float slope, float* stats, rehr_stream stream) {
  if (!x || !y || !x->ptr || !y->ptr || !w_march) return REHR_BAD_SHAPE;
  return launch_march(*x, w_march, bias, *y, y_is_f32, act, slope, stats, (cudaStream_t)stream);
}

}  // extern "C"

// Batched weight re-pack: every 16-bit GEMM-operand copy of a training step in ONE launch.
//
// A PlainConvUNet step refreshes ~57 packed copies (forward / input-gradient layouts of 27 convs, the transposed convs) from the
// fp32 parameters the optimiser has just updated.  As 57 launches of 2-60 us they take ~0.5 ms back to back, sit next to the first
// convs of the step and compete with them for SM slots.  Here the calls are RECORDED (rehr_pack_batch_begin) and issued as one
// kernel whose job table travels in the kernel parameters (<= 32 KB on CUDA 12.1+ / sm_70+): nothing to upload, so the launch is
// capturable in a CUDA graph, and the pointers it bakes in are the stable addresses of the parameters and of the cached copies.
//
// Index maps (identical to the stand-alone kernels they batch; tests/test_pack_batch_gpu.py checks bit equality):
//   kind 0  pack_weight_kernel / pack_weight_runs_kernel (conv_engine.cu):  dst[r][t][c] = src[r*sr + c*sc + t*st]
//   kind 1  pack_march_kernel (conv_march.cu):        dst[ct][khw][chunk][j*Ct + col][BK]
//   kind 2  pack_march_s2dgrad_kernel (conv_march.cu): dst[class][ct][khw][chunk][j*Ct + col][BK]
#include "engine.h"
#include "ptx.cuh"

#include <algorithm>
#include <vector>

namespace rehr {

static constexpr int kPackElemsPerBlock = 4096;  // element-wise kinds: 256 threads x 16
static constexpr int kPackRunC = 64;             // kind 0 with st == 1: block = (r, 64 consecutive c)
static constexpr int kPackRunMaxT = 27;          // largest tap count staged through shared memory (k3); larger kernels go element-wise

struct PackBatch {
  int n;
  int pad_;
  PackJob jobs[kPackBatchMax];
};

__device__ __forceinline__ unsigned short pk16(float v, int f16) { return pack16(v, f16); }  // same rounding as the single packs

__device__ __forceinline__ void pack_generic_elems(const PackJob& j, long long i0, long long i1) {
  unsigned short* dst = reinterpret_cast<unsigned short*>(j.dst);
  for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    const int c = (int)(i % j.C);
    const long long rt = i / j.C;
    const int t = (int)(rt % j.T);
    const int r = (int)(rt / j.T);
    dst[i] = pk16(j.src[r * j.sr + c * j.sc + t * j.st], j.f16);
  }
}

// TC > 0: compile-time tap count (27 = every k3 conv: the divisions below become multiplications); TC = 0: run-time j.T
template <int TC>
__device__ __forceinline__ void pack_generic_runs(const PackJob& j, int lb, float* tile) {
  const int cblocks = (j.C + kPackRunC - 1) / kPackRunC;
  const int r = lb / cblocks, c0 = (lb - r * cblocks) * kPackRunC;
  const int nc = min(kPackRunC, j.C - c0);
  const int T = TC > 0 ? TC : j.T;
  unsigned short* dst = reinterpret_cast<unsigned short*>(j.dst) + (long long)r * T * j.C + c0;
  const float* src = j.src + r * j.sr + (long long)c0 * j.sc;
  if (j.sc == T) {
    // the nc runs are one contiguous block of nc * T floats (forward layout): plain coalesced copy into the padded tile
    for (int i = threadIdx.x; i < nc * T; i += blockDim.x) {
      const int cl = i / T, t = i - cl * T;
      tile[cl * (T + 1) + t] = src[i];
    }
  } else {
    for (int i = threadIdx.x; i < nc * T; i += blockDim.x) {
      const int cl = i / T, t = i - cl * T;
      tile[cl * (T + 1) + t] = src[(long long)cl * j.sc + t];
    }
  }
  __syncthreads();
  if (nc == kPackRunC) {
    for (int i = threadIdx.x; i < kPackRunC * T; i += blockDim.x) {
      const int t = i / kPackRunC, cl = i % kPackRunC;
      dst[(long long)t * j.C + cl] = pk16(tile[cl * (T + 1) + t], j.f16);
    }
  } else {
    for (int i = threadIdx.x; i < nc * T; i += blockDim.x) {
      const int t = i / nc, cl = i - t * nc;
      dst[(long long)t * j.C + cl] = pk16(tile[cl * (T + 1) + t], j.f16);
    }
  }
}

__device__ __forceinline__ void pack_march_elems(const PackJob& j, long long i0, long long i1) {
  unsigned short* dst = reinterpret_cast<unsigned short*>(j.dst);
  const int chunks = j.cin / j.BK;
  const int T = j.kdn * j.ks * j.ks;
  for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    const int k = (int)(i % j.BK);
    long long r = i / j.BK;
    const int rowi = (int)(r % (j.kdn * j.Ct));
    r /= j.kdn * j.Ct;
    const int chunk = (int)(r % chunks);
    r /= chunks;
    const int khw = (int)(r % (j.ks * j.ks));
    const int ct = (int)(r / (j.ks * j.ks));
    const int jj = rowi / j.Ct, col = rowi % j.Ct;
    const int kd = j.kdn - 1 - jj, kh = khw / j.ks, kw = khw % j.ks;
    const int t = (kd * j.ks + kh) * j.ks + kw;
    const int co = ct * j.Ct + col, ci = chunk * j.BK + k;
    dst[i] = co < j.cout ? pk16(j.src[co * j.s_co + ci * j.s_ci + (j.flip ? T - 1 - t : t)], j.f16) : (unsigned short)0;
  }
}

__device__ __forceinline__ void pack_s2dgrad_elems(const PackJob& j, long long i0, long long i1) {
  unsigned short* dst = reinterpret_cast<unsigned short*>(j.dst);
  const int A = j.cin, B = j.cout, Bpad = j.cout_pad;  // A = channels of dy (conv Cout), B = channels of dx (conv Cin)
  const int chunks = A / j.BK;
  const long long per_cls = (long long)Bpad * A * 27;
  for (long long i00 = i0 + threadIdx.x; i00 < i1; i00 += blockDim.x) {
    const int cls = (int)(i00 / per_cls);
    long long i = i00 % per_cls;
    const int rw = cls % j.sw, rh = (cls / j.sw) % j.sh, rd = cls / (j.sw * j.sh);
    const int k = (int)(i % j.BK);
    long long r = i / j.BK;
    const int rowi = (int)(r % (3 * j.Ct));
    r /= 3 * j.Ct;
    const int chunk = (int)(r % chunks);
    r /= chunks;
    const int khw = (int)(r % 9);
    const int ct = (int)(r / 9);
    const int jj = rowi / j.Ct, col = rowi % j.Ct;
    const int u[3] = {2 - jj, khw / 3, khw % 3};
    const int st[3] = {j.sd, j.sh, j.sw}, rr[3] = {rd, rh, rw};
    int kk[3];
    bool ok = true;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      if (st[a] == 1) kk[a] = 2 - u[a];
      else if (rr[a] == 0) { kk[a] = 1; ok = ok && u[a] == 1; }
      else { kk[a] = u[a] == 1 ? 2 : 0; ok = ok && u[a] >= 1; }
    }
    const int b = ct * j.Ct + col, a_ch = chunk * j.BK + k;
    float v = 0.f;
    if (ok && b < B) v = j.src[((long long)a_ch * B + b) * 27 + (kk[0] * 3 + kk[1]) * 3 + kk[2]];
    dst[i00] = pack16(v, 0);
  }
}

__global__ void __launch_bounds__(256) pack_batched_kernel(const __grid_constant__ PackBatch b) {
  __shared__ float tile[kPackRunC * (kPackRunMaxT + 1)];
  // the job this block belongs to: block0 is ascending
  int lo = 0, hi = b.n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (b.jobs[mid].block0 <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const PackJob& j = b.jobs[lo];
  const int lb = (int)blockIdx.x - j.block0;
  if (j.kind == 0 && j.runs) {
    if (j.T == 27) pack_generic_runs<27>(j, lb, tile);
    else pack_generic_runs<0>(j, lb, tile);
    return;
  }
  const long long i0 = (long long)lb * kPackElemsPerBlock;
  const long long i1 = min(j.total, i0 + kPackElemsPerBlock);
  if (j.kind == 0) pack_generic_elems(j, i0, i1);
  else if (j.kind == 1) pack_march_elems(j, i0, i1);
  else pack_s2dgrad_elems(j, i0, i1);
}

// ---- host: thread-local recorder ------------------------------------------------------------------------------------------
static thread_local std::vector<PackJob>* g_rec = nullptr;

bool pack_recording() { return g_rec != nullptr; }

int pack_record(PackJob j) {
  if (!g_rec) return REHR_UNSUPPORTED;
  if (j.kind == 0) {
    j.total = (long long)j.R * j.T * j.C;
    j.runs = (j.st == 1 && j.T <= kPackRunMaxT) ? 1 : 0;
    j.nblocks = j.runs ? j.R * ((j.C + kPackRunC - 1) / kPackRunC) : (int)((j.total + kPackElemsPerBlock - 1) / kPackElemsPerBlock);
  } else {
    j.runs = 0;
    j.nblocks = (int)((j.total + kPackElemsPerBlock - 1) / kPackElemsPerBlock);
  }
  if (j.total <= 0) return REHR_OK;
  g_rec->push_back(j);
  return REHR_OK;
}

}  // namespace rehr

using namespace rehr;

extern "C" {

int rehr_pack_batch_begin(void) {
  if (g_rec) g_rec->clear();
  else g_rec = new std::vector<PackJob>();
  return REHR_OK;
}

int rehr_pack_batch_launch(rehr_stream stream) {
  if (!g_rec) return REHR_BAD_SHAPE;
  std::vector<PackJob>* rec = g_rec;
  g_rec = nullptr;  // recording ends here, whatever happens below
  int rc = REHR_OK;
  for (size_t first = 0; first < rec->size() && rc == REHR_OK; first += kPackBatchMax) {
    PackBatch b;
    b.n = (int)std::min<size_t>(kPackBatchMax, rec->size() - first);
    b.pad_ = 0;
    int blocks = 0;
    for (int i = 0; i < b.n; ++i) {
      b.jobs[i] = (*rec)[first + i];
      b.jobs[i].block0 = blocks;
      blocks += b.jobs[i].nblocks;
    }
    for (int i = b.n; i < kPackBatchMax; ++i) b.jobs[i] = PackJob{};
    if (blocks == 0) continue;
    pack_batched_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(b);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
      g_last_cuda_error = (int)e;
      rc = REHR_CUDA_ERROR;
    }
  }
  delete rec;
  return rc;
}

int rehr_pack_batch_abort(void) {
  delete g_rec;
  g_rec = nullptr;
  return REHR_OK;
}

}  // extern "C"

// HBM-bound helper kernels around the fused tcgen05 path:
//   - skinny weight-gradient reduction (dA, dB)        lora.py:53-58 autograd backward
//   - 16-bit shadow / frozen-weight casts + transposes
//   - global-norm clip + AdamW over the flat LoRA arena  train_lora_dreambooth.py:878-888
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "lora_b200.h"

namespace lb {

__device__ __forceinline__ float2 ld16x2(const uint32_t w, int fmt) {
  if (fmt) {
    __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(&w);
    return __bfloat1622float2(v);
  }
  __half2 v = *reinterpret_cast<const __half2*>(&w);
  return __half22float2(v);
}
__device__ __forceinline__ uint16_t to16(float x, int fmt) {
  if (fmt) {
    __nv_bfloat16 v = __float2bfloat16_rn(x);
    return *reinterpret_cast<uint16_t*>(&v);
  }
  __half v = __float2half_rn(x);
  return *reinterpret_cast<uint16_t*>(&v);
}
__device__ __forceinline__ float from16(uint16_t x, int fmt) {
  if (fmt) return __bfloat162float(*reinterpret_cast<__nv_bfloat16*>(&x));
  return __half2float(*reinterpret_cast<__half*>(&x));
}

// ------------------------------------------------------------------------------- wgrad
// CTA = 8 warps over a [ROWS x 64-column] slab of S. Lane owns 2 adjacent columns (one 32-bit
// load; a warp reads one full 128-B line per row). Warps stride over rows. Each thread keeps
// 16 x 2 fp32 partial sums; warps are combined through shared memory; one atomicAdd per output.
constexpr int WG_COLS = 64;
constexpr int WG_WARPS = 8;
constexpr int WG_ROWS = 256;

template <int RQ>  // number of float4 groups of V actually used: ceil(r/4)
__global__ void __launch_bounds__(WG_WARPS * 32)
wgrad_kernel(const uint32_t* __restrict__ S, const float* __restrict__ V,
             const float* __restrict__ diag, float scale, float* __restrict__ out,
             long long out_js, long long out_cs, int M, int C, int r, int fmt) {
  __shared__ float red[WG_WARPS][RQ * 4][WG_COLS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = blockIdx.x * WG_COLS + lane * 2;
  const int m_begin = blockIdx.y * WG_ROWS;
  const int m_end = min(M, m_begin + WG_ROWS);
  const bool col_ok = c0 < C;  // C is even (C % 8 == 0 enforced by the host)
  const size_t pitch = static_cast<size_t>(C) >> 1;  // row pitch in 32-bit words

  float acc[RQ * 4][2];
#pragma unroll
  for (int j = 0; j < RQ * 4; ++j) acc[j][0] = acc[j][1] = 0.f;

  for (int m = m_begin + warp; m < m_end; m += WG_WARPS) {
    const uint32_t w = col_ok ? __ldg(S + static_cast<size_t>(m) * pitch + (c0 >> 1)) : 0u;
    const float2 x = ld16x2(w, fmt);
    const float4* vrow = reinterpret_cast<const float4*>(V + static_cast<size_t>(m) * 16);
#pragma unroll
    for (int qd = 0; qd < RQ; ++qd) {
      const float4 v = __ldg(vrow + qd);
      acc[qd * 4 + 0][0] += v.x * x.x; acc[qd * 4 + 0][1] += v.x * x.y;
      acc[qd * 4 + 1][0] += v.y * x.x; acc[qd * 4 + 1][1] += v.y * x.y;
      acc[qd * 4 + 2][0] += v.z * x.x; acc[qd * 4 + 2][1] += v.z * x.y;
      acc[qd * 4 + 3][0] += v.w * x.x; acc[qd * 4 + 3][1] += v.w * x.y;
    }
  }
#pragma unroll
  for (int j = 0; j < RQ * 4; ++j) {
    red[warp][j][lane * 2 + 0] = acc[j][0];
    red[warp][j][lane * 2 + 1] = acc[j][1];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < RQ * 4 * WG_COLS; idx += WG_WARPS * 32) {
    const int j = idx / WG_COLS, c = idx % WG_COLS;
    const int cg = blockIdx.x * WG_COLS + c;
    if (j < r && cg < C) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < WG_WARPS; ++w) s += red[w][j][c];
      const float coef = scale * (diag ? diag[j] : 1.f);
      atomicAdd(out + j * out_js + cg * out_cs, coef * s);
    }
  }
}

// ------------------------------------------------------------------------------- casts
__global__ void cast_rows_pad16_kernel(const float* __restrict__ src, long long rs, long long cs,
                                       uint16_t* __restrict__ dst, int r, int C, int fmt) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (c >= C) return;
  const float x = (j < r) ? src[j * rs + c * cs] : 0.f;
  dst[static_cast<size_t>(j) * C + c] = to16(x, fmt);
}

__global__ void refresh_shadows_kernel(const float* __restrict__ p,
                                       const long long* __restrict__ table,
                                       uint16_t* __restrict__ dst_base, int fmt) {
  const long long* e = table + static_cast<size_t>(blockIdx.y) * 7;
  const long long src_off = e[0], rs = e[1], cs = e[2];
  const int r = static_cast<int>(e[3]), C = static_cast<int>(e[4]);
  const long long dst_off = e[5], dst_rs = e[6];
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float x = (j < r) ? p[src_off + j * rs + c * cs] : 0.f;
    dst_base[dst_off + j * dst_rs + c] = to16(x, fmt);
  }
}

// src [R,C] -> dst [R,C] (optional) and dstT [C,R] (optional), 32x32 tiles through smem
template <typename SrcT>
__global__ void cast_weight_kernel(const SrcT* __restrict__ src, int src_fmt,
                                   uint16_t* __restrict__ dst, uint16_t* __restrict__ dstT, int R,
                                   int C, int fmt) {
  __shared__ float tile[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int rr = blockIdx.y * 32 + i;
    float x = 0.f;
    if (rr < R && c < C) {
      if constexpr (sizeof(SrcT) == 4) x = src[static_cast<size_t>(rr) * C + c];
      else x = from16(src[static_cast<size_t>(rr) * C + c], src_fmt);
      if (dst) dst[static_cast<size_t>(rr) * C + c] = to16(x, fmt);
    }
    tile[i][threadIdx.x] = x;
  }
  __syncthreads();
  if (dstT) {
    const int rr = blockIdx.y * 32 + threadIdx.x;  // becomes the contiguous index of dstT
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      const int cc = blockIdx.x * 32 + i;
      if (rr < R && cc < C) dstT[static_cast<size_t>(cc) * R + rr] = to16(tile[threadIdx.x][i], fmt);
    }
  }
}

// ------------------------------------------------------------------------------- clip + AdamW
constexpr int OPT_THREADS = 256;
constexpr int OPT_MAX_PARTIALS = 1024;
constexpr int OPT_MAX_GROUPS = 8;

struct OptGroups {
  long long off[OPT_MAX_GROUPS + 1];
  int n;
};

__device__ __forceinline__ float block_sum(float x, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) sh[warp] = x;
  __syncthreads();
  float t = (threadIdx.x < OPT_THREADS / 32) ? sh[threadIdx.x] : 0.f;
  if (warp == 0) {
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) sh[0] = t;
  }
  __syncthreads();
  const float rsum = sh[0];
  __syncthreads();
  return rsum;
}

// pass 1: per-block partial sums of g^2 (fixed summation order => run-to-run deterministic)
__global__ void __launch_bounds__(OPT_THREADS)
sqnorm_partial_kernel(const float* __restrict__ g, long long n, float* __restrict__ partials,
                      int* __restrict__ step_dev) {
  __shared__ float sh[OPT_THREADS / 32];
  float s = 0.f;
  const long long n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = blockIdx.x * static_cast<long long>(OPT_THREADS) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * OPT_THREADS) {
    const float4 x = g4[i];
    s += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const float x = g[(n4 << 2) + threadIdx.x];
    s += x * x;
  }
  const float tot = block_sum(s, sh);
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = tot;
    if (blockIdx.x == 0) step_dev[0] += 1;  // t for the update pass that follows on the stream
  }
}

// pass 2: every block re-reduces the partials (same order everywhere), then updates its slice
__global__ void __launch_bounds__(OPT_THREADS)
adamw_update_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                    float* __restrict__ v, long long n, OptGroups groups,
                    const float* __restrict__ lr_dev, float beta1, float beta2, float eps,
                    float wd, float max_norm, float inv_world, const int* __restrict__ step_dev,
                    const float* __restrict__ partials, int n_partials,
                    float* __restrict__ gnorm_out) {
  __shared__ float sh[OPT_THREADS / 32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n_partials; i += OPT_THREADS) s += partials[i];
  const float sq = block_sum(s, sh);
  const float total = sqrtf(sq) * inv_world;
  float coef = inv_world;
  if (max_norm > 0.f) coef *= fminf(1.f, max_norm / (total + 1e-6f));
  if (blockIdx.x == 0 && threadIdx.x == 0 && gnorm_out) gnorm_out[0] = total;

  const int t = step_dev[0];
  const float bc1 = static_cast<float>(1.0 - pow(static_cast<double>(beta1), static_cast<double>(t)));
  const float bc2_sqrt = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(beta2), static_cast<double>(t))));

  for (long long i = blockIdx.x * static_cast<long long>(OPT_THREADS) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * OPT_THREADS) {
    int gi = 0;
#pragma unroll
    for (int k = 1; k < OPT_MAX_GROUPS; ++k)
      if (k < groups.n && i >= groups.off[k]) gi = k;
    const float lr = lr_dev[gi];
    const float gg = g[i] * coef;
    float pp = p[i];
    pp *= (1.f - lr * wd);
    const float mm = beta1 * m[i] + (1.f - beta1) * gg;
    const float vv = beta2 * v[i] + (1.f - beta2) * gg * gg;
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    pp -= (lr / bc1) * (mm / denom);
    p[i] = pp;
    m[i] = mm;
    v[i] = vv;
    g[i] = 0.f;
  }
}

}  // namespace lb

// =============================================================================== C ABI
using namespace lb;

extern "C" int lb_abi_version(void) { return 1; }

extern "C" int lb_lora_wgrad(const void* S, const float* V, const float* diag, float scale,
                             float* out, long long out_js, long long out_cs, int M, int C, int r,
                             int in_dtype, void* stream) {
  if (M <= 0 || C <= 0 || (C % 8) != 0) return LB_ERR_SHAPE;
  if (r < 1 || r > 16) return LB_ERR_RANK;
  if (in_dtype != LB_BF16 && in_dtype != LB_F16) return LB_ERR_DTYPE;
  if ((reinterpret_cast<uintptr_t>(S) | reinterpret_cast<uintptr_t>(V)) & 15) return LB_ERR_ALIGN;
  const int fmt = in_dtype == LB_BF16 ? 1 : 0;
  dim3 grid((C + WG_COLS - 1) / WG_COLS, (M + WG_ROWS - 1) / WG_ROWS);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const uint32_t* S32 = reinterpret_cast<const uint32_t*>(S);
  switch ((r + 3) / 4) {
    case 1: wgrad_kernel<1><<<grid, WG_WARPS * 32, 0, st>>>(S32, V, diag, scale, out, out_js, out_cs, M, C, r, fmt); break;
    case 2: wgrad_kernel<2><<<grid, WG_WARPS * 32, 0, st>>>(S32, V, diag, scale, out, out_js, out_cs, M, C, r, fmt); break;
    case 3: wgrad_kernel<3><<<grid, WG_WARPS * 32, 0, st>>>(S32, V, diag, scale, out, out_js, out_cs, M, C, r, fmt); break;
    default: wgrad_kernel<4><<<grid, WG_WARPS * 32, 0, st>>>(S32, V, diag, scale, out, out_js, out_cs, M, C, r, fmt); break;
  }
  return cudaGetLastError() == cudaSuccess ? LB_OK : LB_ERR_CUDA;
}

extern "C" int lb_cast_rows_pad16(const float* src, long long src_rs, long long src_cs,
                                  void* dst16, int r, int C, int out_dtype, void* stream) {
  if (C <= 0) return LB_ERR_SHAPE;
  if (r < 1 || r > 16) return LB_ERR_RANK;
  if (out_dtype != LB_BF16 && out_dtype != LB_F16) return LB_ERR_DTYPE;
  dim3 grid((C + 255) / 256, 16);
  cast_rows_pad16_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      src, src_rs, src_cs, reinterpret_cast<uint16_t*>(dst16), r, C, out_dtype == LB_BF16);
  return cudaGetLastError() == cudaSuccess ? LB_OK : LB_ERR_CUDA;
}

extern "C" int lb_refresh_shadows(const float* p, const long long* table, int n_entries,
                                  int max_C, void* dst16_base, int out_dtype, void* stream) {
  if (n_entries <= 0 || max_C <= 0) return LB_ERR_SHAPE;
  if (out_dtype != LB_BF16 && out_dtype != LB_F16) return LB_ERR_DTYPE;
  dim3 grid((max_C + 255) / 256, n_entries);
  refresh_shadows_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      p, table, reinterpret_cast<uint16_t*>(dst16_base), out_dtype == LB_BF16);
  return cudaGetLastError() == cudaSuccess ? LB_OK : LB_ERR_CUDA;
}

extern "C" int lb_cast_weight(const void* src, int src_dtype, void* dst16, void* dstT16, int R,
                              int C, int out_dtype, void* stream) {
  if (R <= 0 || C <= 0) return LB_ERR_SHAPE;
  if (out_dtype != LB_BF16 && out_dtype != LB_F16) return LB_ERR_DTYPE;
  dim3 grid((C + 31) / 32, (R + 31) / 32), block(32, 8);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint16_t* d = reinterpret_cast<uint16_t*>(dst16);
  uint16_t* dT = reinterpret_cast<uint16_t*>(dstT16);
  const int fmt = out_dtype == LB_BF16;
  if (src_dtype == LB_F32)
    cast_weight_kernel<float><<<grid, block, 0, st>>>(reinterpret_cast<const float*>(src), 0, d, dT, R, C, fmt);
  else if (src_dtype == LB_BF16 || src_dtype == LB_F16)
    cast_weight_kernel<uint16_t><<<grid, block, 0, st>>>(reinterpret_cast<const uint16_t*>(src), src_dtype == LB_BF16, d, dT, R, C, fmt);
  else
    return LB_ERR_DTYPE;
  return cudaGetLastError() == cudaSuccess ? LB_OK : LB_ERR_CUDA;
}

extern "C" int lb_adamw_clip_step(float* p, float* g, float* m, float* v, long long n,
                                  const long long* group_off, int n_groups, const float* lr_dev,
                                  float beta1, float beta2, float eps, float weight_decay,
                                  float max_norm, float inv_world, int* step_dev, float* partials,
                                  float* gnorm_out, void* stream) {
  if (n <= 0 || n_groups < 1 || n_groups > OPT_MAX_GROUPS) return LB_ERR_SHAPE;
  if (reinterpret_cast<uintptr_t>(g) & 15) return LB_ERR_ALIGN;
  OptGroups groups;
  groups.n = n_groups;
  for (int i = 0; i <= n_groups; ++i) groups.off[i] = group_off[i];
  for (int i = n_groups + 1; i <= OPT_MAX_GROUPS; ++i) groups.off[i] = n;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  long long want = (n / 4 + OPT_THREADS - 1) / OPT_THREADS;
  int nblk = static_cast<int>(want < 1 ? 1 : (want > OPT_MAX_PARTIALS ? OPT_MAX_PARTIALS : want));
  sqnorm_partial_kernel<<<nblk, OPT_THREADS, 0, st>>>(g, n, partials, step_dev);
  long long want2 = (n + OPT_THREADS - 1) / OPT_THREADS;
  int nblk2 = static_cast<int>(want2 > 148 * 8 ? 148 * 8 : want2);
  adamw_update_kernel<<<nblk2, OPT_THREADS, 0, st>>>(p, g, m, v, n, groups, lr_dev, beta1, beta2,
                                                     eps, weight_decay, max_norm, inv_world,
                                                     step_dev, partials, nblk, gnorm_out);
  return cudaGetLastError() == cudaSuccess ? LB_OK : LB_ERR_CUDA;
}

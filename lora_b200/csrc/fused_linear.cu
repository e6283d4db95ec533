// Fused frozen-weight GEMM + rank-r LoRA for sm_100a (tcgen05 + TMEM + TMA).
//
//   Y[M,N] = X[M,K] . W[N,K]^T (+ bias) + ((X . D^T) * (scale * diag)) . U^T
//
// D ("down", [r,K]) rides the main K loop as 16 extra rows appended to every W tile in shared
// memory, so ONE tcgen05.mma per K-step (N = BLOCK_N + 16) produces the base accumulator in TMEM
// columns [0,BLOCK_N) and T = X.D^T in columns [BLOCK_N,BLOCK_N+16): X is read from HBM once.
// After the K loop the epilogue warps pull T out of TMEM, scale it, write it back to shared memory
// as a K=16 bf16 operand and the MMA warp issues one more tcgen05.mma (K = 16) against the U tile,
// accumulating the LoRA branch straight into the base accumulator. The sum is written once by TMA.
//
// The same kernel is the backward-dX kernel (dX = gY.W + (s.(gY.B)).A) when called with the
// pre-transposed frozen weight W^T[K,N], down = B^T and up = A^T; see lora_b200.h.
//
// Replaces: LoraInjectedLinear.forward, /root/reference/lora_diffusion/lora.py:53-58 and its
// autograd backward (dX part).
#include <stdio.h>
#include <mutex>
#include <unordered_map>

#include "lora_b200.h"
#include "ptx.cuh"
#include "tmap.h"

namespace lb {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 x 16-bit = 128 B = one SWIZZLE_128B row
constexpr int R_PAD = 16;    // LoRA rank padded to one UMMA K-step
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;  // warp0 TMA, warp1 MMA+TMEM, warps2-5 epilogue
constexpr int EPI_THREADS = 128;

struct LinParams {
  const float* bias;   // [N] or null
  const float* up;     // fp32 LoRA-up factor, element (n, j) at up[n*up_rs + j*up_cs]
  long long up_rs, up_cs;
  const float* diag;   // [r] or null  (selector diagonal)
  float* t_out;        // [M,16] fp32: T = X.D^T (unscaled), written by the n_blk==0 CTAs; or null
  float scale;
  int M, N, K, r;
  int fmt;             // 1 = bf16 operands, 0 = fp16 operands
};

template <int BLOCK_N, int STAGES, typename OutT>
struct Smem {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = (BLOCK_N + R_PAD) * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BOX_COLS = 128 / sizeof(OutT);          // columns per 128-byte store box
  static constexpr int NUM_BOXES = BLOCK_N / BOX_COLS;
  static constexpr int BOX_BYTES = BLOCK_M * 128;
  static constexpr int OFF_OUT = STAGES * STAGE_BYTES;
  static constexpr int OFF_T = OFF_OUT + NUM_BOXES * BOX_BYTES;  // T' operand, 128 x 16 x 2 B
  static constexpr int OFF_UP = OFF_T + BLOCK_M * R_PAD * 2;     // U tile, BLOCK_N x 16 x 2 B
  static constexpr int OFF_BIAS = OFF_UP + BLOCK_N * R_PAD * 2;
  static constexpr int OFF_BAR = OFF_BIAS + BLOCK_N * 4;
  static constexpr int NUM_BARS = 2 * STAGES + 3;
  static constexpr int OFF_TMEM = OFF_BAR + NUM_BARS * 8;
  static constexpr int TOTAL = OFF_TMEM + 16;
  static constexpr int DYN_BYTES = TOTAL + 1024;  // slack for manual 1024-B alignment
  static constexpr int TMEM_COLS = (BLOCK_N + R_PAD <= 128) ? 128 : (BLOCK_N + R_PAD <= 256 ? 256 : 512);
};

template <int BLOCK_N, int STAGES, typename OutT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
fused_lora_linear_kernel(const __grid_constant__ CUtensorMap tmX,
                         const __grid_constant__ CUtensorMap tmW,
                         const __grid_constant__ CUtensorMap tmD,
                         const __grid_constant__ CUtensorMap tmY, const LinParams p) {
  using S = Smem<BLOCK_N, STAGES, OutT>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));  // generic pointer to aligned base

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_blk = blockIdx.x, m_blk = blockIdx.y;
  const int m0 = m_blk * BLOCK_M, n0 = n_blk * BLOCK_N;
  const int num_kb = (p.K + BLOCK_K - 1) / BLOCK_K;

  auto bar_full = [&](int s) { return sbase + S::OFF_BAR + 8 * s; };
  auto bar_empty = [&](int s) { return sbase + S::OFF_BAR + 8 * (STAGES + s); };
  const uint32_t bar_acc = sbase + S::OFF_BAR + 8 * (2 * STAGES + 0);    // main K loop finished
  const uint32_t bar_tready = sbase + S::OFF_BAR + 8 * (2 * STAGES + 1); // T' operand in smem
  const uint32_t bar_final = sbase + S::OFF_BAR + 8 * (2 * STAGES + 2);  // LoRA MMA finished
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sgen + S::OFF_TMEM);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmD);
    tma_prefetch_desc(&tmY);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
    }
    mbar_init(bar_acc, 1);
    mbar_init(bar_tready, EPI_THREADS);
    mbar_init(bar_final, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(sbase + S::OFF_TMEM, S::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(bar_empty(s), ph ^ 1);
        mbar_expect_tx(bar_full(s), S::STAGE_BYTES);
        const uint32_t sa = sbase + s * S::STAGE_BYTES;
        const uint32_t sb = sa + S::A_BYTES;
        tma_load_2d(&tmX, bar_full(s), sa, kb * BLOCK_K, m0);
        tma_load_2d(&tmW, bar_full(s), sb, kb * BLOCK_K, n0);
        tma_load_2d(&tmD, bar_full(s), sb + BLOCK_N * 128, kb * BLOCK_K, 0);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0) {
      const uint32_t idesc_main = umma_idesc_f16(p.fmt, BLOCK_M, BLOCK_N + R_PAD);
      const uint32_t idesc_lora = umma_idesc_f16(p.fmt, BLOCK_M, BLOCK_N);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(bar_full(s), ph);
        tc_fence_after();
        const uint32_t sa = sbase + s * S::STAGE_BYTES;
        const uint32_t sb = sa + S::A_BYTES;
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
          // K-major SWIZZLE_128B: 8-row groups are 1024 B apart; a K-step advances 32 B in the row
          const uint64_t ad = umma_smem_desc(sa + k * UMMA_K * 2, 16, 1024, 2);
          const uint64_t bd = umma_smem_desc(sb + k * UMMA_K * 2, 16, 1024, 2);
          umma_f16_ss(tmem, ad, bd, idesc_main, (kb | k) != 0);
        }
        umma_commit(bar_empty(s));  // frees the smem slot when these MMAs retire
      }
      umma_commit(bar_acc);
      // LoRA up-projection: acc[:, 0:BLOCK_N] += T'[128,16] . U[BLOCK_N,16]^T
      mbar_wait(bar_tready, 0);
      tc_fence_after();
      {
        // no-swizzle K-major: core matrix = 8 rows x 16 B contiguous (128 B);
        // LBO = 128 B between the two K halves, SBO = 256 B between 8-row groups
        const uint64_t ad = umma_smem_desc(sbase + S::OFF_T, 128, 256, 0);
        const uint64_t bd = umma_smem_desc(sbase + S::OFF_UP, 128, 256, 0);
        umma_f16_ss(tmem, ad, bd, idesc_lora, 1);
      }
      umma_commit(bar_final);
    }
  } else {
    // ------------------------------------------------------------ epilogue warps (128 threads)
    const int q = warp & 3;            // TMEM lane quadrant this warp may read
    const int row = q * 32 + lane;     // accumulator row owned by this thread
    const int et = threadIdx.x - 64;   // 0..127
    float* bias_s = reinterpret_cast<float*>(sgen + S::OFF_BIAS);

    // Stage the U tile (fp32 master -> 16-bit, interleaved core-matrix layout) and the bias.
    for (int i = et; i < BLOCK_N; i += EPI_THREADS) {
      const int n = n0 + i;
      float u[R_PAD];
#pragma unroll
      for (int j = 0; j < R_PAD; ++j)
        u[j] = (n < p.N && j < p.r) ? __ldg(p.up + n * p.up_rs + j * p.up_cs) : 0.f;
      const uint32_t dst = sbase + S::OFF_UP + (i >> 3) * 256 + (i & 7) * 16;
      st_shared_v4(dst, pack2(u[0], u[1], p.fmt), pack2(u[2], u[3], p.fmt),
                   pack2(u[4], u[5], p.fmt), pack2(u[6], u[7], p.fmt));
      st_shared_v4(dst + 128, pack2(u[8], u[9], p.fmt), pack2(u[10], u[11], p.fmt),
                   pack2(u[12], u[13], p.fmt), pack2(u[14], u[15], p.fmt));
      bias_s[i] = (p.bias != nullptr && n < p.N) ? __ldg(p.bias + n) : 0.f;
    }
    float coef[R_PAD];
#pragma unroll
    for (int j = 0; j < R_PAD; ++j)
      coef[j] = (j < p.r) ? p.scale * (p.diag ? __ldg(p.diag + j) : 1.f) : 0.f;

    // T = X.D^T out of TMEM -> (optional) global save -> scaled 16-bit operand in smem
    mbar_wait(bar_acc, 0);
    tc_fence_after();
    {
      uint32_t tv[R_PAD];
      tmem_ld16(tmem + (static_cast<uint32_t>(q * 32) << 16) + BLOCK_N, tv);
      tmem_ld_wait();
      float t[R_PAD];
#pragma unroll
      for (int j = 0; j < R_PAD; ++j) t[j] = __uint_as_float(tv[j]);
      if (p.t_out != nullptr && n_blk == 0 && m0 + row < p.M) {
        float4* dst = reinterpret_cast<float4*>(p.t_out + static_cast<size_t>(m0 + row) * R_PAD);
        dst[0] = make_float4(t[0], t[1], t[2], t[3]);
        dst[1] = make_float4(t[4], t[5], t[6], t[7]);
        dst[2] = make_float4(t[8], t[9], t[10], t[11]);
        dst[3] = make_float4(t[12], t[13], t[14], t[15]);
      }
#pragma unroll
      for (int j = 0; j < R_PAD; ++j) t[j] *= coef[j];
      const uint32_t dst = sbase + S::OFF_T + (row >> 3) * 256 + (row & 7) * 16;
      st_shared_v4(dst, pack2(t[0], t[1], p.fmt), pack2(t[2], t[3], p.fmt),
                   pack2(t[4], t[5], p.fmt), pack2(t[6], t[7], p.fmt));
      st_shared_v4(dst + 128, pack2(t[8], t[9], p.fmt), pack2(t[10], t[11], p.fmt),
                   pack2(t[12], t[13], p.fmt), pack2(t[14], t[15], p.fmt));
    }
    tc_fence_before();
    fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor-core proxy
    mbar_arrive(bar_tready);
    named_bar_sync(1, EPI_THREADS);  // bias_s complete for every epilogue thread

    // Final accumulator -> (+bias) -> OutT -> swizzled staging -> TMA store
    mbar_wait(bar_final, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < BLOCK_N / 32; ++c) {
      uint32_t v[32];
      tmem_ld32(tmem + (static_cast<uint32_t>(q * 32) << 16) + c * 32, v);
      tmem_ld_wait();
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) + bias_s[c * 32 + j];
      if constexpr (sizeof(OutT) == 2) {
        const int box = c >> 1;
        const uint32_t rbase = sbase + S::OFF_OUT + box * S::BOX_BYTES + row * 128;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          const int piece = ((c & 1) * 4 + qq) ^ (row & 7);
          st_shared_v4(rbase + piece * 16, pack2(f[qq * 8 + 0], f[qq * 8 + 1], p.fmt),
                       pack2(f[qq * 8 + 2], f[qq * 8 + 3], p.fmt),
                       pack2(f[qq * 8 + 4], f[qq * 8 + 5], p.fmt),
                       pack2(f[qq * 8 + 6], f[qq * 8 + 7], p.fmt));
        }
      } else {
        const uint32_t rbase = sbase + S::OFF_OUT + c * S::BOX_BYTES + row * 128;
#pragma unroll
        for (int qq = 0; qq < 8; ++qq) {
          const int piece = qq ^ (row & 7);
          st_shared_v4(rbase + piece * 16, __float_as_uint(f[qq * 4 + 0]),
                       __float_as_uint(f[qq * 4 + 1]), __float_as_uint(f[qq * 4 + 2]),
                       __float_as_uint(f[qq * 4 + 3]));
        }
      }
    }
    tc_fence_before();
    fence_proxy_async_smem();
    named_bar_sync(1, EPI_THREADS);
    if (et == 0) {
      for (int b = 0; b < S::NUM_BOXES; ++b) {
        const int col = n0 + b * S::BOX_COLS;
        if (col < p.N) tma_store_2d(&tmY, sbase + S::OFF_OUT + b * S::BOX_BYTES, col, m0);
      }
      tma_store_commit();
      tma_store_wait_read0();
    }
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, S::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------ host side
template <int BLOCK_N, int STAGES, typename OutT>
static int launch_cfg(const void* X, const void* W, const void* Dn, void* Y, const LinParams& p,
                      int out_dtype, cudaStream_t stream) {
  using S = Smem<BLOCK_N, STAGES, OutT>;
  auto kern = fused_lora_linear_kernel<BLOCK_N, STAGES, OutT>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::DYN_BYTES);
    if (e != cudaSuccess) return LB_ERR_CUDA;
    attr_set = true;
  }
  const CUtensorMapDataType in_dt = p.fmt ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUtensorMap tmX, tmW, tmD, tmY;
  if (!tmap_2d(&tmX, X, in_dt, 2, p.K, p.M, BLOCK_K, BLOCK_M, true)) return LB_ERR_TMAP;
  if (!tmap_2d(&tmW, W, in_dt, 2, p.K, p.N, BLOCK_K, BLOCK_N, true)) return LB_ERR_TMAP;
  if (!tmap_2d(&tmD, Dn, in_dt, 2, p.K, R_PAD, BLOCK_K, R_PAD, true)) return LB_ERR_TMAP;
  CUtensorMapDataType out_dt = out_dtype == LB_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : in_dt;
  if (!tmap_2d(&tmY, Y, out_dt, sizeof(OutT), p.N, p.M, S::BOX_COLS, BLOCK_M, true)) return LB_ERR_TMAP;

  dim3 grid((p.N + BLOCK_N - 1) / BLOCK_N, (p.M + BLOCK_M - 1) / BLOCK_M, 1);
  kern<<<grid, NUM_THREADS, S::DYN_BYTES, stream>>>(tmX, tmW, tmD, tmY, p);
  return cudaGetLastError() == cudaSuccess ? LB_OK : LB_ERR_CUDA;
}

}  // namespace lb

extern "C" int lb_lora_linear_fwd(const void* X, const void* W, const float* bias,
                                  const void* down16, const float* up, long long up_rs,
                                  long long up_cs, const float* diag, float scale, void* Y,
                                  float* T_out, int M, int K, int N, int r, int in_dtype,
                                  int out_dtype, void* stream) {
  using namespace lb;
  if (M <= 0 || N <= 0 || K <= 0) return LB_ERR_SHAPE;
  if (r < 1 || r > R_PAD) return LB_ERR_RANK;
  if (in_dtype != LB_BF16 && in_dtype != LB_F16) return LB_ERR_DTYPE;
  if (out_dtype != in_dtype && out_dtype != LB_F32) return LB_ERR_DTYPE;
  // TMA: 16-byte aligned base pointers and row pitches
  if ((K % 8) != 0) return LB_ERR_SHAPE;
  if (out_dtype == LB_F32 ? (N % 4) != 0 : (N % 8) != 0) return LB_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(W) |
       reinterpret_cast<uintptr_t>(down16) | reinterpret_cast<uintptr_t>(Y) |
       reinterpret_cast<uintptr_t>(T_out)) & 15)
    return LB_ERR_ALIGN;

  LinParams p;
  p.bias = bias; p.up = up; p.up_rs = up_rs; p.up_cs = up_cs; p.diag = diag; p.t_out = T_out;
  p.scale = scale; p.M = M; p.N = N; p.K = K; p.r = r; p.fmt = (in_dtype == LB_BF16) ? 1 : 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

  const long long tiles128 = static_cast<long long>((M + 127) / 128) * ((N + 127) / 128);
  const bool narrow = tiles128 < 120;  // not enough 128-wide tiles to fill 148 SMs: halve BLOCK_N
  if (out_dtype == LB_F32) {
    return narrow ? launch_cfg<64, 4, float>(X, W, down16, Y, p, out_dtype, st)
                  : launch_cfg<128, 4, float>(X, W, down16, Y, p, out_dtype, st);
  }
  return narrow ? launch_cfg<64, 4, uint16_t>(X, W, down16, Y, p, out_dtype, st)
                : launch_cfg<128, 4, uint16_t>(X, W, down16, Y, p, out_dtype, st);
}

// Persistent variant of the fused LoRA linear kernel (sm_100a): one CTA per SM walks a static
// round-robin list of output tiles; the TMEM accumulator is double-buffered so that the epilogue
// of tile i (TMEM drain, bias, 16-bit convert, TMA store) and the T' round trip of the LoRA branch
// overlap the TMA + tcgen05 main loop of tile i+1. The K loops of the SD1.5 sites are only 5-20
// steps long, so without this overlap every tile pays prologue + drain serially (measured: 18%
// tensor-pipe activity, 26 us for M=4096,K=320,N=2560; profiles/r1a_*).
//
// Per-tile math is identical to fused_core.cuh (G = 1):
//   acc[:, 0:BN] , T = X_tile . [W_tile ; D]^T        (one tcgen05.mma per K-step, N = BN + 16)
//   acc[:, 0:BN] += T' . U_tile^T                       (K = 16, T' = 16-bit(scale*diag*T))
//
// MMA-warp issue order (software pipelined):   main(0) main(1) lora(0) main(2) lora(1) ...
// so the tensor core never waits for the epilogue warps' T' round trip.
#pragma once
#include "fused_core.cuh"

namespace lb {

template <int BLOCK_N, int STAGES, typename OutT>
struct PSmem {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = (BLOCK_N + R_PAD) * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BOX_COLS = 128 / sizeof(OutT);
  static constexpr int NUM_BOXES = BLOCK_N / BOX_COLS;
  static constexpr int BOX_BYTES = BLOCK_M * 128;
  static constexpr int T_BYTES = BLOCK_M * R_PAD * 2;
  static constexpr int UP_BYTES = BLOCK_N * R_PAD * 2;
  static constexpr int OFF_OUT = STAGES * STAGE_BYTES;
  static constexpr int OFF_T = OFF_OUT + NUM_BOXES * BOX_BYTES;   // 2 buffers
  static constexpr int OFF_UP = OFF_T + 2 * T_BYTES;              // 2 buffers
  static constexpr int OFF_BIAS = OFF_UP + 2 * UP_BYTES;          // 2 buffers
  static constexpr int OFF_BAR = OFF_BIAS + 2 * BLOCK_N * 4;
  static constexpr int NUM_BARS = 2 * STAGES + 8;
  static constexpr int OFF_TMEM = OFF_BAR + NUM_BARS * 8;
  static constexpr int TOTAL = OFF_TMEM + 16;
  static constexpr int DYN_BYTES = TOTAL + 1024;
  static constexpr int ACC_STRIDE = BLOCK_N + R_PAD;              // TMEM columns per buffer
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE <= 128 ? 128 : (2 * ACC_STRIDE <= 256 ? 256 : 512);
  static_assert(DYN_BYTES <= 232448, "shared memory budget exceeded");
  static_assert(2 * ACC_STRIDE <= 512, "TMEM budget exceeded");
};

template <int BLOCK_N, int STAGES, typename OutT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
fused_lora_persistent_kernel(const __grid_constant__ CUtensorMap tmX,
                             const __grid_constant__ CUtensorMap tmW,
                             const __grid_constant__ CUtensorMap tmD,
                             const __grid_constant__ CUtensorMap tmY, const FusedParams p) {
  using S = PSmem<BLOCK_N, STAGES, OutT>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int m_tiles = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int total = n_tiles * m_tiles;
  const int num_kb = (p.K + BLOCK_K - 1) / BLOCK_K;

  auto bar_full = [&](int s) { return sbase + S::OFF_BAR + 8 * s; };
  auto bar_empty = [&](int s) { return sbase + S::OFF_BAR + 8 * (STAGES + s); };
  auto bar_acc = [&](int b) { return sbase + S::OFF_BAR + 8 * (2 * STAGES + 0 + b); };     // main loop done
  auto bar_tready = [&](int b) { return sbase + S::OFF_BAR + 8 * (2 * STAGES + 2 + b); };  // T' in smem
  auto bar_final = [&](int b) { return sbase + S::OFF_BAR + 8 * (2 * STAGES + 4 + b); };   // LoRA MMA done
  auto bar_tfree = [&](int b) { return sbase + S::OFF_BAR + 8 * (2 * STAGES + 6 + b); };   // TMEM buffer drained
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
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_acc(b), 1);
      mbar_init(bar_tready(b), EPI_THREADS);
      mbar_init(bar_final(b), 1);
      mbar_init(bar_tfree(b), EPI_THREADS);
    }
    fence_mbar_init();
  }
  __syncthreads();  // barriers initialised
  pdl_launch_dependents();   // the next kernel in the stream may start its own prologue now
  // The TMA producer starts streaming right away; TMEM allocation (a few hundred cycles) proceeds
  // concurrently in warp 1 and is published to the MMA/epilogue warps through named barrier 2.
  uint32_t tmem = 0;
  if (warp != 0) {
    if (warp == 1) {
      tmem_alloc(sbase + S::OFF_TMEM, S::TMEM_COLS);
      tmem_relinquish();
    }
    tc_fence_before();
    named_bar_sync(2, NUM_THREADS - 32);
    tc_fence_after();
    tmem = *tmem_slot;
  }
  pdl_wait();                // predecessor grid complete, its writes visible: global reads may start

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t it = 0;  // global K-step counter: the smem ring runs across tiles
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) * BLOCK_M, n0 = (tile % n_tiles) * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(bar_empty(s), ph ^ 1);
          mbar_expect_tx(bar_full(s), S::STAGE_BYTES);
          const uint32_t sa = sbase + s * S::STAGE_BYTES;
          const uint32_t sb = sa + S::A_BYTES;
          tma_load_2d(&tmX, bar_full(s), sa, kb * BLOCK_K, m0);
          load_w_tile<BLOCK_N>(&tmW, bar_full(s), sb, kb * BLOCK_K, kb, num_kb, n0, p.w_tiled);
          tma_load_2d(&tmD, bar_full(s), sb + BLOCK_N * 128, kb * BLOCK_K, 0);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0) {
      const uint32_t idesc_wide = umma_idesc_f16(p.fmt, BLOCK_M, BLOCK_N + R_PAD);
      const uint32_t idesc_base = umma_idesc_f16(p.fmt, BLOCK_M, BLOCK_N);
      auto lora_mma = [&](int j) {  // acc(j) += T'(j) . U(j)^T, then publish the finished tile
        const int b = j & 1;
        mbar_wait(bar_tready(b), (j >> 1) & 1);
        tc_fence_after();
        const uint64_t ad = umma_smem_desc(sbase + S::OFF_T + b * S::T_BYTES, 128, 256, 0);
        const uint64_t bd = umma_smem_desc(sbase + S::OFF_UP + b * S::UP_BYTES, 128, 256, 0);
        umma_f16_ss(tmem + b * S::ACC_STRIDE, ad, bd, idesc_base, 1);
        umma_commit(bar_final(b));
      };
      uint32_t it = 0;
      int i = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++i) {
        const int b = i & 1;
        mbar_wait(bar_tfree(b), ((i >> 1) & 1) ^ 1);  // epilogue has drained this TMEM buffer
        tc_fence_after();
        const uint32_t acc = tmem + b * S::ACC_STRIDE;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(bar_full(s), ph);
          tc_fence_after();
          const uint32_t sa = sbase + s * S::STAGE_BYTES;
          const uint32_t sb = sa + S::A_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t ad = umma_smem_desc(sa + k * UMMA_K * 2, 16, 1024, 2);
            const uint64_t bd = umma_smem_desc(sb + k * UMMA_K * 2, 16, 1024, 2);
            umma_f16_ss(acc, ad, bd, idesc_wide, (kb | k) != 0);
          }
          umma_commit(bar_empty(s));
        }
        umma_commit(bar_acc(b));
        if (i > 0) lora_mma(i - 1);
      }
      if (i > 0) lora_mma(i - 1);
    }
  } else {
    // ------------------------------------------------------------ epilogue warps (128 threads)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;
    float coef[R_PAD];
#pragma unroll
    for (int j = 0; j < R_PAD; ++j)
      coef[j] = (j < p.r) ? p.scale * (p.diag ? __ldg(p.diag + j) : 1.f) : 0.f;

    int i = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++i) {
      const int b = i & 1;
      const uint32_t ph = (i >> 1) & 1;
      const int m0 = (tile / n_tiles) * BLOCK_M, n0 = (tile % n_tiles) * BLOCK_N;
      const uint32_t acc = tmem + b * S::ACC_STRIDE + (static_cast<uint32_t>(q * 32) << 16);
      float* bias_s = reinterpret_cast<float*>(sgen + S::OFF_BIAS) + b * BLOCK_N;

      // U tile + bias of this tile. Buffer b was last read by lora_mma(i-2), whose completion
      // (bar_final) this warp waited for before draining tile i-2.
      for (int c = et; c < BLOCK_N; c += EPI_THREADS) {
        const int n = n0 + c;
        const bool ok = n < p.N;
        float u[R_PAD];
#pragma unroll
        for (int j = 0; j < R_PAD; ++j)
          u[j] = (ok && j < p.r) ? __ldg(p.up + n * p.up_rs + j * p.up_cs) : 0.f;
        const uint32_t dst = sbase + S::OFF_UP + b * S::UP_BYTES + (c >> 3) * 256 + (c & 7) * 16;
        st_shared_v4(dst, pack2(u[0], u[1], p.fmt), pack2(u[2], u[3], p.fmt),
                     pack2(u[4], u[5], p.fmt), pack2(u[6], u[7], p.fmt));
        st_shared_v4(dst + 128, pack2(u[8], u[9], p.fmt), pack2(u[10], u[11], p.fmt),
                     pack2(u[12], u[13], p.fmt), pack2(u[14], u[15], p.fmt));
        bias_s[c] = (p.bias != nullptr && ok) ? __ldg(p.bias + n) : 0.f;
      }

      // T out of TMEM -> optional global save -> scaled 16-bit operand
      mbar_wait(bar_acc(b), ph);
      tc_fence_after();
      {
        float t[R_PAD];
        if (p.t_in == nullptr) {
          uint32_t tv[R_PAD];
          tmem_ld16(acc + BLOCK_N, tv);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < R_PAD; ++j) t[j] = __uint_as_float(tv[j]);
        } else if (m0 + row < p.M) {
          const float4* src = reinterpret_cast<const float4*>(p.t_in + static_cast<size_t>(m0 + row) * R_PAD);
          const float4 a = __ldg(src), bq = __ldg(src + 1), c = __ldg(src + 2), d = __ldg(src + 3);
          t[0] = a.x; t[1] = a.y; t[2] = a.z; t[3] = a.w; t[4] = bq.x; t[5] = bq.y; t[6] = bq.z; t[7] = bq.w;
          t[8] = c.x; t[9] = c.y; t[10] = c.z; t[11] = c.w; t[12] = d.x; t[13] = d.y; t[14] = d.z; t[15] = d.w;
        } else {
#pragma unroll
          for (int j = 0; j < R_PAD; ++j) t[j] = 0.f;
        }
        if (p.t_out != nullptr && n0 == 0 && m0 + row < p.M) {
          float4* dst = reinterpret_cast<float4*>(p.t_out + static_cast<size_t>(m0 + row) * R_PAD);
          dst[0] = make_float4(t[0], t[1], t[2], t[3]);
          dst[1] = make_float4(t[4], t[5], t[6], t[7]);
          dst[2] = make_float4(t[8], t[9], t[10], t[11]);
          dst[3] = make_float4(t[12], t[13], t[14], t[15]);
        }
#pragma unroll
        for (int j = 0; j < R_PAD; ++j) t[j] *= coef[j];
        const uint32_t dst = sbase + S::OFF_T + b * S::T_BYTES + (row >> 3) * 256 + (row & 7) * 16;
        st_shared_v4(dst, pack2(t[0], t[1], p.fmt), pack2(t[2], t[3], p.fmt),
                     pack2(t[4], t[5], p.fmt), pack2(t[6], t[7], p.fmt));
        st_shared_v4(dst + 128, pack2(t[8], t[9], p.fmt), pack2(t[10], t[11], p.fmt),
                     pack2(t[12], t[13], p.fmt), pack2(t[14], t[15], p.fmt));
      }
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(bar_tready(b));

      // the previous tile's TMA store must have finished READING the staging buffer
      if (et == 0) tma_store_wait_read0();
      named_bar_sync(1, EPI_THREADS);  // also publishes bias_s to every epilogue thread

      mbar_wait(bar_final(b), ph);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(acc + c * 32, v);
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
      mbar_arrive(bar_tfree(b));      // TMEM buffer b may be overwritten by tile i+2
      fence_proxy_async_smem();
      named_bar_sync(1, EPI_THREADS);
      if (et == 0) {
        for (int bx = 0; bx < S::NUM_BOXES; ++bx) {
          const int col = n0 + bx * S::BOX_COLS;
          if (col >= p.N) break;
          tma_store_2d(&tmY, sbase + S::OFF_OUT + bx * S::BOX_BYTES, col, m0);
        }
        tma_store_commit();
      }
    }
    if (et == 0) tma_store_wait_read0();
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, S::TMEM_COLS);
  }
}

}  // namespace lb

// Cluster split-K variant of the fused LoRA linear kernel for the few-tile sites (written at the end of
// round 1, first run in round 2; since visit 8 the planner's choice for K >= 640 sites whose 64-wide
// tiles fill at most half the SMs -- lb_lora_linear_fwd in fused_linear.cu; lb_debug_set_linear_mode
// schedule 3 forces it with 2..4 CTAs). Measured: profiles/r2g_site_table_cluster_splitk.md.
//
// Why: a CTA pulls operands at ~50 B/clk however many CTAs are running (per-SM L2->SMEM port), so a
// 77x768->768 or 256x1280->1280 site -- 12...40 tiles -- streams its whole K loop on a handful of
// SMs while the rest idle. Here SPLIT CTAs of one thread-block cluster share an
// output tile: CTA `rank` reduces K-blocks [rank*nkb/SPLIT, (rank+1)*nkb/SPLIT) into its own TMEM
// accumulator (base columns AND the 16 rank-r T columns -- both are linear in K). The non-leader
// CTAs then stage their fp32 accumulator in shared memory, column-major [BLOCK_N+16][128], and
// push it into the leader's slot `rank-1` with ONE cp.async.bulk over distributed shared memory
// (completion = complete_tx on the leader's mbarrier; no global atomics, no second launch). The
// leader runs the ordinary epilogue with the peers' partials added in registers: T = sum of the T
// partials -> T' -> the K=16 LoRA MMA into its own accumulator; the drain adds the peers' base
// partials and the bias on the way to the staging buffer.
//
// Lifetime rules for distributed shared memory: the leader does not leave before every peer's
// copy has landed (it waits on bar_red); a peer does not leave before its copy has been read out
// of its own shared memory -- the copy signals only the DESTINATION barrier, so the peer learns
// it through the closing cluster barrier, which the leader's epilogue threads arrive at only
// after bar_red completed.
#pragma once
#include "fused_core.cuh"

namespace lb {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// relaxed arrive: nothing is published through the cluster barrier itself -- barrier
// initialisation is released by fence.mbarrier_init, the partials travel by bulk copy + mbarrier
// (a .release arrive would cost a MEMBAR.ALL.GPU per thread)
__device__ __forceinline__ void cluster_arrive_relaxed() {
  asm volatile("barrier.cluster.arrive.relaxed;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait_acquire() {
  asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
// shared::cta address of THIS CTA -> shared::cluster address of the same offset in CTA `rank`
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
// bulk copy own shared memory -> a peer CTA's shared memory; completes on the PEER's mbarrier
__device__ __forceinline__ void dsmem_bulk_copy(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes,
                                                uint32_t bar_cluster) {
  asm volatile(
      "cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst_cluster), "r"(src_cta), "r"(bytes), "r"(bar_cluster)
      : "memory");
}
__device__ __forceinline__ void st_shared_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float ld_shared_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

template <int BLOCK_N, int STAGES, typename OutT, int SPLIT>
struct SplitSmem {
  using Base = Smem<BLOCK_N, STAGES, OutT, 1>;
  static constexpr int A_BYTES = Base::A_BYTES;
  static constexpr int STAGE_BYTES = Base::STAGE_BYTES;
  static constexpr int BOX_COLS = Base::BOX_COLS;
  static constexpr int NUM_BOXES = Base::NUM_BOXES;
  static constexpr int BOX_BYTES = Base::BOX_BYTES;
  static constexpr int ACC_COLS = BLOCK_N + R_PAD;
  static constexpr int SLOT_BYTES = ACC_COLS * BLOCK_M * 4;       // one fp32 partial, column-major
  static constexpr int OFF_SLOTS = STAGES * STAGE_BYTES;          // leader: SPLIT-1 slots; peer: slot 0 = staging
  static constexpr int OFF_EPI = OFF_SLOTS + (SPLIT - 1) * SLOT_BYTES;
  static constexpr int OFF_OUT = OFF_EPI;
  static constexpr int OFF_T = OFF_EPI;
  static constexpr int OFF_UP = OFF_T + Base::T_BYTES;
  static constexpr int OFF_BIAS = OFF_EPI + ((Base::EPI_BYTES + 1023) / 1024) * 1024;
  static constexpr int OFF_BAR = OFF_BIAS + BLOCK_N * 4;
  static constexpr int NUM_BARS = 2 * STAGES + 4;                 // + bar_red
  static constexpr int OFF_TMEM = OFF_BAR + NUM_BARS * 8;
  static constexpr int TOTAL = OFF_TMEM + 16;
  static constexpr int DYN_BYTES = TOTAL + 1024;
  static constexpr int TMEM_COLS = Base::TMEM_COLS;
  static_assert(SPLIT >= 2 && SPLIT <= 4, "cluster split factor");
  static_assert(OFF_SLOTS % 1024 == 0 && SLOT_BYTES % 1024 == 0, "slot alignment");
  static_assert(DYN_BYTES <= 232448, "shared memory budget exceeded");
};

template <int BLOCK_N, int STAGES, typename OutT, int SPLIT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
fused_lora_splitk_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                         const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmY,
                         const FusedParams p) {
  using S = SplitSmem<BLOCK_N, STAGES, OutT, SPLIT>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;   // identical in every CTA of the kernel
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();            // cluster = (1, 1, SPLIT): rank == blockIdx.z
  const bool leader = rank == 0;
  const int n_blk = blockIdx.x, m_blk = blockIdx.y;
  const int n0 = n_blk * BLOCK_N, m0 = m_blk * BLOCK_M;
  const int num_kb = (p.K + BLOCK_K - 1) / BLOCK_K;   // launcher guarantees num_kb >= SPLIT
  const int kb_begin = static_cast<int>((static_cast<long long>(rank) * num_kb) / SPLIT);
  const int kb_end = static_cast<int>((static_cast<long long>(rank + 1) * num_kb) / SPLIT);

  auto bar_full = [&](int s) { return sbase + S::OFF_BAR + 8 * s; };
  auto bar_empty = [&](int s) { return sbase + S::OFF_BAR + 8 * (STAGES + s); };
  const uint32_t bar_acc = sbase + S::OFF_BAR + 8 * (2 * STAGES + 0);
  const uint32_t bar_tready = sbase + S::OFF_BAR + 8 * (2 * STAGES + 1);
  const uint32_t bar_final = sbase + S::OFF_BAR + 8 * (2 * STAGES + 2);
  const uint32_t bar_red = sbase + S::OFF_BAR + 8 * (2 * STAGES + 3);     // peers' partials landed (leader)
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sgen + S::OFF_TMEM);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmD);
    if (leader) tma_prefetch_desc(&tmY);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
    }
    mbar_init(bar_acc, 1);
    mbar_init(bar_tready, EPI_THREADS);
    mbar_init(bar_final, 1);
    mbar_init(bar_red, 1);
    fence_mbar_init();
    if (leader) mbar_expect_tx(bar_red, (SPLIT - 1) * S::SLOT_BYTES);    // the one arrival + all peer bytes
  }
  // cluster-wide: every CTA's barriers exist before any peer signals them
  cluster_arrive_relaxed();
  cluster_wait_acquire();
  pdl_launch_dependents();

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
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (this CTA's K range)
    if (lane == 0) {
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        const int it = kb - kb_begin;
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
    __syncwarp();
    cluster_arrive_relaxed();          // this warp has no distributed-shared-memory business
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc_wide = umma_idesc_f16(p.fmt, BLOCK_M, BLOCK_N + R_PAD);
      const uint32_t idesc_base = umma_idesc_f16(p.fmt, BLOCK_M, BLOCK_N);
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        const int it = kb - kb_begin;
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
          umma_f16_ss(tmem, ad, bd, idesc_wide, (it | k) != 0);
        }
        umma_commit(bar_empty(s));
      }
      umma_commit(bar_acc);
      if (leader) {
        mbar_wait(bar_tready, 0);
        tc_fence_after();
        const uint64_t ad = umma_smem_desc(sbase + S::OFF_T, 128, 256, 0);
        const uint64_t bd = umma_smem_desc(sbase + S::OFF_UP, 128, 256, 0);
        umma_f16_ss(tmem, ad, bd, idesc_base, 1);
        umma_commit(bar_final);
      }
    }
    __syncwarp();
    cluster_arrive_relaxed();
  } else {
    // ------------------------------------------------------------ epilogue warps
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;
    const uint32_t lane_base = tmem + (static_cast<uint32_t>(q * 32) << 16);

    if (!leader) {
      // own partial: TMEM -> fp32 column-major staging (conflict-free: lane = row) -> leader's slot
      mbar_wait(bar_acc, 0);
      tc_fence_after();
      const uint32_t stage = sbase + S::OFF_SLOTS;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(lane_base + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j)
          st_shared_f32(stage + ((c * 32 + j) * BLOCK_M + row) * 4, __uint_as_float(v[j]));
      }
      {
        uint32_t tv[R_PAD];
        tmem_ld16(lane_base + BLOCK_N, tv);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < R_PAD; ++j)
          st_shared_f32(stage + ((BLOCK_N + j) * BLOCK_M + row) * 4, __uint_as_float(tv[j]));
      }
      tc_fence_before();
      fence_proxy_async_smem();              // generic writes -> visible to the bulk-copy (async) proxy
      named_bar_sync(1, EPI_THREADS);
      if (et == 0) {
        const uint32_t dst = mapa_u32(sbase + S::OFF_SLOTS + (rank - 1) * S::SLOT_BYTES, 0);
        dsmem_bulk_copy(dst, stage, S::SLOT_BYTES, mapa_u32(bar_red, 0));
      }
      __syncwarp();
      cluster_arrive_relaxed();
    } else {
      float* bias_s = reinterpret_cast<float*>(sgen + S::OFF_BIAS);
      for (int i = et; i < BLOCK_N; i += EPI_THREADS) {
        const int n = n0 + i;
        const bool ok = n < p.N;
        float u[R_PAD];
#pragma unroll
        for (int j = 0; j < R_PAD; ++j)
          u[j] = (ok && j < p.r) ? __ldg(p.up + n * p.up_rs + j * p.up_cs) : 0.f;
        const uint32_t dst = sbase + S::OFF_UP + (i >> 3) * 256 + (i & 7) * 16;
        st_shared_v4(dst, pack2(u[0], u[1], p.fmt), pack2(u[2], u[3], p.fmt),
                     pack2(u[4], u[5], p.fmt), pack2(u[6], u[7], p.fmt));
        st_shared_v4(dst + 128, pack2(u[8], u[9], p.fmt), pack2(u[10], u[11], p.fmt),
                     pack2(u[12], u[13], p.fmt), pack2(u[14], u[15], p.fmt));
        bias_s[i] = (p.bias != nullptr && ok) ? __ldg(p.bias + n) : 0.f;
      }
      float coef[R_PAD];
#pragma unroll
      for (int j = 0; j < R_PAD; ++j)
        coef[j] = (j < p.r) ? p.scale * (p.diag ? __ldg(p.diag + j) : 1.f) : 0.f;
      const long long grow = (m0 + row < p.M) ? m0 + row : -1;

      mbar_wait(bar_acc, 0);                 // own K range reduced
      mbar_wait(bar_red, 0);                 // every peer's partial is in its slot
      __syncwarp();
      cluster_arrive_relaxed();              // peers may retire: their staging has been read out
      tc_fence_after();
      const uint32_t slots = sbase + S::OFF_SLOTS;
      {
        uint32_t tv[R_PAD];
        tmem_ld16(lane_base + BLOCK_N, tv);
        tmem_ld_wait();
        float t[R_PAD];
#pragma unroll
        for (int j = 0; j < R_PAD; ++j) {
          t[j] = __uint_as_float(tv[j]);
#pragma unroll
          for (int s = 0; s < SPLIT - 1; ++s)
            t[j] += ld_shared_f32(slots + s * S::SLOT_BYTES + ((BLOCK_N + j) * BLOCK_M + row) * 4);
        }
        if (p.t_out != nullptr && n_blk == 0 && grow >= 0) {
          float4* dst = reinterpret_cast<float4*>(p.t_out + grow * R_PAD);
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
      fence_proxy_async_smem();
      mbar_arrive(bar_tready);
      named_bar_sync(1, EPI_THREADS);        // bias_s complete

      mbar_wait(bar_final, 0);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(lane_base + c * 32, v);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          f[j] = __uint_as_float(v[j]) + bias_s[c * 32 + j];
#pragma unroll
          for (int s = 0; s < SPLIT - 1; ++s)
            f[j] += ld_shared_f32(slots + s * S::SLOT_BYTES + ((c * 32 + j) * BLOCK_M + row) * 4);
        }
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
          if (col >= p.N) break;
          tma_store_2d(&tmY, sbase + S::OFF_OUT + b * S::BOX_BYTES, col, m0);
        }
        tma_store_commit();
        tma_store_wait_read0();
      }
    }
  }

  cluster_wait_acquire();   // closes the barrier every thread arrived at above (see header comment)
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, S::TMEM_COLS);
  }
}

}  // namespace lb

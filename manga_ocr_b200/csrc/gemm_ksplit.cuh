// Large-batch decoder program: the residual projections  x += A W^T + bias  (N = 768; K = 768 or 3072) with the K dimension
// split over a CLUSTER OF FOUR CTAs.
//
// Why: with a few hundred rows these GEMMs have only 4 m-tiles, and a tcgen05 MMA costs ~130 cycles whatever its width (its
// 4 KB A operand is read from shared memory per instruction, profiles/r2_gemm_limits.txt), so a CTA that walks the whole K
// issues K/16 MMAs back to back: 12.5 us at K = 3072 however the N dimension is tiled.  Only a K split shortens that chain.
// Here the four CTAs of a cluster share one 128 x 128 output tile: CTA r multiplies k-blocks [r K/4, (r+1) K/4) into its own
// TMEM accumulator, then the tile is reduced through distributed shared memory - CTA r OWNS the 32-column chunk r: every peer
// writes its partial of that chunk straight into r's receive buffer (st.shared::cluster, 128-byte rows) and arrives on r's
// mbarrier; r adds the four partials in rank order (fixed order: results do not depend on timing) and runs the epilogue of
// its chunk - bias, then a TMA reduce-add into the fp32 residual stream, as EPI_F32_ACCUM of gemm_tcgen05.cuh does.
// Reference: BertSelfOutput / BertOutput dense + residual (modeling_bert.py:287-298, 343-356).
#pragma once
#include "gemm_tcgen05.cuh"

namespace mocr {

constexpr int kKsSplit = 4;                                   // CTAs per cluster = K shares = 32-column chunks of the tile
constexpr int kKsBN = 32 * kKsSplit;                          // 128
constexpr int kKsThreads = 192;                               // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue (one per TMEM lane quadrant)
constexpr int kKsStages = 4;
constexpr int kKsStageBytesA = kGemmBM * kGemmBK * 2;         // 16 KB
constexpr int kKsStageBytesB = kKsBN * kGemmBK * 2;           // 16 KB
constexpr int kKsStageBytes = kKsStageBytesA + kKsStageBytesB;
constexpr int kKsRecvSlot = 32 * kGemmBM * 4;                 // one peer's partial of this CTA's chunk: [32 columns][128 rows] fp32
constexpr int kKsRecvBytes = (kKsSplit - 1) * kKsRecvSlot;    // 48 KB
constexpr int kKsSmemBytes = kKsStages * kKsStageBytes + kKsRecvBytes + 4 * kGemmAccumStage + 1024 /*align slack*/ + 256 /*barriers*/;

__device__ __forceinline__ void st_cluster_f32(uint32_t cluster_addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(cluster_addr), "f"(v) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait for arrivals that peers of the cluster made after writing this CTA's shared memory (acquire at cluster scope); bounded
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000LL) {   // ~2 s: a protocol bug becomes a trapped launch, not a hung GPU
      printf("mocr: cluster mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// grid = tiles * 4 (tiles = m_tiles * N / 128, one tile per cluster), block = 192.  args: M, N, K, bias, tmap_out (fp32 [M, N] view of
// the residual stream, box 32 x 32, SWIZZLE_128B), pdl.
__global__ void __cluster_dims__(kKsSplit, 1, 1) __launch_bounds__(kKsThreads, 1)
gemm_ksplit_accum_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ GemmArgs args) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kKsStages * kKsStageBytesA;
  float* recv = reinterpret_cast<float*>(smem + kKsStages * kKsStageBytes);
  uint8_t* stage_tiles = smem + kKsStages * kKsStageBytes + kKsRecvBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_tiles + 4 * kGemmAccumStage);
  uint64_t* full_bar = bars;                     // [kKsStages]  TMA -> MMA
  uint64_t* empty_bar = bars + kKsStages;        // [kKsStages]  MMA -> TMA
  uint64_t* acc_full = bars + 2 * kKsStages;     //              MMA -> epilogue
  uint64_t* recv_bar = bars + 2 * kKsStages + 1; //              peers' epilogue threads -> this CTA's epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kKsStages + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = static_cast<int>(cluster_ctarank());
  const int tile = blockIdx.x / kKsSplit;
  const int n_tiles = args.N / kKsBN;
  const int m0 = (tile / n_tiles) * kGemmBM, n0 = (tile % n_tiles) * kKsBN;
  const int kb_per = args.K / kGemmBK / kKsSplit;
  const int kb0 = rank * kb_per;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < kKsStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(recv_bar, (kKsSplit - 1) * 128);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kKsBN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                            // every CTA of the cluster runs and its barriers exist
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (args.pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    // ------------------------------------------------ TMA producer: weights before the dependency wait, activations after ----
    const int pre = kb_per < kKsStages ? kb_per : kKsStages;
    if (lane == 0) {
      for (int s = 0; s < pre; ++s) {
        mbar_arrive_expect_tx(&full_bar[s], kKsStageBytes);
        tma_load_2d(smem_b + s * kKsStageBytesB, &tmap_b, &full_bar[s], (kb0 + s) * kGemmBK, n0);
      }
    }
    __syncwarp();
    if (args.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < kb_per; ++i) {
      mbar_wait(&empty_bar[stage], phase ^ 1u);
      if (lane == 0) {
        if (i >= pre) {
          mbar_arrive_expect_tx(&full_bar[stage], kKsStageBytes);
          tma_load_2d(smem_b + stage * kKsStageBytesB, &tmap_b, &full_bar[stage], (kb0 + i) * kGemmBK, n0);
        }
        tma_load_2d(smem_a + stage * kKsStageBytesA, &tmap_a, &full_bar[stage], (kb0 + i) * kGemmBK, m0);
      }
      __syncwarp();
      if (++stage == kKsStages) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer ------------
    constexpr uint32_t idesc = umma_idesc_bf16(kGemmBM, kKsBN);
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < kb_per; ++i) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (lane == 0) {
        const uint64_t da = umma_desc_k_sw128(smem_u32(smem_a + stage * kKsStageBytesA));
        const uint64_t db = umma_desc_k_sw128(smem_u32(smem_b + stage * kKsStageBytesB));
#pragma unroll
        for (int k = 0; k < kGemmBK / 16; ++k)
          umma_bf16(tmem_base, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc, static_cast<uint32_t>((i | k) != 0));
        umma_commit(&empty_bar[stage]);
        if (i == kb_per - 1) umma_commit(acc_full);
      }
      __syncwarp();
      if (++stage == kKsStages) { stage = 0; phase ^= 1u; }
    }
  } else {
    // ------------------------------------------------ epilogue: exchange the partial chunks, reduce and store the own one ----
    const int quad = warp & 3;                    // warps 2, 3, 4, 5 -> TMEM lane quadrants 2, 3, 0, 1
    const int row_t = quad * 32 + lane;           // row of the tile this thread owns (= TMEM lane)
    mbar_wait(acc_full, 0u);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < kKsSplit; ++c) {
      if (c == rank) continue;
      uint32_t v[32];
      tmem_ld32(taddr + static_cast<uint32_t>(c * 32), v);
      tmem_ld_wait();
      const int slot = rank < c ? rank : rank - 1;                 // this CTA's slot among the three senders of chunk c
      const uint32_t dst = mapa_u32(recv + static_cast<size_t>(slot) * (kKsRecvSlot / 4) + row_t, static_cast<uint32_t>(c));
#pragma unroll
      for (int j = 0; j < 32; ++j) st_cluster_f32(dst + static_cast<uint32_t>(j * kGemmBM * 4), __uint_as_float(v[j]));   // a warp writes 128 contiguous bytes
      mbar_arrive_cluster(mapa_u32(recv_bar, static_cast<uint32_t>(c)));
    }
    uint32_t own[32];
    tmem_ld32(taddr + static_cast<uint32_t>(rank * 32), own);
    tmem_ld_wait();
    mbar_wait_cluster(recv_bar, 0u);
    const int col0 = n0 + 32 * rank;
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < kKsSplit; ++q) {                          // rank order, whoever this CTA is
        const int slot = q < rank ? q : q - 1;
        const float pq = q == rank ? __uint_as_float(own[j]) : recv[static_cast<size_t>(slot) * (kKsRecvSlot / 4) + j * kGemmBM + row_t];
        s = q == 0 ? pq : s + pq;
      }
      f[j] = s;
    }
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 bb = __ldg(reinterpret_cast<const float4*>(args.bias + col0 + j));
      f[j] += bb.x; f[j + 1] += bb.y; f[j + 2] += bb.z; f[j + 3] += bb.w;
    }
    // the warp's 32 x 32 fp32 tile -> swizzled staging tile -> ONE reduce-add into the residual stream (rows >= M clipped by the map)
    uint8_t* stage_w = stage_tiles + (warp - 2) * kGemmAccumStage;
    const uint32_t dsts = smem_u32(stage_w) + static_cast<uint32_t>(lane * 128);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dsts + static_cast<uint32_t>((j ^ (lane & 7)) << 4)), "f"(f[4 * j]),
                   "f"(f[4 * j + 1]), "f"(f[4 * j + 2]), "f"(f[4 * j + 3])
                   : "memory");
    fence_proxy_async_smem();
    __syncwarp();
    const int row0 = m0 + quad * 32;
    if (lane == 0 && row0 < args.M) {
      tma_reduce_add_2d(&args.tmap_out, stage_w, col0, row0);
      bulk_commit_group();
      bulk_wait_group0();
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                            // nobody leaves while a peer may still write its receive buffer or signal its barrier
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kKsBN);
  }
}

}  // namespace mocr

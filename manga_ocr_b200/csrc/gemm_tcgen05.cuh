// C[M,N] = A[M,K] * W[N,K]^T (+ fused epilogue) on the 5th-gen tensor cores.
//
// One kernel serves every dense contraction on the path: the ViT patch-embed,
// QKV / out-proj / MLP GEMMs (reference: transformers/models/vit/modeling_vit.py
// :151,228-230,266,297-311), the once-per-crop cross-attention K/V projection
// (models/bert/modeling_bert.py:252-267) and the small-M decoder GEMMs incl. the
// LM head (modeling_bert.py:179-181,250,295,340-355,482-500).
//
// Design (sm_100a):
//   * persistent CTAs, static tile schedule (tile = blockIdx.x + i*gridDim.x);
//   * warp 0: TMA producer (cp.async.bulk.tensor, 128-byte swizzle, K-major tiles,
//     out-of-bounds rows zero-filled so ragged M needs no special case);
//   * warp 1: single-thread tcgen05.mma issuer, UMMA 128 x BN x 16, fp32
//     accumulators in TMEM, double-buffered (2 x BN columns) so the epilogue of
//     tile i overlaps the main loop of tile i+1;
//   * warps 2-9: epilogue - two warps per TMEM lane quadrant, each owning half of the tile's
//     columns; tcgen05.ld (32 lanes x 32 columns per instruction), fused
//     bias / erf-GELU / fp32 residual / position-embedding / arg-max, direct
//     vectorised global stores.
#pragma once
#include "common.cuh"

namespace mocr {

enum GemmEpilogue : int {
  EPI_BF16 = 0,        // out(bf16) = acc + bias
  EPI_BF16_GELU = 1,   // out(bf16) = gelu_erf(acc + bias)
  EPI_F32_RESID = 2,   // out(f32)  = acc + bias + resid(f32)      (resid may alias out)
  EPI_PATCH = 3,       // ViT embeddings: out(f32)[b*197+1+p] = acc + bias + pos[1+p]
  EPI_ARGMAX = 4,      // LM head: per-row (max, argmax) over this tile's columns; logits optional
  EPI_F32_GELU = 5,    // out(f32)  = gelu_erf(acc + bias)         (LM-head transform, feeds a LayerNorm)
  EPI_CROSSKV = 6,     // out(bf16) = acc + bias, scattered into the per-head cross-attention K/V cache layout
  EPI_F32_ACCUM = 7,   // out(f32) += acc + bias, in place, by TMA reduce-add from a swizzled smem tile (encoder residual stream)
};

struct GemmArgs {
  int M, N, K;
  const float* bias;     // [N]
  void* out;             // bf16 or f32, row stride ldo elements
  int ldo;
  const float* resid;    // EPI_F32_RESID: [M, ldr] f32
  int ldr;
  const float* pos;      // EPI_PATCH: position table [197, 768] f32
  float* part_max;       // EPI_ARGMAX: [M, 2 * N/BN] (one entry per epilogue warp half)
  int* part_idx;         // EPI_ARGMAX: [M, 2 * N/BN]
  float* logits;         // EPI_ARGMAX: optional f32 tap for parity tests (may be null):
  const int* step;       //   row r writes logits[(r * tap_steps + step[r]) * N ...]
  int tap_steps;
  const int* crop_map;  // EPI_CROSSKV: optional, cache block of the batch's crop c is crop_map[c] (null: c) - admissions into free slots of a session
  int pdl;               // 1: launched with programmatic stream serialization (decoder stage): weights are requested before
                         //    griddepcontrol.wait, the activations after it
  int out_tma;           // EPI_BF16 / EPI_BF16_GELU with BN = 128 or 256: rows leave through tmap_out (bf16 [M, N] view of `out`, box 64 x 32,
                         //    SWIZZLE_128B) as TMA stores from a staging tile instead of per-lane 16-byte stores
#ifdef MOCR_GEMM_DBG
  int dbg;               // timing experiments only (results are garbage): 1 = loads stop after the first ring fill (MMA-side speed),
                         // 2 = no MMA issued (TMA-side speed), 3 = no epilogue work
#endif
  alignas(64) CUtensorMap tmap_out;   // EPI_F32_ACCUM: f32 [M, N] view of `out`, box 32 x 32, SWIZZLE_128B
};

constexpr int kGemmBM = 128;
constexpr int kGemmBK = 64;
constexpr int kGemmThreads = 320;   // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (2 per TMEM lane quadrant)
constexpr int kGemmEpiThreads = 256;
constexpr int kGemmAccumStage = 32 * 128;                      // EPI_F32_ACCUM: one 32-row x 128-byte staging tile per epilogue warp
constexpr int kGemmAccumSmem = 1024 + 8 * kGemmAccumStage;     // (+ padding that keeps the tiles 1024-byte aligned)
constexpr int gemm_smem_bytes(int base, int epi) { return base + ((epi == 7 || epi == 0 || epi == 1) ? kGemmAccumSmem : 0); }   // staging tiles: EPI_F32_ACCUM, EPI_BF16[_GELU]

template <int BN>
struct GemmCfg {
  static constexpr int kStageBytesA = kGemmBM * kGemmBK * 2;
  static constexpr int kStageBytesB = BN * kGemmBK * 2;
  static constexpr int kStageBytes = kStageBytesA + kStageBytesB;
  static constexpr int kStagesFit = (192 * 1024) / kStageBytes;      // + 32 KB of epilogue staging tiles + barriers <= 227 KB
  static constexpr int kStages = kStagesFit > 8 ? 8 : kStagesFit;
  static constexpr int kTmemCols = 2 * BN <= 32 ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

// Epilogue of one accumulator tile for one epilogue warp: TMEM -> registers -> fused bias /
// activation / residual / arg-max -> global.  `taddr` addresses the warp's 32 TMEM lanes at the
// tile's first column; `half` selects which half of the tile's 32-column chunks this warp owns.
// The tile's bias values, one column per lane and chunk, requested BEFORE the wait on the
// accumulator: this only warms L1 for the epilogue's float4 loads (a shuffle-broadcast variant and a
// software-pipelined tcgen05.ld were measured slower: +32 SHFL per chunk made the GELU epilogue
// instruction-bound).
template <int BN>
struct GemmBiasLanes {
  float v[(BN / 32 + 1) / 2];
};
template <int BN>
__device__ __forceinline__ void gemm_load_bias(const GemmArgs& args, int n0, int half, int lane, GemmBiasLanes<BN>& b) {
  constexpr int kChunks = BN / 32;
  const int c_lo = half == 0 ? 0 : (kChunks + 1) / 2, c_hi = half == 0 ? (kChunks + 1) / 2 : kChunks;
#pragma unroll
  for (int i = 0; i < (kChunks + 1) / 2; ++i) b.v[i] = c_lo + i < c_hi ? __ldg(args.bias + n0 + (c_lo + i) * 32 + lane) : 0.f;
}

template <int BN, int EPI>
__device__ __forceinline__ void gemm_epilogue_tile(const GemmArgs& args, uint32_t taddr, int row, bool row_ok, int n0, int nt, int n_tiles,
                                                   int half, const GemmBiasLanes<BN>& bias, uint8_t* stage_w = nullptr) {
  constexpr int kChunks = BN / 32;
  const int c_lo = half == 0 ? 0 : (kChunks + 1) / 2, c_hi = half == 0 ? (kChunks + 1) / 2 : kChunks;
  float best = -INFINITY;
  int best_i = 0x7fffffff;
  int out_row = row;
  const float* extra = nullptr;               // per-row fp32 addend (residual or position row)
  if (EPI == EPI_F32_RESID) extra = args.resid + static_cast<size_t>(row) * args.ldr;
  if (EPI == EPI_PATCH) {
    const int b = row / kPatches, p = row % kPatches;
    out_row = b * kEncTokens + 1 + p;
    extra = args.pos + static_cast<size_t>(1 + p) * kD;
  }
  (void)bias;
  if constexpr ((EPI == EPI_BF16 || EPI == EPI_BF16_GELU) && BN % 128 == 0) {
    if (args.out_tma) {
      // Two 32-column chunks at a time: 64 bf16 = one 128-byte row of the warp's 32 x 128-byte staging tile (16-byte
      // pieces XOR-swizzled with the row: conflict-free st.shared.v4, un-swizzled by the map), then ONE TMA store per
      // warp.  Per-lane row stores cost a warp 32 partial-line transactions per instruction: at 128 x 256 tiles that
      // (plus the GELU arithmetic) made the epilogue slower than the tile's main loop.
      const int lane = static_cast<int>(threadIdx.x & 31u);
      const int row0 = row - lane;
      const uint32_t dst = smem_u32(stage_w) + static_cast<uint32_t>(lane * 128);
#pragma unroll 1
      for (int c = c_lo; c < c_hi; c += 2) {
        uint32_t v0[32], v1[32];
        tmem_ld32(taddr + static_cast<uint32_t>(c * 32), v0);
        tmem_ld32(taddr + static_cast<uint32_t>(c * 32 + 32), v1);
        tmem_ld_wait();
        const int col0 = n0 + c * 32;
        uint32_t pk[32];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(args.bias + col0 + j));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(args.bias + col0 + 32 + j));
          float f0 = __uint_as_float(v0[j]) + b0.x, f1 = __uint_as_float(v0[j + 1]) + b0.y;
          float f2 = __uint_as_float(v0[j + 2]) + b0.z, f3 = __uint_as_float(v0[j + 3]) + b0.w;
          float g0 = __uint_as_float(v1[j]) + b1.x, g1 = __uint_as_float(v1[j + 1]) + b1.y;
          float g2 = __uint_as_float(v1[j + 2]) + b1.z, g3 = __uint_as_float(v1[j + 3]) + b1.w;
          if (EPI == EPI_BF16_GELU) {
            f0 = gelu_erf_fast(f0); f1 = gelu_erf_fast(f1); f2 = gelu_erf_fast(f2); f3 = gelu_erf_fast(f3);
            g0 = gelu_erf_fast(g0); g1 = gelu_erf_fast(g1); g2 = gelu_erf_fast(g2); g3 = gelu_erf_fast(g3);
          }
          pk[j / 2] = pack_bf16(f0, f1);
          pk[j / 2 + 1] = pack_bf16(f2, f3);
          pk[16 + j / 2] = pack_bf16(g0, g1);
          pk[16 + j / 2 + 1] = pack_bf16(g2, g3);
        }
        if (lane == 0) bulk_wait_group_read0();          // the previous pair's tile has been read out
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + static_cast<uint32_t>((j ^ (lane & 7)) << 4)), "r"(pk[4 * j]),
                       "r"(pk[4 * j + 1]), "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3])
                       : "memory");
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && row0 < args.M) {                // rows >= M are clipped by the map
          tma_store_2d(&args.tmap_out, stage_w, col0, row0);
          bulk_commit_group();
        }
      }
      return;
    }
  }
#pragma unroll 1
  for (int c = c_lo; c < c_hi; ++c) {
    uint32_t v[32];
    tmem_ld32(taddr + static_cast<uint32_t>(c * 32), v);
    tmem_ld_wait();
    const int col0 = n0 + c * 32;
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 bb = __ldg(reinterpret_cast<const float4*>(args.bias + col0 + j));
      f[j] = __uint_as_float(v[j]) + bb.x;
      f[j + 1] = __uint_as_float(v[j + 1]) + bb.y;
      f[j + 2] = __uint_as_float(v[j + 2]) + bb.z;
      f[j + 3] = __uint_as_float(v[j + 3]) + bb.w;
    }
    if (EPI == EPI_BF16 || EPI == EPI_BF16_GELU || EPI == EPI_CROSSKV) {
      if (EPI == EPI_BF16_GELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = gelu_erf_fast(f[j]);
      }
      if (row_ok) {
        uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(args.out) + static_cast<size_t>(row) * args.ldo + col0);
        if (EPI == EPI_CROSSKV) {
          // column = (layer * 2 + kv) * 768 + head * 64 + d  ->  cache[crop][layer][kv][head][token][64]: every
          // (crop, layer, head) K or V block is one contiguous 25 KB stream for the decoder's attention
          const int crop = row / kEncTokens, tok = row - crop * kEncTokens;
          const int block = args.crop_map != nullptr ? args.crop_map[crop] : crop;
          const int lkv = col0 / kD, hd = col0 - lkv * kD;
          dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(args.out) +
                                         ((static_cast<size_t>(block) * 4 + lkv) * kHeads + hd / kHeadDim) * (kEncTokens * kHeadDim) +
                                         static_cast<size_t>(tok) * kHeadDim + (hd & (kHeadDim - 1)));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 q;
          q.x = pack_bf16(f[8 * j + 0], f[8 * j + 1]);
          q.y = pack_bf16(f[8 * j + 2], f[8 * j + 3]);
          q.z = pack_bf16(f[8 * j + 4], f[8 * j + 5]);
          q.w = pack_bf16(f[8 * j + 6], f[8 * j + 7]);
          dst[j] = q;
        }
      }
    } else if (EPI == EPI_F32_GELU) {
      if (row_ok) {
        float4* dst = reinterpret_cast<float4*>(static_cast<float*>(args.out) + static_cast<size_t>(row) * args.ldo + col0);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          dst[j] = make_float4(gelu_erf(f[4 * j]), gelu_erf(f[4 * j + 1]), gelu_erf(f[4 * j + 2]), gelu_erf(f[4 * j + 3]));
      }
    } else if (EPI == EPI_F32_ACCUM) {
      // row `lane` of the warp's 32 x 32 f32 tile -> one 128-byte row of the staging tile, 16-byte chunks
      // XOR-swizzled with the row (conflict-free stores; the map un-swizzles), then ONE reduce-add per warp:
      // the residual is never read by the SM and every global access is a full 128-byte line.
      const int lane = static_cast<int>(threadIdx.x & 31u);
      const int row0 = row - lane;
      if (lane == 0) bulk_wait_group_read0();          // the previous chunk's tile has been read out
      __syncwarp();
      const uint32_t dst = smem_u32(stage_w) + static_cast<uint32_t>(lane * 128);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + static_cast<uint32_t>((j ^ (lane & 7)) << 4)), "f"(f[4 * j]),
                     "f"(f[4 * j + 1]), "f"(f[4 * j + 2]), "f"(f[4 * j + 3])
                     : "memory");
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0 && row0 < args.M) {                // rows >= M are clipped by the map
        tma_reduce_add_2d(&args.tmap_out, stage_w, col0, row0);
        bulk_commit_group();
      }
    } else if (EPI == EPI_F32_RESID || EPI == EPI_PATCH) {
      if (row_ok) {
        const float4* ex = reinterpret_cast<const float4*>(extra + col0);
        float4* dst = reinterpret_cast<float4*>(static_cast<float*>(args.out) + static_cast<size_t>(out_row) * args.ldo + col0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 e = ex[j];
          dst[j] = make_float4(f[4 * j] + e.x, f[4 * j + 1] + e.y, f[4 * j + 2] + e.z, f[4 * j + 3] + e.w);
        }
      }
    } else {  // EPI_ARGMAX
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (f[j] > best) { best = f[j]; best_i = col0 + j; }   // strict > keeps the lowest index on ties
      }
      if (args.logits != nullptr && row_ok && args.step[row] < args.tap_steps) {
        float4* dst = reinterpret_cast<float4*>(
            args.logits + (static_cast<size_t>(row) * args.tap_steps + args.step[row]) * args.N + col0);
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
      }
    }
  }
  if (EPI == EPI_ARGMAX && row_ok) {
    args.part_max[(static_cast<size_t>(row) * n_tiles + nt) * 2 + half] = best;
    args.part_idx[(static_cast<size_t>(row) * n_tiles + nt) * 2 + half] = best_i;
  }
}

template <int BN, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ GemmArgs args) {
  using Cfg = GemmCfg<BN>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);   // SWIZZLE_128B tiles need 1024-B alignment
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * Cfg::kStageBytesA;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* full_bar = bars;                    // [kStages]  TMA -> MMA
  uint64_t* empty_bar = bars + kStages;         // [kStages]  MMA -> TMA
  uint64_t* acc_full = bars + 2 * kStages;      // [2]        MMA -> epilogue
  uint64_t* acc_empty = bars + 2 * kStages + 2; // [2]        epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (args.M + kGemmBM - 1) / kGemmBM;
  const int n_tiles = args.N / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int k_blocks = args.K / kGemmBK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], kGemmEpiThreads);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (args.pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (warp == 0) {
    // ------------------------------------------------ TMA producer ----------
    int stage = 0;
    uint32_t phase = 0;
    int pre = 0;        // k-blocks of the first tile whose weight tiles were requested before the dependency wait
    if (args.pdl) {
      pre = k_blocks < kStages ? k_blocks : kStages;
      if (lane == 0) {
        const int n0 = (static_cast<int>(blockIdx.x) % n_tiles) * BN;
        for (int s = 0; s < pre; ++s) {
          mbar_arrive_expect_tx(&full_bar[s], Cfg::kStageBytes);
          tma_load_2d(smem_b + s * Cfg::kStageBytesB, &tmap_b, &full_bar[s], s * kGemmBK, n0);
        }
      }
      __syncwarp();
      asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / n_tiles) * kGemmBM;
      const int n0 = (tile % n_tiles) * BN;
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
#ifdef MOCR_GEMM_DBG
        if (args.dbg == 1 && (tile != static_cast<int>(blockIdx.x) || kb >= kStages)) {
          if (lane == 0) mbar_arrive(&full_bar[stage]);
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
          continue;
        }
#endif
        if (lane == 0) {
          if (tile == static_cast<int>(blockIdx.x) && kb < pre) {     // B is on its way already
            tma_load_2d(smem_a + stage * Cfg::kStageBytesA, &tmap_a, &full_bar[stage], kb * kGemmBK, m0);
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
            tma_load_2d(smem_a + stage * Cfg::kStageBytesA, &tmap_a, &full_bar[stage], kb * kGemmBK, m0);
            tma_load_2d(smem_b + stage * Cfg::kStageBytesB, &tmap_b, &full_bar[stage], kb * kGemmBK, n0);
          }
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer ------------
    constexpr uint32_t idesc = umma_idesc_bf16(kGemmBM, BN);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1u;
      mbar_wait(&acc_empty[as], aphase ^ 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * BN);
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint64_t da = umma_desc_k_sw128(smem_u32(smem_a + stage * Cfg::kStageBytesA));
          const uint64_t db = umma_desc_k_sw128(smem_u32(smem_b + stage * Cfg::kStageBytesB));
#ifdef MOCR_GEMM_DBG
          if (args.dbg != 2)
#endif
#pragma unroll
          for (int k = 0; k < kGemmBK / 16; ++k) {
#ifdef MOCR_GEMM_DBG
            if (args.dbg == 4 && k > 0) break;           // one MMA per k-block
#endif
            // advance 16 bf16 = 32 B along K inside the 128-B swizzle atom: +2 in 16-B units
            umma_bf16(tmem_d, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc,
                      static_cast<uint32_t>((kb | k) != 0));
          }
          umma_commit(&empty_bar[stage]);                 // smem slot free once these MMAs retire
          if (kb == k_blocks - 1) umma_commit(&acc_full[as]);  // accumulator complete
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    // ------------------------------------------------ epilogue --------------
    const int quad = warp & 3;                    // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;             // which half of the tile's 32-column chunks it owns
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1u;
      const int m0 = (tile / n_tiles) * kGemmBM;
      const int nt = tile % n_tiles;
      const int n0 = nt * BN;
      const int row = m0 + quad * 32 + lane;
      const bool row_ok = row < args.M;
      GemmBiasLanes<BN> bias;
      gemm_load_bias<BN>(args, n0, half, lane, bias);
      mbar_wait(&acc_full[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(as * BN);
#ifdef MOCR_GEMM_DBG
      if (args.dbg != 3)
#endif
      gemm_epilogue_tile<BN, EPI>(args, taddr, row, row_ok, n0, nt, n_tiles, half, bias,
                                  smem + kStages * Cfg::kStageBytes + 1024 + (warp - 2) * kGemmAccumStage);
      tc_fence_before();
      mbar_arrive(&acc_empty[as]);
    }
    if ((EPI == EPI_F32_ACCUM || EPI == EPI_BF16 || EPI == EPI_BF16_GELU) && lane == 0) bulk_wait_group0();   // every reduce-add / TMA store of this warp has landed
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace mocr

// ViT self-attention over the 197 encoder tokens of one (crop, head) on the 5th-gen tensor cores:
//   S = Q K^T  (tcgen05.mma, fp32 accumulators in TMEM)  ->  softmax in registers (fp32)
//   O = P V    (tcgen05.mma, P staged in shared memory as the K-major A operand, V as an MN-major B)
// Reference: transformers/models/vit/modeling_vit.py:199-251 (non-causal, no mask); 1/sqrt(64) is
// folded into W_q at load time.
//
// One CTA = one 128-row query tile of a (head, crop): 128 threads, every thread owns one query row (= one TMEM lane).
// Q (128 rows), K and V (208 rows each) of the head are fetched straight from the [n*197, 2304] QKV buffer by three TMA
// loads (128-byte swizzle; rows past the crop's 197 belong to the next crop or are zero-filled - they are masked).
// Thread 0 issues the MMAs.  Two passes over the 208 score columns (max, then exp / sum); P is written as bf16 into the
// swizzled K-major layout UMMA expects, INTO THE SHARED MEMORY OF Q AND K (dead once S is complete), and the
// normalisation by 1 / sum is applied to the 64 output columns in the epilogue.
// Footprint: 92 KB of shared memory and 256 TMEM columns, so TWO CTAs share an SM and the serial chain of one
// (TMA -> S -> softmax -> P V -> epilogue) overlaps the other's; the first version (both query tiles in one 218 KB,
// 512-column CTA, one per SM) ran the same chain unoverlapped: 54 us per layer at 64 crops.
#pragma once
#include "common.cuh"

namespace mocr {

constexpr int kAtcThreads = 128;                    // one warpgroup = one 128-row query tile
constexpr int kAtcKeys = 208;                       // 197 keys padded to a multiple of 16 (UMMA N and K granularity)
constexpr int kAtcQBytes = 128 * 128;               // one 128-row Q tile, 64 bf16 = 128 B per row
constexpr int kAtcKVBytes = kAtcKeys * 128;         // 26 624 B = 26 swizzle atoms of 1 KB
constexpr int kAtcPBytes = 4 * 128 * 128;           // P of the tile: 4 K-atoms (256 keys) x 128 rows x 128 B; aliases Q and K
constexpr int kAtcSmemBytes = kAtcPBytes + kAtcKVBytes + 1024 /*align*/ + 64 /*barriers*/;
static_assert(kAtcQBytes + kAtcKVBytes <= kAtcPBytes, "P reuses the shared memory of Q and K");
constexpr int kAtcTmemCols = 256;                   // S: 208 fp32 columns; O overwrites its first 64 once the softmax has consumed it

// kind::f16 instruction descriptor with an MN-major B operand (bit 16)
__host__ __device__ constexpr uint32_t umma_idesc_bf16_bmn(int m, int n) { return umma_idesc_bf16(m, n) | (1u << 16); }

// grid = (2 query tiles, 12 heads, n crops), block = 128
__global__ void __launch_bounds__(kAtcThreads, 2)
encoder_attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_q /*box 64 x 128*/, const __grid_constant__ CUtensorMap tmap_kv /*box 64 x 208*/,
                            __nv_bfloat16* __restrict__ ctx /*[n*197, 768]*/) {
  extern __shared__ uint8_t atc_raw[];
  const uint32_t raw_addr = smem_u32(atc_raw);
  uint8_t* smem = atc_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sP = smem;                    // [4 atoms][128 rows][128 B], written after S is complete ...
  uint8_t* sQ = smem;                    // ... over Q
  uint8_t* sK = sQ + kAtcQBytes;         // ... and K
  uint8_t* sV = smem + kAtcPBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kAtcKVBytes);
  uint64_t* bar_load = bars;            // TMA -> issuer
  uint64_t* bar_s = bars + 1;           // S tile complete
  uint64_t* bar_o = bars + 2;           // P V complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

  const int t = blockIdx.x, head = blockIdx.y, crop = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int r = tid;                    // row within the tile = TMEM lane
  const int row0 = crop * kEncTokens;

  if (tid == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_kv);
    mbar_init(bar_load, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, kAtcTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (tid == 0) {
    mbar_arrive_expect_tx(bar_load, kAtcQBytes + 2 * kAtcKVBytes);
    tma_load_2d(sQ, &tmap_q, bar_load, head * kHeadDim, row0 + t * 128);
    tma_load_2d(sK, &tmap_kv, bar_load, kD + head * kHeadDim, row0);
    tma_load_2d(sV, &tmap_kv, bar_load, 2 * kD + head * kHeadDim, row0);
    mbar_wait(bar_load, 0);
    tc_fence_after();
    // S = Q K^T : M = 128, N = 208, K = 64 (4 steps of 16), both operands K-major
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, kAtcKeys);
    const uint64_t dk = umma_desc_k_sw128(smem_u32(sK));
    const uint64_t dq = umma_desc_k_sw128(smem_u32(sQ));
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, dq + static_cast<uint64_t>(2 * k), dk + static_cast<uint64_t>(2 * k), idesc_s, k != 0);
    umma_commit(bar_s);
  }

  constexpr float kLog2e = 1.4426950408889634f;
  const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;      // this warp's TMEM lane quadrant
  const int q = t * 128 + r;                           // query row owned by this thread
  const uint32_t ts = tmem_base + lane_sel;
  uint8_t* sPt = sP;
  {
    mbar_wait(bar_s, 0);                               // S is complete: Q and K have been read, their memory is free for P
    tc_fence_after();
    // A warp whose 32 query rows all lie past the 197th token (the last warp of the second tile) has nothing to compute: its P rows
    // may hold anything (an output row depends on its own P row only, and these are never stored).
    const bool warp_live = t * 128 + (warp & 3) * 32 < kEncTokens;
    // ---- pass 1: row maximum over the 197 valid keys (two TMEM loads in flight per wait)
    float mx = -INFINITY;
    if (warp_live) {
#pragma unroll 1
      for (int c = 0; c < 7; c += 2) {
        uint32_t v[32], w2[32];
        tmem_ld32(ts + static_cast<uint32_t>(c * 32), v);
        if (c + 1 < 7) tmem_ld32(ts + static_cast<uint32_t>(c * 32 + 32), w2);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c * 32 + j < kEncTokens) mx = fmaxf(mx, __uint_as_float(v[j]));
        if (c + 1 < 7) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c * 32 + 32 + j < kEncTokens) mx = fmaxf(mx, __uint_as_float(w2[j]));
        }
      }
    }
    // ---- pass 2: p = exp(s - max), row sum, P (bf16) into the swizzled K-major A-operand layout
    const float mxl = mx * kLog2e;
    float sum = 0.f;
    uint8_t* prow = sPt + r * 128;
    // one 32-key chunk: exp, row sum, 4 x 8 bf16 (16 B) into the P tile; key k lives in atom k / 64, chunk (k % 64) / 8, XOR-swizzled
    // with row % 8
    auto emit_chunk = [&](const uint32_t (&v)[32], int c) {
      float p[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float e = exp2f(fmaf(__uint_as_float(v[j]), kLog2e, -mxl));
        p[j] = c * 32 + j < kEncTokens ? e : 0.f;
        sum += p[j];
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int k0 = c * 32 + g * 8;
        if (k0 < kAtcKeys) {
          uint4 w;
          w.x = pack_bf16(p[8 * g], p[8 * g + 1]);
          w.y = pack_bf16(p[8 * g + 2], p[8 * g + 3]);
          w.z = pack_bf16(p[8 * g + 4], p[8 * g + 5]);
          w.w = pack_bf16(p[8 * g + 6], p[8 * g + 7]);
          const int atom = k0 >> 6, chunk = (k0 & 63) >> 3;
          *reinterpret_cast<uint4*>(prow + atom * (128 * 128) + ((chunk ^ (r & 7)) << 4)) = w;
        }
      }
    };
    if (warp_live) {
#pragma unroll 1
      for (int c = 0; c < 7; c += 2) {                 // two TMEM loads in flight per wait
        uint32_t v[32], w2[32];
        tmem_ld32(ts + static_cast<uint32_t>(c * 32), v);
        if (c + 1 < 7) tmem_ld32(ts + static_cast<uint32_t>(c * 32 + 32), w2);
        tmem_ld_wait();
        emit_chunk(v, c);
        if (c + 1 < 7) emit_chunk(w2, c + 1);
      }
    }
    fence_proxy_async_smem();                          // generic-proxy writes of P -> visible to the tensor core
    tc_fence_before();
    __syncthreads();                                   // the P tile is complete, S fully read
    if (r == 0) {
      tc_fence_after();
      // O_t = P V : M = 128, N = 64, K = 208 (13 steps of 16); A = P (K-major), B = V as stored [key][d] = MN-major:
      // 8 key rows of 128 B form one 1 KB swizzle atom (SBO = 1024), a 16-key step advances 2 KB.
      // O goes to the first 64 columns of S (consumed above).
      constexpr uint32_t idesc_o = umma_idesc_bf16_bmn(128, kHeadDim);
#pragma unroll
      for (int s2 = 0; s2 < kAtcKeys / 16; ++s2) {
        const uint64_t dp = umma_desc_k_sw128(smem_u32(sPt + (s2 >> 2) * (128 * 128))) + static_cast<uint64_t>(2 * (s2 & 3));
        const uint64_t dv = umma_desc_k_sw128(smem_u32(sV + s2 * 2048));
        umma_bf16(tmem_base, dp, dv, idesc_o, s2 != 0);
      }
      umma_commit(bar_o);
    }
    mbar_wait(bar_o, 0);
    tc_fence_after();
    // ---- epilogue: O / sum -> bf16 -> ctx[row, head * 64 ..]
    const float inv = warp_live ? __fdividef(1.0f, sum) : 0.f;   // sum >= 1 (the maximum contributes exp(0))
    uint32_t o0[32], o1[32];
    tmem_ld32(ts, o0);
    tmem_ld32(ts + 32, o1);
    tmem_ld_wait();
    if (q < kEncTokens) {
      uint4* dst = reinterpret_cast<uint4*>(ctx + static_cast<size_t>(row0 + q) * kD + head * kHeadDim);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 w;
        w.x = pack_bf16(__uint_as_float(o0[8 * g]) * inv, __uint_as_float(o0[8 * g + 1]) * inv);
        w.y = pack_bf16(__uint_as_float(o0[8 * g + 2]) * inv, __uint_as_float(o0[8 * g + 3]) * inv);
        w.z = pack_bf16(__uint_as_float(o0[8 * g + 4]) * inv, __uint_as_float(o0[8 * g + 5]) * inv);
        w.w = pack_bf16(__uint_as_float(o0[8 * g + 6]) * inv, __uint_as_float(o0[8 * g + 7]) * inv);
        dst[g] = w;
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 w;
        w.x = pack_bf16(__uint_as_float(o1[8 * g]) * inv, __uint_as_float(o1[8 * g + 1]) * inv);
        w.y = pack_bf16(__uint_as_float(o1[8 * g + 2]) * inv, __uint_as_float(o1[8 * g + 3]) * inv);
        w.z = pack_bf16(__uint_as_float(o1[8 * g + 4]) * inv, __uint_as_float(o1[8 * g + 5]) * inv);
        w.w = pack_bf16(__uint_as_float(o1[8 * g + 6]) * inv, __uint_as_float(o1[8 * g + 7]) * inv);
        dst[4 + g] = w;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kAtcTmemCols);
  }
}

}  // namespace mocr

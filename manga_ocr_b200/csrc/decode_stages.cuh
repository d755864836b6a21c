// Greedy decoder: the per-token program of stage kernels (one launch per stage, CUDA graph + programmatic
// dependent launch).
//
// Replaces the per-step host loop of GenerationMixin._sample (transformers/generation/utils.py:
// 2743-2805: ~120 library launches and a host sync per token).  A token step is a short program:
//   BertLayer x2 [QKV -> self-attn -> out+LN -> cross-q -> cross-attn -> out+LN -> FFN1 -> FFN2 -> LN]
//   (modeling_bert.py:143-421), LM head (:471-501), arg-max / EOS / append / embed (utils.py:2793-2805).
//
// Design rules that came out of measurements on the B200 (tools/microbench*.cu, tools/decode_timeline.py):
//  * a token is latency/bandwidth-bound (M = batch rows, 46 MFLOP per row), so the small-M GEMMs use
//    warp-level mma.sync fed straight from L2 with 16-byte loads in a k-permuted fragment order
//    that needs no shared-memory staging;
//  * what costs is the broadcast of the activation to all CTAs (98 KB per CTA = 1.3 us at the L2's
//    ~11 TB/s aggregate) and the ~1 us every kernel boundary adds to the dependent chain; the chain is
//    therefore kept short: the projections that feed a LayerNorm run as 16-CTA clusters that own
//    complete rows (16 x 48 columns), exchange the row statistics through distributed shared memory
//    and write the normalised row themselves (no separate LayerNorm stage);
//  * weights and the encoder K/V do not depend on the previous stage, so they are requested BEFORE
//    the dependency wait (griddepcontrol.wait); attention K/V travel through cp.async into
//    shared-memory staging (no registers held, double-buffered across units).
#pragma once
#include "common.cuh"
#include "rowops.cuh"

namespace mocr {

constexpr int kPdThreads = 256;              // row-stage kernels: 8 warps per CTA at most
constexpr int kPdWarps = kPdThreads / 32;
constexpr int kPdRowsPerBlock = 64;          // activation rows per pass (4 m-tiles of 16)
constexpr int kPdKeySlots = 8;               // 8 * 16 = 128 keys per staged block (a 197-key unit = 2 blocks); each warp owns 32 consecutive keys
constexpr int kPdStageBytes = 2 * kPdKeySlots * 128 * 16;          // K and V of one block, one group
constexpr int kPdMaxNT = 48;
constexpr int kPdMaxStages = 32;
constexpr int kPdVocabTiles = kVocab / kPdMaxNT;
constexpr int kPdMaxPartials = 192;          // arg-max partials per row the next-token stage can merge
constexpr int kPdSplit = 3;                  // CTA-level K split of the N = 768 projections

struct PdLinear {
  const __nv_bfloat16* w;   // [N, K]
  const float* bias;        // [N]
};
struct PdLn {
  const float* g;
  const float* b;
};
struct PdLayer {
  PdLinear qkv, self_out, cross_q, cross_out, fc1, fc2;
  PdLn ln_self, ln_cross, ln_ffn;
  __nv_bfloat16* self_k;    // [B, cache_len, 768]
  __nv_bfloat16* self_v;
};

struct PdParams {
  int B;                    // rows
  int max_len;              // this decode's max_length
  int cache_len;            // self-KV cache capacity per row (tokens)
  int n_partials;           // per-row (max, arg-max) partials of the vocabulary GEMM: kPdVocabTiles, or 2 * tiles of the tcgen05 kernel
  int kv_div;               // decoder rows per crop (1; num_beams in beam mode: the beams of a crop share its cross-attention K/V)
  int logits_cur;           // 1: the logits tap holds the CURRENT step only, [B, 6144] (beam mode)
  int kv_evict_first;       // 1: encoder K/V are streamed through L2 with an evict-first policy (the per-step weights stay resident)
  int fuse_ln;              // 1: the projections that feed a LayerNorm run as 16-CTA clusters that normalise the rows themselves
  int big_accum;             // large-batch program: x += projection in place (TMA reduce-add epilogue), LayerNorm reads x
  int kv_prefetch;          // 1: a layer's encoder K/V are requested into L2 (bulk prefetch) by the layer's first stage
  int big;                  // 1: large-batch program - every Linear on the tcgen05 GEMM (128-row tiles, weights read once for all rows)
  int eos_id;
  PdLayer layer[kDecLayers];
  PdLinear head_t, head_dec;
  PdLn head_ln;
  EmbedWeights emb;
  const __nv_bfloat16* crosskv;   // [crop][layer][K|V][head][197][64]
  // state
  const int* queue_slots;   // session mode (admission into a running decode): ring [n_crops] of crop SLOTS in publication order; the queue
                            // counters then count publications, and publication i decodes into slot queue_slots[i % n_crops]
  int idle_start;           // session mode: every row starts idle (no crop is ready yet)
  int ext_queue;            // 1: the queue is initialised and fed by the host's encoder stream (crops become ready while the decode runs)
  int n_crops;              // crops of this decode (>= B: with more crops than rows, a row that finishes takes the next waiting crop)
  int* ids;                 // [n_crops, max_len]
  int* lens;                // [n_crops] valid ids per crop, written when the crop finishes
  int* slot_crop;           // [B] crop each row is decoding (-1: idle)
  const int* kv_row;        // [B] physical self-attention cache row of each row (beam search: the cache follows the beams), or null: row r uses row r
  int* queue;               // [0] next waiting crop, [1] crops whose encoder K/V are ready, [2] crops finished
  int* pos;                 // [B]
  int* finished;            // [B]
  const int* forced;        // teacher forcing or null
  float* x;                 // [B, 768] post-LN hidden (fp32 residual)
  __nv_bfloat16* xb;        // [B, 768] bf16 copy (GEMM A operand)
  __nv_bfloat16* tb;        // [B, 768] LM-head transform output (A operand of the vocabulary projection)
  __nv_bfloat16* q;         // [B, 768] cross-attention query (large-batch program)
  float* y;                 // [3, B, 768] split-K partials of the projections that feed a LayerNorm
  float* yq;                // [3, B, 768] split-K partials of the cross-attention query
  __nv_bfloat16* qkv;       // [B, 2304]
  __nv_bfloat16* ctx;       // [B, 768]
  __nv_bfloat16* ffn;       // [B, 3072]
  float* part_max;          // [B, kPdVocabTiles]
  int* part_idx;
  float* logits;            // tap or null: [B, max_len-1, 6144]
  long long* prof;          // optional [4096] stage timeline of CTA 0 (debug / tuning), or null
};

enum PdStageType { PD_GEMM16 = 0, PD_GEMM32 = 1, PD_GEMM48 = 2, PD_ATTN_SELF = 3, PD_ATTN_CROSS = 4, PD_LN = 5, PD_NEXT = 6, PD_PROJ_LN = 7, PD_TC = 8 };
// the decoder's Linear layers, for the host (tensor maps of the tcgen05 stages)
enum PdLinId { PD_LIN_QKV = 0, PD_LIN_SELF_OUT = 1, PD_LIN_CROSS_Q = 2, PD_LIN_CROSS_OUT = 3, PD_LIN_FC1 = 4, PD_LIN_FC2 = 5, PD_LIN_PER_LAYER = 6,
               PD_LIN_HEAD_T = 2 * PD_LIN_PER_LAYER, PD_LIN_HEAD_DEC = 2 * PD_LIN_PER_LAYER + 1 };
enum PdEpi { PD_BF16 = 0, PD_BF16_GELU = 1, PD_F32_PARTIAL = 2, PD_ARGMAX = 3 };

// One entry of the per-token program.
struct PdStage {
  int type;
  int epi;                  // GEMM: PdEpi.  LN / PROJ_LN: bit 0 = GELU before the norm (LM-head transform).  TC: GemmEpilogue
  int lin;                  // TC: PdLinId of the weights
  int N, K, ksplit, ldo;
  int parts;                // LN / cross-attention query: number of split-K partials to add
  int layer;                // attention: decoder layer
  const __nv_bfloat16* A;   // GEMM A operand [B, K]; cross-attention: the query rows [B, 768] when they are not split-K partials
  const __nv_bfloat16* W;   // GEMM weights [N, K]
  const float* bias;        // GEMM bias (unsplit) / LN: bias of the producing projection (nullable) / cross: query bias
  __nv_bfloat16* ob;        // GEMM bf16 out / LN bf16 out
  float* of;                // GEMM fp32 partial out / LN fp32 out (nullable)
  const float* src;         // LN: partials [parts][B][768]; cross-attention: partials of the query projection
  const float* resid;       // LN: residual rows (nullable)
  const float* g;           // LN gamma / beta
  const float* b;
  const __nv_bfloat16* pf;  // optional region to request into L2 before the dependency wait (the layer's encoder K/V), or null
  long long pf_bytes;       // bytes per row block of that region
  long long pf_stride;      // bytes between the row blocks
  int pf_blocks;
};

// ------------------------------------------------------------------ primitives ---

__device__ __forceinline__ uint4 ldg_nc16(const void* p) {   // read-only data (weights, cross K/V)
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ldg_cg16(const void* p) {   // data written by other CTAs during this launch: L2 only
  uint4 r;
  asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ float ldg_cg_f32(const float* p) {
  float r;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(r) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ int ldg_cg_s32(const int* p) {
  int r;
  asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ float4 ldg_cg_f4(const float* p) {
  float4 r;
  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void group_sync(int group) {      // 128-thread named barrier (ids 1..)
  asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory");
}
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16_cg(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }

// Dependency of a stage kernel on its predecessor.  With programmatic dependent launch the next stage
// kernel is already resident while this one runs: everything before wait() (weight / encoder-K/V
// prefetch, index arithmetic) overlaps the previous stage, and wait() returns once the previous grid
// has completed and flushed.
__device__ __forceinline__ long long global_timer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
struct StageDep {
  long long* prof;      // optional timeline (option decode_prof): prof[0] = entry count, then (tag, t_entry, t_ready, t_done) records
  int tag;
  long long t_entry, t_ready;
  __device__ __forceinline__ void begin(long long* prof_, int tag_) {
    prof = prof_;
    tag = tag_;
    if (prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0) t_entry = global_timer_ns();
  }
  __device__ __forceinline__ void arrive() {
    if (prof != nullptr && blockIdx.x == 0) {
      __syncthreads();
      if (threadIdx.x == 0) {
        const int i = static_cast<int>(atomicAdd(reinterpret_cast<unsigned long long*>(prof), 1ull));
        if (i < 1000) {
          prof[1 + 4 * i] = tag;
          prof[2 + 4 * i] = t_entry;
          prof[3 + 4 * i] = t_ready;
          prof[4 + 4 * i] = global_timer_ns();
        }
      }
    }
  }
  __device__ __forceinline__ void wait() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0) t_ready = global_timer_ns();
  }
};
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Request a stage's prefetch region (PdStage::pf: the layer's encoder K/V, which never change during a decode) into
// L2: 16 KB bulk prefetches, spread over the grid, one instruction per thread.  Issued BEFORE the dependency wait
// by the first stage of a layer, so that HBM streams the K/V while the layer's GEMM stages (whose weights are L2
// hits) run; the cross-attention stage, four stages later, then reads L2 instead of HBM.
__device__ __forceinline__ void pd_prefetch_region(const void* base, long long block_bytes, long long block_stride, int blocks) {
  if (base == nullptr) return;
  constexpr long long kChunk = 16384;
  const int per_block = static_cast<int>((block_bytes + kChunk - 1) / kChunk);
  const int total = per_block * blocks;
  for (int c = blockIdx.x + gridDim.x * threadIdx.x; c < total; c += gridDim.x * blockDim.x) {
    const int b = c / per_block, k = c - b * per_block;
    const long long off = static_cast<long long>(k) * kChunk;
    const unsigned bytes = static_cast<unsigned>(block_bytes - off < kChunk ? block_bytes - off : kChunk);
    const char* ptr = static_cast<const char*>(base) + b * block_stride + off;
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ptr), "r"(bytes) : "memory");
  }
}

// ------------------------------------------------------------------ small-M GEMM stage ---
//
// out[r, n] = epilogue( sum_k A[r,k] * W[n,k] (+ bias[n]) ),  r < B; tile = (N-slice of NT columns,
// K-split kq of ksplit).
//
// mma.m16n8k16 fragments want, per thread (g = lane/4, t = lane%4), logical k slots {2t,2t+1,2t+8,2t+9}
// of a 16-wide k-step.  The sum over k is order-free, so a 32-wide chunk of physical k is mapped
// onto two k-steps such that thread t owns the 8 CONSECUTIVE physical elements [8t, 8t+8): one
// 16-byte load per row gives a0/a2 (or b0/b1) of both k-steps.  A and W use the same mapping, so
// the product is exact, every 32-byte sector fetched is fully used, and nothing is staged in smem.
// Inside the CTA the warps are 4 m-tiles x KS K-slices, reduced through smem in a fixed order
// (KS = 4: 512 threads keep ~130 KB in flight).
template <int NT, int CH, int KS, class Bar>
__device__ __forceinline__ void pd_gemm_stage(Bar& bar, float* red, const PdParams& p, const PdStage& st) {
  constexpr int NTL = NT / 8;
  constexpr int kThreads = 128 * KS;           // 4 m-tiles x KS K-slices of one warp each
  constexpr int kIters = (kPdRowsPerBlock * NT + kThreads - 1) / kThreads;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int mt = warp & 3, ks = warp >> 2;
  const int B = p.B, K = st.K, N = st.N, ksplit = st.ksplit, epi = st.epi;
  const int kslice = K / (KS * ksplit);
  const int n_chunks = kslice / 32;
  const int n_tiles = (N / NT) * ksplit;
  bool waited = false;

  for (int tile = blockIdx.x; tile < n_tiles || !waited; tile += gridDim.x) {
    const bool has_tile = tile < n_tiles;
    const int n0 = (tile / ksplit) * NT;
    const int kq = tile % ksplit;
    const int k0 = (kq * KS + ks) * kslice;
    // Two K batches are in flight: the fragments of batch n+1 (A and W) are requested before the
    // MMAs of batch n are issued.  Batch 0's weights do not depend on the previous stage and are
    // requested before the grid barrier.
    uint4 wf[2][CH][NTL], af[2][CH][2];
    const __nv_bfloat16* wbase = st.W + static_cast<size_t>(n0 + g) * K + k0 + 8 * t;
    float bias_v[kIters];
#pragma unroll
    for (int it = 0; it < kIters; ++it) bias_v[it] = 0.f;
    if (has_tile) {
#pragma unroll
      for (int c = 0; c < CH; ++c)
#pragma unroll
        for (int j = 0; j < NTL; ++j) wf[0][c][j] = ldg_nc16(wbase + static_cast<size_t>(8 * j) * K + c * 32);
      if (epi != PD_F32_PARTIAL) {       // the bias vector is a constant too (a fresh kernel starts with a cold L1)
#pragma unroll
        for (int it = 0; it < kIters; ++it) {
          const int i = tid + it * kThreads;
          if (i < kPdRowsPerBlock * NT) bias_v[it] = __ldg(st.bias + n0 + i % NT);
        }
      }
    }
    if (!waited) {
      pd_prefetch_region(st.pf, st.pf_bytes, st.pf_stride, st.pf_blocks);
      bar.wait();
      waited = true;
    }
    if (!has_tile) break;
    const int n_batches = n_chunks / CH;       // even by construction (K / 32 / (2 ksplit) / CH)
#pragma unroll 1
    for (int rb = 0; rb < B; rb += kPdRowsPerBlock) {
      const int r_lo = rb + mt * 16 + g, r_hi = r_lo + 8;
      const bool lo_ok = r_lo < B, hi_ok = r_hi < B;
      const __nv_bfloat16* a_lo = st.A + static_cast<size_t>(lo_ok ? r_lo : 0) * K + k0 + 8 * t;
      const __nv_bfloat16* a_hi = st.A + static_cast<size_t>(hi_ok ? r_hi : 0) * K + k0 + 8 * t;
      float acc[NTL][4];
#pragma unroll
      for (int j = 0; j < NTL; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
      if (rb > 0) {
#pragma unroll
        for (int c = 0; c < CH; ++c)
#pragma unroll
          for (int j = 0; j < NTL; ++j) wf[0][c][j] = ldg_nc16(wbase + static_cast<size_t>(8 * j) * K + c * 32);
      }
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        af[0][c][0] = ldg_cg16(a_lo + c * 32);
        af[0][c][1] = ldg_cg16(a_hi + c * 32);
      }
#pragma unroll 1
      for (int bt = 0; bt < n_batches; bt += 2) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int nb = bt + half + 1;          // batch to request now (into the other buffer)
          if (nb < n_batches) {
#pragma unroll
            for (int c = 0; c < CH; ++c) {
              af[half ^ 1][c][0] = ldg_cg16(a_lo + (nb * CH + c) * 32);
              af[half ^ 1][c][1] = ldg_cg16(a_hi + (nb * CH + c) * 32);
#pragma unroll
              for (int j = 0; j < NTL; ++j)
                wf[half ^ 1][c][j] = ldg_nc16(wbase + static_cast<size_t>(8 * j) * K + (nb * CH + c) * 32);
            }
          }
          if (bt + half < n_batches) {
#pragma unroll
            for (int c = 0; c < CH; ++c) {
              uint4 a0 = af[half][c][0], a1 = af[half][c][1];
              if (!lo_ok) a0 = make_uint4(0, 0, 0, 0);
              if (!hi_ok) a1 = make_uint4(0, 0, 0, 0);
#pragma unroll
              for (int j = 0; j < NTL; ++j) {
                mma16816(acc[j], a0.x, a1.x, a0.y, a1.y, wf[half][c][j].x, wf[half][c][j].y);
                mma16816(acc[j], a0.z, a1.z, a0.w, a1.w, wf[half][c][j].z, wf[half][c][j].w);
              }
            }
          }
        }
      }
      // fixed-order K reduction through shared memory
      if (rb > 0) __syncthreads();       // previous row block's readers are done with `red`
      {
        float* rr = red + (ks * kPdRowsPerBlock + mt * 16) * (NT + 1);
#pragma unroll
        for (int j = 0; j < NTL; ++j) {
          rr[g * (NT + 1) + 8 * j + 2 * t] = acc[j][0];
          rr[g * (NT + 1) + 8 * j + 2 * t + 1] = acc[j][1];
          rr[(g + 8) * (NT + 1) + 8 * j + 2 * t] = acc[j][2];
          rr[(g + 8) * (NT + 1) + 8 * j + 2 * t + 1] = acc[j][3];
        }
      }
      __syncthreads();
      float vals[kIters];
#pragma unroll
      for (int it = 0; it < kIters; ++it) {
        const int i = tid + it * kThreads;
        const int rl = i / NT, c = i - rl * NT;
        float v = bias_v[it];
        if (i < kPdRowsPerBlock * NT) {
#pragma unroll
          for (int sl = 0; sl < KS; ++sl) v += red[(sl * kPdRowsPerBlock + rl) * (NT + 1) + c];
        }
        vals[it] = v;
      }
      if (epi == PD_ARGMAX) {
        __syncthreads();                  // all partial sums are read: reuse slice 0 of `red` for the logits
#pragma unroll
        for (int it = 0; it < kIters; ++it) {
          const int i = tid + it * kThreads;
          if (i < kPdRowsPerBlock * NT) red[(i / NT) * (NT + 1) + i % NT] = vals[it];
        }
        __syncthreads();
        if (tid < kPdRowsPerBlock && rb + tid < B) {
          const int r = rb + tid;
          float best = -INFINITY;
          int best_i = 0;
          float* lg = nullptr;
          if (p.logits != nullptr) {
            const int stp = ldg_cg_s32(p.pos + r);
            if (p.logits_cur) lg = p.logits + static_cast<size_t>(r) * N + n0;
            else if (stp < p.max_len - 1) lg = p.logits + (static_cast<size_t>(r) * (p.max_len - 1) + stp) * N + n0;
          }
#pragma unroll 4
          for (int cc = 0; cc < NT; ++cc) {
            const float v = red[tid * (NT + 1) + cc];
            if (v > best) { best = v; best_i = n0 + cc; }     // strict >: lowest index wins ties
            if (lg != nullptr) lg[cc] = v;
          }
          p.part_max[static_cast<size_t>(r) * p.n_partials + tile] = best;
          p.part_idx[static_cast<size_t>(r) * p.n_partials + tile] = best_i;
        }
      } else {
#pragma unroll
        for (int it = 0; it < kIters; ++it) {
          const int i = tid + it * kThreads;
          const int r = rb + i / NT, c = i % NT;
          if (i < kPdRowsPerBlock * NT && r < B) {
            float v = vals[it];
            if (epi == PD_BF16_GELU) v = gelu_erf(v);
            if (epi == PD_F32_PARTIAL) st.of[(static_cast<size_t>(kq) * B + r) * st.ldo + n0 + c] = v;
            else st.ob[static_cast<size_t>(r) * st.ldo + n0 + c] = __float2bfloat16(v);
          }
        }
      }
    }
    __syncthreads();                      // `red` is free for the next tile
  }
  bar.arrive();
}

// ------------------------------------------------------------------ row stages (one warp per row) ---

__device__ __forceinline__ void pd_ln_row_warp(const float (&v)[24], const float* g, const float* b, float* x, __nv_bfloat16* xb, int lane) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 24; ++i) s += v[i];
  const float mean = warp_sum(s) * (1.0f / kD);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 24; ++i) q += (v[i] - mean) * (v[i] - mean);
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / kD) + kLnEps);
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const int c = (lane + 32 * i) * 4;
    const float4 gm = __ldg(reinterpret_cast<const float4*>(g + c));
    const float4 bt = __ldg(reinterpret_cast<const float4*>(b + c));
    float4 yv;
    yv.x = (v[4 * i] - mean) * rstd * gm.x + bt.x;
    yv.y = (v[4 * i + 1] - mean) * rstd * gm.y + bt.y;
    yv.z = (v[4 * i + 2] - mean) * rstd * gm.z + bt.z;
    yv.w = (v[4 * i + 3] - mean) * rstd * gm.w + bt.w;
    if (x != nullptr) *reinterpret_cast<float4*>(x + c) = yv;
    uint2 pk;
    pk.x = pack_bf16(yv.x, yv.y);
    pk.y = pack_bf16(yv.z, yv.w);
    *reinterpret_cast<uint2*>(xb + c) = pk;
  }
}

// x, xb = LayerNorm( [gelu]( sum_parts src[part] + bias ) + resid )   (BERT post-LN, modeling_bert.py:
// 297, 355, 484).  The split-K partials of the producing projection are added here in a fixed order.
template <class Bar>
__device__ __forceinline__ void pd_ln_stage(Bar& bar, const PdParams& p, const PdStage& st, int n_ctas) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int B = p.B;
  const int wpc = blockDim.x >> 5;
  const int r0 = blockIdx.x * wpc + warp;
  // gamma, beta and the producing projection's bias are constants: request them before the
  // dependency wait (a fresh kernel starts with a cold L1, each would cost an L2 round trip after it)
  float4 gm[6], bt[6], bs[6];
  if (r0 < B) {
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const int c = (lane + 32 * i) * 4;
      gm[i] = __ldg(reinterpret_cast<const float4*>(st.g + c));
      bt[i] = __ldg(reinterpret_cast<const float4*>(st.b + c));
      bs[i] = st.bias != nullptr ? __ldg(reinterpret_cast<const float4*>(st.bias + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  bar.wait();
  for (int r = r0; r < B; r += n_ctas * wpc) {
    // every load is requested before the first add (the .cg loads are ordered volatile asm: a load
    // placed after a dependent add would cost a full L2 round trip each)
    float4 part[kPdSplit][6], rs[6];
#pragma unroll
    for (int pt = 0; pt < kPdSplit; ++pt)
#pragma unroll
      for (int i = 0; i < 6; ++i)
        part[pt][i] = pt < st.parts ? ldg_cg_f4(st.src + (static_cast<size_t>(pt) * B + r) * kD + (lane + 32 * i) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (st.resid != nullptr) {
#pragma unroll
      for (int i = 0; i < 6; ++i) rs[i] = ldg_cg_f4(st.resid + static_cast<size_t>(r) * kD + (lane + 32 * i) * 4);
    }
    float v[24];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      float4 f = bs[i];
#pragma unroll
      for (int pt = 0; pt < kPdSplit; ++pt) { f.x += part[pt][i].x; f.y += part[pt][i].y; f.z += part[pt][i].z; f.w += part[pt][i].w; }
      if (st.epi & 1) { f.x = gelu_erf(f.x); f.y = gelu_erf(f.y); f.z = gelu_erf(f.z); f.w = gelu_erf(f.w); }
      if (st.resid != nullptr) { f.x += rs[i].x; f.y += rs[i].y; f.z += rs[i].z; f.w += rs[i].w; }
      v[4 * i] = f.x; v[4 * i + 1] = f.y; v[4 * i + 2] = f.z; v[4 * i + 3] = f.w;
      s += (f.x + f.y) + (f.z + f.w);
    }
    const float mean = warp_sum(s) * (1.0f / kD);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 24; ++i) q += (v[i] - mean) * (v[i] - mean);
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / kD) + kLnEps);
    float* x = st.of != nullptr ? st.of + static_cast<size_t>(r) * kD : nullptr;
    __nv_bfloat16* xb = st.ob + static_cast<size_t>(r) * kD;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const int c = (lane + 32 * i) * 4;
      float4 yv;
      yv.x = (v[4 * i] - mean) * rstd * gm[i].x + bt[i].x;
      yv.y = (v[4 * i + 1] - mean) * rstd * gm[i].y + bt[i].y;
      yv.z = (v[4 * i + 2] - mean) * rstd * gm[i].z + bt[i].z;
      yv.w = (v[4 * i + 3] - mean) * rstd * gm[i].w + bt[i].w;
      if (x != nullptr) *reinterpret_cast<float4*>(x + c) = yv;
      uint2 pk;
      pk.x = pack_bf16(yv.x, yv.y);
      pk.y = pack_bf16(yv.z, yv.w);
      *reinterpret_cast<uint2*>(xb + c) = pk;
    }
  }
  bar.arrive();
}

struct PdEmbedConsts {
  float4 gm[6], bt[6], ty[6];
};
__device__ __forceinline__ void pd_embed_consts(const PdParams& p, int lane, PdEmbedConsts& k) {
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const int c = (lane + 32 * i) * 4;
    k.gm[i] = __ldg(reinterpret_cast<const float4*>(p.emb.gamma + c));
    k.bt[i] = __ldg(reinterpret_cast<const float4*>(p.emb.beta + c));
    k.ty[i] = __ldg(reinterpret_cast<const float4*>(p.emb.type0 + c));
  }
}

// x = LayerNorm(word[tok] + type[0] + pos[position])   (modeling_bert.py:102-111), one warp per row
__device__ __forceinline__ void pd_embed_row_warp(const PdParams& p, const PdEmbedConsts& k, int r, int tok, int position, int lane) {
  float4 a[6], d[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const int c = (lane + 32 * i) * 4;
    a[i] = __ldg(reinterpret_cast<const float4*>(p.emb.word + static_cast<size_t>(tok) * kD + c));
    d[i] = __ldg(reinterpret_cast<const float4*>(p.emb.posemb + static_cast<size_t>(position) * kD + c));
  }
  float v[24];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    v[4 * i] = a[i].x + k.ty[i].x + d[i].x; v[4 * i + 1] = a[i].y + k.ty[i].y + d[i].y;
    v[4 * i + 2] = a[i].z + k.ty[i].z + d[i].z; v[4 * i + 3] = a[i].w + k.ty[i].w + d[i].w;
    s += (v[4 * i] + v[4 * i + 1]) + (v[4 * i + 2] + v[4 * i + 3]);
  }
  const float mean = warp_sum(s) * (1.0f / kD);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 24; ++i) q += (v[i] - mean) * (v[i] - mean);
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / kD) + kLnEps);
  float* x = p.x + static_cast<size_t>(r) * kD;
  __nv_bfloat16* xb = p.xb + static_cast<size_t>(r) * kD;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const int c = (lane + 32 * i) * 4;
    float4 yv;
    yv.x = (v[4 * i] - mean) * rstd * k.gm[i].x + k.bt[i].x;
    yv.y = (v[4 * i + 1] - mean) * rstd * k.gm[i].y + k.bt[i].y;
    yv.z = (v[4 * i + 2] - mean) * rstd * k.gm[i].z + k.bt[i].z;
    yv.w = (v[4 * i + 3] - mean) * rstd * k.gm[i].w + k.bt[i].w;
    *reinterpret_cast<float4*>(x + c) = yv;
    uint2 pk;
    pk.x = pack_bf16(yv.x, yv.y);
    pk.y = pack_bf16(yv.z, yv.w);
    *reinterpret_cast<uint2*>(xb + c) = pk;
  }
}

// Greedy step tail (generation/utils.py:2793-2805, stopping_criteria.py:76,470): final arg-max over
// the vocabulary tiles, EOS / max_length rules, append, embed the next input token.
//
// In-flight slot refill: the reference pads a finished row with [PAD] and keeps stepping it until EVERY row of the batch
// has finished (utils.py:2797-2805).  Rows are independent, so here a row that finishes hands its slot to the next
// waiting crop of the decode (p.queue, crops beyond the first B) and restarts at [CLS] - the ids of every crop are
// exactly those of the padded scheme, but the step count follows the total number of tokens instead of
// (crops / rows) x the longest sequence.
__device__ __forceinline__ int pd_pop_waiting_crop(const PdParams& p) {      // one lane; -1: nothing is waiting
  unsigned int head = static_cast<unsigned int>(ldg_cg_s32(p.queue));
  while (true) {
    int avail;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(avail) : "l"(p.queue + 1) : "memory");
    if (static_cast<int>(head) >= avail) return -1;
    const unsigned int seen = atomicCAS(reinterpret_cast<unsigned int*>(p.queue), head, head + 1u);
    if (seen == head) return p.queue_slots != nullptr ? ldg_cg_s32(p.queue_slots + head % static_cast<unsigned int>(p.n_crops)) : static_cast<int>(head);
    head = seen;
  }
}

template <class Bar>
__device__ __forceinline__ void pd_next_token_stage(Bar& bar, const PdParams& p, int n_ctas) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wpc = blockDim.x >> 5;
  const int r0 = blockIdx.x * wpc + warp;
  PdEmbedConsts k;
  if (r0 < p.B) pd_embed_consts(p, lane, k);         // constants: before the dependency wait
  bar.wait();
  for (int r = r0; r < p.B; r += n_ctas * wpc) {
    // one round trip: state and the 128 per-tile (max, arg-max) partials together
    const int ps = ldg_cg_s32(p.pos + r);
    const int fin0 = ldg_cg_s32(p.finished + r);
    const int crop = ldg_cg_s32(p.slot_crop + r);
    float pv[kPdMaxPartials / 32];
    int pi[kPdMaxPartials / 32];
#pragma unroll
    for (int j = 0; j < kPdMaxPartials / 32; ++j) {
      const int i = lane + 32 * j;
      pv[j] = -INFINITY;
      pi[j] = 0x7fffffff;
      if (i < p.n_partials) {
        pv[j] = ldg_cg_f32(p.part_max + static_cast<size_t>(r) * p.n_partials + i);
        pi[j] = ldg_cg_s32(p.part_idx + static_cast<size_t>(r) * p.n_partials + i);
      }
    }
    const bool was_finished = fin0 != 0;
    if (was_finished && (p.forced == nullptr || ps >= p.max_len - 1)) {
      // an idle row (its crop is done and nothing was waiting then): crops whose encoder K/V arrive later are picked up here
      if (p.forced == nullptr && (p.n_crops > p.B || p.queue_slots != nullptr)) {
        int c = lane == 0 ? pd_pop_waiting_crop(p) : 0;
        c = __shfl_sync(0xffffffffu, c, 0);
        if (c >= 0) {
          if (lane == 0) {
            p.slot_crop[r] = c;
            p.pos[r] = 0;
            p.finished[r] = 0;
          }
          pd_embed_row_warp(p, k, r, 2, 0, lane);
        }
      }
      continue;
    }
    float bv = -INFINITY;
    int bi = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < kPdMaxPartials / 32; ++j)
      if (pv[j] > bv || (pv[j] == bv && pi[j] < bi)) { bv = pv[j]; bi = pi[j]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (bi < 0 || bi >= kVocab) bi = 1;    // all-NaN logits: [UNK] rather than an out-of-range index
    int tok = bi;
    const int np = ps + 1;
    int fin = was_finished ? 1 : 0;
    if (tok == p.eos_id && p.forced == nullptr) fin = 1;
    if (np >= p.max_len - 1) fin = 1;
    int next_crop = -1;
    if (lane == 0) {
      if (np < p.max_len) p.ids[static_cast<size_t>(crop) * p.max_len + np] = tok;
      if (fin && !was_finished) {            // the crop is complete
        p.lens[crop] = np + 1 < p.max_len ? np + 1 : p.max_len;
        atomicAdd(p.queue + 2, 1);
        if (p.forced == nullptr && (p.n_crops > p.B || p.queue_slots != nullptr)) next_crop = pd_pop_waiting_crop(p);
      }
      if (next_crop >= 0) {                  // the slot goes to the next waiting crop
        p.slot_crop[r] = next_crop;
        p.finished[r] = 0;
        p.pos[r] = 0;
      } else {
        p.finished[r] = fin;
        p.pos[r] = np;
      }
    }
    next_crop = __shfl_sync(0xffffffffu, next_crop, 0);
    if (next_crop >= 0) {
      pd_embed_row_warp(p, k, r, 2, 0, lane);        // [CLS] at position 0 (generation/utils.py:806-863)
      continue;
    }
    if (p.forced != nullptr && np < p.max_len) tok = p.forced[static_cast<size_t>(crop) * p.max_len + np];
    if (np >= p.max_len - 1 || np >= kMaxPos) continue;
    pd_embed_row_warp(p, k, r, tok, np, lane);
  }
  bar.arrive();
}

// ------------------------------------------------------------------ projection + residual + LayerNorm, one launch ---
//
// x, xb = LayerNorm( [gelu]( A W^T + bias ) + resid )  for the N = 768 projections that feed a LayerNorm
// (BertSelfOutput / BertOutput: modeling_bert.py:287-298, 343-356; LM-head transform :471-486).
//
// A cluster of 16 CTAs owns 16 complete rows: CTA c computes columns [48 c, 48 c + 48) over the whole K
// (8 or 16 warps = K slices, fixed-order reduction through shared memory), adds bias and residual, and the
// row statistics are combined across the cluster through distributed shared memory (one exchange of
// (sum, M2) per CTA and row, merged with Chan's formula: the two-pass variance of the reference, exactly).
// Every CTA then normalises and writes its own 48 columns.  Compared with split-K partials + a separate
// LayerNorm stage this removes a dependent launch (~1 us) and the partials' round trip through L2 per
// LayerNorm; the CTA's whole weight slab (74 KB at K = 768) is requested into registers BEFORE the
// dependency wait, so that only 16 rows of the activation (24 KB) are on the critical path.
constexpr int kPlCluster = 16;
constexpr int kPlCols = kD / kPlCluster;     // 48
constexpr int kPlRows = 16;                  // one m16 tile per cluster
constexpr int kPlNTL = kPlCols / 8;          // 6 n-tiles
constexpr int pd_proj_ln_smem_bytes(int warps) { return warps * kPlRows * (kPlCols + 1) * 4 + kPlCluster * kPlRows * 8; }

__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void st_cluster_f2(uint32_t cluster_addr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(cluster_addr), "f"(a), "f"(b) : "memory");
}

// WARPS K-slices per CTA; CH 32-wide chunks per load batch; the K slice of a warp is NB batches: K = WARPS * NB * CH * 32.
template <int WARPS, int CH, int NB>
__global__ void __cluster_dims__(kPlCluster, 1, 1) __launch_bounds__(WARPS * 32, 1)
pd_proj_ln_kernel(const __grid_constant__ PdParams p, const __grid_constant__ PdStage st) {
  extern __shared__ __align__(16) float pl_smem[];
  float* red = pl_smem;                                                             // [WARPS][16][49]
  float2* stats = reinterpret_cast<float2*>(pl_smem + WARPS * kPlRows * (kPlCols + 1));   // [16 ranks][16 rows] (sum, M2)
  StageDep bar;
  bar.begin(p.prof, st.type * 100 + (st.K > 1000 ? 1 : 0));
#ifdef MOCR_PL_STAMPS
#define PL_STAMP(i) do { if (p.prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0) p.prof[3000 + (i)] = global_timer_ns(); } while (0)
#else
#define PL_STAMP(i) do { } while (0)
#endif
  PL_STAMP(0);
  pdl_launch_dependents();
  cluster_arrive_relaxed();                    // phase 0: every CTA of the cluster is running (its shared memory exists)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const uint32_t rank = cluster_ctarank();
  const int rg = blockIdx.x / kPlCluster;      // row group
  const int n0 = static_cast<int>(rank) * kPlCols;
  constexpr int K = WARPS * NB * CH * 32;
  constexpr int kslice = K / WARPS;            // 96 (K = 768, 8 warps) or 192 (K = 3072, 16 warps)
  const int B = p.B;
  const int k0 = warp * kslice;
  const int r_lo = rg * kPlRows + g, r_hi = r_lo + 8;
  const bool lo_ok = r_lo < B, hi_ok = r_hi < B;

  // ---- before the dependency wait: the first batch of weight fragments, bias / gamma / beta of this CTA's columns
  const __nv_bfloat16* wbase = st.W + static_cast<size_t>(n0 + g) * K + k0 + 8 * t;
  uint4 wf[2][CH][kPlNTL], af[2][CH][2];
#pragma unroll
  for (int c = 0; c < CH; ++c)
#pragma unroll
    for (int j = 0; j < kPlNTL; ++j) wf[0][c][j] = ldg_nc16(wbase + static_cast<size_t>(8 * j) * K + c * 32);
  // epilogue mapping: warp w owns rows {w, w + WARPS, ..} < 16 of the tile, lane l columns l and l + 32 (< 48)
  const int c0 = lane, c1 = lane + 32;
  const bool c1_ok = c1 < kPlCols;
  const float bias0 = __ldg(st.bias + n0 + c0), bias1 = c1_ok ? __ldg(st.bias + n0 + c1) : 0.f;
  const float gm0 = __ldg(st.g + n0 + c0), gm1 = c1_ok ? __ldg(st.g + n0 + c1) : 0.f;
  const float bt0 = __ldg(st.b + n0 + c0), bt1 = c1_ok ? __ldg(st.b + n0 + c1) : 0.f;
  pd_prefetch_region(st.pf, st.pf_bytes, st.pf_stride, st.pf_blocks);
  PL_STAMP(1);
  bar.wait();
  PL_STAMP(2);

  // ---- the activation rows of this cluster and the residual elements this thread will need
  const __nv_bfloat16* a_lo = st.A + static_cast<size_t>(lo_ok ? r_lo : 0) * K + k0 + 8 * t;
  const __nv_bfloat16* a_hi = st.A + static_cast<size_t>(hi_ok ? r_hi : 0) * K + k0 + 8 * t;
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    af[0][c][0] = ldg_cg16(a_lo + c * 32);
    af[0][c][1] = ldg_cg16(a_hi + c * 32);
  }
  constexpr int kRowsPerWarp = (kPlRows + WARPS - 1) / WARPS;     // 2 (8 warps) or 1 (16 warps)
  float rs0[kRowsPerWarp], rs1[kRowsPerWarp];
#pragma unroll
  for (int i = 0; i < kRowsPerWarp; ++i) {
    const int rl = warp + i * WARPS, r = rg * kPlRows + rl;
    rs0[i] = rs1[i] = 0.f;
    if (st.resid != nullptr && rl < kPlRows && r < B) {
      rs0[i] = ldg_cg_f32(st.resid + static_cast<size_t>(r) * kD + n0 + c0);
      if (c1_ok) rs1[i] = ldg_cg_f32(st.resid + static_cast<size_t>(r) * kD + n0 + c1);
    }
  }
  float acc[kPlNTL][4];
#pragma unroll
  for (int j = 0; j < kPlNTL; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
  for (int bt = 0; bt < NB; ++bt) {
    const int cur = bt & 1, nxt = cur ^ 1;
    if (bt + 1 < NB) {                         // the next batch is requested before the MMAs of this one are issued
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        af[nxt][c][0] = ldg_cg16(a_lo + ((bt + 1) * CH + c) * 32);
        af[nxt][c][1] = ldg_cg16(a_hi + ((bt + 1) * CH + c) * 32);
#pragma unroll
        for (int j = 0; j < kPlNTL; ++j) wf[nxt][c][j] = ldg_nc16(wbase + static_cast<size_t>(8 * j) * K + ((bt + 1) * CH + c) * 32);
      }
    }
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      uint4 a0 = af[cur][c][0], a1 = af[cur][c][1];
      if (!lo_ok) a0 = make_uint4(0, 0, 0, 0);
      if (!hi_ok) a1 = make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int j = 0; j < kPlNTL; ++j) {
        mma16816(acc[j], a0.x, a1.x, a0.y, a1.y, wf[cur][c][j].x, wf[cur][c][j].y);
        mma16816(acc[j], a0.z, a1.z, a0.w, a1.w, wf[cur][c][j].z, wf[cur][c][j].w);
      }
    }
  }
  // ---- fixed-order K reduction through shared memory
  PL_STAMP(3);
  {
    float* rr = red + warp * kPlRows * (kPlCols + 1);
#pragma unroll
    for (int j = 0; j < kPlNTL; ++j) {
      rr[g * (kPlCols + 1) + 8 * j + 2 * t] = acc[j][0];
      rr[g * (kPlCols + 1) + 8 * j + 2 * t + 1] = acc[j][1];
      rr[(g + 8) * (kPlCols + 1) + 8 * j + 2 * t] = acc[j][2];
      rr[(g + 8) * (kPlCols + 1) + 8 * j + 2 * t + 1] = acc[j][3];
    }
  }
  __syncthreads();
  float v0[kRowsPerWarp], v1[kRowsPerWarp], sum_c[kRowsPerWarp], m2_c[kRowsPerWarp];
#pragma unroll
  for (int i = 0; i < kRowsPerWarp; ++i) {
    const int rl = warp + i * WARPS;
    float a = bias0, b = bias1;
    if (rl < kPlRows) {
#pragma unroll
      for (int sl = 0; sl < WARPS; ++sl) {
        a += red[(sl * kPlRows + rl) * (kPlCols + 1) + c0];
        if (c1_ok) b += red[(sl * kPlRows + rl) * (kPlCols + 1) + c1];
      }
    }
    if (st.epi & 1) { a = gelu_erf(a); b = gelu_erf(b); }
    a += rs0[i];
    b += rs1[i];
    if (!c1_ok) b = 0.f;
    v0[i] = a;
    v1[i] = b;
    // statistics of this CTA's 48 columns of the row: sum and M2 about their own mean
    const float s = warp_sum(a + b);
    const float mc = s * (1.0f / kPlCols);
    const float d0 = a - mc, d1 = c1_ok ? b - mc : 0.f;
    sum_c[i] = s;
    m2_c[i] = warp_sum(d0 * d0 + d1 * d1);
  }
  PL_STAMP(4);
  cluster_wait_acquire();                      // phase 0 complete: every peer's shared memory may be written
  PL_STAMP(5);
#pragma unroll
  for (int i = 0; i < kRowsPerWarp; ++i) {
    const int rl = warp + i * WARPS;
    if (rl < kPlRows && lane < kPlCluster)     // lane j hands this CTA's (sum, M2) of the row to CTA j
      st_cluster_f2(mapa_u32(&stats[rank * kPlRows + rl], static_cast<uint32_t>(lane)), sum_c[i], m2_c[i]);
  }
  cluster_arrive_release();                    // phase 1: the statistics of all 16 column slices have been delivered
  cluster_wait_acquire();
  PL_STAMP(6);
#pragma unroll
  for (int i = 0; i < kRowsPerWarp; ++i) {
    const int rl = warp + i * WARPS, r = rg * kPlRows + rl;
    if (rl >= kPlRows) continue;
    const float2 sj = stats[(lane & (kPlCluster - 1)) * kPlRows + rl];   // lanes j and j + 16 read slice j
    float tot = sj.x;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    const float mean = tot * (1.0f / kD);
    const float dm = sj.x * (1.0f / kPlCols) - mean;
    float m2 = sj.y + kPlCols * dm * dm;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) m2 += __shfl_xor_sync(0xffffffffu, m2, o);
    const float rstd = rsqrtf(m2 * (1.0f / kD) + kLnEps);
    if (r < B) {
      const float y0 = (v0[i] - mean) * rstd * gm0 + bt0;
      if (st.of != nullptr) st.of[static_cast<size_t>(r) * kD + n0 + c0] = y0;
      st.ob[static_cast<size_t>(r) * kD + n0 + c0] = __float2bfloat16(y0);
      if (c1_ok) {
        const float y1 = (v1[i] - mean) * rstd * gm1 + bt1;
        if (st.of != nullptr) st.of[static_cast<size_t>(r) * kD + n0 + c1] = y1;
        st.ob[static_cast<size_t>(r) * kD + n0 + c1] = __float2bfloat16(y1);
      }
    }
  }
  PL_STAMP(7);
  bar.arrive();
}

// ------------------------------------------------------------------ attention stage ---
// Single-query attention for (row, head) units, 4 warps per unit.  Key j of a 208-key block is
// owned by (warp gw, lane quarter sub, slot i): j = 16 i + 4 gw + sub; the 8 lanes of a quarter
// hold 8 channels each.  Every thread requests ALL its K and V rows of a block with cp.async into
// 16-byte staging slots that only it reads back (no barrier, no registers held while the loads
// fly), keeps a private online-softmax state (m, l, acc[8]); the 16 partial states of a unit are
// merged at the end in a fixed order.  The staging area is double-buffered: the next unit's rows
// are requested before the current unit is reduced, and for cross-attention the first unit is
// requested BEFORE the grid barrier - the encoder K/V never change during a decode, so the
// dominant HBM stream of the token overlaps the barrier.
// self: the key at index pos[b] is this token's K/V row, taken straight from the QKV buffer and
// appended to the cache here (modeling_bert.py:190 re-concatenates the whole cache instead).

__device__ __forceinline__ void pd_bf16x8(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(p[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

struct PdAttnUnit {
  const __nv_bfloat16* kc;     // this thread's 16-byte column of the unit's key rows
  const __nv_bfloat16* vc;
  const __nv_bfloat16* nk;     // self: this token's K / V row in the QKV buffer
  const __nv_bfloat16* nv;
  int b, h, n_keys, ps;
  int crop;                    // cross: the crop whose encoder K/V the row reads
  bool skip;
};

__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void cp_async16_cg_hint(void* smem_dst, const void* gsrc, uint64_t pol) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async16_zero(void* smem_dst, const void* any_valid_gsrc) {   // src-size 0: writes 16 zero bytes
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, 0;" ::"r"(smem_u32(smem_dst)), "l"(any_valid_gsrc) : "memory");
}
// Stage one 128-key block: warp w of the group fetches - and later consumes, so a __syncwarp is all the
// synchronisation the stream needs - keys [32w, 32w+32) of the block.  A key row is 128 bytes (64 dims);
// its 16-byte chunks are XOR-swizzled with the key index so that ldmatrix (8 rows x 16 bytes) is
// conflict-free.  V rows past the last key are zero-filled (their probabilities are 0, but 0 x garbage
// could be NaN inside the MMA).  self: the row at index pos is this token's K/V, taken from the QKV buffer
// and appended to the cache here (modeling_bert.py:190 re-concatenates the whole cache instead).
// KEYS = key rows of the stage's K region (V follows it): 128 in the four-warp kernel, 32 in the warp-per-unit kernel (gt = lane).
template <bool SELF, bool HINT, int KEYS = 16 * kPdKeySlots>
__device__ __forceinline__ void pd_attn_request_t(uint4* stage, const PdAttnUnit& a, int key_stride, int j0, int gt, uint64_t pol) {
  const int gw = gt >> 5, sub = (gt & 31) >> 3, ch = gt & 7;
  const int kl0 = 32 * gw + sub;                        // this lane's first key of the block; its others follow every 4 keys
  const int left = a.n_keys - (j0 + 32 * gw);           // keys of this warp's range that exist (warp-uniform)
  if (left <= 0) return;
  // (kl & 7) = (sub + 4 i) & 7 = sub + 4 (i & 1): two swizzled chunk positions per lane, fixed strides between the slots
  uint4* kd0 = stage + kl0 * 8;
  const int sw0 = ch ^ sub, sw1 = ch ^ (sub + 4);
  const __nv_bfloat16* ks0 = a.kc + static_cast<size_t>(j0 + kl0) * key_stride;
  const __nv_bfloat16* vs0 = a.vc + static_cast<size_t>(j0 + kl0) * key_stride;
  const int mine = left - sub;                          // slot i exists iff 4 i < mine
  int fresh_i = -1;
#pragma unroll
  for (int i = 0; i < kPdKeySlots; ++i) {
    uint4* kd = kd0 + i * 32 + ((i & 1) ? sw1 : sw0);
    uint4* vd = kd + KEYS * 8;
    if (4 * i < mine) {
      const __nv_bfloat16* ks = ks0 + static_cast<size_t>(4 * i) * key_stride;
      const __nv_bfloat16* vs = vs0 + static_cast<size_t>(4 * i) * key_stride;
      if (SELF && j0 + kl0 + 4 * i == a.ps) {           // this token's row comes from the QKV buffer, not from the cache
        ks = a.nk;
        vs = a.nv;
        fresh_i = i;
      }
      if (HINT) {
        cp_async16_cg_hint(kd, ks, pol);
        cp_async16_cg_hint(vd, vs, pol);
      } else {
        cp_async16_cg(kd, ks);
        cp_async16_cg(vd, vs);
      }
    } else {
      cp_async16_zero(vd, a.vc);
    }
  }
  if (SELF && fresh_i >= 0) {    // append this token's row to the cache (8 lanes, 16 bytes each, per K and V) once the stream is issued
    const uint4 kq = ldg_cg16(a.nk), vq = ldg_cg16(a.nv);
    const size_t fresh_j = static_cast<size_t>(j0 + kl0 + 4 * fresh_i);
    *reinterpret_cast<uint4*>(const_cast<__nv_bfloat16*>(a.kc) + fresh_j * key_stride) = kq;
    *reinterpret_cast<uint4*>(const_cast<__nv_bfloat16*>(a.vc) + fresh_j * key_stride) = vq;
  }
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}

template <bool SELF, class Bar>
__device__ __forceinline__ void pd_attention_stage(Bar& bar, uint8_t* smem, const PdParams& p, const PdStage& st) {
  constexpr int kBlockKeys = 16 * kPdKeySlots;
  const int tid = threadIdx.x;
  const int n_groups = blockDim.x >> 7;       // 128-thread groups per CTA (1 in the stage kernel)
  const int group = tid >> 7, gt = tid & 127;
  const int gw = gt >> 5, lane = gt & 31;
  const int ch = lane & 7;
  uint4* stage0 = reinterpret_cast<uint4*>(smem + (group * 2) * kPdStageBytes);
  uint4* stage1 = reinterpret_cast<uint4*>(smem + (group * 2 + 1) * kPdStageBytes);
  const PdLayer& L = p.layer[st.layer];
  // self: cache rows [B][t][768], a head is a 64-column slice; cross: one contiguous [197][64] block per (crop, layer, K|V, head)
  const __nv_bfloat16* kbase = SELF ? L.self_k : p.crosskv + static_cast<size_t>(st.layer * 2) * kHeads * kEncTokens * kHeadDim;
  const __nv_bfloat16* vbase = SELF ? L.self_v : p.crosskv + static_cast<size_t>(st.layer * 2 + 1) * kHeads * kEncTokens * kHeadDim;
  const long long b_stride = SELF ? static_cast<long long>(p.cache_len) * kD : static_cast<long long>(kEncTokens) * 4 * kD;
  const int key_stride = SELF ? kD : kHeadDim;
  const int head_stride = SELF ? kHeadDim : kEncTokens * kHeadDim;
  const int units = p.B * kHeads;
  const int ustride = gridDim.x * n_groups;
  const int u0 = blockIdx.x * n_groups + group;

  // Row state (finished flag, position) of the first kPre units of this group is requested in one go
  // right after the dependency wait; together with the first unit's query they cost ONE L2 round trip.
  constexpr int kPre = 4;
  int pre_fin[kPre] = {0, 0, 0, 0}, pre_pos[kPre] = {0, 0, 0, 0}, pre_kvr[kPre] = {0, 0, 0, 0};
  auto make_unit = [&](int u, int fin, int pos, bool state_visible, int kvr) {     // pos: self = the row's position; cross = the crop the row decodes
    PdAttnUnit a;
    a.b = u / kHeads;
    a.h = u - a.b * kHeads;
    a.crop = SELF ? 0 : pos;
    // self: the row's cache row (kvr: its own, or the one the beam-search row table gives); cross: the beams of a crop read the same K/V
    const int kvb = SELF ? kvr : (pos < 0 ? 0 : pos) / p.kv_div;
    a.kc = kbase + static_cast<size_t>(kvb) * b_stride + static_cast<size_t>(a.h) * head_stride + ch * 8;
    a.vc = vbase + static_cast<size_t>(kvb) * b_stride + static_cast<size_t>(a.h) * head_stride + ch * 8;
    a.nk = a.nv = nullptr;
    a.n_keys = kEncTokens;
    a.ps = -1;
    a.skip = false;
    if (state_visible) {
      a.skip = fin != 0 && p.forced == nullptr;     // group-uniform
      if (SELF) {
        a.ps = pos;
        a.n_keys = a.ps + 1;
        a.nk = p.qkv + static_cast<size_t>(a.b) * 3 * kD + kD + a.h * kHeadDim + ch * 8;
        a.nv = a.nk + kD;
      }
      if (a.skip) a.n_keys = 0;                     // a finished row contributes no blocks
    }
    return a;
  };
  auto unit_state = [&](int u, int k) {             // k = index of u in this group's unit sequence
    int fin, pos, kvr;
    if (k < kPre) {
      fin = pre_fin[0];
      pos = pre_pos[0];
      kvr = pre_kvr[0];
#pragma unroll
      for (int i = 1; i < kPre; ++i) {              // (select chain: the arrays stay in registers)
        if (k == i) {
          fin = pre_fin[i];
          pos = pre_pos[i];
          kvr = pre_kvr[i];
        }
      }
    } else {
      fin = ldg_cg_s32(p.finished + u / kHeads);
      pos = ldg_cg_s32((SELF ? p.pos : p.slot_crop) + u / kHeads);
      kvr = (SELF && p.kv_row != nullptr) ? ldg_cg_s32(p.kv_row + u / kHeads) : u / kHeads;
    }
    return make_unit(u, fin, pos, true, kvr);
  };
  // The query of a unit, requested (raw) one unit ahead and summed when the unit starts:
  // self: bf16 row of the QKV buffer; cross: bias + split-K partials of the cross-q projection.
  struct QRaw {
    uint4 qb;
    float4 e0[kPdSplit], e1[kPdSplit];
  };
  auto issue_q = [&](int u, QRaw& raw) {
    const int b = u / kHeads, h = u - b * kHeads;
    if (SELF) {
      raw.qb = ldg_cg16(p.qkv + static_cast<size_t>(b) * 3 * kD + h * kHeadDim + ch * 8);
    } else if (st.A != nullptr) {                   // complete bf16 query rows (large-batch program)
      raw.qb = ldg_cg16(st.A + static_cast<size_t>(b) * kD + h * kHeadDim + ch * 8);
    } else {
      const int c = h * kHeadDim + ch * 8;
#pragma unroll
      for (int pt = 0; pt < kPdSplit; ++pt) {
        const float* src = st.src + (static_cast<size_t>(pt) * p.B + b) * kD + c;
        raw.e0[pt] = ldg_cg_f4(src);
        raw.e1[pt] = ldg_cg_f4(src + 4);
      }
    }
  };
  auto finish_q = [&](int u, const QRaw& raw, float (&qq)[8]) {
    if (SELF || st.A != nullptr) {
      pd_bf16x8(raw.qb, qq);
    } else {
      const int c = (u % kHeads) * kHeadDim + ch * 8;
      float4 lo = __ldg(reinterpret_cast<const float4*>(st.bias + c)), hi = __ldg(reinterpret_cast<const float4*>(st.bias + c + 4));
#pragma unroll
      for (int pt = 0; pt < kPdSplit; ++pt) {       // fixed order
        lo.x += raw.e0[pt].x; lo.y += raw.e0[pt].y; lo.z += raw.e0[pt].z; lo.w += raw.e0[pt].w;
        hi.x += raw.e1[pt].x; hi.y += raw.e1[pt].y; hi.z += raw.e1[pt].z; hi.w += raw.e1[pt].w;
      }
      qq[0] = lo.x; qq[1] = lo.y; qq[2] = lo.z; qq[3] = lo.w;
      qq[4] = hi.x; qq[5] = hi.y; qq[6] = hi.z; qq[7] = hi.w;
    }
  };

  // The group's work is one stream of 112-key blocks (unit after unit) flowing through a two-deep
  // ring of staging buffers: block n+1 is requested before block n is reduced.
  const bool stream_kv = SELF ? (p.kv_evict_first & 2) != 0 : (p.kv_evict_first & 1) != 0;
  const uint64_t pol = stream_kv ? l2_policy_evict_first() : 0ull;
  auto request = [&](uint4* stage, const PdAttnUnit& a, int j0) {
    if (stream_kv) pd_attn_request_t<SELF, true>(stage, a, key_stride, j0, gt, pol);
    else pd_attn_request_t<SELF, false>(stage, a, key_stride, j0, gt, 0ull);
  };
  // cross: which crop the first unit's row decodes is read speculatively before the dependency wait (it changes only when
  // the row finishes and takes the next waiting crop) and checked again after it
  const int spec_crop = (!SELF && u0 < units) ? ldg_cg_s32(p.slot_crop + u0 / kHeads) : 0;
  PdAttnUnit cur = make_unit(u0 < units ? u0 : 0, 0, SELF ? 0 : spec_crop, false, 0);
  bool pre2 = false;       // both blocks of the first unit were requested before the dependency wait
  if (!SELF) {   // encoder K/V never change during a decode: request the first unit (2 blocks) before the wait
    if (u0 < units && spec_crop >= 0) {
      request(stage0, cur, 0);
      cp_async_commit();
      request(stage1, cur, kBlockKeys);
      pre2 = true;
    }
    cp_async_commit();
  }
  bar.wait();
  QRaw raw_cur, raw_nxt;
#pragma unroll
  for (int k = 0; k < kPre; ++k) {
    const int uk = u0 + k * ustride;
    pre_fin[k] = 0;
    pre_pos[k] = 0;
    if (uk < units) {
      pre_fin[k] = ldg_cg_s32(p.finished + uk / kHeads);
      pre_pos[k] = ldg_cg_s32((SELF ? p.pos : p.slot_crop) + uk / kHeads);
      pre_kvr[k] = (SELF && p.kv_row != nullptr) ? ldg_cg_s32(p.kv_row + uk / kHeads) : uk / kHeads;
    }
  }
  if (u0 < units) issue_q(u0, raw_cur);
  int u = u0, uk_idx = 0;
  if (u < units) {
    cur = unit_state(u, uk_idx);
    while (cur.n_keys == 0 && u + ustride < units) {   // skip finished rows
      u += ustride;
      ++uk_idx;
      cur = unit_state(u, uk_idx);
      issue_q(u, raw_cur);
    }
    if (SELF || u != u0 || !pre2 || cur.crop != spec_crop) {
      if (!SELF) cp_async_wait_group<0>();          // (the prefetched blocks of a finished row, or of the crop the row has left, are dropped)
      pre2 = false;
      if (cur.n_keys > 0) request(stage0, cur, 0);
      cp_async_commit();
    }
  }
  int par = 0, j0 = 0;
  // Per-warp online-softmax state.  The attention of ONE query is a GEMV; it runs on mma.sync.m16n8k16 with the
  // query in row 0 of the A operand (rows 1-15 are zero): 12x fewer instructions than the per-thread FMA/shuffle
  // form, which had made the stage issue-bound.  Row 0 lives in lanes 0-3 (t = lane): score / output columns
  // 2t, 2t+1 of every 8-wide tile.
  float q[8], m = -INFINITY, l = 0.f;
  float o[8][4];
  uint32_t qa[4][2];                     // A fragments (a0, a2) of the four 16-dim k-steps; a1 = a3 = 0
#pragma unroll 1
  while (u < units && cur.n_keys > 0) {
    uint4* stage = par ? stage1 : stage0;
    // ---- locate and request the next block of the stream
    PdAttnUnit nxt = cur;
    int nu = u, nj0 = j0 + kBlockKeys, nk_idx = uk_idx;
    if (nj0 >= cur.n_keys) {
      nj0 = 0;
      nu = u + ustride;
      nk_idx = uk_idx + 1;
      bool found = false;
      while (nu < units) {
        nxt = unit_state(nu, nk_idx);
        if (nxt.n_keys > 0) { found = true; break; }
        nu += ustride;
        ++nk_idx;
      }
      if (!found) nu = units;
    }
    if (pre2) pre2 = false;                       // (u0, block 1) is already on its way
    else if (nu < units) request(par ? stage0 : stage1, nxt, nj0);
    cp_async_commit();
    // ---- first block of a unit: the query (requested one unit ahead) and a fresh online-softmax state
    if (j0 == 0) {
      finish_q(u, raw_cur, q);
      m = -INFINITY;
      l = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
      // lane c (< 8) holds the 8-dim chunk c of q; the row-0 quad needs, per 16-dim k-step ks,
      // a0 = q[16ks + 2t, +1] (chunk 2ks, pair t) and a2 = q[16ks + 8 + 2t, +1] (chunk 2ks+1, pair t)
      uint32_t pk[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) pk[e] = pack_bf16(q[2 * e], q[2 * e + 1]);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint32_t f = 0u;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint32_t v = __shfl_sync(0xffffffffu, pk[e], c);
          if (lane == e) f = v;
        }
        qa[c >> 1][c & 1] = f;                          // lanes >= 4 keep 0: rows 1-15 of the A operand
      }
    }
    if (nu != u && nu < units) issue_q(nu, raw_nxt);      // the next unit's query travels while this block is reduced
    cp_async_wait_group<1>();             // everything but the newest group (the next block) has landed
    __syncwarp();                         // ... for every lane of this warp (a warp consumes only what it staged)
    {
      const int wk0 = 32 * gw;
      const int valid = cur.n_keys - (j0 + wk0);       // keys of this warp's range that exist (warp-uniform)
      if (valid > 0) {
        const uint32_t kaddr = smem_u32(stage), vaddr = kaddr + kPdKeySlots * 128 * 16;
        const int t2 = 2 * (lane & 3);
        // ---- scores of up to 32 keys: S[0, key] = q . K[key]   (4 key tiles x 4 k-steps)
        float sc[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
          if (8 * nt < valid) {
            const int kr = wk0 + 8 * nt + (lane & 7);          // the key row whose address this lane supplies
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {                   // dims 32 hf .. 32 hf + 31
              uint32_t b[4];
              ldmatrix_x4(b, kaddr + static_cast<uint32_t>(kr * 128 + (((4 * hf + (lane >> 3)) ^ (kr & 7)) << 4)));
              mma16816(sc[nt], qa[2 * hf][0], 0u, qa[2 * hf][1], 0u, b[0], b[1]);
              mma16816(sc[nt], qa[2 * hf + 1][0], 0u, qa[2 * hf + 1][1], 0u, b[2], b[3]);
            }
          }
        }
        float bm = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            if (8 * nt + t2 + e >= valid) sc[nt][e] = -INFINITY;
            bm = fmaxf(bm, sc[nt][e]);
          }
        }
        bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 1));
        bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 2));   // key 0 of the range exists: finite on the row-0 quad
        // ---- one rescale of the running state per block, then P = exp(S - m) as the A operand of O += P V
        const float mn = fmaxf(m, bm);
        const float cs = __expf(m - mn);                       // m = -inf on the first block: 0
        l *= cs;
#pragma unroll
        for (int i = 0; i < 8; ++i) { o[i][0] *= cs; o[i][1] *= cs; }
        m = mn;
        uint32_t pa[2][2];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const float p0 = __expf(sc[2 * ks][0] - mn), p1 = __expf(sc[2 * ks][1] - mn);
          const float p2 = __expf(sc[2 * ks + 1][0] - mn), p3 = __expf(sc[2 * ks + 1][1] - mn);
          l += (p0 + p1) + (p2 + p3);
          pa[ks][0] = lane < 4 ? pack_bf16(p0, p1) : 0u;       // rows 1-15 of P are unused: keep them exactly zero
          pa[ks][1] = lane < 4 ? pack_bf16(p2, p3) : 0u;
        }
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          if (16 * ks < valid) {
            const int kr = wk0 + 16 * ks + ((lane >> 3) & 1) * 8 + (lane & 7);
#pragma unroll
            for (int dp = 0; dp < 4; ++dp) {                   // dims 16 dp .. 16 dp + 15 (two output tiles)
              uint32_t b[4];
              ldmatrix_x4_trans(b, vaddr + static_cast<uint32_t>(kr * 128 + (((2 * dp + (lane >> 4)) ^ (kr & 7)) << 4)));
              mma16816(o[2 * dp], pa[ks][0], 0u, pa[ks][1], 0u, b[0], b[1]);
              mma16816(o[2 * dp + 1], pa[ks][0], 0u, pa[ks][1], 0u, b[2], b[3]);
            }
          }
        }
      }
    }
    // ---- last block of a unit: merge the 4 warps' states in a fixed order and write the context
    if (j0 + kBlockKeys >= cur.n_keys) {
      l += __shfl_xor_sync(0xffffffffu, l, 1);
      l += __shfl_xor_sync(0xffffffffu, l, 2);
      // (each warp parks its state in its own, fully consumed, 4 KB of the current block's K rows)
      float* s_part = reinterpret_cast<float*>(stage);
      constexpr int kPartStride = 32 * 128 / 4;          // floats between two warps' regions
      __syncwarp();
      if (lane < 4) {
        float* dst = s_part + gw * kPartStride;
        if (lane == 0) { dst[0] = m; dst[1] = l; }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          dst[2 + 8 * i + 2 * lane] = o[i][0];
          dst[2 + 8 * i + 2 * lane + 1] = o[i][1];
        }
      }
      group_sync(group);
      if (gt < kHeadDim) {
        float mx = -INFINITY;
#pragma unroll
        for (int w = 0; w < 4; ++w) mx = fmaxf(mx, s_part[w * kPartStride]);
        float lt = 0.f, ot = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const float mw = s_part[w * kPartStride];
          const float cw = mw == -INFINITY ? 0.f : __expf(mw - mx);
          lt += s_part[w * kPartStride + 1] * cw;
          ot += s_part[w * kPartStride + 2 + gt] * cw;
        }
        p.ctx[static_cast<size_t>(cur.b) * kD + cur.h * kHeadDim + gt] = __float2bfloat16(__fdividef(ot, lt));   // lt >= 1
      }
      group_sync(group);     // s_part is reused by the next unit
    }
    if (nu != u) raw_cur = raw_nxt;
    cur = nxt;
    u = nu;
    uk_idx = nk_idx;
    j0 = nj0;
    par ^= 1;
  }
  cp_async_wait_group<0>();   // nothing may still be landing in the staging area: the next stage reuses it
  bar.arrive();
}

// ------------------------------------------------------------------ the kernel ---


__host__ __device__ inline PdStage pd_gemm_desc(int type, int epi, const __nv_bfloat16* A, int K, const PdLinear& lin, int N, int ksplit,
                                                __nv_bfloat16* ob, float* of, int ldo) {
  PdStage s{};
  s.type = type; s.epi = epi; s.A = A; s.K = K; s.W = lin.w; s.bias = lin.bias; s.N = N; s.ksplit = ksplit; s.ob = ob; s.of = of; s.ldo = ldo;
  return s;
}
__host__ __device__ inline PdStage pd_ln_desc(const float* src, int parts, const float* bias, int gelu, const float* resid, const PdLn& ln,
                                              float* x, __nv_bfloat16* xb) {
  PdStage s{};
  s.type = PD_LN; s.epi = gelu; s.src = src; s.parts = parts; s.bias = bias; s.resid = resid; s.g = ln.g; s.b = ln.b; s.of = x; s.ob = xb;
  return s;
}
// x, xb = LayerNorm([gelu](A W^T + bias) + resid) in one launch (pd_proj_ln_kernel)
__host__ __device__ inline PdStage pd_proj_ln_desc(const __nv_bfloat16* A, int K, const PdLinear& lin, int gelu, const float* resid, const PdLn& ln,
                                                   float* x, __nv_bfloat16* xb) {
  PdStage s{};
  s.type = PD_PROJ_LN; s.epi = gelu; s.A = A; s.K = K; s.N = kD; s.W = lin.w; s.bias = lin.bias; s.resid = resid; s.g = ln.g; s.b = ln.b; s.of = x; s.ob = xb;
  return s;
}
// a Linear on the tcgen05 GEMM (large-batch program); epi is a GemmEpilogue value
__host__ __device__ inline PdStage pd_tc_desc(int lin, int epi, const __nv_bfloat16* A, int K, int N, __nv_bfloat16* ob, float* of, int ldo,
                                              const float* resid) {
  PdStage s{};
  s.type = PD_TC; s.lin = lin; s.epi = epi; s.A = A; s.K = K; s.N = N; s.ob = ob; s.of = of; s.ldo = ldo; s.resid = resid;
  return s;
}

// The per-token program; the host launches it stage by stage (captured into a CUDA graph).
//   small batches (p.big = 0): warp-level mma.sync stages; the N = 768 projections in front of a LayerNorm either as
//     16-CTA clusters that normalise the rows themselves (p.fuse_ln) or as split-K partials + a LayerNorm stage;
//   large batches (p.big = 1): every Linear on the tcgen05 kernel with 128-row tiles, LayerNorm as a row stage.
__host__ __device__ inline int pd_build_program(const PdParams& p, PdStage* prog) {
  int n = 0;
  const long long kv_layer_elems = 2ll * kHeads * kEncTokens * kHeadDim;      // K and V of one layer of one crop
  for (int l = 0; l < kDecLayers; ++l) {
    const PdLayer& L = p.layer[l];
    const int lin0 = l * PD_LIN_PER_LAYER;
    PdStage a{};
    // self-attention block (modeling_bert.py:143-207, 287-298)
    if (p.big) prog[n++] = pd_tc_desc(lin0 + PD_LIN_QKV, 0 /*EPI_BF16*/, p.xb, kD, 3 * kD, p.qkv, nullptr, 3 * kD, nullptr);
    else prog[n++] = pd_gemm_desc(PD_GEMM16, PD_BF16, p.xb, kD, L.qkv, 3 * kD, 1, p.qkv, nullptr, 3 * kD);
    if (p.kv_prefetch && !p.big) {           // this layer's encoder K/V start streaming into L2 now
      PdStage& q = prog[n - 1];
      q.pf = p.crosskv + static_cast<size_t>(l) * kv_layer_elems;
      q.pf_bytes = kv_layer_elems * 2;
      q.pf_stride = 2 * kDecLayers * kHeads * kEncTokens * kHeadDim * 2ll;
      q.pf_blocks = (p.B + p.kv_div - 1) / p.kv_div;
    }
    a = PdStage{}; a.type = PD_ATTN_SELF; a.layer = l; prog[n++] = a;
    if (p.big) {
      if (p.big_accum) {
        prog[n++] = pd_tc_desc(lin0 + PD_LIN_SELF_OUT, 7 /*EPI_F32_ACCUM*/, p.ctx, kD, kD, nullptr, p.x, kD, nullptr);
        prog[n++] = pd_ln_desc(p.x, 1, nullptr, 0, nullptr, L.ln_self, p.x, p.xb);
      } else {
        prog[n++] = pd_tc_desc(lin0 + PD_LIN_SELF_OUT, 2 /*EPI_F32_RESID*/, p.ctx, kD, kD, nullptr, p.y, kD, p.x);
        prog[n++] = pd_ln_desc(p.y, 1, nullptr, 0, nullptr, L.ln_self, p.x, p.xb);
      }
    } else if (p.fuse_ln) {
      prog[n++] = pd_proj_ln_desc(p.ctx, kD, L.self_out, 0, p.x, L.ln_self, p.x, p.xb);
    } else {
      prog[n++] = pd_gemm_desc(PD_GEMM16, PD_F32_PARTIAL, p.ctx, kD, L.self_out, kD, kPdSplit, nullptr, p.y, kD);
      prog[n++] = pd_ln_desc(p.y, kPdSplit, L.self_out.bias, 0, p.x, L.ln_self, p.x, p.xb);
    }
    // cross-attention block (:210-284)
    a = PdStage{}; a.type = PD_ATTN_CROSS; a.layer = l;
    if (p.big) {
      prog[n++] = pd_tc_desc(lin0 + PD_LIN_CROSS_Q, 0 /*EPI_BF16*/, p.xb, kD, kD, p.q, nullptr, kD, nullptr);
      a.A = p.q;
    } else {
      prog[n++] = pd_gemm_desc(PD_GEMM16, PD_F32_PARTIAL, p.xb, kD, L.cross_q, kD, kPdSplit, nullptr, p.yq, kD);
      a.src = p.yq; a.parts = kPdSplit; a.bias = L.cross_q.bias;
    }
    prog[n++] = a;
    if (p.big) {
      if (p.big_accum) {
        prog[n++] = pd_tc_desc(lin0 + PD_LIN_CROSS_OUT, 7 /*EPI_F32_ACCUM*/, p.ctx, kD, kD, nullptr, p.x, kD, nullptr);
        prog[n++] = pd_ln_desc(p.x, 1, nullptr, 0, nullptr, L.ln_cross, p.x, p.xb);
      } else {
        prog[n++] = pd_tc_desc(lin0 + PD_LIN_CROSS_OUT, 2 /*EPI_F32_RESID*/, p.ctx, kD, kD, nullptr, p.y, kD, p.x);
        prog[n++] = pd_ln_desc(p.y, 1, nullptr, 0, nullptr, L.ln_cross, p.x, p.xb);
      }
    } else if (p.fuse_ln) {
      prog[n++] = pd_proj_ln_desc(p.ctx, kD, L.cross_out, 0, p.x, L.ln_cross, p.x, p.xb);
    } else {
      prog[n++] = pd_gemm_desc(PD_GEMM16, PD_F32_PARTIAL, p.ctx, kD, L.cross_out, kD, kPdSplit, nullptr, p.y, kD);
      prog[n++] = pd_ln_desc(p.y, kPdSplit, L.cross_out.bias, 0, p.x, L.ln_cross, p.x, p.xb);
    }
    // feed-forward (:330-356)
    if (p.big) {
      prog[n++] = pd_tc_desc(lin0 + PD_LIN_FC1, 1 /*EPI_BF16_GELU*/, p.xb, kD, kFFN, p.ffn, nullptr, kFFN, nullptr);
      if (p.big_accum) {
        prog[n++] = pd_tc_desc(lin0 + PD_LIN_FC2, 7 /*EPI_F32_ACCUM*/, p.ffn, kFFN, kD, nullptr, p.x, kD, nullptr);
        prog[n++] = pd_ln_desc(p.x, 1, nullptr, 0, nullptr, L.ln_ffn, p.x, p.xb);
      } else {
        prog[n++] = pd_tc_desc(lin0 + PD_LIN_FC2, 2 /*EPI_F32_RESID*/, p.ffn, kFFN, kD, nullptr, p.y, kD, p.x);
        prog[n++] = pd_ln_desc(p.y, 1, nullptr, 0, nullptr, L.ln_ffn, p.x, p.xb);
      }
    } else {
      prog[n++] = pd_gemm_desc(PD_GEMM32, PD_BF16_GELU, p.xb, kD, L.fc1, kFFN, 1, p.ffn, nullptr, kFFN);
      prog[n++] = pd_gemm_desc(PD_GEMM16, PD_F32_PARTIAL, p.ffn, kFFN, L.fc2, kD, kPdSplit, nullptr, p.y, kD);
      prog[n++] = pd_ln_desc(p.y, kPdSplit, L.fc2.bias, 0, p.x, L.ln_ffn, p.x, p.xb);
    }
  }
  // LM head (:471-501): dense -> GELU -> LayerNorm, then the vocabulary projection fused with the
  // per-tile arg-max: logits never leave the SM unless the parity tap is on
  if (p.big) {
    prog[n++] = pd_tc_desc(PD_LIN_HEAD_T, 5 /*EPI_F32_GELU*/, p.xb, kD, kD, nullptr, p.y, kD, nullptr);
    prog[n++] = pd_ln_desc(p.y, 1, nullptr, 0, nullptr, p.head_ln, nullptr, p.tb);
  } else if (p.fuse_ln) {
    prog[n++] = pd_proj_ln_desc(p.xb, kD, p.head_t, 1, nullptr, p.head_ln, nullptr, p.tb);
  } else {
    prog[n++] = pd_gemm_desc(PD_GEMM16, PD_F32_PARTIAL, p.xb, kD, p.head_t, kD, kPdSplit, nullptr, p.y, kD);
    prog[n++] = pd_ln_desc(p.y, kPdSplit, p.head_t.bias, 1, nullptr, p.head_ln, nullptr, p.tb);
  }
  prog[n++] = pd_gemm_desc(PD_GEMM48, PD_ARGMAX, p.tb, kD, p.head_dec, kVocab, 1, nullptr, nullptr, 0);
  PdStage nx{};
  nx.type = PD_NEXT;
  prog[n++] = nx;
  return n;
}

// ------------------------------------------------------------------ stage kernels ---
// One launch per stage; grid = number of tiles / units so that every CTA has exactly one piece of
// work and several CTAs share an SM.

constexpr int kPdStageKS = 4;                  // K-slices (warps per m-tile) of the stage-kernel GEMMs: 512 threads
constexpr int pd_gemm_smem_bytes(int nt) { return kPdStageKS * kPdRowsPerBlock * (nt + 1) * 4; }
template <int NT, int CH>
__global__ void __launch_bounds__(128 * kPdStageKS, 1) pd_gemm_kernel(const __grid_constant__ PdParams p, const __grid_constant__ PdStage st) {
  extern __shared__ __align__(16) float red[];      // kPdStageKS * 64 * (NT + 1) floats (pd_gemm_smem_bytes)
  StageDep bar;
  bar.begin(p.prof, st.type * 100 + st.ksplit * 10 + (st.K > 1000 ? 1 : 0));
  pdl_launch_dependents();
  pd_gemm_stage<NT, CH, kPdStageKS>(bar, red, p, st);
}
template <bool SELF>
__global__ void __launch_bounds__(128) pd_attention_kernel(const __grid_constant__ PdParams p, const __grid_constant__ PdStage st) {
  extern __shared__ __align__(128) uint8_t pd_smem[];
  StageDep bar;
  bar.begin(p.prof, st.type * 100);
  pdl_launch_dependents();
  pd_attention_stage<SELF>(bar, pd_smem, p, st);
}
// ---- large-batch attention: ONE WARP per (row, head) unit ------------------------------------------------------
// With hundreds of rows there are thousands of units per stage, and the four-warp form above pays per unit what it
// was built to shorten per unit: two named barriers, a shared-memory merge of four partial states and a 128-key
// block granularity that a short self-attention context leaves three quarters empty.  Here a warp owns a unit from
// its query to its context row: 32-key chunks (K and V, 8 KB) flow through the warp's own two-deep cp.async ring,
// the online-softmax state never leaves registers, the only synchronisation is __syncwarp, and the first chunk of
// the warp's next unit is requested while the last chunk of this one is reduced.  Same arithmetic as the four-warp
// form (m16n8k16 with the query in row 0), one sequential softmax chain per unit: results do not depend on the
// batch size or on the grid.  (modeling_bert.py:143-207 self, :210-284 cross.)
constexpr int kPdRowsChunkBytes = 2 * 32 * 128;                     // K and V rows of 32 keys
constexpr int kPdAttnRowsSmemBytes = 4 * 2 * kPdRowsChunkBytes;     // 4 warps x 2 chunks = 64 KB (three CTAs per SM)

template <bool SELF>
__global__ void __launch_bounds__(128) pd_attention_rows_kernel(const __grid_constant__ PdParams p, const __grid_constant__ PdStage st) {
  extern __shared__ __align__(128) uint8_t pd_smem[];
  StageDep bar;
  bar.begin(p.prof, st.type * 100 + 1);
  pdl_launch_dependents();
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, ch = lane & 7;
  uint4* const ring = reinterpret_cast<uint4*>(pd_smem + w * 2 * kPdRowsChunkBytes);
  const uint32_t ring_addr = smem_u32(ring);
  const PdLayer& L = p.layer[st.layer];
  const __nv_bfloat16* kbase = SELF ? L.self_k : p.crosskv + static_cast<size_t>(st.layer * 2) * kHeads * kEncTokens * kHeadDim;
  const __nv_bfloat16* vbase = SELF ? L.self_v : p.crosskv + static_cast<size_t>(st.layer * 2 + 1) * kHeads * kEncTokens * kHeadDim;
  const long long b_stride = SELF ? static_cast<long long>(p.cache_len) * kD : static_cast<long long>(kEncTokens) * 4 * kD;
  const int key_stride = SELF ? kD : kHeadDim;
  const int head_stride = SELF ? kHeadDim : kEncTokens * kHeadDim;
  const int units = p.B * kHeads;
  const int ustride = gridDim.x * 4;
  const int u0 = blockIdx.x * 4 + w;
  const bool stream_kv = SELF ? (p.kv_evict_first & 2) != 0 : (p.kv_evict_first & 1) != 0;
  const uint64_t pol = stream_kv ? l2_policy_evict_first() : 0ull;
  const __nv_bfloat16* qsrc = SELF ? p.qkv : st.A;                  // complete bf16 query rows
  const int q_ld = SELF ? 3 * kD : kD;

  auto make_unit = [&](int u, int fin, int pos, int kvr) {          // pos: self = the row's position; cross = the crop the row decodes
    PdAttnUnit a;
    a.b = u / kHeads;
    a.h = u - a.b * kHeads;
    a.crop = SELF ? 0 : pos;
    const int kvb = SELF ? kvr : (pos < 0 ? 0 : pos) / p.kv_div;
    a.kc = kbase + static_cast<size_t>(kvb) * b_stride + static_cast<size_t>(a.h) * head_stride + ch * 8;
    a.vc = vbase + static_cast<size_t>(kvb) * b_stride + static_cast<size_t>(a.h) * head_stride + ch * 8;
    a.nk = a.nv = nullptr;
    a.ps = -1;
    a.n_keys = kEncTokens;
    if (SELF) {
      a.ps = pos;
      a.n_keys = pos + 1;
      a.nk = p.qkv + static_cast<size_t>(a.b) * 3 * kD + kD + a.h * kHeadDim + ch * 8;
      a.nv = a.nk + kD;
    }
    a.skip = fin != 0 && p.forced == nullptr;
    if (a.skip) a.n_keys = 0;
    return a;
  };
  auto request = [&](int par, const PdAttnUnit& a, int j0) {
    uint4* stage = ring + par * (kPdRowsChunkBytes / 16);
    if (stream_kv) pd_attn_request_t<SELF, true, 32>(stage, a, key_stride, j0, lane, pol);
    else pd_attn_request_t<SELF, false, 32>(stage, a, key_stride, j0, lane, 0ull);
  };
  // The query as A fragments, read straight into the row-0 quad (lane t < 4): per 16-dim k-step ks,
  // a0 = q[16ks + 2t, +1] and a2 = q[16ks + 8 + 2t, +1]; lanes >= 4 keep 0 (rows 1-15 of the A operand).
  struct QFrag { uint32_t a[4][2]; };
  auto issue_q = [&](int u, QFrag& f) {
    const int b = u / kHeads, h = u - b * kHeads;
    const int* src = reinterpret_cast<const int*>(qsrc + static_cast<size_t>(b) * q_ld + h * kHeadDim) + lane;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      f.a[ks][0] = lane < 4 ? static_cast<uint32_t>(ldg_cg_s32(src + 8 * ks)) : 0u;
      f.a[ks][1] = lane < 4 ? static_cast<uint32_t>(ldg_cg_s32(src + 8 * ks + 4)) : 0u;
    }
  };

  bar.wait();
  // lane k holds the row state of the warp's k-th unit: one round trip covers (up to) 32 units
  int s_fin = 0, s_pos = 0, s_kvr = 0;
  auto load_states = [&](int k0) {
    const int uk = u0 + (k0 + lane) * ustride;
    s_fin = 1; s_pos = 0; s_kvr = 0;
    if (uk < units) {
      const int b = uk / kHeads;
      s_fin = ldg_cg_s32(p.finished + b);
      s_pos = ldg_cg_s32((SELF ? p.pos : p.slot_crop) + b);
      s_kvr = (SELF && p.kv_row != nullptr) ? ldg_cg_s32(p.kv_row + b) : b;
    }
  };
  load_states(0);
  QFrag qa{}, q_nxt{};
  if (u0 < units) issue_q(u0, qa);
  bool have0 = false;        // chunk 0 of the current unit is already on its way (requested during the previous unit)
  int par = 0, k = 0;
  const int t2 = 2 * (lane & 3);
#pragma unroll 1
  for (int u = u0; u < units; u += ustride, ++k) {
    if (k != 0 && (k & 31) == 0) load_states(k);
    const PdAttnUnit cur = make_unit(u, __shfl_sync(0xffffffffu, s_fin, k & 31), __shfl_sync(0xffffffffu, s_pos, k & 31),
                                     __shfl_sync(0xffffffffu, s_kvr, k & 31));
    const int nu = u + ustride;
    PdAttnUnit nxt = cur;
    bool nxt_live = false;
    if (nu < units) {
      issue_q(nu, q_nxt);
      if (((k + 1) & 31) != 0) {                  // (its state is in this batch of 32; otherwise it is requested when it starts)
        nxt = make_unit(nu, __shfl_sync(0xffffffffu, s_fin, (k + 1) & 31), __shfl_sync(0xffffffffu, s_pos, (k + 1) & 31),
                        __shfl_sync(0xffffffffu, s_kvr, (k + 1) & 31));
        nxt_live = nxt.n_keys > 0;
      }
    }
    if (cur.n_keys == 0) {                        // a finished row: nothing to read, its context row is not used
      qa = q_nxt;
      continue;
    }
    if (!have0) {
      request(par, cur, 0);
      cp_async_commit();
    }
    float m = -INFINITY, l = 0.f;
    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
#pragma unroll 1
    for (int j0 = 0; j0 < cur.n_keys; j0 += 32) {
      const bool last = j0 + 32 >= cur.n_keys;
      if (!last) request(par ^ 1, cur, j0 + 32);
      else if (nxt_live) request(par ^ 1, nxt, 0);
      cp_async_commit();
      cp_async_wait_group<1>();
      __syncwarp();
      const int valid = cur.n_keys - j0;
      const uint32_t kaddr = ring_addr + par * kPdRowsChunkBytes, vaddr = kaddr + 32 * 128;
      float sc[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
        if (8 * nt < valid) {
          const int kr = 8 * nt + (lane & 7);
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            uint32_t b[4];
            ldmatrix_x4(b, kaddr + static_cast<uint32_t>(kr * 128 + (((4 * hf + (lane >> 3)) ^ (lane & 7)) << 4)));
            mma16816(sc[nt], qa.a[2 * hf][0], 0u, qa.a[2 * hf][1], 0u, b[0], b[1]);
            mma16816(sc[nt], qa.a[2 * hf + 1][0], 0u, qa.a[2 * hf + 1][1], 0u, b[2], b[3]);
          }
        }
      }
      float bm = -INFINITY;
      if (valid < 32) {                           // the tail chunk: keys that do not exist score -inf
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          if (8 * nt + t2 >= valid) sc[nt][0] = -INFINITY;
          if (8 * nt + t2 + 1 >= valid) sc[nt][1] = -INFINITY;
        }
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) bm = fmaxf(bm, fmaxf(sc[nt][0], sc[nt][1]));
      bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 1));
      bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 2));
      const float mn = fmaxf(m, bm);
      const float cs = __expf(m - mn);
      l *= cs;
#pragma unroll
      for (int i = 0; i < 8; ++i) { o[i][0] *= cs; o[i][1] *= cs; }
      m = mn;
      uint32_t pa[2][2];
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const float p0 = __expf(sc[2 * ks][0] - mn), p1 = __expf(sc[2 * ks][1] - mn);
        const float p2 = __expf(sc[2 * ks + 1][0] - mn), p3 = __expf(sc[2 * ks + 1][1] - mn);
        l += (p0 + p1) + (p2 + p3);
        pa[ks][0] = lane < 4 ? pack_bf16(p0, p1) : 0u;
        pa[ks][1] = lane < 4 ? pack_bf16(p2, p3) : 0u;
      }
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        if (16 * ks < valid) {
          const int kr = 16 * ks + ((lane >> 3) & 1) * 8 + (lane & 7);
#pragma unroll
          for (int dp = 0; dp < 4; ++dp) {
            uint32_t b[4];
            ldmatrix_x4_trans(b, vaddr + static_cast<uint32_t>(kr * 128 + (((2 * dp + (lane >> 4)) ^ (lane & 7)) << 4)));
            mma16816(o[2 * dp], pa[ks][0], 0u, pa[ks][1], 0u, b[0], b[1]);
            mma16816(o[2 * dp + 1], pa[ks][0], 0u, pa[ks][1], 0u, b[2], b[3]);
          }
        }
      }
      par ^= 1;
    }
    // the context row of this (row, head): the row-0 quad holds dims 8 i + 2 t, + 1
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    if (lane < 4) {
      const float inv = __fdividef(1.f, l);       // l >= 1
      uint32_t* dst = reinterpret_cast<uint32_t*>(p.ctx + static_cast<size_t>(cur.b) * kD + cur.h * kHeadDim + t2);
#pragma unroll
      for (int i = 0; i < 8; ++i) dst[4 * i] = pack_bf16(o[i][0] * inv, o[i][1] * inv);
    }
    have0 = nxt_live;
    qa = q_nxt;
  }
  cp_async_wait_group<0>();
  bar.arrive();
}
__global__ void __launch_bounds__(kPdThreads) pd_ln_kernel(const __grid_constant__ PdParams p, const __grid_constant__ PdStage st) {
  StageDep bar;
  bar.begin(p.prof, st.type * 100);
  pdl_launch_dependents();
  pd_ln_stage(bar, p, st, gridDim.x);
}
__global__ void __launch_bounds__(kPdThreads) pd_next_kernel(const __grid_constant__ PdParams p) {
  StageDep bar;
  bar.begin(p.prof, PD_NEXT * 100);
  pdl_launch_dependents();
  pd_next_token_stage(bar, p, gridDim.x);
}
// Every crop: ids[c][0] = [CLS], [PAD] elsewhere.  Rows 0..B-1 start on crops 0..B-1 (pos = 0, x = embed([CLS], 0)); the
// crops beyond wait in the queue.
__global__ void __launch_bounds__(kPdThreads) pd_begin_kernel(const __grid_constant__ PdParams p) {
  pdl_launch_dependents();
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (blockIdx.x == 0 && threadIdx.x == 0 && !p.ext_queue) {
    p.queue[0] = p.B;
    p.queue[1] = p.n_crops;
    p.queue[2] = 0;
  }
  for (int c = blockIdx.x * kPdWarps + warp; c < p.n_crops; c += gridDim.x * kPdWarps) {
    for (int i = lane; i < p.max_len; i += 32) p.ids[static_cast<size_t>(c) * p.max_len + i] = i == 0 ? 2 : 0;
    if (lane == 0) p.lens[c] = p.max_len <= 1 ? 1 : 0;
    if (c >= p.B) continue;
    if (p.idle_start) {          // session mode: the row waits for a published crop
      if (lane == 0) {
        p.slot_crop[c] = 0;
        p.pos[c] = 0;
        p.finished[c] = 1;
      }
      continue;
    }
    if (lane == 0) {
      p.slot_crop[c] = c;
      p.pos[c] = 0;
      p.finished[c] = p.max_len <= 1 ? 1 : 0;
    }
    PdEmbedConsts ek;
    pd_embed_consts(p, lane, ek);
    pd_embed_row_warp(p, ek, c, 2, 0, lane);
  }
}

// Encoder stream -> decoder: the encoder K/V of crops [0, ready) are complete (the launch is ordered after the kernels that
// wrote them); rows that finish, or idle rows, may now take those crops.
__global__ void pd_publish_kernel(int* queue, int head, int ready) {
  if (head >= 0) {           // first publication of a decode: reset the queue
    queue[0] = head;
    queue[2] = 0;
    __threadfence();
    queue[1] = ready;
  } else {
    __threadfence();
    atomicMax(queue + 1, ready);
  }
}

// Session mode (admission into a running decode): crops whose encoder K/V have just been written into the given SLOTS become
// available to idle rows.  The slots' id rows and lengths are reset here (a slot is reused once the host has fetched its previous
// result), then the slot numbers enter the ring in publication order and the "ready" counter moves.  One CTA of 256 threads.
struct PdSlotList {
  int n;
  int slot[64];
};
// The slot list as a device table (the cross-K/V epilogue of the admission's encoder pass scatters crop c into block slot[c]).
__global__ void pd_slot_map_kernel(int* map, const PdSlotList l) {
  if (threadIdx.x < l.n) map[threadIdx.x] = l.slot[threadIdx.x];
}
__global__ void __launch_bounds__(256) pd_publish_slots_kernel(int* queue, int* ring, int cap, int first_pub, int* ids, int* lens, int max_len,
                                                               const PdSlotList l) {
  for (int i = threadIdx.x; i < l.n * max_len; i += blockDim.x) {
    const int k = i / max_len, c = i - k * max_len;
    ids[static_cast<size_t>(l.slot[k]) * max_len + c] = c == 0 ? 2 : 0;
  }
  if (threadIdx.x < l.n) lens[l.slot[threadIdx.x]] = 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 0; i < l.n; ++i) ring[(first_pub + i) % cap] = l.slot[i];
    __threadfence();
    atomicMax(queue + 1, first_pub + l.n);
  }
}

// one 128-thread group per CTA.  Exactly 64 KB: three CTAs (+1 KB each reserved by the system) fit the 196 KB shared-memory
// carve-out; one more byte selects the 228 KB split, and a kernel whose L1/smem split differs from its neighbours' cannot
// overlap them under programmatic dependent launch (measured: +13 us per token step).
constexpr int kPdAttnSmemBytes = 2 * kPdStageBytes;

}  // namespace mocr

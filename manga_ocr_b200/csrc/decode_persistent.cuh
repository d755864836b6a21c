// Persistent greedy decoder: ONE cooperative launch runs every decode step of a batch.
//
// Replaces the per-step host loop of GenerationMixin._sample (transformers/generation/utils.py:
// 2743-2805: ~120 library launches and a host sync per token).  One CTA per SM stays resident and
// interprets a small stage table (26 entries per token):
//   BertLayer x2 [QKV -> self-attn -> out -> LN -> cross-q -> cross-attn -> out -> LN -> FFN1 -> FFN2 -> LN]
//   (modeling_bert.py:143-421), LM head (:471-501), arg-max / EOS / append / embed (utils.py:2793-2805).
// Stages are separated by a grid-wide barrier (one release-atomic + a poll, 1.1 us measured) and the
// loop ends on the device when every row has produced [SEP] or reached max_length - no host round
// trip per token.
//
// Design rules that came out of measurements on the B200 (tools/microbench.cu, tools/decode_prof.py):
//  * a token is latency/bandwidth-bound (M = batch rows, 46 MFLOP per row), so the small-M GEMMs use
//    warp-level mma.sync fed straight from L2 with 16-byte loads in a k-permuted fragment order
//    that needs no shared-memory staging;
//  * what costs is the broadcast of the activation to all CTAs (98 KB per CTA = 1.3 us at the L2's
//    ~11 TB/s aggregate): the N = 768 projections are therefore split over K as well as N
//    (48 x 3 tiles), each CTA reads a third of the activation, and the raw fp32 partials are summed
//    in a fixed order (no atomics: results are deterministic) by the LayerNorm / attention stage
//    that consumes them, together with bias, GELU and residual;
//  * weights and the encoder K/V do not depend on the previous stage, so they are requested BEFORE
//    waiting on the barrier; attention K/V travel through cp.async into per-thread shared-memory
//    staging slots (no registers held, double-buffered across units);
//  * the code is kept small (one instance per stage TYPE, rolled loops): the stage sequence does not
//    fit in the instruction caches, and an earlier fully inlined version (247 KB of SASS) lost more
//    to instruction fetch than it gained.
#pragma once
#include "common.cuh"
#include "rowops.cuh"

namespace mocr {

constexpr int kPdThreads = 256;              // 8 warps, one CTA per SM
constexpr int kPdWarps = kPdThreads / 32;
constexpr int kPdRowsPerBlock = 64;          // activation rows per pass (4 m-tiles of 16)
constexpr int kPdKSlices = 2;                // warps per m-tile: K is split in two inside a CTA
constexpr int kPdGroups = kPdThreads / 128;  // attention groups of 4 warps
constexpr int kPdKeySlots = 8;               // 8 * 16 = 128 keys per staged block (a 197-key unit = 2 blocks); each warp owns 32 consecutive keys
constexpr int kPdStageBytes = 2 * kPdKeySlots * 128 * 16;          // K and V of one block, one group
constexpr int kPdMaxNT = 48;
constexpr int kPdRedFloats = kPdKSlices * kPdRowsPerBlock * (kPdMaxNT + 1);
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
  int eos_id;
  PdLayer layer[kDecLayers];
  PdLinear head_t, head_dec;
  PdLn head_ln;
  EmbedWeights emb;
  const __nv_bfloat16* crosskv;   // [B][layer][K|V][head][197][64]
  // state
  int* ids;                 // [B, max_len]
  int* pos;                 // [B]
  int* finished;            // [B]
  const int* forced;        // teacher forcing or null
  float* x;                 // [B, 768] post-LN hidden (fp32 residual)
  __nv_bfloat16* xb;        // [B, 768] bf16 copy (GEMM A operand)
  float* y;                 // [3, B, 768] split-K partials of the projections that feed a LayerNorm
  float* yq;                // [3, B, 768] split-K partials of the cross-attention query
  __nv_bfloat16* qkv;       // [B, 2304]
  __nv_bfloat16* ctx;       // [B, 768]
  __nv_bfloat16* ffn;       // [B, 3072]
  float* part_max;          // [B, kPdVocabTiles]
  int* part_idx;
  float* logits;            // tap or null: [B, max_len-1, 6144]
  unsigned int* barrier;    // grid barrier counter (zeroed by the host before launch)
  int* steps_done;          // out
  long long* prof;          // optional [4096] stage timeline of CTA 0 (debug / tuning), or null
};

enum PdStageType { PD_GEMM16 = 0, PD_GEMM32 = 1, PD_GEMM48 = 2, PD_ATTN_SELF = 3, PD_ATTN_CROSS = 4, PD_LN = 5, PD_NEXT = 6 };
enum PdEpi { PD_BF16 = 0, PD_BF16_GELU = 1, PD_F32_PARTIAL = 2, PD_ARGMAX = 3 };

// One entry of the per-token program, built once in shared memory.
struct PdStage {
  int type;
  int epi;                  // GEMM: PdEpi.  LN: bit 0 = GELU before the norm (LM-head transform)
  int N, K, ksplit, ldo;
  int parts;                // LN: number of split-K partials to add
  int layer;                // attention: decoder layer
  const __nv_bfloat16* A;   // GEMM A operand [B, K]
  const __nv_bfloat16* W;   // GEMM weights [N, K]
  const float* bias;        // GEMM bias (unsplit) / LN: bias of the producing projection / cross: query bias
  __nv_bfloat16* ob;        // GEMM bf16 out / LN bf16 out
  float* of;                // GEMM fp32 partial out / LN fp32 out (nullable)
  const float* src;         // LN: partials [parts][B][768]
  const float* resid;       // LN: residual rows (nullable)
  const float* g;           // LN gamma / beta
  const float* b;
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
// Barrier poll.  Deliberately relaxed: an acquire load at gpu scope makes ptxas invalidate the
// SM's whole L1 (CCTL.IVALL) at every barrier, evicting the cached LayerNorm / bias vectors.
// Every datum that another CTA writes during the launch is read with ld.global.cg (L2, the
// coherence point) and the writer publishes with red.release.gpu, so no L1 line can be stale; the
// consuming loads are issued after the poll loop exits (control dependence + bar.sync).
__device__ __forceinline__ unsigned int ld_poll_u32(const unsigned int* p) {
  unsigned int r;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
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

// Grid barrier, split so that independent loads can be issued between arrive and wait.
struct GridBarrier {
  unsigned int* ctr;
  unsigned int target;
  long long* prof;        // optional stage timeline of CTA 0 (clock64 at every wait-exit and arrive), or null
  int prof_n;
  __device__ __forceinline__ void stamp() {
    if (prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0 && prof_n < 4096) prof[prof_n] = clock64();
    ++prof_n;
  }
  __device__ __forceinline__ void arrive() {
    __syncthreads();                       // every thread's writes of this stage are done
    stamp();
    if (threadIdx.x == 0) {
      // release at gpu scope: the CTA's writes (ordered before this by the barrier above) become
      // visible to whoever observes the counter
      asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ctr), "r"(1u) : "memory");
    }
    target += gridDim.x;
  }
  __device__ __forceinline__ void wait() {
    if (threadIdx.x == 0) {
      const long long t0 = clock64();
      while (ld_poll_u32(ctr) < target) {
        if (clock64() - t0 > 8000000000LL) __trap();   // a lost arrival must not hang the GPU box
      }
    }
    __syncthreads();
    stamp();
  }
};

// Stage kernels launched one by one (CUDA graph) use the same stage code with this no-op barrier:
// stream order provides the dependency.
// With programmatic dependent launch the next stage kernel is already resident while this one
// runs: everything before wait() (weight / encoder-K/V prefetch, index arithmetic) overlaps the
// previous stage, and wait() returns once the previous grid has completed and flushed.
__device__ __forceinline__ long long global_timer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
struct NullBarrier {
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
// (KS = 2 in the persistent kernel, 4 in the stage kernels: 512 threads keep ~130 KB in flight).
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
      bs[i] = __ldg(reinterpret_cast<const float4*>(st.bias + c));
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
      for (int i = 0; i < 6; ++i) part[pt][i] = ldg_cg_f4(st.src + (static_cast<size_t>(pt) * B + r) * kD + (lane + 32 * i) * 4);
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
    if (was_finished && (p.forced == nullptr || ps >= p.max_len - 1)) continue;
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
    if (lane == 0) {
      if (np < p.max_len) p.ids[static_cast<size_t>(r) * p.max_len + np] = tok;
      p.finished[r] = fin;
      p.pos[r] = np;
    }
    if (p.forced != nullptr && np < p.max_len) tok = p.forced[static_cast<size_t>(r) * p.max_len + np];
    if (np >= p.max_len - 1 || np >= kMaxPos) continue;
    pd_embed_row_warp(p, k, r, tok, np, lane);
  }
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
template <bool HINT>
__device__ __forceinline__ void pd_attn_request_t(uint4* stage, const PdAttnUnit& a, int key_stride, int j0, int gt, uint64_t pol) {
  const int gw = gt >> 5, sub = (gt & 31) >> 3, ch = gt & 7;
  if (j0 + 32 * gw >= a.n_keys) return;                 // warp-uniform: nothing of this warp's range exists
  int fresh_j = -1;
#pragma unroll 1
  for (int i = 0; i < kPdKeySlots; ++i) {
    const int kl = 32 * gw + 4 * i + sub;               // key index inside the block
    const int j = j0 + kl;
    uint4* kd = stage + kl * 8 + (ch ^ (kl & 7));
    uint4* vd = kd + kPdKeySlots * 128;
    if (j < a.n_keys) {
      const bool fresh = j == a.ps;
      const __nv_bfloat16* ks = fresh ? a.nk : a.kc + static_cast<size_t>(j) * key_stride;
      const __nv_bfloat16* vs = fresh ? a.nv : a.vc + static_cast<size_t>(j) * key_stride;
      if (HINT) {
        cp_async16_cg_hint(kd, ks, pol);
        cp_async16_cg_hint(vd, vs, pol);
      } else {
        cp_async16_cg(kd, ks);
        cp_async16_cg(vd, vs);
      }
      if (fresh) fresh_j = j;
    } else {
      cp_async16_zero(vd, a.vc);
    }
  }
  if (fresh_j >= 0) {    // append this token's row to the cache (8 lanes, 16 bytes each, per K and V) once the stream is issued
    const uint4 kq = ldg_cg16(a.nk), vq = ldg_cg16(a.nv);
    *reinterpret_cast<uint4*>(const_cast<__nv_bfloat16*>(a.kc) + static_cast<size_t>(fresh_j) * key_stride) = kq;
    *reinterpret_cast<uint4*>(const_cast<__nv_bfloat16*>(a.vc) + static_cast<size_t>(fresh_j) * key_stride) = vq;
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
  const int n_groups = blockDim.x >> 7;       // 2 in the persistent kernel, 1 in the stage kernel
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
  int pre_fin[kPre] = {0, 0, 0, 0}, pre_pos[kPre] = {0, 0, 0, 0};
  auto make_unit = [&](int u, int fin, int pos, bool state_visible) {
    PdAttnUnit a;
    a.b = u / kHeads;
    a.h = u - a.b * kHeads;
    const int kvb = SELF ? a.b : a.b / p.kv_div;                    // cross: the beams of a crop read the same K/V
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
    int fin, pos;
    if (k < kPre) {
      fin = pre_fin[0];
      pos = pre_pos[0];
#pragma unroll
      for (int i = 1; i < kPre; ++i) {              // (select chain: the arrays stay in registers)
        if (k == i) {
          fin = pre_fin[i];
          pos = pre_pos[i];
        }
      }
    } else {
      fin = ldg_cg_s32(p.finished + u / kHeads);
      pos = SELF ? ldg_cg_s32(p.pos + u / kHeads) : 0;
    }
    return make_unit(u, fin, pos, true);
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
    } else {
      const int c = h * kHeadDim + ch * 8;
#pragma unroll
      for (int pt = 0; pt < kPdSplit; ++pt) {
        const float* src = p.yq + (static_cast<size_t>(pt) * p.B + b) * kD + c;
        raw.e0[pt] = ldg_cg_f4(src);
        raw.e1[pt] = ldg_cg_f4(src + 4);
      }
    }
  };
  auto finish_q = [&](int u, const QRaw& raw, float (&qq)[8]) {
    if (SELF) {
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
    if (stream_kv) pd_attn_request_t<true>(stage, a, key_stride, j0, gt, pol);
    else pd_attn_request_t<false>(stage, a, key_stride, j0, gt, 0ull);
  };
  PdAttnUnit cur = make_unit(u0 < units ? u0 : 0, 0, 0, false);
  bool pre2 = false;       // both blocks of the first unit were requested before the dependency wait
  if (!SELF) {   // encoder K/V never change during a decode: request the first unit (2 blocks) before the wait
    if (u0 < units) {
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
      if (SELF) pre_pos[k] = ldg_cg_s32(p.pos + uk / kHeads);
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
    if (SELF || u != u0) {
      if (!SELF) cp_async_wait_group<0>();          // (the prefetched blocks of a finished row are dropped)
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

constexpr int kPdSmemBytes = kPdGroups * 2 * kPdStageBytes + kPdMaxStages * static_cast<int>(sizeof(PdStage)) + 128;
static_assert(kPdRedFloats * 4 <= kPdGroups * 2 * kPdStageBytes, "reduction scratch must fit in the staging area it aliases");

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

// The per-token program (shared by the persistent kernel, which builds it in shared memory, and
// by the host, which launches it stage by stage in the CUDA-graph mode).
__host__ __device__ inline int pd_build_program(const PdParams& p, PdStage* prog) {
  int n = 0;
  for (int l = 0; l < kDecLayers; ++l) {
    const PdLayer& L = p.layer[l];
    PdStage a{};
    // self-attention block (modeling_bert.py:143-207, 287-298)
    prog[n++] = pd_gemm_desc(PD_GEMM16, PD_BF16, p.xb, kD, L.qkv, 3 * kD, 1, p.qkv, nullptr, 3 * kD);
    a = PdStage{}; a.type = PD_ATTN_SELF; a.layer = l; prog[n++] = a;
    prog[n++] = pd_gemm_desc(PD_GEMM16, PD_F32_PARTIAL, p.ctx, kD, L.self_out, kD, kPdSplit, nullptr, p.y, kD);
    prog[n++] = pd_ln_desc(p.y, kPdSplit, L.self_out.bias, 0, p.x, L.ln_self, p.x, p.xb);
    // cross-attention block (:210-284)
    prog[n++] = pd_gemm_desc(PD_GEMM16, PD_F32_PARTIAL, p.xb, kD, L.cross_q, kD, kPdSplit, nullptr, p.yq, kD);
    a = PdStage{}; a.type = PD_ATTN_CROSS; a.layer = l; a.bias = L.cross_q.bias; prog[n++] = a;
    prog[n++] = pd_gemm_desc(PD_GEMM16, PD_F32_PARTIAL, p.ctx, kD, L.cross_out, kD, kPdSplit, nullptr, p.y, kD);
    prog[n++] = pd_ln_desc(p.y, kPdSplit, L.cross_out.bias, 0, p.x, L.ln_cross, p.x, p.xb);
    // feed-forward (:330-356)
    prog[n++] = pd_gemm_desc(PD_GEMM32, PD_BF16_GELU, p.xb, kD, L.fc1, kFFN, 1, p.ffn, nullptr, kFFN);
    prog[n++] = pd_gemm_desc(PD_GEMM16, PD_F32_PARTIAL, p.ffn, kFFN, L.fc2, kD, kPdSplit, nullptr, p.y, kD);
    prog[n++] = pd_ln_desc(p.y, kPdSplit, L.fc2.bias, 0, p.x, L.ln_ffn, p.x, p.xb);
  }
  // LM head (:471-501): dense -> GELU -> LayerNorm, then the vocabulary projection fused with the
  // per-tile arg-max: logits never leave the SM unless the parity tap is on
  prog[n++] = pd_gemm_desc(PD_GEMM16, PD_F32_PARTIAL, p.xb, kD, p.head_t, kD, kPdSplit, nullptr, p.y, kD);
  prog[n++] = pd_ln_desc(p.y, kPdSplit, p.head_t.bias, 1, nullptr, p.head_ln, nullptr, p.xb);
  prog[n++] = pd_gemm_desc(PD_GEMM48, PD_ARGMAX, p.xb, kD, p.head_dec, kVocab, 1, nullptr, nullptr, 0);
  PdStage nx{};
  nx.type = PD_NEXT;
  prog[n++] = nx;
  return n;
}

// One out-of-line instance per stage type keeps the persistent kernel's code small.
template <int NT, int CH>
__device__ __noinline__ void pd_gemm_call(GridBarrier& bar, float* red, const PdParams& p, const PdStage& st) { pd_gemm_stage<NT, CH, kPdKSlices>(bar, red, p, st); }
template <bool SELF>
__device__ __noinline__ void pd_attention_call(GridBarrier& bar, uint8_t* smem, const PdParams& p, const PdStage& st) { pd_attention_stage<SELF>(bar, smem, p, st); }
__device__ __noinline__ void pd_ln_call(GridBarrier& bar, const PdParams& p, const PdStage& st) { pd_ln_stage(bar, p, st, gridDim.x); }
__device__ __noinline__ void pd_next_call(GridBarrier& bar, const PdParams& p) { pd_next_token_stage(bar, p, gridDim.x); }

__global__ void __launch_bounds__(kPdThreads, 1) decode_persistent_kernel(const __grid_constant__ PdParams p) {
  extern __shared__ __align__(128) uint8_t pd_smem[];
  float* red = reinterpret_cast<float*>(pd_smem);     // aliases the attention staging area (stages never overlap)
  PdStage* prog = reinterpret_cast<PdStage*>(pd_smem + kPdGroups * 2 * kPdStageBytes);
  __shared__ int s_nstages;
  GridBarrier bar{p.barrier, 0u, p.prof, 0};
  const int B = p.B;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) s_nstages = pd_build_program(p, prog);   // the per-token program

  // step 0 input: ids[b][0] = [CLS] (generation/utils.py:806-863), PAD elsewhere; x = embed([CLS], 0)
  for (int r = blockIdx.x * kPdWarps + warp; r < B; r += gridDim.x * kPdWarps) {
    for (int i = lane; i < p.max_len; i += 32) p.ids[static_cast<size_t>(r) * p.max_len + i] = i == 0 ? 2 : 0;
    if (lane == 0) {
      p.pos[r] = 0;
      p.finished[r] = p.max_len <= 1 ? 1 : 0;
    }
    PdEmbedConsts ek;
    pd_embed_consts(p, lane, ek);
    pd_embed_row_warp(p, ek, r, 2, 0, lane);
  }
  bar.arrive();       // (also orders the program table: arrive() starts with __syncthreads)
  const int n_stages = s_nstages;

  const int max_steps = p.max_len - 1;
  int step = 0;
#pragma unroll 1
  for (; step < max_steps; ++step) {
#pragma unroll 1
    for (int si = 0; si < n_stages; ++si) {
      const PdStage& st = prog[si];
      switch (st.type) {
        case PD_GEMM16: pd_gemm_call<16, 4>(bar, red, p, st); break;
        case PD_GEMM32: pd_gemm_call<32, 2>(bar, red, p, st); break;
        case PD_GEMM48: pd_gemm_call<48, 2>(bar, red, p, st); break;
        case PD_ATTN_SELF: pd_attention_call<true>(bar, pd_smem, p, st); break;
        case PD_ATTN_CROSS: pd_attention_call<false>(bar, pd_smem, p, st); break;
        case PD_LN: pd_ln_call(bar, p, st); break;
        default: pd_next_call(bar, p); break;
      }
    }
    // device-side termination (generation/utils.py:2805 does this with a host sync per token)
    bar.wait();
    int live = 0;
    for (int r = threadIdx.x; r < B; r += kPdThreads) live |= (ldg_cg_s32(p.finished + r) == 0) ? 1 : 0;
    live = __syncthreads_or(live);      // no arrive: the next stage's wait() passes at once (same target)
    if (!live && p.forced == nullptr) { ++step; break; }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *p.steps_done = step;
}


// ------------------------------------------------------------------ stage kernels (CUDA-graph mode) ---
// The same stage code, one launch per stage; grid = number of tiles / units so that every CTA has
// exactly one piece of work and several CTAs share an SM (latency hiding the persistent kernel,
// with 8 warps per SM, does not have).

constexpr int kPdStageKS = 4;                  // K-slices (warps per m-tile) of the stage-kernel GEMMs: 512 threads
constexpr int pd_gemm_smem_bytes(int nt) { return kPdStageKS * kPdRowsPerBlock * (nt + 1) * 4; }
// `tail` (PD_LN or PD_NEXT, or type < 0 for none) is the row stage that consumes this GEMM: it is
// fused into the same launch.  Every CTA publishes its tile with a release-add on `counter`; the
// first `tail_ctas` CTAs then wait until all tiles have arrived and run the row stage.  All CTAs
// of the grid are co-resident (grid <= SM count, one 512-thread CTA per SM), so the wait cannot
// deadlock; it replaces a 3.5 us dependent launch by a ~1 us counter poll.  The counters are
// zeroed by the first kernel of every token step (flag kPdZeroCounters).
constexpr int kPdCounters = 16;
constexpr int kPdZeroCounters = 0x100;          // PdStage::epi flag of the step's first GEMM

template <int NT, int CH>
__global__ void __launch_bounds__(128 * kPdStageKS, 1) pd_gemm_kernel(const __grid_constant__ PdParams p, const __grid_constant__ PdStage st,
                                                                       const __grid_constant__ PdStage tail, unsigned int* counters, int slot,
                                                                       int tail_ctas) {
  extern __shared__ __align__(16) float red[];      // kPdStageKS * 64 * (NT + 1) floats (pd_gemm_smem_bytes)
  NullBarrier bar;
  bar.begin(p.prof, st.type * 100 + st.ksplit * 10 + (st.K > 1000 ? 1 : 0));
  pdl_launch_dependents();
  PdStage mine = st;
  mine.epi = st.epi & 0xff;
  pd_gemm_stage<NT, CH, kPdStageKS>(bar, red, p, mine);
  if ((st.epi & kPdZeroCounters) && blockIdx.x == 0 && threadIdx.x < kPdCounters && threadIdx.x != slot) counters[threadIdx.x] = 0u;
  if (tail.type < 0) return;
  __syncthreads();                                   // the whole tile is written
  if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(counters + slot), "r"(1u) : "memory");
  if (blockIdx.x >= tail_ctas) return;
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    while (ld_poll_u32(counters + slot) < gridDim.x) {
      if (clock64() - t0 > 8000000000LL) __trap();
    }
  }
  __syncthreads();
  struct Passed { __device__ __forceinline__ void arrive() {} __device__ __forceinline__ void wait() {} } done;
  if (tail.type == PD_LN) pd_ln_stage(done, p, tail, tail_ctas);
  else pd_next_token_stage(done, p, tail_ctas);
}
template <bool SELF>
__global__ void __launch_bounds__(128) pd_attention_kernel(const __grid_constant__ PdParams p, const __grid_constant__ PdStage st) {
  extern __shared__ __align__(128) uint8_t pd_smem[];
  NullBarrier bar;
  bar.begin(p.prof, st.type * 100);
  pdl_launch_dependents();
  pd_attention_stage<SELF>(bar, pd_smem, p, st);
}
__global__ void __launch_bounds__(kPdThreads) pd_ln_kernel(const __grid_constant__ PdParams p, const __grid_constant__ PdStage st) {
  NullBarrier bar;
  bar.begin(p.prof, st.type * 100);
  pdl_launch_dependents();
  pd_ln_stage(bar, p, st, gridDim.x);
}
__global__ void __launch_bounds__(kPdThreads) pd_next_kernel(const __grid_constant__ PdParams p) {
  NullBarrier bar;
  bar.begin(p.prof, PD_NEXT * 100);
  pdl_launch_dependents();
  pd_next_token_stage(bar, p, gridDim.x);
}
// ids[b][0] = [CLS], pos = 0, finished = 0, x = embed([CLS], 0)
__global__ void __launch_bounds__(kPdThreads) pd_begin_kernel(const __grid_constant__ PdParams p) {
  pdl_launch_dependents();
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = blockIdx.x * kPdWarps + warp; r < p.B; r += gridDim.x * kPdWarps) {
    for (int i = lane; i < p.max_len; i += 32) p.ids[static_cast<size_t>(r) * p.max_len + i] = i == 0 ? 2 : 0;
    if (lane == 0) {
      p.pos[r] = 0;
      p.finished[r] = p.max_len <= 1 ? 1 : 0;
    }
    PdEmbedConsts ek;
    pd_embed_consts(p, lane, ek);
    pd_embed_row_warp(p, ek, r, 2, 0, lane);
  }
}

// one 128-thread group per CTA.  Exactly 64 KB: three CTAs (+1 KB each reserved by the system) fit the 196 KB shared-memory
// carve-out; one more byte selects the 228 KB split, and a kernel whose L1/smem split differs from its neighbours' cannot
// overlap them under programmatic dependent launch (measured: +13 us per token step).
constexpr int kPdAttnSmemBytes = 2 * kPdStageBytes;

}  // namespace mocr

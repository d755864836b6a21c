// Device side of the beam-search mode.  A beam step = the greedy
// path's stage kernels over n * num_beams rows (all but the next-token stage), then:
//   beam_topk_kernel    log_softmax over the row's 6144 logits (generation/utils.py:3249-3250), the
//                       n-gram ban of NoRepeatNGramLogitsProcessor (logits_process.py:1133-1134: banned
//                       tokens get -inf AFTER the log-softmax), and the row's K best continuations
//   (host)              BeamSearch::step (beam_search.h)
//   beam_kv_gather      the self-attention cache follows the surviving beams (Cache.reorder_cache, :3345-3350)
//   beam_advance_kernel position += 1, x = embed(next token)
#pragma once
#include "decode_stages.cuh"

namespace mocr {

constexpr int kBeamMaxK = 32;

// grid = rows, block = 256.  cand_lp / cand_tok: [rows][K], descending log-probability, ties -> lowest token id.
__global__ void __launch_bounds__(256)
beam_topk_kernel(const float* __restrict__ logits, const int* __restrict__ ban_tok, const int* __restrict__ ban_cnt, int ban_cap, int K,
                 float* __restrict__ cand_lp, int* __restrict__ cand_tok) {
  constexpr int kPer = kVocab / 256;   // 24
  __shared__ float s_val[8];
  __shared__ int s_idx[8];
  __shared__ int s_bidx;
  const int r = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* lg = logits + static_cast<size_t>(r) * kVocab;
  float v[kPer];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < kPer; ++i) {
    v[i] = lg[tid + 256 * i];
    mx = fmaxf(mx, v[i]);
  }
  mx = warp_max(mx);
  if (lane == 0) s_val[warp] = mx;
  __syncthreads();
  mx = s_val[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, s_val[w]);
  __syncthreads();
  float se = 0.f;
#pragma unroll
  for (int i = 0; i < kPer; ++i) se += expf(v[i] - mx);
  se = warp_sum(se);
  if (lane == 0) s_val[warp] = se;
  __syncthreads();
  se = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) se += s_val[w];   // fixed order
  const float lse = mx + logf(se);
  __syncthreads();
  // n-gram ban: the banned tokens leave the candidate set (their log-probability is -inf)
  const int nb = min(ban_cnt[r], ban_cap);
  for (int k = 0; k < nb; ++k) {
    const int t = ban_tok[static_cast<size_t>(r) * ban_cap + k];
    if (t >= 0 && t < kVocab && (t & 255) == tid) {
#pragma unroll
      for (int i = 0; i < kPer; ++i)
        if (i == (t >> 8)) v[i] = -INFINITY;
    }
  }
  // K rounds of (max, lowest index) over the block
  for (int k = 0; k < K; ++k) {
    float best = -INFINITY;
    int bi = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      const int idx = tid + 256 * i;
      if (v[i] > best || (v[i] == best && idx < bi)) { best = v[i]; bi = idx; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) { s_val[warp] = best; s_idx[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
      float b = s_val[0];
      int ix = s_idx[0];
      for (int w = 1; w < 8; ++w)
        if (s_val[w] > b || (s_val[w] == b && s_idx[w] < ix)) { b = s_val[w]; ix = s_idx[w]; }
      s_bidx = ix;
      cand_lp[static_cast<size_t>(r) * K + k] = b - lse;
      cand_tok[static_cast<size_t>(r) * K + k] = ix;
    }
    __syncthreads();
    const int ix = s_bidx;
    if ((ix & 255) == tid) {
#pragma unroll
      for (int i = 0; i < kPer; ++i)
        if (i == (ix >> 8)) v[i] = -INFINITY;      // taken
    }
    __syncthreads();
  }
}

// dst[r][0..len) = src[parent[r]][0..len) for K and V of both layers.  grid = (rows, 2 * kDecLayers), block = 256.
struct BeamCaches {
  const __nv_bfloat16* src[2 * kDecLayers];
  __nv_bfloat16* dst[2 * kDecLayers];
};
__global__ void __launch_bounds__(256)
beam_kv_gather_kernel(BeamCaches c, const int* __restrict__ parent, int len, int cache_len) {
  const int r = blockIdx.x, which = blockIdx.y;
  const uint4* s = reinterpret_cast<const uint4*>(c.src[which] + static_cast<size_t>(parent[r]) * cache_len * kD);
  uint4* d = reinterpret_cast<uint4*>(c.dst[which] + static_cast<size_t>(r) * cache_len * kD);
  const int n16 = len * (kD * 2 / 16);
  for (int i = threadIdx.x; i < n16; i += 256) d[i] = s[i];
}

// One warp per row: the row consumes `next_tok[r]` at position `position`.
__global__ void __launch_bounds__(256)
beam_advance_kernel(const __grid_constant__ PdParams p, const int* __restrict__ next_tok, int position) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;
  if (r >= p.B) return;
  PdEmbedConsts ek;
  pd_embed_consts(p, lane, ek);
  const int tok = next_tok[r];
  if (lane == 0) {
    p.pos[r] = position;
    p.finished[r] = 0;
    if (position < p.max_len) p.ids[static_cast<size_t>(r) * p.max_len + position] = tok;
  }
  pd_embed_row_warp(p, ek, r, tok, position, lane);
}

}  // namespace mocr

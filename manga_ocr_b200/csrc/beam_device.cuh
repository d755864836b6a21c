// Device-resident beam search: the per-step selection of transformers'
// GenerationMixin._beam_search (generation/utils.py:3076-3370) as kernels, so that a beam step needs no host round trip
// and a run of steps replays from a CUDA graph.
//
// The arithmetic is beam_search.h (the host bookkeeping, pinned exactly against transformers by tests/test_beam_host.py)
// restated for one CTA per crop: thread 0 runs the O(beams^2) selection in float32 in the same order of operations,
// all threads copy the token rows.  What changes is where the state lives:
//   * running / finished token rows, scores and flags stay in device memory (two parities of the row buffers, so a step
//     reads the old rows while it writes the new ones);
//   * the n-gram ban (NoRepeatNGramLogitsProcessor, logits_process.py:1012-1136) is computed from the running rows
//     inside the top-k kernel;
//   * the self-attention cache follows the beams through a row table (logical beam row -> physical cache row): a beam
//     that survives keeps its row, only a parent chosen by SEVERAL children is copied (into the row of a dropped beam) -
//     the reference instead gathers the whole cache every step (Cache.reorder_cache, :3345-3350);
//   * the loop state (current length, "unfinished", parity) is a small control block the kernels read, so that the
//     kernels of a graph replayed past the end of the search are no-ops.
#pragma once
#include "beam_kernels.cuh"

namespace mocr {

constexpr int kBeamMaxBeams = kBeamMaxK / 2;      // 16

enum BeamCtl { BC_CUR_LEN = 0, BC_UNFINISHED = 1, BC_PARITY = 2, BC_ALL_HIT = 3, BC_ANY_UNSAT = 4, BC_ALL_FIN = 5, BC_ARRIVED = 6, BC_STEPS = 7 };

struct BeamDev {
  int n, beams, K, T, ngram, early;      // early: 0 False, 1 True, 2 "never"
  float length_penalty;
  int eos, fill, prompt_len, start_token;
  int* run[2];          // [n*beams][T] running token rows, two parities
  int* fin[2];          // [n*beams][T] finished hypotheses, two parities
  float* run_score;     // [n*beams]
  float* fin_score;     // [n*beams]
  int* fin_len;         // [n*beams] generated tokens
  int* is_fin;          // [n*beams]
  int* unsat;           // [n] the early-stop heuristic still allows an improvement
  int* ctl;             // [8] BeamCtl
  float* cand_lp;       // [n*beams][K]
  int* cand_tok;        // [n*beams][K]
  int* next;            // [n*beams] token each row consumes next
  int* parent;          // [n*beams] row each row continues
  int* phys;            // [n*beams] physical self-attention cache row of each beam row
  int* copy_src;        // [n*beams] cache row to copy from (-1: none) ...
  int* copy_dst;        // ... into this row, before the next step
};

__global__ void __launch_bounds__(256) beam_init_kernel(BeamDev d) {
  const int R = d.n * d.beams;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < R * d.T; i += gridDim.x * blockDim.x) {
    const int v = (i % d.T) == 0 ? d.start_token : d.fill;
    d.run[0][i] = v; d.run[1][i] = v; d.fin[0][i] = v; d.fin[1][i] = v;
  }
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < R; r += gridDim.x * blockDim.x) {
    d.run_score[r] = (r % d.beams) == 0 ? 0.f : -1e9f;
    d.fin_score[r] = -1e9f;
    d.fin_len[r] = 0;
    d.is_fin[r] = 0;
    d.phys[r] = r;
    d.copy_src[r] = d.copy_dst[r] = -1;
    if (r < d.n) d.unsat[r] = 1;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    d.ctl[BC_CUR_LEN] = d.prompt_len;
    d.ctl[BC_UNFINISHED] = d.T > d.prompt_len ? 1 : 0;
    d.ctl[BC_PARITY] = 0;
    d.ctl[BC_ALL_HIT] = 1;
    d.ctl[BC_ANY_UNSAT] = 0;
    d.ctl[BC_ALL_FIN] = 1;
    d.ctl[BC_ARRIVED] = 0;
    d.ctl[BC_STEPS] = 0;
  }
}

// log-softmax + n-gram ban + the row's K best continuations (beam_topk_kernel with the ban list computed here from the
// row's running tokens).  grid = rows, block = 256.
__global__ void __launch_bounds__(256) beam_topk_dev_kernel(BeamDev d, const float* __restrict__ logits) {
  pdl_launch_dependents();
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (d.ctl[BC_UNFINISHED] == 0) return;
  constexpr int kPer = kVocab / 256;   // 24
  __shared__ float s_val[8];
  __shared__ float s_rv[2][8];
  __shared__ int s_ri[2][8];
  __shared__ int s_ban[kMaxPos];
  __shared__ int s_nban;
  const int r = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cur_len = d.ctl[BC_CUR_LEN], K = d.K;
  const int* seq = d.run[d.ctl[BC_PARITY]] + static_cast<size_t>(r) * d.T;
  const float* lg = logits + static_cast<size_t>(r) * kVocab;
  if (tid == 0) s_nban = 0;
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
  // every w such that the n-gram (last ngram-1 tokens, w) already occurs in the row (prompt included)
  if (d.ngram > 0 && cur_len + 1 >= d.ngram) {
    const int m = d.ngram - 1;
    for (int i = tid; i + d.ngram <= cur_len; i += 256) {
      bool same = true;
      for (int k = 0; k < m && same; ++k) same = seq[i + k] == seq[cur_len - m + k];
      if (same) {
        const int slot = atomicAdd(&s_nban, 1);
        if (slot < kMaxPos) s_ban[slot] = seq[i + m];
      }
    }
  }
  __syncthreads();
  se = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) se += s_val[w];   // fixed order
  const float lse = mx + logf(se);
  const int nb = min(s_nban, kMaxPos);
  for (int k = 0; k < nb; ++k) {
    const int t = s_ban[k];
    if (t >= 0 && t < kVocab && (t & 255) == tid) {
#pragma unroll
      for (int i = 0; i < kPer; ++i)
        if (i == (t >> 8)) v[i] = -INFINITY;
    }
  }
  __syncthreads();
  // K rounds of (max, lowest index) over the block; the warps' partial results are double-buffered by round, every thread
  // merges them itself: one barrier per round
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
    if (lane == 0) { s_rv[k & 1][warp] = best; s_ri[k & 1][warp] = bi; }
    __syncthreads();
    float bb = s_rv[k & 1][0];
    int ix = s_ri[k & 1][0];
#pragma unroll
    for (int w = 1; w < 8; ++w) {
      const float wv = s_rv[k & 1][w];
      const int wi = s_ri[k & 1][w];
      if (wv > bb || (wv == bb && wi < ix)) { bb = wv; ix = wi; }
    }
    if (tid == 0) {
      d.cand_lp[static_cast<size_t>(r) * K + k] = bb - lse;
      d.cand_tok[static_cast<size_t>(r) * K + k] = ix;
    }
    if ((ix & 255) == tid) {
#pragma unroll
      for (int i = 0; i < kPer; ++i)
        if (i == (ix >> 8)) v[i] = -INFINITY;      // taken
    }
  }
}

// One selection step for crop blockIdx.x (BeamSearch::step of beam_search.h) and the embedding of the tokens its beams
// consume next; the last CTA to finish closes the step (global "unfinished", current length, parity).  block = 128.
// The three sorts of the host code (stable, by descending value) are done as rank computations: element x goes to position
// #{y : y sorts before x}, every thread ranking its own elements - the comparators are strict total orders, so the result is
// the stable sort's.
__global__ void __launch_bounds__(128) beam_select_kernel(const __grid_constant__ PdParams p, BeamDev d) {
  pdl_launch_dependents();
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (d.ctl[BC_UNFINISHED] == 0) return;
  constexpr int MB = kBeamMaxBeams, MK = kBeamMaxK;
  __shared__ float s_acc[MB * MK], s_top_lp[MK], s_run_lp[MK], s_mscore[MB + MK], s_denom;
  __shared__ int s_tok[MB * MK], s_order[MK], s_top_beam[MK], s_top_tok[MK], s_hit[MK], s_sel[MB], s_midx[MB], s_mfin[MB + MK], s_mlen[MB + MK];
  __shared__ int s_full, s_unsat;
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int beams = d.beams, K = d.K, T = d.T, nc = beams * K;
  const int cur_len = d.ctl[BC_CUR_LEN], par = d.ctl[BC_PARITY];
  const int* run_old = d.run[par];
  int* run_new = d.run[par ^ 1];
  const int* fin_old = d.fin[par];
  int* fin_new = d.fin[par ^ 1];
  const int row0 = b * beams;
  // accumulated log-probabilities of the beams * K candidates
  for (int i = tid; i < nc; i += blockDim.x) {
    const int j = i / K;
    s_acc[i] = d.cand_lp[static_cast<size_t>(row0 + j) * K + i % K] + d.run_score[row0 + j];
    s_tok[i] = d.cand_tok[static_cast<size_t>(row0 + j) * K + i % K];
  }
  if (tid == 0) {
    s_denom = static_cast<float>(pow(static_cast<double>(cur_len + 1 - d.prompt_len), static_cast<double>(d.length_penalty)));
    bool full = d.early == 1;
    for (int j = 0; j < beams; ++j) full = full && d.is_fin[row0 + j] != 0;
    s_full = full ? 1 : 0;
    s_unsat = d.unsat[b] != 0 ? 1 : 0;
  }
  __syncthreads();
  // top K of them: (value desc, beam asc, token asc) = torch.topk over the flattened [beams * vocab] scores
  for (int x = tid; x < nc; x += blockDim.x) {
    int rank = 0;
    const float ax = s_acc[x];
    for (int y = 0; y < nc; ++y) {
      const float ay = s_acc[y];
      bool before;
      if (ay != ax) before = ay > ax;
      else if (y / K != x / K) before = y / K < x / K;
      else before = s_tok[y] < s_tok[x];
      rank += (y != x && before) ? 1 : 0;
    }
    if (rank < K) s_order[rank] = x;
  }
  __syncthreads();
  if (tid < K) {
    const int c = s_order[tid];
    s_top_lp[tid] = s_acc[c];
    s_top_beam[tid] = c / K;
    s_top_tok[tid] = s_tok[c];
    const int hit = ((cur_len + 1 >= T) || s_tok[c] == d.eos) ? 1 : 0;
    s_hit[tid] = hit;
    s_run_lp[tid] = s_acc[c] + (hit ? 1.0f : 0.0f) * -1.0e9f;
  }
  __syncthreads();
  if (tid < K) {
    // running beams of the next iteration: the best `beams` candidates that did not stop (stable by candidate rank)
    int rank = 0;
    for (int y = 0; y < K; ++y) rank += (s_run_lp[y] > s_run_lp[tid] || (s_run_lp[y] == s_run_lp[tid] && y < tid)) ? 1 : 0;
    if (rank < beams) s_sel[rank] = tid;
    // finished beams: candidates among the first `beams` that stopped, length-penalised
    const bool did = s_hit[tid] && tid < beams;
    float v = s_top_lp[tid] / s_denom;
    v += (s_full ? 1.0f : 0.0f) * -1.0e9f;
    v += (s_unsat ? 0.0f : 1.0f) * -1.0e9f;
    v += (did ? 0.0f : 1.0f) * -1.0e9f;
    s_mscore[beams + tid] = v;
    s_mfin[beams + tid] = did ? 1 : 0;
    s_mlen[beams + tid] = cur_len + 1 - d.prompt_len;
  } else if (tid >= 64 && tid < 64 + beams) {        // ... merged with the previous finished set
    const int j = tid - 64;
    s_mscore[j] = d.fin_score[row0 + j];
    s_mfin[j] = d.is_fin[row0 + j];
    s_mlen[j] = d.fin_len[row0 + j];
  }
  __syncthreads();
  if (tid < beams + K) {
    int rank = 0;
    for (int y = 0; y < beams + K; ++y) rank += (s_mscore[y] > s_mscore[tid] || (s_mscore[y] == s_mscore[tid] && y < tid)) ? 1 : 0;
    if (rank < beams) s_midx[rank] = tid;
  }
  __syncthreads();
  // ---- token rows (read the old parity, write the new one): one warp per row, the whole row in registers before the first
  //      store (row after row with interleaved loads and stores, every piece waited for the previous stores: ~24 dependent
  //      L2 round trips per step)
  for (int rowi = warp; rowi < 2 * beams; rowi += blockDim.x >> 5) {
    const int j = rowi >> 1;
    const int* src;
    int* dst;
    int tok = -1;                                   // token written at position cur_len (-1: plain copy)
    if ((rowi & 1) == 0) {                          // finished set, slot j
      const int i = s_midx[j];
      dst = fin_new + static_cast<size_t>(row0 + j) * T;
      if (i < beams) {
        src = fin_old + static_cast<size_t>(row0 + i) * T;
      } else {
        src = run_old + static_cast<size_t>(row0 + s_top_beam[i - beams]) * T;
        tok = s_top_tok[i - beams];
      }
    } else {                                        // running beam j
      const int k = s_sel[j];
      src = run_old + static_cast<size_t>(row0 + s_top_beam[k]) * T;
      dst = run_new + static_cast<size_t>(row0 + j) * T;
      tok = s_top_tok[k];
    }
    int v[16];                                      // T <= 512
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int c = lane + 32 * i;
      v[i] = c < T ? src[c] : 0;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int c = lane + 32 * i;
      if (c < T) dst[c] = (tok >= 0 && c == cur_len) ? tok : v[i];
    }
  }
  // ---- every running row of the crop consumes its next token at position cur_len (one warp per row)
  {
    PdEmbedConsts ek;
    pd_embed_consts(p, lane, ek);
    for (int j = warp; j < beams; j += blockDim.x >> 5) {
      const int r = row0 + j;
      if (lane == 0) {
        p.pos[r] = cur_len;
        p.finished[r] = 0;
      }
      pd_embed_row_warp(p, ek, r, s_top_tok[s_sel[j]], cur_len, lane);
    }
  }
  if (tid == 0) {
    float fs[MB], new_score[MB];
    int ff[MB], old_phys[MB], new_phys[MB], claimed[MB];
    bool all_hit = true;
    for (int k = 0; k < K; ++k) all_hit = all_hit && s_hit[k] != 0;
    if (!all_hit) atomicAnd(d.ctl + BC_ALL_HIT, 0);
    for (int j = 0; j < beams; ++j) {
      const int i = s_midx[j];
      fs[j] = s_mscore[i];
      ff[j] = s_mfin[i];
      d.fin_score[row0 + j] = fs[j];
      d.is_fin[row0 + j] = ff[j];
      d.fin_len[row0 + j] = s_mlen[i];
    }
    for (int j = 0; j < beams; ++j) {
      const int k = s_sel[j];
      new_score[j] = s_run_lp[k];
      d.run_score[row0 + j] = new_score[j];
      d.next[row0 + j] = s_top_tok[k];
      d.parent[row0 + j] = row0 + s_top_beam[k];
    }
    // ---- cache rows: a surviving parent hands its physical row to its first child; further children of the same parent
    //      get the row of a beam nobody continues, and a copy of the parent's cache
    for (int j = 0; j < beams; ++j) { old_phys[j] = d.phys[row0 + j]; claimed[j] = 0; }
    for (int j = 0; j < beams; ++j) {
      const int pj = s_top_beam[s_sel[j]];
      if (!claimed[pj]) { new_phys[j] = old_phys[pj]; claimed[pj] = 1; d.copy_src[row0 + j] = -1; d.copy_dst[row0 + j] = -1; }
      else new_phys[j] = -1;
    }
    int fr = 0;
    for (int j = 0; j < beams; ++j) {
      if (new_phys[j] >= 0) continue;
      while (claimed[fr]) ++fr;            // as many unclaimed parents as duplicated ones
      new_phys[j] = old_phys[fr];
      claimed[fr] = 1;
      d.copy_src[row0 + j] = old_phys[s_top_beam[s_sel[j]]];
      d.copy_dst[row0 + j] = new_phys[j];
    }
    for (int j = 0; j < beams; ++j) d.phys[row0 + j] = new_phys[j];
    // ---- can the open beams still beat the finished ones?  (with the length after this step; :2876-2921)
    const int len1 = cur_len + 1;
    const int best_len = (d.early == 2 && d.length_penalty > 0.0f) ? T - d.prompt_len : len1 - d.prompt_len;
    const float best_possible = new_score[0] / static_cast<float>(pow(static_cast<double>(best_len), static_cast<double>(d.length_penalty)));
    float mn = fs[0];
    for (int j = 1; j < beams; ++j) mn = fminf(mn, fs[j]);
    bool any = false, all_fin = true;
    for (int j = 0; j < beams; ++j) {
      const float worst = ff[j] ? mn : -1.0e9f;
      any = any || best_possible > worst;
      all_fin = all_fin && ff[j] != 0;
    }
    const int us = (s_unsat && any) ? 1 : 0;
    d.unsat[b] = us;
    if (us) atomicOr(d.ctl + BC_ANY_UNSAT, 1);
    if (!all_fin) atomicAnd(d.ctl + BC_ALL_FIN, 0);
  }
  __syncthreads();                                          // every thread's writes of this crop are issued
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(d.ctl + BC_ARRIVED, 1) == d.n - 1) {      // the last crop closes the step
      __threadfence();
      const int all_hit = atomicAdd(d.ctl + BC_ALL_HIT, 0), any_unsat = atomicAdd(d.ctl + BC_ANY_UNSAT, 0), allf = atomicAdd(d.ctl + BC_ALL_FIN, 0);
      d.ctl[BC_UNFINISHED] = (any_unsat && !(allf && d.early == 1) && !all_hit) ? 1 : 0;
      d.ctl[BC_CUR_LEN] = cur_len + 1;
      d.ctl[BC_PARITY] = par ^ 1;
      d.ctl[BC_STEPS] = d.ctl[BC_STEPS] + 1;
      d.ctl[BC_ALL_HIT] = 1;
      d.ctl[BC_ANY_UNSAT] = 0;
      d.ctl[BC_ALL_FIN] = 1;
      d.ctl[BC_ARRIVED] = 0;
    }
  }
}

// The children of a parent that several beams continue get a copy of its cache rows [0, len).
// grid = (rows, 2 * layers, kBeamCopySplit): a row's copy (up to 460 KB per layer and K|V) is spread over several CTAs, four
// 16-byte loads in flight per thread (one CTA per row copy took 46 us at 250 cached tokens).
constexpr int kBeamCopySplit = 8;
__global__ void __launch_bounds__(256) beam_kv_copy_kernel(BeamDev d, BeamCaches c, int cache_len) {
  pdl_launch_dependents();
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (d.ctl[BC_UNFINISHED] == 0) return;
  const int r = blockIdx.x, which = blockIdx.y;
  const int src_row = d.copy_src[r], dst_row = d.copy_dst[r];
  if (src_row < 0 || dst_row < 0) return;
  const int len = d.ctl[BC_CUR_LEN] - 1;             // positions whose K/V exist
  const uint4* s = reinterpret_cast<const uint4*>(c.src[which] + static_cast<size_t>(src_row) * cache_len * kD);
  uint4* t = reinterpret_cast<uint4*>(c.dst[which] + static_cast<size_t>(dst_row) * cache_len * kD);
  const int n16 = len * (kD * 2 / 16);
  const int per = (n16 + kBeamCopySplit - 1) / kBeamCopySplit;
  const int lo = blockIdx.z * per, hi = min(n16, lo + per);
  int i = lo + threadIdx.x;
  for (; i + 3 * 256 < hi; i += 4 * 256) {
    const uint4 a = s[i], b = s[i + 256], e = s[i + 512], f = s[i + 768];
    t[i] = a;
    t[i + 256] = b;
    t[i + 512] = e;
    t[i + 768] = f;
  }
  for (; i < hi; i += 256) t[i] = s[i];
}

}  // namespace mocr

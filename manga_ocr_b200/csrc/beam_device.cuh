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
  if (d.ctl[BC_UNFINISHED] == 0) return;
  constexpr int kPer = kVocab / 256;   // 24
  __shared__ float s_val[8];
  __shared__ int s_idx[8];
  __shared__ int s_bidx;
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
      d.cand_lp[static_cast<size_t>(r) * K + k] = b - lse;
      d.cand_tok[static_cast<size_t>(r) * K + k] = ix;
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

// One selection step for crop blockIdx.x (BeamSearch::step of beam_search.h); the last CTA to finish closes the step
// (global "unfinished", current length, parity).  block = 128.
__global__ void __launch_bounds__(128) beam_select_kernel(BeamDev d) {
  if (d.ctl[BC_UNFINISHED] == 0) return;
  constexpr int MB = kBeamMaxBeams, MK = kBeamMaxK;
  __shared__ int s_top_beam[MK], s_top_tok[MK], s_sel[MB], s_midx[MB];
  __shared__ float s_run_lp[MK], s_mscore[MB + MK];
  __shared__ int s_mfin[MB + MK], s_mlen[MB + MK];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int beams = d.beams, K = d.K, T = d.T;
  const int cur_len = d.ctl[BC_CUR_LEN], par = d.ctl[BC_PARITY];
  const int* run_old = d.run[par];
  int* run_new = d.run[par ^ 1];
  const int* fin_old = d.fin[par];
  int* fin_new = d.fin[par ^ 1];
  const int row0 = b * beams;
  if (tid == 0) {
    float acc[MB * MK];
    int order[MB * MK];
    for (int j = 0; j < beams; ++j)
      for (int k = 0; k < K; ++k) acc[j * K + k] = d.cand_lp[(static_cast<size_t>(row0) + j) * K + k] + d.run_score[row0 + j];
    // stable sort by (value desc, beam asc, token asc): torch.topk over the flattened [beams * vocab] scores
    const int nc = beams * K;
    for (int i = 0; i < nc; ++i) {
      int x = i, p = i;
      while (p > 0) {
        const int y = order[p - 1];
        bool before;     // does x come before y?
        if (acc[x] != acc[y]) before = acc[x] > acc[y];
        else if (x / K != y / K) before = x / K < y / K;
        else before = d.cand_tok[(static_cast<size_t>(row0) + x / K) * K + x % K] < d.cand_tok[(static_cast<size_t>(row0) + y / K) * K + y % K];
        if (!before) break;
        order[p] = y;
        --p;
      }
      order[p] = x;
    }
    float top_lp[MK], fin_lp[MK];
    int hit[MK];
    bool all_hit = true;
    for (int k = 0; k < K; ++k) {
      const int c = order[k];
      top_lp[k] = acc[c];
      s_top_beam[k] = c / K;
      s_top_tok[k] = d.cand_tok[(static_cast<size_t>(row0) + c / K) * K + c % K];
      hit[k] = (cur_len + 1 >= T) || s_top_tok[k] == d.eos;
      all_hit = all_hit && hit[k];
      s_run_lp[k] = top_lp[k] + (hit[k] ? 1.0f : 0.0f) * -1.0e9f;
    }
    // running beams of the next iteration: the best `beams` candidates that did not stop (stable by candidate rank)
    int top[MK];
    for (int i = 0; i < K; ++i) {
      int p = i;
      while (p > 0 && s_run_lp[i] > s_run_lp[top[p - 1]]) { top[p] = top[p - 1]; --p; }
      top[p] = i;
    }
    for (int j = 0; j < beams; ++j) s_sel[j] = top[j];
    // finished beams: candidates among the first `beams` that stopped, length-penalised, merged with the previous set
    const float denom = static_cast<float>(pow(static_cast<double>(cur_len + 1 - d.prompt_len), static_cast<double>(d.length_penalty)));
    bool full = d.early == 1;
    for (int j = 0; j < beams; ++j) full = full && d.is_fin[row0 + j] != 0;
    const bool unsat_b = d.unsat[b] != 0;
    for (int k = 0; k < K; ++k) {
      const bool did = hit[k] && k < beams;
      float v = top_lp[k] / denom;
      v += (full ? 1.0f : 0.0f) * -1.0e9f;
      v += (unsat_b ? 0.0f : 1.0f) * -1.0e9f;
      v += (did ? 0.0f : 1.0f) * -1.0e9f;
      fin_lp[k] = v;
    }
    for (int j = 0; j < beams; ++j) {
      s_mscore[j] = d.fin_score[row0 + j];
      s_mfin[j] = d.is_fin[row0 + j];
      s_mlen[j] = d.fin_len[row0 + j];
    }
    for (int k = 0; k < K; ++k) {
      s_mscore[beams + k] = fin_lp[k];
      s_mfin[beams + k] = (hit[k] && k < beams) ? 1 : 0;
      s_mlen[beams + k] = cur_len + 1 - d.prompt_len;
    }
    int midx[MB + MK];
    for (int i = 0; i < beams + K; ++i) {
      int p = i;
      while (p > 0 && s_mscore[i] > s_mscore[midx[p - 1]]) { midx[p] = midx[p - 1]; --p; }
      midx[p] = i;
    }
    for (int j = 0; j < beams; ++j) s_midx[j] = midx[j];
    if (!all_hit) atomicAnd(d.ctl + BC_ALL_HIT, 0);
  }
  __syncthreads();
  // ---- token rows (read the old parity, write the new one)
  for (int j = 0; j < beams; ++j) {
    const int i = s_midx[j];
    int* dst = fin_new + static_cast<size_t>(row0 + j) * T;
    if (i < beams) {
      const int* src = fin_old + static_cast<size_t>(row0 + i) * T;
      for (int c = tid; c < T; c += blockDim.x) dst[c] = src[c];
    } else {
      const int k = i - beams;
      const int* src = run_old + static_cast<size_t>(row0 + s_top_beam[k]) * T;
      for (int c = tid; c < T; c += blockDim.x) dst[c] = c == cur_len ? s_top_tok[k] : src[c];
    }
    const int k = s_sel[j];
    const int* rsrc = run_old + static_cast<size_t>(row0 + s_top_beam[k]) * T;
    int* rdst = run_new + static_cast<size_t>(row0 + j) * T;
    for (int c = tid; c < T; c += blockDim.x) rdst[c] = c == cur_len ? s_top_tok[k] : rsrc[c];
  }
  __syncthreads();
  if (tid == 0) {
    float fs[MB];
    int ff[MB], fl[MB], old_phys[MB], new_phys[MB], claimed[MB];
    for (int j = 0; j < beams; ++j) {
      const int i = s_midx[j];
      fs[j] = s_mscore[i];
      ff[j] = s_mfin[i];
      fl[j] = s_mlen[i];
    }
    for (int j = 0; j < beams; ++j) {
      d.fin_score[row0 + j] = fs[j];
      d.is_fin[row0 + j] = ff[j];
      d.fin_len[row0 + j] = fl[j];
    }
    float new_score[MB];
    for (int j = 0; j < beams; ++j) new_score[j] = s_run_lp[s_sel[j]];
    for (int j = 0; j < beams; ++j) {
      const int k = s_sel[j];
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
    const int us = (d.unsat[b] != 0 && any) ? 1 : 0;
    d.unsat[b] = us;
    if (us) atomicOr(d.ctl + BC_ANY_UNSAT, 1);
    if (!all_fin) atomicAnd(d.ctl + BC_ALL_FIN, 0);
    __threadfence();
    if (atomicAdd(d.ctl + BC_ARRIVED, 1) == d.n - 1) {      // the last crop closes the step
      __threadfence();
      const int all_hit = atomicAdd(d.ctl + BC_ALL_HIT, 0), any_unsat = atomicAdd(d.ctl + BC_ANY_UNSAT, 0), allf = atomicAdd(d.ctl + BC_ALL_FIN, 0);
      d.ctl[BC_UNFINISHED] = (any_unsat && !(allf && d.early == 1) && !all_hit) ? 1 : 0;
      d.ctl[BC_CUR_LEN] = len1;
      d.ctl[BC_PARITY] = par ^ 1;
      d.ctl[BC_STEPS] = d.ctl[BC_STEPS] + 1;
      d.ctl[BC_ALL_HIT] = 1;
      d.ctl[BC_ANY_UNSAT] = 0;
      d.ctl[BC_ALL_FIN] = 1;
      d.ctl[BC_ARRIVED] = 0;
    }
  }
}

// The children of a parent that several beams continue get a copy of its cache rows [0, len).  grid = (rows, 2 * layers).
__global__ void __launch_bounds__(256) beam_kv_copy_kernel(BeamDev d, BeamCaches c, int cache_len) {
  if (d.ctl[BC_UNFINISHED] == 0) return;
  const int r = blockIdx.x, which = blockIdx.y;
  const int src_row = d.copy_src[r], dst_row = d.copy_dst[r];
  if (src_row < 0 || dst_row < 0) return;
  const int len = d.ctl[BC_CUR_LEN] - 1;             // positions whose K/V exist
  const uint4* s = reinterpret_cast<const uint4*>(c.src[which] + static_cast<size_t>(src_row) * cache_len * kD);
  uint4* t = reinterpret_cast<uint4*>(c.dst[which] + static_cast<size_t>(dst_row) * cache_len * kD);
  const int n16 = len * (kD * 2 / 16);
  for (int i = threadIdx.x; i < n16; i += 256) t[i] = s[i];
}

// Every running row consumes its next token at position cur_len - 1.  One warp per row.
__global__ void __launch_bounds__(256) beam_advance_dev_kernel(const __grid_constant__ PdParams p, BeamDev d) {
  if (d.ctl[BC_UNFINISHED] == 0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;
  if (r >= p.B) return;
  const int position = d.ctl[BC_CUR_LEN] - 1;
  PdEmbedConsts ek;
  pd_embed_consts(p, lane, ek);
  const int tok = d.next[r];
  if (lane == 0) {
    p.pos[r] = position;
    p.finished[r] = 0;
  }
  pd_embed_row_warp(p, ek, r, tok, position, lane);
}

}  // namespace mocr

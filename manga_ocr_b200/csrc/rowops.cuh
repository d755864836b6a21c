// Row-wise kernels: LayerNorm (fp32 statistics), ViT CLS rows, decoder embeddings and
// the greedy next-token step.  All are HBM/latency-bound vector kernels: one warp (or
// one small CTA) per 768-wide row, float4 / 16-byte accesses, no shared-memory staging.
#pragma once
#include "common.cuh"

namespace mocr {

// y = LayerNorm(x) * gamma + beta over rows of 768 fp32; writes bf16 (GEMM A operand)
// and optionally fp32 (residual stream / parity tap).  One warp per row.
// Reference: ViT layernorm_before/after + final layernorm (modeling_vit.py:333,340,455),
// BERT post-LN (modeling_bert.py:110,297,355,484).  eps = 1e-12.
__global__ void __launch_bounds__(256)
layernorm_rows_kernel(const float* __restrict__ x, int rows, const float* __restrict__ gamma, const float* __restrict__ beta,
                      __nv_bfloat16* __restrict__ out_bf16, float* __restrict__ out_f32) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * kD);
  float4 v[6];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    v[i] = xr[lane + 32 * i];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.0f / kD);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / kD) + kLnEps);
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const int c4 = lane + 32 * i;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c4);
    float4 y;
    y.x = (v[i].x - mean) * rstd * g.x + b.x;
    y.y = (v[i].y - mean) * rstd * g.y + b.y;
    y.z = (v[i].z - mean) * rstd * g.z + b.z;
    y.w = (v[i].w - mean) * rstd * g.w + b.w;
    if (out_bf16) {
      uint2 p;
      p.x = pack_bf16(y.x, y.y);
      p.y = pack_bf16(y.z, y.w);
      reinterpret_cast<uint2*>(out_bf16 + static_cast<size_t>(row) * kD)[c4] = p;
    }
    if (out_f32) reinterpret_cast<float4*>(out_f32 + static_cast<size_t>(row) * kD)[c4] = y;
  }
}

// ViT embeddings, CLS row: h[b*197 + 0] = cls_token + position_embeddings[0]
// (modeling_vit.py:117-124).  grid = n_crops, block = 192 (float4 each).
__global__ void __launch_bounds__(192)
cls_rows_kernel(float* __restrict__ h, const float* __restrict__ cls, const float* __restrict__ pos) {
  const float4 c = reinterpret_cast<const float4*>(cls)[threadIdx.x];
  const float4 p = reinterpret_cast<const float4*>(pos)[threadIdx.x];
  reinterpret_cast<float4*>(h + static_cast<size_t>(blockIdx.x) * kEncTokens * kD)[threadIdx.x] =
      make_float4(c.x + p.x, c.y + p.y, c.z + p.z, c.w + p.w);
}

struct DecodeState {
  int* ids;            // [B, max_len]  ids[b][0] = [CLS]
  int* pos;            // [B] index of the current input token (= number of cached keys)
  int* finished;       // [B]
  const int* forced;   // optional teacher-forcing ids [B, max_len] (parity tests), else null
  int max_len;
  float* x;            // [B, 768] fp32 hidden state
  __nv_bfloat16* xb;   // [B, 768] bf16 copy (GEMM A operand)
};

struct EmbedWeights {
  const float* word;   // [6144, 768]
  const float* posemb; // [512, 768]
  const float* type0;  // [768]  token_type_embeddings[0]
  const float* gamma;  // embeddings.LayerNorm
  const float* beta;
};

// x = LayerNorm(word[tok] + type[0] + pos[p])   (modeling_bert.py:102-111)
// block = 256 threads, 3 elements each.
__device__ __forceinline__ void embed_ln_row(int tok, int p, const EmbedWeights& w, float* x, __nv_bfloat16* xb, float* red /*[16]*/) {
  const int tid = threadIdx.x;
  float v[3];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int c = tid + 256 * i;
    v[i] = w.word[static_cast<size_t>(tok) * kD + c] + w.type0[c] + w.posemb[static_cast<size_t>(p) * kD + c];
    s += v[i];
  }
  s = warp_sum(s);
  if ((tid & 31) == 0) red[tid >> 5] = s;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[i];
  const float mean = tot * (1.0f / kD);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) q += (v[i] - mean) * (v[i] - mean);
  q = warp_sum(q);
  if ((tid & 31) == 0) red[8 + (tid >> 5)] = q;
  __syncthreads();
  float qt = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) qt += red[8 + i];
  const float rstd = rsqrtf(qt * (1.0f / kD) + kLnEps);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int c = tid + 256 * i;
    const float y = (v[i] - mean) * rstd * w.gamma[c] + w.beta[c];
    x[c] = y;
    xb[c] = __float2bfloat16(y);
  }
}

// Start of a decode: ids[b][0] = [CLS], pos = 0, finished = 0, x = embed([CLS], 0).
// grid = B, block = 256.  (generation/utils.py:806-863: decoder_input_ids = [[2]])
__global__ void __launch_bounds__(256)
decode_begin_kernel(DecodeState st, EmbedWeights w, int start_id, int pad_id) {
  __shared__ float red[16];
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < st.max_len; i += 256)
    st.ids[static_cast<size_t>(b) * st.max_len + i] = i == 0 ? start_id : pad_id;
  if (threadIdx.x == 0) {
    st.pos[b] = 0;
    st.finished[b] = st.max_len <= 1 ? 1 : 0;
  }
  embed_ln_row(start_id, 0, w, st.x + static_cast<size_t>(b) * kD, st.xb + static_cast<size_t>(b) * kD, red);
}

// Greedy step tail, fused: reduce the LM head's per-tile (max, argmax) pairs, apply the
// finished/EOS/max_length rules of GenerationMixin._sample (generation/utils.py:2793-2805,
// stopping_criteria.py:76,470), append the token, and embed it for the next step.
// Ties resolve to the lowest index like torch.argmax.  grid = B, block = 256.
__global__ void __launch_bounds__(256)
next_token_kernel(DecodeState st, EmbedWeights w, const float* __restrict__ part_max, const int* __restrict__ part_idx,
                  int n_parts, int eos_id) {
  __shared__ float red[16];
  __shared__ float s_val[8];
  __shared__ int s_idx[8];
  __shared__ int s_tok;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int p = st.pos[b];
  const bool was_finished = st.finished[b] != 0;
  // A finished row keeps the [PAD] fill of decode_begin (generation/utils.py:2797) and is
  // never read again; a teacher-forced row only stops at max_length.
  if (was_finished && (st.forced == nullptr || p >= st.max_len - 1)) return;

  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = tid; i < n_parts; i += 256) {
    const float v = part_max[static_cast<size_t>(b) * n_parts + i];
    const int ix = part_idx[static_cast<size_t>(b) * n_parts + i];
    if (v > bv || (v == bv && ix < bi)) { bv = v; bi = ix; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  if ((tid & 31) == 0) { s_val[tid >> 5] = bv; s_idx[tid >> 5] = bi; }
  __syncthreads();
  if (tid == 0) {
    for (int i = 1; i < 8; ++i)
      if (s_val[i] > bv || (s_val[i] == bv && s_idx[i] < bi)) { bv = s_val[i]; bi = s_idx[i]; }
    if (bi < 0 || bi >= kVocab) bi = 1;      // all-NaN logits: emit [UNK] rather than index out of range
    int tok = bi;                            // teacher-forced runs record the raw argmax of every step
    const int np = p + 1;
    if (np < st.max_len) st.ids[static_cast<size_t>(b) * st.max_len + np] = tok;
    int fin = was_finished;
    if (tok == eos_id && st.forced == nullptr) fin = 1;
    if (np >= st.max_len - 1) fin = 1;       // length reaches max_length after this append
    st.finished[b] = fin;
    st.pos[b] = np;
    if (st.forced != nullptr && np < st.max_len) tok = st.forced[static_cast<size_t>(b) * st.max_len + np];
    s_tok = tok;
  }
  __syncthreads();
  const int np = p + 1;
  if (np >= st.max_len - 1 || np >= kMaxPos) return;   // no further step will read x
  embed_ln_row(s_tok, np, w, st.x + static_cast<size_t>(b) * kD, st.xb + static_cast<size_t>(b) * kD, red);
}

}  // namespace mocr

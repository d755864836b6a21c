// Single-query attention for one decode step: self-attention over the device-resident
// KV cache (causal => every cached key is visible; modeling_bert.py:143-207) and
// cross-attention over the 197 encoder keys projected once per crop
// (modeling_bert.py:210-284).  1/sqrt(64) is folded into the query weights.
//
// HBM-bound: per (crop, head) the kernel streams n_keys x 64 bf16 of K and of V exactly
// once with 16-byte loads (8 lanes cover one 128-byte key row, 4 rows per warp
// instruction, 16 rows per CTA iteration), fp32 math, no tensor cores.
#pragma once
#include "common.cuh"

namespace mocr {

constexpr int kDecAttnThreads = 128;
constexpr int kDecAttnMaxKeys = 512;

struct DecodeAttnArgs {
  const __nv_bfloat16* q;     // [B, ldq], head h at column h*64
  int ldq;
  __nv_bfloat16* kcache;      // key j of crop b: kcache + b*b_stride + j*key_stride + h*64
  __nv_bfloat16* vcache;
  long long b_stride;         // elements
  int key_stride;             // elements
  int head_stride;            // elements between heads (0: heads are 64-column slices of a key row)
  const __nv_bfloat16* new_k; // self-attention: this step's K/V rows [B, ld_new] to append at pos[b]; null for cross
  const __nv_bfloat16* new_v;
  int ld_new;
  const int* pos;             // self: n_keys = pos[b] + 1
  int fixed_keys;             // cross: 197
  const int* finished;        // rows to skip
  __nv_bfloat16* ctx;         // [B, 768]
};

__device__ __forceinline__ void bf16x8_to_f32(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(p[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

// grid = (12, B), block = 128
__global__ void __launch_bounds__(kDecAttnThreads)
decode_attention_kernel(const DecodeAttnArgs a) {
  __shared__ float s_score[kDecAttnMaxKeys];
  __shared__ float s_red[8];
  __shared__ float s_out[4][kHeadDim];
  const int h = blockIdx.x, b = blockIdx.y;
  if (a.finished[b]) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int sub = lane >> 3, ch = lane & 7;
  const int hs = a.head_stride > 0 ? a.head_stride : kHeadDim;
  __nv_bfloat16* kc = a.kcache + static_cast<size_t>(b) * a.b_stride + static_cast<size_t>(h) * hs;
  __nv_bfloat16* vc = a.vcache + static_cast<size_t>(b) * a.b_stride + static_cast<size_t>(h) * hs;
  int n_keys = a.fixed_keys;
  if (a.new_k != nullptr) {
    const int p = a.pos[b];
    n_keys = p + 1;
    if (tid < 16) {   // append this step's key and value rows (64 bf16 = 8 x 16 B each)
      const __nv_bfloat16* src = (tid < 8 ? a.new_k : a.new_v) + static_cast<size_t>(b) * a.ld_new + h * kHeadDim + (tid & 7) * 8;
      __nv_bfloat16* dst = (tid < 8 ? kc : vc) + static_cast<size_t>(p) * a.key_stride + (tid & 7) * 8;
      *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
    }
    __syncthreads();
  }
  if (n_keys > kDecAttnMaxKeys) n_keys = kDecAttnMaxKeys;

  float q[8];
  bf16x8_to_f32(*reinterpret_cast<const uint4*>(a.q + static_cast<size_t>(b) * a.ldq + h * kHeadDim + ch * 8), q);

  // scores
  float lmax = -INFINITY;
  for (int jb = warp * 4; jb < n_keys; jb += 64) {   // warp-uniform trip count (shuffles inside)
    const int j0 = jb + sub;
    uint4 kv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + 16 * u;
      kv[u] = j < n_keys ? *reinterpret_cast<const uint4*>(kc + static_cast<size_t>(j) * a.key_stride + ch * 8) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + 16 * u;
      float f[8];
      bf16x8_to_f32(kv[u], f);
      float d = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) d = fmaf(q[i], f[i], d);
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      d += __shfl_xor_sync(0xffffffffu, d, 2);
      d += __shfl_xor_sync(0xffffffffu, d, 4);
      if (j < n_keys) {
        if (ch == 0) s_score[j] = d;
        lmax = fmaxf(lmax, d);
      }
    }
  }
  lmax = warp_max(lmax);
  if (lane == 0) s_red[warp] = lmax;
  __syncthreads();
  const float gmax = fmaxf(fmaxf(s_red[0], s_red[1]), fmaxf(s_red[2], s_red[3]));
  float lsum = 0.f;
  for (int j = tid; j < n_keys; j += kDecAttnThreads) {
    const float e = __expf(s_score[j] - gmax);
    s_score[j] = e;
    lsum += e;
  }
  lsum = warp_sum(lsum);
  if (lane == 0) s_red[4 + warp] = lsum;
  __syncthreads();
  const float inv = 1.0f / (s_red[4] + s_red[5] + s_red[6] + s_red[7]);

  // weighted sum of values
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int jb = warp * 4; jb < n_keys; jb += 64) {   // warp-uniform trip count (shuffles inside)
    const int j0 = jb + sub;
    uint4 vv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + 16 * u;
      vv[u] = j < n_keys ? *reinterpret_cast<const uint4*>(vc + static_cast<size_t>(j) * a.key_stride + ch * 8) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + 16 * u;
      const float p = j < n_keys ? s_score[j] : 0.f;
      float f[8];
      bf16x8_to_f32(vv[u], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(p, f[i], acc[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
    acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
  }
  if (sub == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) s_out[warp][ch * 8 + i] = acc[i];
  }
  __syncthreads();
  if (tid < kHeadDim) {
    const float o = (s_out[0][tid] + s_out[1][tid] + s_out[2][tid] + s_out[3][tid]) * inv;
    a.ctx[static_cast<size_t>(b) * kD + h * kHeadDim + tid] = __float2bfloat16(o);
  }
}

}  // namespace mocr

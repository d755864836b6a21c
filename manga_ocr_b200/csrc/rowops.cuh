// Row-wise kernels of the encoder: LayerNorm (fp32 statistics) and the ViT CLS rows; HBM-bound vector kernels:
// one warp (or one small CTA) per 768-wide row, float4 / 16-byte accesses, no shared-memory staging.
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

struct EmbedWeights {
  const float* word;   // [6144, 768]
  const float* posemb; // [512, 768]
  const float* type0;  // [768]  token_type_embeddings[0]
  const float* gamma;  // embeddings.LayerNorm
  const float* beta;
};

}  // namespace mocr

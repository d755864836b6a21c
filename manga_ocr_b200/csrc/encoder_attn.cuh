// ViT self-attention over the 197 encoder tokens of one crop, all 12 heads:
//   ctx = softmax(Q K^T) V          (1/sqrt(64) is folded into W_q at load time)
// Reference: transformers/models/vit/modeling_vit.py:199-251 (non-causal, no mask).
//
// One CTA = (64-query tile, head, crop).  K and V of the head (197 x 64 bf16 each)
// are staged once in shared memory; each of the 4 warps owns 16 query rows, keeps the
// whole 16 x 208 score tile in registers (no online rescaling needed at S = 197),
// does the softmax in fp32 and feeds P straight back as the A operand of P V.
// Tensor-core path: warp-level mma.sync m16n8k16 (bf16 in, fp32 accumulate).
// Attention is 4 % of the encoder FLOPs.
#pragma once
#include "common.cuh"

namespace mocr {

constexpr int kAttnKeysPad = 208;     // 197 keys padded to 13 x 16
constexpr int kAttnPitch = 72;        // bf16 elements per smem row (64 + 8 pad: conflict-free fragments)
constexpr int kAttnThreads = 128;     // 4 warps x 16 query rows; 3 CTAs per SM overlap K/V staging with the MMAs
constexpr int kAttnQTile = kAttnThreads / 32 * 16;
constexpr int kAttnQTiles = (kEncTokens + kAttnQTile - 1) / kAttnQTile;
constexpr int kAttnSmemBytes = 2 * kAttnKeysPad * kAttnPitch * 2;

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// grid = (kAttnQTiles, 12, n_crops), block = 128
__global__ void __launch_bounds__(kAttnThreads, 3)
encoder_attention_kernel(const __nv_bfloat16* __restrict__ qkv /*[n*197, 2304]*/, __nv_bfloat16* __restrict__ ctx /*[n*197, 768]*/) {
  extern __shared__ __align__(16) uint8_t attn_smem[];
  __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(attn_smem);
  __nv_bfloat16* Vs = Ks + kAttnKeysPad * kAttnPitch;
  const int qt = blockIdx.x, head = blockIdx.y, crop = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const size_t row0 = static_cast<size_t>(crop) * kEncTokens;
  const int ld = 3 * kD;

  // Stage K and V (rows >= 197 zero-filled): 8 threads x 16 B per 128-B row, through cp.async so
  // that all 26 requests of a thread are in flight at once (a load->store loop costs one L2 round
  // trip per iteration).
  for (int i = threadIdx.x; i < kAttnKeysPad * 8; i += kAttnThreads) {
    const int r = i >> 3, ch = i & 7;
    __nv_bfloat16* kd = Ks + r * kAttnPitch + ch * 8;
    __nv_bfloat16* vd = Vs + r * kAttnPitch + ch * 8;
    if (r < kEncTokens) {
      const __nv_bfloat16* src = qkv + (row0 + r) * ld + head * kHeadDim + ch * 8;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(kd)), "l"(src + kD) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(vd)), "l"(src + 2 * kD) : "memory");
    } else {
      *reinterpret_cast<uint4*>(kd) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(vd) = make_uint4(0, 0, 0, 0);
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");

  // Q fragments of this warp's 16 rows straight from global memory.
  const int qrow = qt * kAttnQTile + warp * 16;
  uint32_t qa[4][4];
  {
    const int r_lo = qrow + g, r_hi = qrow + g + 8;
    const __nv_bfloat16* q_lo = qkv + (row0 + r_lo) * ld + head * kHeadDim;
    const __nv_bfloat16* q_hi = qkv + (row0 + r_hi) * ld + head * kHeadDim;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const int c = ks * 16 + 2 * t;
      qa[ks][0] = r_lo < kEncTokens ? *reinterpret_cast<const uint32_t*>(q_lo + c) : 0u;
      qa[ks][1] = r_hi < kEncTokens ? *reinterpret_cast<const uint32_t*>(q_hi + c) : 0u;
      qa[ks][2] = r_lo < kEncTokens ? *reinterpret_cast<const uint32_t*>(q_lo + c + 8) : 0u;
      qa[ks][3] = r_hi < kEncTokens ? *reinterpret_cast<const uint32_t*>(q_hi + c + 8) : 0u;
    }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  if (qrow >= kEncTokens) return;      // warp-uniform: this warp's rows are all padding

  // S = Q K^T : 26 key tiles of 8.
  float s[26][4];
#pragma unroll
  for (int nt = 0; nt < 26; ++nt) {
    s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
    const __nv_bfloat16* kr = Ks + (nt * 8 + g) * kAttnPitch + 2 * t;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const uint32_t b0 = *reinterpret_cast<const uint32_t*>(kr + ks * 16);
      const uint32_t b1 = *reinterpret_cast<const uint32_t*>(kr + ks * 16 + 8);
      mma_bf16_16816(s[nt], qa[ks], b0, b1);
    }
  }
  // Softmax over the 197 valid keys, fp32.
  constexpr float kLog2e = 1.4426950408889634f;
  float m_lo = -INFINITY, m_hi = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < 26; ++nt) {
    const int c = nt * 8 + 2 * t;
    if (c >= kEncTokens) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
    if (c + 1 >= kEncTokens) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
    m_lo = fmaxf(m_lo, fmaxf(s[nt][0], s[nt][1]));
    m_hi = fmaxf(m_hi, fmaxf(s[nt][2], s[nt][3]));
  }
  m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 1));
  m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 2));
  m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 1));
  m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 2));
  const float ml = m_lo * kLog2e, mh = m_hi * kLog2e;
  float sum_lo = 0.f, sum_hi = 0.f;
#pragma unroll
  for (int nt = 0; nt < 26; ++nt) {
    s[nt][0] = exp2f(s[nt][0] * kLog2e - ml);
    s[nt][1] = exp2f(s[nt][1] * kLog2e - ml);
    s[nt][2] = exp2f(s[nt][2] * kLog2e - mh);
    s[nt][3] = exp2f(s[nt][3] * kLog2e - mh);
    sum_lo += s[nt][0] + s[nt][1];
    sum_hi += s[nt][2] + s[nt][3];
  }
  sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 1);
  sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 2);
  sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 1);
  sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 2);

  // O = P V : 13 key steps of 16, 8 output tiles of 8 channels.
  float o[8][4];
#pragma unroll
  for (int nd = 0; nd < 8; ++nd) o[nd][0] = o[nd][1] = o[nd][2] = o[nd][3] = 0.f;
  // ldmatrix.x4.trans: lanes 0-7 / 8-15 address key rows k0..k0+7 / k0+8..k0+15 at channel nd*8,
  // lanes 16-23 / 24-31 the same rows at channel (nd+1)*8.
  const uint32_t v_lane = smem_u32(Vs + (lane & 15) * kAttnPitch + (lane >> 4) * 8);
#pragma unroll
  for (int kt = 0; kt < 13; ++kt) {
    uint32_t pa[4];
    pa[0] = pack_bf16(s[2 * kt][0], s[2 * kt][1]);
    pa[1] = pack_bf16(s[2 * kt][2], s[2 * kt][3]);
    pa[2] = pack_bf16(s[2 * kt + 1][0], s[2 * kt + 1][1]);
    pa[3] = pack_bf16(s[2 * kt + 1][2], s[2 * kt + 1][3]);
#pragma unroll
    for (int nd = 0; nd < 8; nd += 2) {
      uint32_t b0, b1, b2, b3;
      const uint32_t addr = v_lane + static_cast<uint32_t>((kt * 16 * kAttnPitch + nd * 8) * 2);
      asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                   : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3) : "r"(addr));
      mma_bf16_16816(o[nd], pa, b0, b1);
      mma_bf16_16816(o[nd + 1], pa, b2, b3);
    }
  }
  const float inv_lo = 1.f / sum_lo, inv_hi = 1.f / sum_hi;
  const int r_lo = qrow + g, r_hi = qrow + g + 8;
#pragma unroll
  for (int nd = 0; nd < 8; ++nd) {
    const int c = head * kHeadDim + nd * 8 + 2 * t;
    if (r_lo < kEncTokens)
      *reinterpret_cast<uint32_t*>(ctx + (row0 + r_lo) * kD + c) = pack_bf16(o[nd][0] * inv_lo, o[nd][1] * inv_lo);
    if (r_hi < kEncTokens)
      *reinterpret_cast<uint32_t*>(ctx + (row0 + r_hi) * kD + c) = pack_bf16(o[nd][2] * inv_hi, o[nd][3] * inv_hi);
  }
}

}  // namespace mocr

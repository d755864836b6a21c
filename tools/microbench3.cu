// Upper-bound experiment for a row-partitioned cluster decoder: 8 clusters x 16 CTAs, every cluster
// owns 8 rows and streams ALL decoder weights each token step (its CTAs 1/16 each), stages separated
// by cluster barriers, activations exchanged through L2.  No math: loads are XOR-reduced.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__device__ __forceinline__ void cluster_sync_() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint4 ld_nc(const void* p) { uint4 r; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); return r; }
__device__ __forceinline__ uint4 ld_cg(const void* p) { uint4 r; asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory"); return r; }

template <int DEPTH, bool NC>
__device__ __forceinline__ unsigned stream(const uint4* base, int n16) {
  unsigned acc = 0;
  for (int b = threadIdx.x; b < n16; b += blockDim.x * DEPTH) {
    uint4 v[DEPTH];
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) { int i = b + d * blockDim.x; v[d] = i < n16 ? (NC ? ld_nc(base + i) : ld_cg(base + i)) : make_uint4(0, 0, 0, 0); }
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) acc += v[d].x ^ v[d].y ^ v[d].z ^ v[d].w;
  }
  return acc;
}

struct Stage { int w_kb, kv_kb, act_kb, out_b; };
__constant__ Stage c_prog[32];

template <int DEPTH>
__global__ void k_decoder(const uint4* weights, const uint4* kv, uint4* act, int n_stages, int steps, long long* out) {
  cg::cluster_group cl = cg::this_cluster();
  const unsigned rank = cl.block_rank();
  const unsigned cluster = blockIdx.x / cl.num_blocks();
  unsigned acc = 0;
  cluster_sync_();
  long long t0 = clock64();
  for (int s = 0; s < steps; ++s) {
    size_t woff = 0, kvoff = 0;
    for (int i = 0; i < n_stages; ++i) {
      const Stage st = c_prog[i];
      // activations of this cluster (written by its 16 CTAs in the previous stage)
      acc += stream<4, false>(act + (size_t)cluster * 8192, st.act_kb * 64);
      // weights: the same slice for the same rank in every cluster (L2 hits after the first reader)
      acc += stream<DEPTH, true>(weights + (woff + (size_t)rank * st.w_kb * 64), st.w_kb * 64);
      woff += (size_t)16 * st.w_kb * 64;
      // K/V of this CTA's (row, head) units: private per CTA
      acc += stream<DEPTH, true>(kv + (kvoff + (size_t)blockIdx.x * st.kv_kb * 64), st.kv_kb * 64);
      kvoff += (size_t)gridDim.x * st.kv_kb * 64;
      for (int j = threadIdx.x; j < st.out_b / 16; j += blockDim.x) act[(size_t)cluster * 8192 + rank * 64 + j] = make_uint4(acc, s, i, j);
      cluster_sync_();
    }
  }
  long long dt = clock64() - t0;
  if (threadIdx.x == 0) out[1 + blockIdx.x] = acc;
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = dt / steps;
}

int main() {
  // per-layer stages (KB per CTA): qkv, self-attn, out, cross-q, cross-attn, cross-out, fc1, fc2 ; head_t, vocab, next
  Stage layer[8] = {{221, 0, 12, 2304}, {0, 230, 37, 768}, {74, 0, 12, 1536}, {74, 0, 24, 768}, {0, 300, 12, 768}, {74, 0, 12, 1536}, {295, 0, 24, 3072}, {295, 0, 48, 1536}};
  Stage prog[32]; int n = 0;
  for (int l = 0; l < 2; ++l) for (int i = 0; i < 8; ++i) prog[n++] = layer[i];
  prog[n++] = {74, 0, 24, 1536}; prog[n++] = {590, 0, 24, 1024}; prog[n++] = {0, 0, 16, 1536};
  CK(cudaMemcpyToSymbol(c_prog, prog, sizeof(Stage) * n));
  size_t wbytes = 0, kvbytes = 0;
  for (int i = 0; i < n; ++i) { wbytes += (size_t)16 * prog[i].w_kb * 1024; kvbytes += (size_t)128 * prog[i].kv_kb * 1024; }
  printf("stages %d, weights %.1f MB per cluster per step, K/V %.1f MB total per step\n", n, wbytes / 1e6, kvbytes / 1e6);
  uint4 *w, *kv, *act; long long* out;
  CK(cudaMalloc(&w, wbytes)); CK(cudaMalloc(&kv, kvbytes)); CK(cudaMalloc(&act, 8 * 8192 * 16 * 2)); CK(cudaMalloc(&out, 8 * 1024));
  CK(cudaMemset(w, 1, wbytes)); CK(cudaMemset(kv, 1, kvbytes)); CK(cudaMemset(act, 0, 8 * 8192 * 16 * 2));
  for (int depth : {8, 16}) for (int threads : {256, 512}) {
    auto kern = depth == 8 ? k_decoder<8> : k_decoder<16>;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(128); cfg.blockDim = dim3(threads);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 16; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    int steps = 50;
    const uint4* wc = w; const uint4* kvc = kv;
    CK(cudaLaunchKernelEx(&cfg, kern, wc, kvc, act, n, steps, out));
    CK(cudaDeviceSynchronize());
    long long cyc; CK(cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost));
    printf("depth %2d, %d threads: %.1f us per token step (%d stages)\n", depth, threads, cyc / 1965.0, n);
  }
  return 0;
}

// Micro-benchmarks for a cluster-partitioned decoder (run on a GPU box):
//  1. can 8/9 clusters of 16 CTAs be co-resident?   2. cluster barrier latency
//  3. DSMEM broadcast (every CTA stores S bytes into all 16 peers) + barrier
//  4. weight streaming: every cluster streams the same W (L2-resident), each CTA its 1/16 slice
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint4 ld_nc(const void* p) { uint4 r; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); return r; }

extern __shared__ __align__(16) uint8_t smem[];

__global__ void k_cbar(int iters, long long* out) {
  cluster_arrive(); cluster_wait();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) { cluster_arrive(); cluster_wait(); }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = (clock64() - t0) / iters;
}

__global__ void k_dsmem(int iters, int bytes, long long* out) {
  cg::cluster_group cl = cg::this_cluster();
  const unsigned rank = cl.block_rank(), n = cl.num_blocks();
  uint4* mine = reinterpret_cast<uint4*>(smem);
  cluster_arrive(); cluster_wait();
  long long t0 = clock64();
  const int n16 = bytes / 16;
  for (int it = 0; it < iters; ++it) {
    for (unsigned pr = 0; pr < n; ++pr) {
      uint4* dst = cl.map_shared_rank(mine, pr) + rank * n16;
      for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = make_uint4(it, i, rank, pr);
    }
    cluster_arrive(); cluster_wait();
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = (clock64() - t0) / iters;
  if (threadIdx.x == 0) out[1 + blockIdx.x] = mine[0].x;
}

template <int DEPTH>
__global__ void k_stream(const uint4* w, long long n16_total, int iters, long long* out) {
  cg::cluster_group cl = cg::this_cluster();
  const unsigned rank = cl.block_rank(), n = cl.num_blocks();
  const long long per = n16_total / n;
  const uint4* mine = w + rank * per;
  unsigned acc = 0;
  cluster_arrive(); cluster_wait();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    for (long long base = threadIdx.x; base < per; base += (long long)blockDim.x * DEPTH) {
      uint4 v[DEPTH];
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) { long long i = base + (long long)d * blockDim.x; v[d] = i < per ? ld_nc(mine + i) : make_uint4(0, 0, 0, 0); }
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) acc += v[d].x ^ v[d].y ^ v[d].z ^ v[d].w;
    }
  }
  long long dt = clock64() - t0;
  if (threadIdx.x == 0) out[1 + blockIdx.x] = acc;
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = dt / iters;
}

template <typename K, typename... A>
int launch(K kern, int clusters, int csize, int threads, size_t smem_bytes, A... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * csize); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem_bytes;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  int maxc = 0;
  CK(cudaOccupancyMaxActiveClusters(&maxc, kern, &cfg));
  printf("    [csize %d, %d thr, %zu B smem] max active clusters = %d\n", csize, threads, smem_bytes, maxc);
  if (maxc < clusters) { printf("    cannot co-schedule %d clusters\n", clusters); return 0; }
  CK(cudaLaunchKernelEx(&cfg, kern, args...));
  CK(cudaDeviceSynchronize());
  return 0;
}

int main() {
  long long* out; CK(cudaMalloc(&out, 8 * 1024));
  long long cyc;
  for (int cs : {8, 16}) {
    int clusters = cs == 16 ? 8 : 18;
    printf("cluster size %d x %d clusters:\n", cs, clusters);
    if (launch(k_cbar, clusters, cs, 256, 100 * 1024, 2000, out)) return 1;
    CK(cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost));
    printf("  cluster barrier: %lld cycles = %.2f us\n", cyc, cyc / 1965.0);
    for (int bytes : {768, 2304, 6144}) {
      if (launch(k_dsmem, clusters, cs, 256, 100 * 1024, 500, bytes, out)) return 1;
      CK(cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost));
      printf("  DSMEM broadcast of %d B per CTA to all peers + barrier: %lld cycles = %.2f us\n", bytes, cyc, cyc / 1965.0);
    }
    uint4* w; const long long wbytes = 44ll << 20; CK(cudaMalloc(&w, wbytes)); CK(cudaMemset(w, 1, wbytes));
    if (launch(k_stream<8>, clusters, cs, 256, 100 * 1024, (const uint4*)w, wbytes / 16, 20, out)) return 1;
    CK(cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost));
    printf("  stream 44 MB per cluster (depth 8): %.1f us per pass -> %.1f GB/s per SM, %.2f TB/s aggregate\n", cyc / 1965.0,
           wbytes / (double)cs / (cyc / 1965.0) / 1e3, wbytes * (double)clusters / (cyc / 1965.0) / 1e6);
    if (launch(k_stream<16>, clusters, cs, 256, 100 * 1024, (const uint4*)w, wbytes / 16, 20, out)) return 1;
    CK(cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost));
    printf("  stream 44 MB per cluster (depth 16): %.1f us per pass -> %.1f GB/s per SM, %.2f TB/s aggregate\n", cyc / 1965.0,
           wbytes / (double)cs / (cyc / 1965.0) / 1e3, wbytes * (double)clusters / (cyc / 1965.0) / 1e6);
    cudaFree(w);
  }
  return 0;
}

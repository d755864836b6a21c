// Micro-benchmarks behind the persistent decoder's design (run on a GPU box):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/microbench.cu -o /tmp/mb && /tmp/mb
// 1. grid barrier latency (red.release + relaxed poll)      2. broadcast read: every CTA reads the
// same L2-resident buffer (just written by all CTAs) with N 16-byte loads in flight per thread.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) { unsigned r; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(r) : "l"(p) : "memory"); return r; }
__device__ __forceinline__ uint4 ld_cg(const void* p) { uint4 r; asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory"); return r; }
__device__ __forceinline__ uint4 ld_ca(const void* p) { uint4 r; asm volatile("ld.global.ca.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory"); return r; }
__device__ __forceinline__ uint4 ld_nc(const void* p) { uint4 r; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); return r; }

struct Bar {
  unsigned* ctr; unsigned target;
  __device__ void sync() {
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ctr), "r"(1u) : "memory");
    target += gridDim.x;
    if (threadIdx.x == 0) while (ld_relaxed(ctr) < target) {}
    __syncthreads();
  }
};

__global__ void k_barrier(unsigned* ctr, int iters, long long* out) {
  Bar b{ctr, 0};
  b.sync();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) b.sync();
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = (clock64() - t0) / iters;
}

// MODE 0: ld.cg, 1: ld.ca, 2: ld.nc ; DEPTH loads in flight per thread; bytes = buffer size read by every CTA
template <int MODE, int DEPTH>
__global__ void k_bcast(unsigned* ctr, uint4* buf, int n16, int iters, long long* out, int write) {
  Bar b{ctr, 0};
  b.sync();
  long long t_read = 0;
  unsigned acc = 0;
  for (int it = 0; it < iters; ++it) {
    if (write) {   // every CTA dirties its slice so the data really comes from other SMs
      for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x) buf[i] = make_uint4(it, i, acc, 1);
    }
    b.sync();
    long long t0 = clock64();
    for (int base = threadIdx.x; base < n16; base += blockDim.x * DEPTH) {
      uint4 v[DEPTH];
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) {
        const int i = base + d * blockDim.x;
        if (i < n16) v[d] = MODE == 0 ? ld_cg(buf + i) : (MODE == 1 ? ld_ca(buf + i) : ld_nc(buf + i));
        else v[d] = make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) acc += v[d].x ^ v[d].y ^ v[d].z ^ v[d].w;
    }
    __syncthreads();
    t_read += clock64() - t0;
    b.sync();
  }
  if (threadIdx.x == 0) out[1 + blockIdx.x] = acc;
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t_read / iters;
}

template <int MODE, int DEPTH>
int run_bcast(const char* name, unsigned* ctr, uint4* buf, int n16, long long* out, int sms, int threads, int write) {
  CK(cudaMemset(ctr, 0, 4));
  void* args[] = {&ctr, &buf, &n16, nullptr, &out, &write};
  int iters = 200;
  args[3] = &iters;
  CK(cudaLaunchCooperativeKernel((void*)k_bcast<MODE, DEPTH>, dim3(sms), dim3(threads), args, 0, 0));
  CK(cudaDeviceSynchronize());
  long long cyc;
  CK(cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost));
  double us = cyc / 1965.0;
  printf("  %-34s %7.2f us  -> %6.1f GB/s per SM, %6.2f TB/s aggregate\n", name, us, n16 * 16.0 / us / 1e3, n16 * 16.0 * sms / us / 1e6);
  return 0;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  unsigned* ctr; long long* out; uint4* buf;
  CK(cudaMalloc(&ctr, 4)); CK(cudaMalloc(&out, 8 * 1024)); CK(cudaMalloc(&buf, 8 << 20));
  for (int threads : {128, 384}) {
    CK(cudaMemset(ctr, 0, 4));
    int iters = 2000;
    void* args[] = {&ctr, &iters, &out};
    CK(cudaLaunchCooperativeKernel((void*)k_barrier, dim3(sms), dim3(threads), args, 0, 0));
    CK(cudaDeviceSynchronize());
    long long cyc; CK(cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost));
    printf("grid barrier, %d CTAs x %d threads: %lld cycles = %.2f us\n", sms, threads, cyc, cyc / 1965.0);
  }
  for (int kb : {24, 98, 196, 393}) {
    const int n16 = kb * 1024 / 16;
    printf("broadcast read of %d KB by every CTA (384 threads):\n", kb);
    run_bcast<0, 4>("ld.cg depth 4, freshly written", ctr, buf, n16, out, sms, 384, 1);
    run_bcast<0, 8>("ld.cg depth 8, freshly written", ctr, buf, n16, out, sms, 384, 1);
    run_bcast<0, 16>("ld.cg depth 16, freshly written", ctr, buf, n16, out, sms, 384, 1);
    run_bcast<0, 16>("ld.cg depth 16, not rewritten", ctr, buf, n16, out, sms, 384, 0);
    run_bcast<1, 16>("ld.ca depth 16, not rewritten", ctr, buf, n16, out, sms, 384, 0);
    run_bcast<2, 16>("ld.nc depth 16, not rewritten", ctr, buf, n16, out, sms, 384, 0);
  }
  return 0;
}

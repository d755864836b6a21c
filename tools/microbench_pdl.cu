// Floor of a dependent stage chain: N kernels per "step" launched with programmatic dependent launch inside a CUDA graph,
// each doing (a) nothing, (b) one dependent L2 round trip (read what the previous kernel wrote, write for the next),
// (c) the same as 16-CTA clusters.  Prints us per kernel.   nvcc -arch=sm_100a -O3 -o /tmp/mb_pdl tools/microbench_pdl.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void k_empty(float* buf, int n) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
__global__ void k_rt(float* buf, int n) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float v;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(buf + (i * 37 + 11) % n) : "memory");
  buf[n + i] = v + 1.f;
  float w;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(w) : "l"(buf + n + (i * 53 + 7) % n) : "memory");
  buf[i] = w;
}
__global__ void __cluster_dims__(16, 1, 1) k_rt_cluster(float* buf, int n) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float v;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(buf + (i * 37 + 11) % n) : "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  buf[i] = v + 1.f;
}

template <typename K>
int run(const char* name, K kern, int grid, int block, int per_step, bool pdl, float* buf, int n, cudaStream_t s) {
  cudaGraph_t g; cudaGraphExec_t ex;
  CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  for (int i = 0; i < per_step; ++i) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, kern, buf, n));
  }
  CK(cudaStreamEndCapture(s, &g));
  CK(cudaGraphInstantiate(&ex, g, 0));
  for (int i = 0; i < 5; ++i) CK(cudaGraphLaunch(ex, s));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0, s));
  const int reps = 50;
  for (int i = 0; i < reps; ++i) CK(cudaGraphLaunch(ex, s));
  CK(cudaEventRecord(e1, s)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  printf("%-28s grid %4d x %3d, pdl %d: %.2f us per kernel\n", name, grid, block, pdl ? 1 : 0, ms * 1e3 / (reps * per_step));
  cudaGraphExecDestroy(ex); cudaGraphDestroy(g);
  return 0;
}

int main() {
  cudaStream_t s; CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  const int n = 1 << 20;
  float* buf; CK(cudaMalloc(&buf, 2 * n * sizeof(float))); CK(cudaMemset(buf, 0, 2 * n * sizeof(float)));
  CK(cudaFuncSetAttribute(k_rt_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  const int per_step = 26 * 13;
  for (int pdl = 1; pdl >= 0; --pdl) {
    if (run("empty", k_empty, 144, 128, per_step, pdl, buf, n, s)) return 1;
    if (run("empty", k_empty, 144, 512, per_step, pdl, buf, n, s)) return 1;
    if (run("empty", k_empty, 32, 64, per_step, pdl, buf, n, s)) return 1;
    if (run("2 dependent L2 round trips", k_rt, 144, 512, per_step, pdl, buf, n, s)) return 1;
    if (run("2 dependent L2 round trips", k_rt, 32, 64, per_step, pdl, buf, n, s)) return 1;
    if (run("cluster16 load+barrier+store", k_rt_cluster, 64, 256, per_step, pdl, buf, n, s)) return 1;
  }
  return 0;
}

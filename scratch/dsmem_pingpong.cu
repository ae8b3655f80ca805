// Round-trip latency of remote mbarrier arrives inside a CTA pair (cluster of 2): rank 0 arrives on rank 1's barrier,
// rank 1 waits and arrives on rank 0's barrier, rank 0 waits.  Prints clk per round trip (2 one-way hops).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scratch/dsmem_pingpong scratch/dsmem_pingpong.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__global__ void __cluster_dims__(2, 1, 1) k(int iters, long long* out, int local_only) {
  __shared__ uint64_t bar;
  uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (threadIdx.x == 0) {
    uint32_t peer;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peer) : "r"(smem_u32(&bar)), "r"(rank ^ 1));
    long long t0 = clock64();
    if (local_only) {               // baseline: arrive on the own barrier and wait for it
      for (int i = 0; i < iters; ++i) {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");
        wait(&bar, i & 1);
      }
    } else if (rank == 0) {
      for (int i = 0; i < iters; ++i) {
        asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(peer) : "memory");
        wait(&bar, i & 1);
      }
    } else {
      for (int i = 0; i < iters; ++i) {
        wait(&bar, i & 1);
        asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(peer) : "memory");
      }
    }
    long long t1 = clock64();
    if (rank == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
int main() {
  long long* d; cudaMalloc(&d, 8);
  const int iters = 10000;
  for (int local = 0; local < 2; ++local)
    for (int grid : {2, 148}) {
      k<<<grid, 32>>>(iters, d, local);
      cudaError_t e = cudaDeviceSynchronize();
      long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
      printf("%s grid=%3d: %.0f clk per round (%s)\n", local ? "local arrive+wait " : "remote ping-pong  ", grid, (double)c / iters, cudaGetErrorString(e));
    }
  return 0;
}

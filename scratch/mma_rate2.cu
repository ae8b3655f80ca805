// Microbenchmark: cycles per tcgen05.mma.cta_group::2 (kind::tf32, M = 256), SS and TS, N = 128 / 256.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scratch/mma_rate2 scratch/mma_rate2.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
// MODE 0: SS, 1: TS ; PASSES: 1 = plain, 3 = the 3xTF32 issue pattern (lo*hi, hi*lo, hi*hi on the same accumulator)
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k(int n_mma, int N, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) {
    uint32_t h = (i + 1) * 2654435761u + blockIdx.x * 40503u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    float f = ((h >> 8) * (1.0f / 8388608.0f)) - 1.0f;
    ((uint32_t*)smem)[i] = __float_as_uint(f);
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tptr;
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    const uint32_t sa = smem_u32(smem), sb = sa + 16384;
    long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      const uint32_t koff = (i & 3) * 32;
      const uint64_t da = make_desc(sa + koff, 16, 1024, 2), db = make_desc(sb + koff, 16, 1024, 2);
      const uint32_t acc = i > 0;
      if (MODE == 0)
        asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;}" ::"r"(tb), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
      else
        asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;}" ::"r"(tb), "r"(tb + 256 + (i & 3) * 8), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
    long long t1 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; }
  }
  {  // both CTAs wait for the multicast commit
    uint32_t done = 0;
    if (threadIdx.x == 0)
      while (!done)
        asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p;}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = clock64();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}
// LOOP: per iteration 12 MMAs (N=128, the 3xTF32 pattern) [+ 2 multicast commits] [+ 2 try_waits on completed barriers]
template <int COMMITS, int WAITS, int POLL>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(448, 1) kloop(int iters, long long* out, int vary) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bars[8];
  __shared__ uint32_t tptr;
  uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  for (int i = threadIdx.x; i < 200000 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3f000000u + i;
  if (threadIdx.x == 0) {
    for (int b = 0; b < 8; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[b])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tptr;
  if (threadIdx.x == 0 && rank == 0) {
    const int N = 128;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    const uint32_t sb = smem_u32(smem) + 16384;
    // pre-complete phase 0 of bars[4..5] so that waits on parity 0 succeed immediately, forever (never re-armed)
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[4])) : "memory");
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[5])) : "memory");
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (WAITS) {
        for (int b = 4; b < 6; ++b) {
          uint32_t done = 0;
          while (!done)
            asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p;}" : "=r"(done) : "r"(smem_u32(&bars[b])) : "memory");
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t sbv = sb + ((vary & 1) ? (uint32_t)(it % 6) * 32768u : 0u);
        const uint32_t dt = tb + ((vary & 2) ? (uint32_t)((it >> 3) & 1) * 128u : 0u);
        const uint64_t db = make_desc(sbv + k * 32, 16, 1024, 2), dbl = make_desc(sbv + 8192 + k * 32, 16, 1024, 2);
        const uint32_t a_hi = tb + 256 + (it & 3) * 64 + k * 8, a_lo = a_hi + 32;
        const uint32_t acc = ((vary & 4) ? ((it & 7) > 0 || k > 0) : (it > 0 || k > 0));
        asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;}" ::"r"(dt), "r"(a_lo), "l"(db), "r"(idesc), "r"(acc) : "memory");
        asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;}" ::"r"(dt), "r"(a_hi), "l"(dbl), "r"(idesc), "r"(1u) : "memory");
        asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;}" ::"r"(dt), "r"(a_hi), "l"(db), "r"(idesc), "r"(1u) : "memory");
      }
      for (int c = 0; c < COMMITS; ++c)
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bars[c])), "h"((uint16_t)3) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bars[7])), "h"((uint16_t)3) : "memory");
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  // everybody else polls the final barrier, like the waiting roles of the real kernel do
  if (threadIdx.x == 0 || (POLL == 1 && threadIdx.x >= 64) || (POLL == 2 && threadIdx.x >= 64 && (threadIdx.x & 31) == 0)) {
    uint32_t done = 0;
    while (!done)
      asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p;}" : "=r"(done) : "r"(smem_u32(&bars[7])) : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}
template <int COMMITS, int WAITS, int POLL>
void runloop(int grid, int vary) {
  long long* d; cudaMalloc(&d, 16);
  const int iters = 4096;
  cudaFuncSetAttribute(kloop<COMMITS, WAITS, POLL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210000);
  kloop<COMMITS, WAITS, POLL><<<grid, 448, 210000>>>(iters, d, vary);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  kloop<COMMITS, WAITS, POLL><<<grid, 448, 210000>>>(iters, d, vary);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
  printf("loop commits=%d waits=%d poll=%d vary=%d grid=%3d: %.0f ns per iteration of 12 MMAs (ideal 420), issue-side %.0f clk/iter (%s)\n", COMMITS, WAITS, POLL, vary, grid,
         ms * 1e6 / iters, (double)c / iters, cudaGetErrorString(e));
  cudaFree(d);
}
template <int MODE>
void run(const char* name, int N, int grid) {
  long long* d; cudaMalloc(&d, 16);
  const int n = 32768;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
  k<MODE><<<grid, 128, 60000>>>(n, N, d);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<grid, 128, 60000>>>(n, N, d);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("%-8s M=256 N=%3d grid=%3d: kernel %.3f ms -> %6.1f ns per MMA, %.0f TFLOP/s chip (%s)\n", name, N, grid, ms,
         ms * 1e6 / n, 2.0 * 256 * N * 8.0 * n * (grid / 2) / (ms * 1e-3) / 1e12, cudaGetErrorString(e));
  cudaFree(d);
}
int main() {
  for (int v : {0, 1, 2, 4, 7}) runloop<2, 1, 0>(148, v);
  return 0;
  for (int grid : {2, 148}) {
    run<0>("pair SS", 256, grid); run<1>("pair TS", 256, grid); run<0>("pair SS", 128, grid); run<1>("pair TS", 128, grid);
    run<1>("pair TS", 64, grid);
  }
  return 0;
}

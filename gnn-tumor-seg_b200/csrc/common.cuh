// common.cuh — shared helpers for libgts.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include "../../include/gts.h"

namespace gts {

void set_error(const char* fmt, ...);

#define GTS_CHECK_ARG(cond, ...)                \
  do {                                          \
    if (!(cond)) {                              \
      ::gts::set_error(__VA_ARGS__);            \
      return GTS_ERR_INVALID;                   \
    }                                           \
  } while (0)

#define GTS_CUDA(expr)                                                              \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      ::gts::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,         \
                       cudaGetErrorString(_e));                                     \
      return GTS_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

#define GTS_LAUNCH_CHECK() GTS_CUDA(cudaGetLastError())

inline cudaStream_t as_stream(gts_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();   // cached SM count of the current device

// out[c, r] = in[r, c] for up to kMaxJobs small matrices in one launch (gemm_simt.cu)
struct TransposeJob { const float* in; float* out; int64_t ldin, ldout; int32_t rows, cols; };
struct TransposeBatch {
  static constexpr int kMaxJobs = 96;
  int n = 0;
  TransposeJob job[kMaxJobs];
};
int launch_transpose_batch(const TransposeBatch& b, cudaStream_t st);

// Deferred split-K reductions (gemm_simt.cu): while a batch is installed on the calling thread, every reduction that
// qualifies for the wide kernel is queued instead of launched, and launch_splitk_reduce_batch runs all of them in ONE
// launch (same kernel body and summation order as the immediate launches: identical bits).  The caller owns the
// partial-product buffers until that launch — every deferred GEMM needs its own workspace slice.
struct ReduceJob {
  const float4* partial; float4* C; float4* tail;
  int64_t stride4, main4, total4;
  int32_t splits, blk0;              // blk0: first block of this job inside the batched grid
};
struct ReduceBatch {
  static constexpr int kMaxJobs = 48;
  int n = 0, blocks = 0;
  ReduceJob job[kMaxJobs];
};
void splitk_defer_set(ReduceBatch* batch);      // nullptr: launch immediately (the default)
ReduceBatch* splitk_defer_target();
int launch_splitk_reduce_batch(ReduceBatch& batch, cudaStream_t st);   // launches and empties the batch

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) { return (a + b - 1) / b; }

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// 128-bit read-only global load that does not allocate in L1 (streaming rows).
__device__ __forceinline__ float4 ldg_nc_na(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
// 128-bit read-only global load through L1 (gathered rows that neighbouring
// warps re-read).
__device__ __forceinline__ float4 ldg_nc(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg_na(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void stg_na(int4* p, const int4& v) {
  asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

}  // namespace gts

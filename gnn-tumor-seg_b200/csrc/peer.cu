// peer.cu — data-parallel gradient exchange over NVLink peer memory, fused with AdamW.
// The reference trains on one device (model/gnn_model.py:23,41-47); whole graphs per rank is the natural shard, and
// the only cross-rank step is the sum of the 5 MB gradient arena.  Instead of NCCL launches between CUDA-graph
// segments, every rank publishes its arena into an IPC-mapped staging buffer and ONE kernel per rank reads all
// ranks' buffers over NVLink, sums them in rank order and applies the optimiser — both launches sit inside the
// captured step.  Protocol and layout: include/gts.h ("Data-parallel gradient exchange").
#include "common.cuh"

namespace gts {

constexpr int kHeaderBytes = 256;
// header words (uint32): [0, GTS_MAX_PEERS) = flag of rank s (last epoch s has published), then local-only words
constexpr int kEpoch = 32, kError = 33, kCtrPublish = 34, kCtrReduce = 35;
constexpr unsigned long long kSpinLimitNs = 10ull * 1000ull * 1000ull * 1000ull;

struct PeerView {
  int32_t rank, world;
  int64_t n4;
  char* base[GTS_MAX_PEERS];
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_volatile(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// peer rows change from step to step: never serve them from a stale L1 line
__device__ __forceinline__ float4 ld_peer(const float4* p) { return __ldcv(p); }

__device__ __forceinline__ float4* stage_of(char* base, uint32_t epoch, int64_t n4) {
  return reinterpret_cast<float4*>(base + kHeaderBytes) + (int64_t)(epoch & 1u) * n4;
}

// copy the arena into staging buffer (epoch & 1); the last CTA to finish raises this rank's flag at every peer
__global__ void __launch_bounds__(256) peer_publish_kernel(PeerView c, const float4* __restrict__ src) {
  uint32_t* hdr = reinterpret_cast<uint32_t*>(c.base[c.rank]);
  const uint32_t e = ld_volatile(hdr + kEpoch) + 1u;      // changes only at the end of the reduce kernel
  float4* dst = stage_of(c.base[c.rank], e, c.n4);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < c.n4; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = src[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(hdr + kCtrPublish, 1u);
    if (prev == gridDim.x - 1) {
      hdr[kCtrPublish] = 0u;
      __threadfence_system();
      for (int r = 0; r < c.world; ++r) st_release_sys(reinterpret_cast<uint32_t*>(c.base[r]) + c.rank, e);
    }
  }
}

// wait for every rank's flag, sum the staged arenas in rank order, write the sums back, AdamW on sum / denominator
__global__ void __launch_bounds__(256) peer_allreduce_adamw_kernel(PeerView c, float4* __restrict__ grads, int64_t np4,
                                                                    float4* __restrict__ p, float4* __restrict__ m,
                                                                    float4* __restrict__ v, float* __restrict__ hyper,
                                                                    int64_t denom_index, int apply) {
  __shared__ float s_inv_denom;
  uint32_t* hdr = reinterpret_cast<uint32_t*>(c.base[c.rank]);
  const uint32_t e = ld_volatile(hdr + kEpoch) + 1u;
  if ((int)threadIdx.x < c.world) {
    const unsigned long long t0 = global_ns();
    while ((int32_t)(ld_acquire_sys(hdr + threadIdx.x) - e) < 0) {
      if (global_ns() - t0 > kSpinLimitNs) { atomicExch(hdr + kError, 1u); break; }
      __nanosleep(64);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float d = 1.f;
    if (denom_index >= 0) {
      d = 0.f;
      for (int r = 0; r < c.world; ++r)
        d += __ldcv(reinterpret_cast<const float*>(stage_of(c.base[r], e, c.n4)) + denom_index);
    }
    s_inv_denom = 1.f / d;
  }
  __syncthreads();
  float lr = 0.f, b1 = 0.f, b2 = 0.f, eps = 0.f, wd = 0.f, bc1 = 1.f, bc2_sqrt = 1.f, stepf = 0.f;
  if (apply) {
    lr = hyper[0]; b1 = hyper[1]; b2 = hyper[2]; eps = hyper[3]; wd = hyper[4];
    stepf = hyper[5] + 1.f;                               // written back by the last CTA
    bc1 = (float)(1.0 - pow((double)b1, (double)stepf));
    bc2_sqrt = (float)sqrt(1.0 - pow((double)b2, (double)stepf));
  }
  const float gs = s_inv_denom;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < c.n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r0 = 0; r0 < c.world; r0 += 8) {             // eight peers' loads in flight, summed in rank order
      float4 x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (r0 + j < c.world) x[j] = ld_peer(stage_of(c.base[r0 + j], e, c.n4) + i);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (r0 + j < c.world) {
          if (r0 + j == 0) acc = x[j];
          else { acc.x += x[j].x; acc.y += x[j].y; acc.z += x[j].z; acc.w += x[j].w; }
        }
    }
    grads[i] = acc;
    if (apply && i < np4) {
      float4 pi = p[i], mi = m[i], vi = v[i];
      const float g4[4] = {acc.x * gs, acc.y * gs, acc.z * gs, acc.w * gs};
      float* pp = reinterpret_cast<float*>(&pi);
      float* mm = reinterpret_cast<float*>(&mi);
      float* vv = reinterpret_cast<float*>(&vi);
#pragma unroll
      for (int k = 0; k < 4; ++k) {                       // same arithmetic as adamw_dev_kernel (loss_optim.cu)
        const float grad = g4[k];
        float param = pp[k] * (1.f - lr * wd);
        const float mk = b1 * mm[k] + (1.f - b1) * grad;
        const float vk = b2 * vv[k] + (1.f - b2) * grad * grad;
        mm[k] = mk;
        vv[k] = vk;
        const float denom = sqrtf(vk) / bc2_sqrt + eps;
        param -= (lr / bc1) * (mk / denom);
        pp[k] = param;
      }
      p[i] = pi; m[i] = mi; v[i] = vi;
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(hdr + kCtrReduce, 1u);
    if (prev == gridDim.x - 1) {                          // every CTA has read the epoch / hyper[5]: advance them
      hdr[kCtrReduce] = 0u;
      hdr[kEpoch] = e;
      if (apply) hyper[5] = stepf;
    }
  }
}

static int view_of(const gts_peer_comm* c, PeerView& v, const char* who) {
  GTS_CHECK_ARG(c != nullptr, "%s: null comm", who);
  GTS_CHECK_ARG(c->world >= 1 && c->world <= GTS_MAX_PEERS && c->rank >= 0 && c->rank < c->world,
                "%s: rank %d of %d (at most %d peers)", who, c->rank, c->world, GTS_MAX_PEERS);
  GTS_CHECK_ARG(c->n >= 4 && c->n % 4 == 0, "%s: n must be a positive multiple of 4", who);
  v.rank = c->rank; v.world = c->world; v.n4 = c->n / 4;
  for (int r = 0; r < GTS_MAX_PEERS; ++r) v.base[r] = nullptr;
  for (int r = 0; r < c->world; ++r) {
    GTS_CHECK_ARG(c->base[r] != nullptr, "%s: base[%d] is null", who, r);
    v.base[r] = reinterpret_cast<char*>(c->base[r]);
  }
  return GTS_OK;
}

// one resident wave: every CTA spins on the flags at its start
static inline int peer_grid(int64_t n4) {
  int64_t b = ceil_div<int64_t>(n4, 256);
  const int64_t cap = (int64_t)sm_count() * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace gts

using namespace gts;

extern "C" {

size_t gts_peer_buffer_bytes(int64_t n) {
  if (n < 0) return 0;
  const size_t n4 = ((size_t)n + 3) / 4;
  return (size_t)kHeaderBytes + 2 * n4 * 16;
}

int gts_peer_alloc(size_t bytes, void** dptr, unsigned char* handle) {
  GTS_CHECK_ARG(dptr && handle, "gts_peer_alloc: null pointer");
  GTS_CHECK_ARG(bytes >= (size_t)kHeaderBytes, "gts_peer_alloc: buffer smaller than its header");
  static_assert(sizeof(cudaIpcMemHandle_t) == GTS_PEER_HANDLE_BYTES, "IPC handle size");
  void* p = nullptr;
  GTS_CUDA(cudaMalloc(&p, bytes));
  cudaError_t e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    set_error("gts_peer_alloc: %s", cudaGetErrorString(e));
    cudaFree(p);
    return GTS_ERR_CUDA;
  }
  for (int i = 0; i < GTS_PEER_HANDLE_BYTES; ++i) handle[i] = reinterpret_cast<const unsigned char*>(&h)[i];
  *dptr = p;
  return GTS_OK;
}

int gts_peer_free(void* dptr) {
  if (dptr) GTS_CUDA(cudaFree(dptr));
  return GTS_OK;
}

int gts_peer_open(const unsigned char* handle, void** dptr) {
  GTS_CHECK_ARG(dptr && handle, "gts_peer_open: null pointer");
  cudaIpcMemHandle_t h;
  for (int i = 0; i < GTS_PEER_HANDLE_BYTES; ++i) reinterpret_cast<unsigned char*>(&h)[i] = handle[i];
  void* p = nullptr;
  GTS_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *dptr = p;
  return GTS_OK;
}

int gts_peer_close(void* dptr) {
  if (dptr) GTS_CUDA(cudaIpcCloseMemHandle(dptr));
  return GTS_OK;
}

int gts_peer_publish(const gts_peer_comm* comm, const float* src, gts_stream_t stream) {
  PeerView v;
  int rc = view_of(comm, v, "gts_peer_publish");
  if (rc != GTS_OK) return rc;
  GTS_CHECK_ARG(src && (reinterpret_cast<uintptr_t>(src) & 15) == 0, "gts_peer_publish: src must be 16-byte aligned");
  peer_publish_kernel<<<peer_grid(v.n4), 256, 0, as_stream(stream)>>>(v, reinterpret_cast<const float4*>(src));
  GTS_LAUNCH_CHECK();
  return GTS_OK;
}

int gts_peer_allreduce_adamw(const gts_peer_comm* comm, float* grads, int64_t n_params, float* param, float* exp_avg,
                             float* exp_avg_sq, float* hyper, int64_t denom_index, int32_t apply, gts_stream_t stream) {
  PeerView v;
  int rc = view_of(comm, v, "gts_peer_allreduce_adamw");
  if (rc != GTS_OK) return rc;
  GTS_CHECK_ARG(grads && (reinterpret_cast<uintptr_t>(grads) & 15) == 0, "gts_peer_allreduce_adamw: grads must be 16-byte aligned");
  GTS_CHECK_ARG(denom_index < comm->n, "gts_peer_allreduce_adamw: denom_index outside the exchanged vector");
  if (apply) {
    GTS_CHECK_ARG(param && exp_avg && exp_avg_sq && hyper, "gts_peer_allreduce_adamw: null optimiser state");
    GTS_CHECK_ARG(n_params >= 0 && n_params <= comm->n && n_params % 4 == 0,
                  "gts_peer_allreduce_adamw: n_params must be a multiple of 4 within the exchanged vector");
    GTS_CHECK_ARG(((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(exp_avg) |
                    reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0,
                  "gts_peer_allreduce_adamw: optimiser arenas must be 16-byte aligned");
  }
  peer_allreduce_adamw_kernel<<<peer_grid(v.n4), 256, 0, as_stream(stream)>>>(
      v, reinterpret_cast<float4*>(grads), apply ? n_params / 4 : 0, reinterpret_cast<float4*>(param),
      reinterpret_cast<float4*>(exp_avg), reinterpret_cast<float4*>(exp_avg_sq), hyper, denom_index, apply ? 1 : 0);
  GTS_LAUNCH_CHECK();
  return GTS_OK;
}

int gts_peer_status(const gts_peer_comm* comm, uint32_t* epoch, uint32_t* error) {
  PeerView v;
  int rc = view_of(comm, v, "gts_peer_status");
  if (rc != GTS_OK) return rc;
  uint32_t w[4] = {0, 0, 0, 0};
  GTS_CUDA(cudaMemcpy(w, v.base[v.rank] + 4 * kEpoch, sizeof(w), cudaMemcpyDeviceToHost));
  if (epoch) *epoch = w[0];
  if (error) *error = w[1];
  return GTS_OK;
}

}  // extern "C"

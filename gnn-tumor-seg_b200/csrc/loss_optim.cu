// loss_optim.cu — K8 weighted-mean cross entropy (reference
// model/gnn_model.py:30,42: torch.nn.CrossEntropyLoss(weight=w)) and the AdamW
// step on a flat arena (model/gnn_model.py:28,46).
#include "common.cuh"

namespace gts {

constexpr int kMaxClasses = 32;

__global__ void ce_weighted_kernel(const float* __restrict__ logits, int64_t ld, const int64_t* __restrict__ labels,
                                   const float* __restrict__ class_w, int32_t N, int32_t C,
                                   float* __restrict__ sums, float* __restrict__ dlogits, int64_t ldd) {
  float loss_part = 0.f, w_part = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    float z[kMaxClasses];
    float mx = -INFINITY;
#pragma unroll 4
    for (int c = 0; c < C; ++c) { z[c] = logits[i * ld + c]; mx = fmaxf(mx, z[c]); }
    float se = 0.f;
#pragma unroll 4
    for (int c = 0; c < C; ++c) se += expf(z[c] - mx);
    const float lse = mx + logf(se);
    const int64_t y = labels[i];
    const bool valid = (y >= 0 && y < C);
    const float w = valid ? class_w[y] : 0.f;
    if (valid) { loss_part += w * (lse - z[y]); w_part += w; }
    else if (y != -100) loss_part = __int_as_float(0x7fc00000);   // torch raises on such a label: poison the loss
    if (dlogits) {
#pragma unroll 4
      for (int c = 0; c < C; ++c) {
        const float p = expf(z[c] - lse);
        dlogits[i * ldd + c] = w * (p - ((int64_t)c == y ? 1.f : 0.f));
      }
    }
  }
  // block reduction -> one atomic pair per block
  __shared__ float s_l[32], s_w[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    loss_part += __shfl_xor_sync(0xffffffffu, loss_part, o);
    w_part += __shfl_xor_sync(0xffffffffu, w_part, o);
  }
  if (lane == 0) { s_l[warp] = loss_part; s_w[warp] = w_part; }
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    float l = lane < nw ? s_l[lane] : 0.f, w = lane < nw ? s_w[lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      l += __shfl_xor_sync(0xffffffffu, l, o);
      w += __shfl_xor_sync(0xffffffffu, w, o);
    }
    if (lane == 0) { atomicAdd(&sums[0], l); atomicAdd(&sums[1], w); }
  }
}

__global__ void scale_by_inv_kernel(float* __restrict__ x, int64_t n, float alpha, const float* __restrict__ denom) {
  const float s = alpha / *denom;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] *= s;
}

__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps, float wd,
                             float bc1, float bc2_sqrt, float gscale, const float* __restrict__ gdenom) {
  const float gs = gdenom ? gscale / *gdenom : gscale;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float grad = g[i] * gs;
    float param = p[i] * (1.f - lr * wd);
    const float mi = b1 * m[i] + (1.f - b1) * grad;
    const float vi = b2 * v[i] + (1.f - b2) * grad * grad;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    param -= (lr / bc1) * (mi / denom);
    p[i] = param;
  }
}

// hyper = {lr, beta1, beta2, eps, weight_decay, step}: bias corrections from the device-side step count
__global__ void adamw_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, int64_t n, const float* __restrict__ hyper, float gscale,
                                 const float* __restrict__ gdenom) {
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4];
  const double step = (double)hyper[5];
  const float bc1 = (float)(1.0 - pow((double)b1, step));
  const float bc2_sqrt = (float)sqrt(1.0 - pow((double)b2, step));
  const float gs = gdenom ? gscale / *gdenom : gscale;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float grad = g[i] * gs;
    float param = p[i] * (1.f - lr * wd);
    const float mi = b1 * m[i] + (1.f - b1) * grad;
    const float vi = b2 * v[i] + (1.f - b2) * grad * grad;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    param -= (lr / bc1) * (mi / denom);
    p[i] = param;
  }
}
__global__ void bump_step_kernel(float* hyper) { hyper[5] += 1.f; }

static inline int ew_grid(int64_t n, int threads) {
  int64_t b = ceil_div<int64_t>(n, threads);
  const int64_t cap = (int64_t)sm_count() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace gts

using namespace gts;

extern "C" {

int gts_ce_weighted(const float* logits, int64_t ld, const int64_t* labels, const float* class_w,
                    int32_t n_nodes, int32_t n_classes, float* sums, float* dlogits, int64_t ldd,
                    gts_stream_t stream) {
  GTS_CHECK_ARG(n_nodes >= 0, "gts_ce_weighted: negative size");
  GTS_CHECK_ARG(n_classes >= 1 && n_classes <= kMaxClasses, "gts_ce_weighted: n_classes must be in [1,%d]", kMaxClasses);
  if (n_nodes == 0) return GTS_OK;
  GTS_CHECK_ARG(logits && labels && class_w && sums, "gts_ce_weighted: null pointer");
  ce_weighted_kernel<<<ew_grid(n_nodes, 256), 256, 0, as_stream(stream)>>>(logits, ld, labels, class_w, n_nodes, n_classes,
                                                                          sums, dlogits, ldd);
  GTS_LAUNCH_CHECK();
  return GTS_OK;
}

int gts_scale_by_inv(float* x, int64_t n, float alpha, const float* denom, gts_stream_t stream) {
  GTS_CHECK_ARG(n >= 0, "gts_scale_by_inv: negative size");
  if (n == 0) return GTS_OK;
  GTS_CHECK_ARG(x && denom, "gts_scale_by_inv: null pointer");
  scale_by_inv_kernel<<<ew_grid(n, 256), 256, 0, as_stream(stream)>>>(x, n, alpha, denom);
  GTS_LAUNCH_CHECK();
  return GTS_OK;
}

int gts_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                   float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step,
                   float grad_scale, const float* grad_denom, gts_stream_t stream) {
  GTS_CHECK_ARG(n >= 0 && step >= 1, "gts_adamw_step: n >= 0 and step >= 1 required");
  if (n == 0) return GTS_OK;
  GTS_CHECK_ARG(param && grad && exp_avg && exp_avg_sq, "gts_adamw_step: null pointer");
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  adamw_kernel<<<ew_grid(n, 256), 256, 0, as_stream(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                               weight_decay, (float)bc1, (float)sqrt(bc2), grad_scale, grad_denom);
  GTS_LAUNCH_CHECK();
  return GTS_OK;
}

int gts_adamw_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                       float* hyper, float grad_scale, const float* grad_denom, gts_stream_t stream) {
  GTS_CHECK_ARG(n >= 0, "gts_adamw_step_dev: negative size");
  GTS_CHECK_ARG(hyper, "gts_adamw_step_dev: null hyper-parameter block");
  bump_step_kernel<<<1, 1, 0, as_stream(stream)>>>(hyper);
  GTS_LAUNCH_CHECK();
  if (n == 0) return GTS_OK;
  GTS_CHECK_ARG(param && grad && exp_avg && exp_avg_sq, "gts_adamw_step_dev: null pointer");
  adamw_dev_kernel<<<ew_grid(n, 256), 256, 0, as_stream(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, hyper, grad_scale, grad_denom);
  GTS_LAUNCH_CHECK();
  return GTS_OK;
}

}  // extern "C"

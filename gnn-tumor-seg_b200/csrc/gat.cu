// gat.cu — K5: GATConv edge score + edge softmax + weighted aggregation fused
// into one kernel per layer over the in-edge CSR (forward), and a two-pass
// deterministic backward (by destination, then by source over the out-edge
// CSC).  Replaces DGL GATConv.forward's apply_edges(u_add_v) / leaky_relu /
// edge_softmax / update_all(u_mul_e, sum) chain and its autograd (invoked at
// reference model/networks.py:63,65).
//
// Work unit: one warp per (node, head).  Lanes own float4 chunks of the F-wide
// head row (F = 256 -> 2 chunks per lane); per-edge scalars live one edge per
// lane and are broadcast by shuffle.  No E x H x F intermediate is ever
// materialised; the only per-edge array is dt[E,H] in the backward.
// HBM-bound: algorithmic bytes per layer 4*(2*N*H*F + 6*N*H) + 4*(N+1+E).
#include "common.cuh"
#include <cstdlib>

namespace gts {

constexpr int kGatThreads = 256;
// (A contiguous-(node, head)-range-per-SM distribution of the edge kernels, which pays off for the seg-max kernel, was
// measured SLOWER here: 550 -> 678 us forward at 4 x 256 — only 8 nodes are in flight per SM, too few for L1 reuse.)
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float warp_max(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x = fmaxf(x, __shfl_xor_sync(kFull, x, o));
  return x;
}
__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(kFull, x, o);
  return x;
}
__device__ __forceinline__ float lrelu(float t, float slope) { return t > 0.f ? t : t * slope; }

// Row accessor: VEC float4 chunks per lane when F % 4 == 0 (VECTOR), else
// scalar columns lane, lane+32, ... (up to VEC*4 of them).
template <int VEC, bool VECTOR>
struct RowFrag {
  float v[VEC * 4];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < VEC * 4; ++i) v[i] = 0.f;
  }
  __device__ __forceinline__ void load(const float* row, int F, int lane) {
    if (VECTOR) {
#pragma unroll
      for (int c = 0; c < VEC; ++c) {
        const int chunk = lane + 32 * c;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (chunk * 4 < F) x = ldg_nc(reinterpret_cast<const float4*>(row) + chunk);
        v[4 * c] = x.x; v[4 * c + 1] = x.y; v[4 * c + 2] = x.z; v[4 * c + 3] = x.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < VEC * 4; ++i) {
        const int f = lane + 32 * i;
        v[i] = f < F ? __ldg(row + f) : 0.f;
      }
    }
  }
  __device__ __forceinline__ void store(float* row, int F, int lane) const {
    if (VECTOR) {
#pragma unroll
      for (int c = 0; c < VEC; ++c) {
        const int chunk = lane + 32 * c;
        if (chunk * 4 < F)
          *(reinterpret_cast<float4*>(row) + chunk) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < VEC * 4; ++i) {
        const int f = lane + 32 * i;
        if (f < F) row[f] = v[i];
      }
    }
  }
  __device__ __forceinline__ void fma(float a, const RowFrag& o) {
#pragma unroll
    for (int i = 0; i < VEC * 4; ++i) v[i] = fmaf(a, o.v[i], v[i]);
  }
  __device__ __forceinline__ float dot(const RowFrag& o) const {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VEC * 4; ++i) s = fmaf(v[i], o.v[i], s);
    return s;
  }
};

__global__ void gat_scores_kernel(const float* __restrict__ Z, int64_t ldz, const float* __restrict__ al,
                                  const float* __restrict__ ar, int64_t NH, int H, int F,
                                  float* __restrict__ el, float* __restrict__ er) {
  const int lane = threadIdx.x & 31;
  const int64_t w0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t w = w0; w < NH; w += nw) {
    const int64_t u = w / H;
    const int h = (int)(w - u * H);
    const float* row = Z + u * ldz + (int64_t)h * F;
    float sl = 0.f, sr = 0.f;
    for (int f = lane; f < F; f += 32) {
      const float z = row[f];
      sl = fmaf(z, al[h * F + f], sl);
      sr = fmaf(z, ar[h * F + f], sr);
    }
    sl = warp_sum(sl);
    sr = warp_sum(sr);
    if (lane == 0) { el[w] = sl; er[w] = sr; }
  }
}

template <int VEC, bool VECTOR, bool RANGE>
__global__ void __launch_bounds__(RANGE ? 1024 : kGatThreads, 1)
gat_fwd_kernel(const float* __restrict__ Z, int64_t ldz, const float* __restrict__ el, const float* __restrict__ er,
               const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices, int64_t NH, int H, int F,
               float slope, const float* __restrict__ res, int64_t ldres, const float* __restrict__ bias, int act,
               float* __restrict__ out, int64_t ldo, float* __restrict__ rowmax, float* __restrict__ rowsum,
               int32_t* __restrict__ err, int sweep_piece) {
  const int lane = threadIdx.x & 31;
  // RANGE: one 1024-thread CTA per SM owns a contiguous node range and walks it HEAD-MAJOR (all 32 warps on the same
  // head of 32 consecutive nodes): consecutive supervoxels share most of their neighbours, so the 1 KB head slices of
  // the neighbour rows are re-read from L1 (node-major order keeps 8 nodes x 4 heads x 15 rows in flight and thrashes it)
  // The id space is swept in `steps` rounds of gridDim.x nearly equal pieces (piece s * gridDim.x + blockIdx.x at
  // round s) instead of one long range per CTA: the neighbour rows a CTA shares with the CTAs next to it in id space
  // (the +-z supervoxels, one z-slab of ids away) are then touched by everybody at about the same time and come from
  // L2 instead of DRAM (see segmax.cu / profiles/r01_segmax_variants.md).  sweep_piece = target nodes per piece.
  const int64_t n_total = NH / H;
  const int64_t steps = (RANGE && sweep_piece > 0)
      ? max((int64_t)1, (n_total + (int64_t)gridDim.x * sweep_piece / 2) / ((int64_t)gridDim.x * sweep_piece)) : 1;
  const int64_t n_pieces = steps * gridDim.x;
  for (int64_t sweep = 0; sweep < steps; ++sweep) {
  const int64_t piece = sweep * gridDim.x + blockIdx.x;
  const int64_t r_beg = RANGE ? piece * n_total / n_pieces : 0;
  const int64_t r_cnt = RANGE ? (piece + 1) * n_total / n_pieces - r_beg : 0;
  const int64_t i_beg = RANGE ? (int64_t)(threadIdx.x >> 5) : ((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
  const int64_t i_end = RANGE ? r_cnt * H : NH;
  const int64_t i_step = RANGE ? (int64_t)(blockDim.x >> 5) : (((int64_t)gridDim.x * blockDim.x) >> 5);
  for (int64_t i = i_beg; i < i_end; i += i_step) {
    const int h = RANGE ? (int)(i / r_cnt) : (int)(i % H);
    const int64_t v = RANGE ? r_beg + (i - (int64_t)h * r_cnt) : i / H;
    const int64_t w = v * H + h;
    const int32_t beg = indptr[v], end = indptr[v + 1];
    const float erv = er[w];
    // pass A: softmax statistics over the row's edges (one edge per lane)
    float m = -INFINITY, l = 0.f;
    for (int32_t base = beg; base < end; base += 32) {
      const bool on = base + lane < end;
      float s = -INFINITY;
      if (on) s = lrelu(el[(int64_t)indices[base + lane] * H + h] + erv, slope);
      const float m_new = fmaxf(m, warp_max(s));
      const float part = warp_sum(on ? expf(s - m_new) : 0.f);
      l = l * expf(m - m_new) + part;
      m = m_new;
    }
    // pass B: weighted gather of neighbour rows
    RowFrag<VEC, VECTOR> acc;
    acc.zero();
    const float inv_l = end > beg ? 1.f / l : 0.f;
    for (int32_t base = beg; base < end; base += 32) {
      const bool on = base + lane < end;
      int32_t u_l = 0;
      float a_l = 0.f;
      if (on) {
        u_l = indices[base + lane];
        a_l = expf(lrelu(el[(int64_t)u_l * H + h] + erv, slope) - m) * inv_l;
      }
      const int cnt = min(32, end - base);
      for (int j = 0; j < cnt; ++j) {
        const int32_t u = __shfl_sync(kFull, u_l, j);
        const float a = __shfl_sync(kFull, a_l, j);
        RowFrag<VEC, VECTOR> z;
        z.load(Z + (int64_t)u * ldz + (int64_t)h * F, F, lane);
        acc.fma(a, z);
      }
    }
    if (end == beg && lane == 0) *err = 1;
    // epilogue: + residual + bias, activation
    if (res) {
      RowFrag<VEC, VECTOR> r;
      r.load(res + v * ldres + (int64_t)h * F, F, lane);
      acc.fma(1.f, r);
    }
    if (bias) {
      RowFrag<VEC, VECTOR> b;
      b.load(bias + (int64_t)h * F, F, lane);
      acc.fma(1.f, b);
    }
    if (act == 1) {
#pragma unroll
      for (int i = 0; i < VEC * 4; ++i) acc.v[i] = acc.v[i] > 0.f ? acc.v[i] : expm1f(acc.v[i]);
    }
    acc.store(out + v * ldo + (int64_t)h * F, F, lane);
    if (lane == 0) { rowmax[w] = m; rowsum[w] = l; }
  }
  }
}

__global__ void gat_act_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ out, int64_t n, int act,
                                   float* __restrict__ dpre) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float g = dout[i];
    if (act == 1) {
      const float o = out[i];
      dpre[i] = o > 0.f ? g : g * (o + 1.f);    // ELU'(x) = elu(x) + 1 for x <= 0
    } else {
      dpre[i] = g;
    }
  }
}

// Backward pass A, by destination.
template <int VEC, bool VECTOR, bool RANGE>
__global__ void __launch_bounds__(RANGE ? 1024 : kGatThreads, 1)
gat_bwd_dst_kernel(const float* __restrict__ Z, int64_t ldz, const float* __restrict__ el, const float* __restrict__ er,
                   const float* __restrict__ rowmax, const float* __restrict__ rowsum,
                   const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                   const float* __restrict__ dO, int64_t lddo, int64_t NH, int H, int F, float slope,
                   float* __restrict__ dt_edge, float* __restrict__ der, int sweep_piece) {
  const int lane = threadIdx.x & 31;
  // RANGE: one 1024-thread CTA per SM owns a contiguous node range and walks it HEAD-MAJOR (all 32 warps on the same
  // head of 32 consecutive nodes): consecutive supervoxels share most of their neighbours, so the 1 KB head slices of
  // the neighbour rows are re-read from L1 (node-major order keeps 8 nodes x 4 heads x 15 rows in flight and thrashes it)
  // The id space is swept in `steps` rounds of gridDim.x nearly equal pieces (piece s * gridDim.x + blockIdx.x at
  // round s) instead of one long range per CTA: the neighbour rows a CTA shares with the CTAs next to it in id space
  // (the +-z supervoxels, one z-slab of ids away) are then touched by everybody at about the same time and come from
  // L2 instead of DRAM (see segmax.cu / profiles/r01_segmax_variants.md).  sweep_piece = target nodes per piece.
  const int64_t n_total = NH / H;
  const int64_t steps = (RANGE && sweep_piece > 0)
      ? max((int64_t)1, (n_total + (int64_t)gridDim.x * sweep_piece / 2) / ((int64_t)gridDim.x * sweep_piece)) : 1;
  const int64_t n_pieces = steps * gridDim.x;
  for (int64_t sweep = 0; sweep < steps; ++sweep) {
  const int64_t piece = sweep * gridDim.x + blockIdx.x;
  const int64_t r_beg = RANGE ? piece * n_total / n_pieces : 0;
  const int64_t r_cnt = RANGE ? (piece + 1) * n_total / n_pieces - r_beg : 0;
  const int64_t i_beg = RANGE ? (int64_t)(threadIdx.x >> 5) : ((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
  const int64_t i_end = RANGE ? r_cnt * H : NH;
  const int64_t i_step = RANGE ? (int64_t)(blockDim.x >> 5) : (((int64_t)gridDim.x * blockDim.x) >> 5);
  for (int64_t i = i_beg; i < i_end; i += i_step) {
    const int h = RANGE ? (int)(i / r_cnt) : (int)(i % H);
    const int64_t v = RANGE ? r_beg + (i - (int64_t)h * r_cnt) : i / H;
    const int64_t w = v * H + h;
    const int32_t beg = indptr[v], end = indptr[v + 1];
    const float erv = er[w], m = rowmax[w];
    const float inv_l = end > beg ? 1.f / rowsum[w] : 0.f;
    RowFrag<VEC, VECTOR> g;
    g.load(dO + v * lddo + (int64_t)h * F, F, lane);
    // sweep 1: dalpha_e = <dO_v, Z_u>, delta = sum alpha_e dalpha_e
    float delta_part = 0.f;
    for (int32_t base = beg; base < end; base += 32) {
      const bool on = base + lane < end;
      const int32_t u_l = on ? indices[base + lane] : 0;
      const int cnt = min(32, end - base);
      float my_da = 0.f;
      for (int j = 0; j < cnt; ++j) {
        const int32_t u = __shfl_sync(kFull, u_l, j);
        RowFrag<VEC, VECTOR> z;
        z.load(Z + (int64_t)u * ldz + (int64_t)h * F, F, lane);
        const float d = warp_sum(g.dot(z));
        if (lane == j) my_da = d;
      }
      if (on) {
        const float alpha = expf(lrelu(el[(int64_t)u_l * H + h] + erv, slope) - m) * inv_l;
        delta_part = fmaf(alpha, my_da, delta_part);
        dt_edge[(int64_t)(base + lane) * H + h] = my_da;    // parked; finalised in sweep 2 by the same lane
      }
    }
    const float delta = warp_sum(delta_part);
    // sweep 2: dt_e = alpha_e (dalpha_e - delta) * lrelu'(t_e); der[v,h] = sum dt_e
    float der_part = 0.f;
    for (int32_t base = beg; base < end; base += 32) {
      if (base + lane < end) {
        const int32_t u = indices[base + lane];
        const float t = el[(int64_t)u * H + h] + erv;
        const float alpha = expf(lrelu(t, slope) - m) * inv_l;
        const int64_t idx = (int64_t)(base + lane) * H + h;
        const float ds = alpha * (dt_edge[idx] - delta);
        const float dt = t > 0.f ? ds : ds * slope;
        dt_edge[idx] = dt;
        der_part += dt;
      }
    }
    der_part = warp_sum(der_part);
    if (lane == 0) der[w] = der_part;
  }
  }
}

// Backward pass B, by source over the out-edge CSC.
template <int VEC, bool VECTOR, bool RANGE>
__global__ void __launch_bounds__(RANGE ? 1024 : kGatThreads, 1)
gat_bwd_src_kernel(const float* __restrict__ el, const float* __restrict__ er, const float* __restrict__ rowmax,
                   const float* __restrict__ rowsum, const int32_t* __restrict__ cptr, const int32_t* __restrict__ cidx,
                   const int32_t* __restrict__ csc2csr, const float* __restrict__ dO, int64_t lddo,
                   const float* __restrict__ dt_edge, const float* __restrict__ der, const float* __restrict__ al,
                   const float* __restrict__ ar, int64_t NH, int H, int F, float slope,
                   float* __restrict__ dZ, int64_t lddz, float* __restrict__ del, int sweep_piece) {
  const int lane = threadIdx.x & 31;
  // RANGE: one 1024-thread CTA per SM owns a contiguous node range and walks it HEAD-MAJOR (all 32 warps on the same
  // head of 32 consecutive nodes): consecutive supervoxels share most of their neighbours, so the 1 KB head slices of
  // the neighbour rows are re-read from L1 (node-major order keeps 8 nodes x 4 heads x 15 rows in flight and thrashes it)
  // The id space is swept in `steps` rounds of gridDim.x nearly equal pieces (piece s * gridDim.x + blockIdx.x at
  // round s) instead of one long range per CTA: the neighbour rows a CTA shares with the CTAs next to it in id space
  // (the +-z supervoxels, one z-slab of ids away) are then touched by everybody at about the same time and come from
  // L2 instead of DRAM (see segmax.cu / profiles/r01_segmax_variants.md).  sweep_piece = target nodes per piece.
  const int64_t n_total = NH / H;
  const int64_t steps = (RANGE && sweep_piece > 0)
      ? max((int64_t)1, (n_total + (int64_t)gridDim.x * sweep_piece / 2) / ((int64_t)gridDim.x * sweep_piece)) : 1;
  const int64_t n_pieces = steps * gridDim.x;
  for (int64_t sweep = 0; sweep < steps; ++sweep) {
  const int64_t piece = sweep * gridDim.x + blockIdx.x;
  const int64_t r_beg = RANGE ? piece * n_total / n_pieces : 0;
  const int64_t r_cnt = RANGE ? (piece + 1) * n_total / n_pieces - r_beg : 0;
  const int64_t i_beg = RANGE ? (int64_t)(threadIdx.x >> 5) : ((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
  const int64_t i_end = RANGE ? r_cnt * H : NH;
  const int64_t i_step = RANGE ? (int64_t)(blockDim.x >> 5) : (((int64_t)gridDim.x * blockDim.x) >> 5);
  for (int64_t i = i_beg; i < i_end; i += i_step) {
    const int h = RANGE ? (int)(i / r_cnt) : (int)(i % H);
    const int64_t u = RANGE ? r_beg + (i - (int64_t)h * r_cnt) : i / H;
    const int64_t w = u * H + h;
    const int32_t beg = cptr[u], end = cptr[u + 1];
    const float elu_ = el[w];
    RowFrag<VEC, VECTOR> acc;
    acc.zero();
    float del_part = 0.f;
    for (int32_t base = beg; base < end; base += 32) {
      const bool on = base + lane < end;
      int32_t v_l = 0;
      float a_l = 0.f;
      if (on) {
        v_l = cidx[base + lane];
        const int64_t vh = (int64_t)v_l * H + h;
        a_l = expf(lrelu(elu_ + er[vh], slope) - rowmax[vh]) / rowsum[vh];
        del_part += dt_edge[(int64_t)csc2csr[base + lane] * H + h];
      }
      const int cnt = min(32, end - base);
      for (int j = 0; j < cnt; ++j) {
        const int32_t v = __shfl_sync(kFull, v_l, j);
        const float a = __shfl_sync(kFull, a_l, j);
        RowFrag<VEC, VECTOR> g;
        g.load(dO + (int64_t)v * lddo + (int64_t)h * F, F, lane);
        acc.fma(a, g);
      }
    }
    const float del_u = warp_sum(del_part);
    RowFrag<VEC, VECTOR> a;
    a.load(al + (int64_t)h * F, F, lane);
    acc.fma(del_u, a);
    a.load(ar + (int64_t)h * F, F, lane);
    acc.fma(der[w], a);
    acc.store(dZ + u * lddz + (int64_t)h * F, F, lane);
    if (lane == 0) del[w] = del_u;
  }
  }
}

// dattn[h,f] = sum_u coef[u,h] Z[u,h,f]: pass 1 per row-chunk partials.
__global__ void gat_attn_grad_partial_kernel(const float* __restrict__ Z, int64_t ldz, const float* __restrict__ coef,
                                             int64_t N, int H, int F, int64_t rows_per_block, float* __restrict__ partial) {
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < N ? r0 + rows_per_block : N;
  const int HF = H * F;
  for (int c = threadIdx.x; c < HF; c += blockDim.x) {
    const int h = c / F;
    float s = 0.f;
    for (int64_t r = r0; r < r1; ++r) s = fmaf(coef[r * H + h], Z[r * ldz + c], s);
    partial[(int64_t)blockIdx.x * HF + c] = s;
  }
}
// Both attention-vector gradients in ONE pass over Z: dal[h,f] = sum_u del[u,h] Z[u,h,f], dar likewise with der.
// Thread = 4 consecutive columns (float4), 8 rows in flight; HF % 4 == 0 and 16-byte aligned rows.
__global__ void __launch_bounds__(256)
gat_attn_grad2_partial_kernel(const float* __restrict__ Z, int64_t ldz, const float* __restrict__ cl,
                              const float* __restrict__ cr, int64_t N, int H, int F, int64_t rows_per_block,
                              float* __restrict__ partial_l, float* __restrict__ partial_r) {
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < N ? r0 + rows_per_block : N;
  const int HF4 = H * F / 4;
  for (int c4 = threadIdx.x; c4 < HF4; c4 += blockDim.x) {
    const int h = c4 * 4 / F;
    float4 sl = make_float4(0.f, 0.f, 0.f, 0.f), sr = sl;
    int64_t r = r0;
    for (; r + 8 <= r1; r += 8) {
      float4 z[8];
      float a[8], b[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        z[j] = ldg_nc_na(reinterpret_cast<const float4*>(Z + (r + j) * ldz) + c4);
        a[j] = __ldg(cl + (r + j) * H + h);
        b[j] = __ldg(cr + (r + j) * H + h);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sl.x = fmaf(a[j], z[j].x, sl.x); sl.y = fmaf(a[j], z[j].y, sl.y); sl.z = fmaf(a[j], z[j].z, sl.z); sl.w = fmaf(a[j], z[j].w, sl.w);
        sr.x = fmaf(b[j], z[j].x, sr.x); sr.y = fmaf(b[j], z[j].y, sr.y); sr.z = fmaf(b[j], z[j].z, sr.z); sr.w = fmaf(b[j], z[j].w, sr.w);
      }
    }
    for (; r < r1; ++r) {
      const float4 z = ldg_nc_na(reinterpret_cast<const float4*>(Z + r * ldz) + c4);
      const float a = __ldg(cl + r * H + h), b = __ldg(cr + r * H + h);
      sl.x = fmaf(a, z.x, sl.x); sl.y = fmaf(a, z.y, sl.y); sl.z = fmaf(a, z.z, sl.z); sl.w = fmaf(a, z.w, sl.w);
      sr.x = fmaf(b, z.x, sr.x); sr.y = fmaf(b, z.y, sr.y); sr.z = fmaf(b, z.z, sr.z); sr.w = fmaf(b, z.w, sr.w);
    }
    reinterpret_cast<float4*>(partial_l + (int64_t)blockIdx.x * H * F)[c4] = sl;
    reinterpret_cast<float4*>(partial_r + (int64_t)blockIdx.x * H * F)[c4] = sr;
  }
}
__global__ void gat_attn_grad_final_kernel(const float* __restrict__ partial, int n_blocks, int HF, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= HF) return;
  float s = 0.f;
  for (int b = 0; b < n_blocks; ++b) s += partial[(int64_t)b * HF + c];
  out[c] = s;
}

static inline int gat_grid(int64_t warps) {
  int64_t b = ceil_div<int64_t>(warps, kGatThreads / 32);
  const int64_t cap = (int64_t)sm_count() * 32;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}
static inline int attn_blocks(int64_t N) {
  int64_t b = ceil_div<int64_t>(N, 128);
  const int64_t cap = (int64_t)sm_count() * 4;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}
// nodes per sweep piece of the RANGE kernels (A/B: GTS_GAT_PIECE; 0 = one contiguous range per CTA)
static inline int gat_sweep_piece() {
  static const int v = getenv("GTS_GAT_PIECE") ? atoi(getenv("GTS_GAT_PIECE")) : 256;   // measured: 18.3 (one range) -> 17.4 ms per GAT step
  return v;
}
static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Dispatch on (F, alignment): float4 lanes when possible.
#define GAT_DISPATCH_R(KERNEL, R, TH, vec_ok, F, ...)                                               \
  do {                                                                                              \
    if ((vec_ok) && (F) <= 128)       KERNEL<1, true, R><<<grid, TH, 0, st>>>(__VA_ARGS__);         \
    else if ((vec_ok) && (F) <= 256)  KERNEL<2, true, R><<<grid, TH, 0, st>>>(__VA_ARGS__);         \
    else if ((vec_ok) && (F) <= 512)  KERNEL<4, true, R><<<grid, TH, 0, st>>>(__VA_ARGS__);         \
    else if ((F) <= 128)              KERNEL<1, false, R><<<grid, TH, 0, st>>>(__VA_ARGS__);        \
    else if ((F) <= 256)              KERNEL<2, false, R><<<grid, TH, 0, st>>>(__VA_ARGS__);        \
    else if ((F) <= 512)              KERNEL<4, false, R><<<grid, TH, 0, st>>>(__VA_ARGS__);        \
    else { set_error("GAT kernels support F <= 512 (got %d)", (int)(F)); return GTS_ERR_UNSUPPORTED; } \
  } while (0)
// head-major contiguous ranges (one 1024-thread CTA per SM) when there is at least one full sweep of nodes per SM;
// GTS_GAT_RANGE=0 keeps the grid-stride distribution (A/B switch)
#define GAT_DISPATCH(KERNEL, vec_ok, F, ...)                                                        \
  do {                                                                                              \
    static const bool range_off = getenv("GTS_GAT_RANGE") && atoi(getenv("GTS_GAT_RANGE")) == 0;    \
    if (!range_off && NH / H >= (int64_t)sm_count() * 32) {                                         \
      const int grid = sm_count();                                                                  \
      GAT_DISPATCH_R(KERNEL, true, 1024, vec_ok, F, __VA_ARGS__);                                   \
    } else {                                                                                        \
      const int grid = gat_grid(NH);                                                                \
      GAT_DISPATCH_R(KERNEL, false, kGatThreads, vec_ok, F, __VA_ARGS__);                           \
    }                                                                                               \
  } while (0)

}  // namespace gts

using namespace gts;

extern "C" {

int gts_gat_scores(const float* Z, int64_t ldz, const float* attn_l, const float* attn_r,
                   int32_t n_nodes, int32_t H, int32_t F, float* el, float* er, gts_stream_t stream) {
  GTS_CHECK_ARG(n_nodes >= 0 && H >= 1 && F >= 1, "gts_gat_scores: bad size");
  if (n_nodes == 0) return GTS_OK;
  GTS_CHECK_ARG(Z && attn_l && attn_r && el && er, "gts_gat_scores: null pointer");
  const int64_t NH = (int64_t)n_nodes * H;
  gat_scores_kernel<<<gat_grid(NH), kGatThreads, 0, as_stream(stream)>>>(Z, ldz, attn_l, attn_r, NH, H, F, el, er);
  GTS_LAUNCH_CHECK();
  return GTS_OK;
}

int gts_gat_fwd(const float* Z, int64_t ldz, const float* el, const float* er,
                const int32_t* indptr, const int32_t* indices,
                int32_t n_nodes, int32_t H, int32_t F, float slope,
                const float* res, int64_t ldres, const float* bias, int32_t act,
                float* out, int64_t ldo, float* rowmax, float* rowsum,
                int32_t* err_flag, gts_stream_t stream) {
  GTS_CHECK_ARG(n_nodes >= 0 && H >= 1 && F >= 1, "gts_gat_fwd: bad size");
  GTS_CHECK_ARG(act == 0 || act == 1, "gts_gat_fwd: act must be 0 (none) or 1 (ELU)");
  if (n_nodes == 0) return GTS_OK;
  GTS_CHECK_ARG(Z && el && er && indptr && out && rowmax && rowsum && err_flag, "gts_gat_fwd: null pointer");
  cudaStream_t st = as_stream(stream);
  const int64_t NH = (int64_t)n_nodes * H;
  const bool vec_ok = (F % 4 == 0) && (ldz % 4 == 0) && (ldo % 4 == 0) && al16(Z) && al16(out) &&
                      (!res || ((ldres % 4 == 0) && al16(res))) && (!bias || al16(bias));
  GAT_DISPATCH(gat_fwd_kernel, vec_ok, F, Z, ldz, el, er, indptr, indices, NH, H, F, slope, res, ldres, bias, act,
               out, ldo, rowmax, rowsum, err_flag, gat_sweep_piece());
  GTS_LAUNCH_CHECK();
  return GTS_OK;
}

int gts_gat_act_bwd(const float* dout, const float* out, int64_t n, int32_t act, float* dpre, gts_stream_t stream) {
  GTS_CHECK_ARG(n >= 0, "gts_gat_act_bwd: negative size");
  if (n == 0) return GTS_OK;
  GTS_CHECK_ARG(dout && dpre && (act == 0 || out), "gts_gat_act_bwd: null pointer");
  int64_t b = ceil_div<int64_t>(n, 256);
  const int64_t cap = (int64_t)sm_count() * 16;
  if (b > cap) b = cap;
  gat_act_bwd_kernel<<<(int)b, 256, 0, as_stream(stream)>>>(dout, out, n, act, dpre);
  GTS_LAUNCH_CHECK();
  return GTS_OK;
}

int gts_gat_bwd_dst(const float* Z, int64_t ldz, const float* el, const float* er,
                    const float* rowmax, const float* rowsum,
                    const int32_t* indptr, const int32_t* indices,
                    const float* dO, int64_t lddo,
                    int32_t n_nodes, int32_t H, int32_t F, float slope,
                    float* dt_edge, float* der, gts_stream_t stream) {
  GTS_CHECK_ARG(n_nodes >= 0 && H >= 1 && F >= 1, "gts_gat_bwd_dst: bad size");
  if (n_nodes == 0) return GTS_OK;
  GTS_CHECK_ARG(Z && el && er && rowmax && rowsum && indptr && dO && dt_edge && der, "gts_gat_bwd_dst: null pointer");
  cudaStream_t st = as_stream(stream);
  const int64_t NH = (int64_t)n_nodes * H;
  const bool vec_ok = (F % 4 == 0) && (ldz % 4 == 0) && (lddo % 4 == 0) && al16(Z) && al16(dO);
  GAT_DISPATCH(gat_bwd_dst_kernel, vec_ok, F, Z, ldz, el, er, rowmax, rowsum, indptr, indices, dO, lddo, NH, H, F,
               slope, dt_edge, der, gat_sweep_piece());
  GTS_LAUNCH_CHECK();
  return GTS_OK;
}

int gts_gat_bwd_src(const float* el, const float* er, const float* rowmax, const float* rowsum,
                    const int32_t* csc_indptr, const int32_t* csc_indices, const int32_t* csc2csr,
                    const float* dO, int64_t lddo, const float* dt_edge, const float* der,
                    const float* attn_l, const float* attn_r,
                    int32_t n_nodes, int32_t H, int32_t F, float slope,
                    float* dZ, int64_t lddz, float* del, gts_stream_t stream) {
  GTS_CHECK_ARG(n_nodes >= 0 && H >= 1 && F >= 1, "gts_gat_bwd_src: bad size");
  if (n_nodes == 0) return GTS_OK;
  GTS_CHECK_ARG(el && er && rowmax && rowsum && csc_indptr && dO && dt_edge && der && attn_l && attn_r && dZ && del,
                "gts_gat_bwd_src: null pointer");
  cudaStream_t st = as_stream(stream);
  const int64_t NH = (int64_t)n_nodes * H;
  const bool vec_ok = (F % 4 == 0) && (lddo % 4 == 0) && (lddz % 4 == 0) && al16(dO) && al16(dZ) && al16(attn_l) && al16(attn_r);
  GAT_DISPATCH(gat_bwd_src_kernel, vec_ok, F, el, er, rowmax, rowsum, csc_indptr, csc_indices, csc2csr, dO, lddo,
               dt_edge, der, attn_l, attn_r, NH, H, F, slope, dZ, lddz, del, gat_sweep_piece());
  GTS_LAUNCH_CHECK();
  return GTS_OK;
}

size_t gts_gat_attn_grad_workspace_bytes(int32_t n_nodes, int32_t H, int32_t F) {
  if (n_nodes <= 0 || H <= 0 || F <= 0) return 256;
  return align_up((size_t)attn_blocks(n_nodes) * (size_t)H * (size_t)F * sizeof(float), 256);
}

int gts_gat_attn_grad(const float* Z, int64_t ldz, const float* coef, int32_t n_nodes,
                      int32_t H, int32_t F, float* dattn, void* workspace, size_t workspace_bytes,
                      gts_stream_t stream) {
  GTS_CHECK_ARG(n_nodes >= 0 && H >= 1 && F >= 1, "gts_gat_attn_grad: bad size");
  GTS_CHECK_ARG(dattn != nullptr, "gts_gat_attn_grad: dattn is null");
  cudaStream_t st = as_stream(stream);
  if (n_nodes == 0) {
    GTS_CUDA(cudaMemsetAsync(dattn, 0, sizeof(float) * (size_t)H * F, st));
    return GTS_OK;
  }
  GTS_CHECK_ARG(Z && coef && workspace, "gts_gat_attn_grad: null pointer");
  const size_t need = gts_gat_attn_grad_workspace_bytes(n_nodes, H, F);
  if (workspace_bytes < need) {
    set_error("gts_gat_attn_grad: workspace %zu < required %zu", workspace_bytes, need);
    return GTS_ERR_WORKSPACE;
  }
  const int nb = attn_blocks(n_nodes);
  const int64_t rpb = ceil_div<int64_t>(n_nodes, nb);
  float* partial = reinterpret_cast<float*>(workspace);
  gat_attn_grad_partial_kernel<<<nb, 256, 0, st>>>(Z, ldz, coef, n_nodes, H, F, rpb, partial);
  GTS_LAUNCH_CHECK();
  gat_attn_grad_final_kernel<<<(H * F + 127) / 128, 128, 0, st>>>(partial, nb, H * F, dattn);
  GTS_LAUNCH_CHECK();
  return GTS_OK;
}

int gts_gat_attn_grad2(const float* Z, int64_t ldz, const float* coef_l, const float* coef_r, int32_t n_nodes,
                       int32_t H, int32_t F, float* dattn_l, float* dattn_r, void* workspace, size_t workspace_bytes,
                       gts_stream_t stream) {
  GTS_CHECK_ARG(n_nodes >= 0 && H >= 1 && F >= 1, "gts_gat_attn_grad2: bad size");
  GTS_CHECK_ARG(dattn_l && dattn_r, "gts_gat_attn_grad2: null output");
  const size_t one = gts_gat_attn_grad_workspace_bytes(n_nodes, H, F);
  const bool vec_ok = n_nodes > 0 && F % 4 == 0 && ldz % 4 == 0 && al16(Z);
  if (!vec_ok) {       // generic shapes: two passes of the scalar kernel
    int rc = gts_gat_attn_grad(Z, ldz, coef_l, n_nodes, H, F, dattn_l, workspace, workspace_bytes, stream);
    if (rc != GTS_OK) return rc;
    return gts_gat_attn_grad(Z, ldz, coef_r, n_nodes, H, F, dattn_r, workspace, workspace_bytes, stream);
  }
  GTS_CHECK_ARG(Z && coef_l && coef_r && workspace, "gts_gat_attn_grad2: null pointer");
  if (workspace_bytes < 2 * one) {
    set_error("gts_gat_attn_grad2: workspace %zu < required %zu", workspace_bytes, 2 * one);
    return GTS_ERR_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  const int nb = attn_blocks(n_nodes);
  const int64_t rpb = ceil_div<int64_t>(n_nodes, nb);
  float* pl = reinterpret_cast<float*>(workspace);
  float* pr = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + one);
  gat_attn_grad2_partial_kernel<<<nb, 256, 0, st>>>(Z, ldz, coef_l, coef_r, n_nodes, H, F, rpb, pl, pr);
  GTS_LAUNCH_CHECK();
  gat_attn_grad_final_kernel<<<(H * F + 127) / 128, 128, 0, st>>>(pl, nb, H * F, dattn_l);
  GTS_LAUNCH_CHECK();
  gat_attn_grad_final_kernel<<<(H * F + 127) / 128, 128, 0, st>>>(pr, nb, H * F, dattn_r);
  GTS_LAUNCH_CHECK();
  return GTS_OK;
}

}  // extern "C"

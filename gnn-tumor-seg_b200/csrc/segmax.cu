// segmax.cu — K2 / K2b: neighbour max-aggregation over the in-edge CSR and its
// backward.  Replaces DGL update_all(copy_src, max) (gspmm copy_lhs/max with
// arg outputs) and the scatter_add_ of its autograd inside SAGEConv('pool')
// (reference model/networks.py:35).
//
// Forward: one warp (or a 2^k-lane slice of a warp for narrow rows) per
// destination node; lane l owns float4 column chunks l, l+LPN, ...; the row's
// neighbour ids are read coalesced once and broadcast by shuffle; neighbour
// rows are gathered with 128-bit loads, four neighbours in flight, and folded
// in CSR order with a strictly-greater test, so the FIRST maximum wins
// (bit-exact arg-max vs the oracle).  HBM-bound: algorithmic bytes
// 4*(N*D read + N*D write + N*D argmax + (N+1) + E).
#include "common.cuh"
#include <algorithm>
#include <cfloat>
#include <cstdlib>

namespace gts {

constexpr int kSegThreads = 256;

// The arg-max fold is `if (v > best) { best = v; arg = u; }` per element.  Written plainly it compiles to
// FSETP + FSEL + SEL, all three on the half-rate ALU pipe, and the kernel is bound by that pipe (ncu: pipe_alu 82 %,
// pipe_fma 9 %).  Here the two selects are predicated FFMA / IMAD with operands the compiler cannot fold
// (v * 1.0f + -0.0f == v and arg * 0 + u == u exactly, for every v incl. signed zeros, infinities, denormals; the
// constants arrive as kernel parameters): one ALU instruction and two on the otherwise idle FMA pipe per element.
struct FoldConst { float one, neg_zero; int zero; };
static inline FoldConst fold_const() { return FoldConst{1.0f, -0.0f, 0}; }

__device__ __forceinline__ void fold_one(float& best, int& arg, float v, int u, const FoldConst& k) {
  asm("{\n\t.reg .pred p;\n\t"
      "setp.gt.f32 p, %2, %0;\n\t"
      "@p fma.rn.f32 %0, %2, %4, %5;\n\t"
      "@p mad.lo.s32 %1, %1, %6, %3;\n\t}"
      : "+f"(best), "+r"(arg) : "f"(v), "r"(u), "f"(k.one), "f"(k.neg_zero), "r"(k.zero));
}
// Select forms for the double-buffered kernel (template parameter FV): 0 = the FFMA / IMAD form above;
// 3 = FADD / IMAD: v + (-0.0f) == v exactly for every v, and its one opaque operand comes straight from the constant
// bank, so the two float constants need no registers.  (Rejected by SASS inspection: plain predicated movs become
// FSEL + SEL, both ALU pipe; `u * 1 + 0` is hoisted out of the column loop as a common subexpression and the select
// comes back as an ALU-pipe SEL.)
template <int FV>
__device__ __forceinline__ void fold_one_p(float& best, int& arg, float v, int u, const FoldConst& k) {
  if (FV == 0) {
    fold_one(best, arg, v, u, k);
  } else {
    // ONE opaque register z (bits 0: +0.0f and int 0) serves both selects: v - (+0.0f) == v exactly for every v
    // (incl. -0.0f, infinities, denormals; the negation is an operand modifier) and arg * 0 + u == u
    asm("{\n\t.reg .pred p;\n\t"
        "setp.gt.f32 p, %2, %0;\n\t"
        "@p sub.rn.f32 %0, %2, %4;\n\t"
        "@p mad.lo.s32 %1, %1, %5, %3;\n\t}"
        : "+f"(best), "+r"(arg) : "f"(v), "r"(u), "f"(__int_as_float(k.zero)), "r"(k.zero));
  }
}
template <int FV>
__device__ __forceinline__ void fold_max_p(float4& best, int4& arg, const float4& v, int u, const FoldConst& k) {
  fold_one_p<FV>(best.x, arg.x, v.x, u, k);
  fold_one_p<FV>(best.y, arg.y, v.y, u, k);
  fold_one_p<FV>(best.z, arg.z, v.z, u, k);
  fold_one_p<FV>(best.w, arg.w, v.w, u, k);
}
// inference (no arg-max wanted): the value alone is one FMNMX per element
__device__ __forceinline__ void fold_val(float4& best, const float4& v) {
  best.x = fmaxf(best.x, v.x); best.y = fmaxf(best.y, v.y); best.z = fmaxf(best.z, v.z); best.w = fmaxf(best.w, v.w);
}
__device__ __forceinline__ void fold_max(float4& best, int4& arg, const float4& v, int u, const FoldConst& k) {
  fold_one(best.x, arg.x, v.x, u, k);
  fold_one(best.y, arg.y, v.y, u, k);
  fold_one(best.z, arg.z, v.z, u, k);
  fold_one(best.w, arg.w, v.w, u, k);
}

// LPN lanes per node (power of two <= 32), VEC float4 chunks per lane:
// covers D4 = D/4 <= LPN*VEC.
// Work distribution: each CTA owns ONE contiguous range of nodes_per_cta destination
// nodes and walks it with all its warps side by side.  Supervoxel ids are spatially
// coherent (SLIC numbers its clusters in raster order), so the neighbour rows of
// consecutive nodes overlap heavily: keeping an SM on one sliding window of ids turns
// most of the E x D gather traffic into L1 hits instead of L2 round trips.
template <int LPN, int VEC, bool WRITE_ARG, int THREADS>
__global__ void __launch_bounds__(THREADS)
segmax_fwd_vec_kernel(const float* __restrict__ P, int64_t ldp, const int32_t* __restrict__ indptr,
                      const int32_t* __restrict__ indices, int32_t N, int32_t D4,
                      float* __restrict__ neigh, int64_t ldn, int32_t* __restrict__ argmax, int64_t ldarg,
                      int64_t nodes_per_cta, const FoldConst fc) {
  constexpr int NPW = 32 / LPN;   // nodes per warp
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPN;
  const int slot = lane / LPN;
  const unsigned full = 0xffffffffu;
  const int64_t v_begin = (int64_t)blockIdx.x * nodes_per_cta;
  const int64_t v_end = v_begin + nodes_per_cta < N ? v_begin + nodes_per_cta : N;
  const int64_t warp_stride = (int64_t)(THREADS / 32) * NPW;

  for (int64_t v0 = v_begin + (int64_t)(threadIdx.x >> 5) * NPW; v0 < v_end; v0 += warp_stride) {
    const int64_t v = v0 + slot;
    const bool live = v < v_end;
    int32_t beg = 0, end = 0;
    if (live) { beg = indptr[v]; end = indptr[v + 1]; }
    float4 best[VEC];
    int4 arg[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      best[j] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      arg[j] = make_int4(-1, -1, -1, -1);
    }
    // longest row among the nodes sharing this warp (uniform loop bound)
    int32_t deg = end - beg;
    int32_t max_deg = deg;
    if (NPW > 1) {
#pragma unroll
      for (int o = 16; o >= LPN; o >>= 1) max_deg = max(max_deg, __shfl_xor_sync(full, max_deg, o));
    }
    for (int32_t base = 0; base < max_deg; base += LPN) {
      // each LPN-lane group reads LPN neighbour ids coalesced
      const int32_t my_idx = (base + sub < deg) ? indices[beg + base + sub] : -1;
      const int32_t cnt = min(LPN, max_deg - base);
      for (int32_t j = 0; j < cnt; j += 4) {
        int32_t u[4];
        float4 r[4][VEC];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          // j+q < LPN holds whenever LPN >= 4; LPN < 4 wraps harmlessly onto -1 lanes via the cnt test
          u[q] = __shfl_sync(full, my_idx, (slot * LPN) + ((j + q) % LPN));
          if (j + q >= cnt) u[q] = -1;
#pragma unroll
          for (int c = 0; c < VEC; ++c) {
            const int chunk = sub + c * LPN;
            if (u[q] >= 0 && chunk < D4)
              r[q][c] = ldg_nc(reinterpret_cast<const float4*>(P + (int64_t)u[q] * ldp) + chunk);
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (u[q] >= 0) {
#pragma unroll
            for (int c = 0; c < VEC; ++c)
              if (sub + c * LPN < D4) fold_max(best[c], arg[c], r[q][c], u[q], fc);
          }
        }
      }
    }
    if (live) {
#pragma unroll
      for (int c = 0; c < VEC; ++c) {
        const int chunk = sub + c * LPN;
        if (chunk < D4) {
          float4 o = best[c];
          if (arg[c].x < 0) o.x = 0.f;
          if (arg[c].y < 0) o.y = 0.f;
          if (arg[c].z < 0) o.z = 0.f;
          if (arg[c].w < 0) o.w = 0.f;
          stg_na(reinterpret_cast<float4*>(neigh + v * ldn) + chunk, o);
          if (WRITE_ARG) stg_na(reinterpret_cast<int4*>(argmax + v * ldarg) + chunk, arg[c]);
        }
      }
    }
  }
}


// Wide rows with D == 128 * VEC exactly (the 256-wide hidden layers): one warp per
// destination node, lane l owns float4 chunks l, l+32, ....  The first version of this
// kernel was ISSUE-bound, not memory-bound (ncu: ~29 SASS instructions per 128-bit load,
// most of them predicate bookkeeping), so this form keeps the inner loop branch-free:
// whole groups of four neighbours with no per-neighbour or per-column predicates, the
// (< 4) remainder handled once per row.  One 768-thread CTA per SM on a contiguous id
// range keeps the L1 hit rate of the gathers at ~65 % (see the note above).
template <int VEC, bool WRITE_ARG, int kWideWarps>
__global__ void __launch_bounds__(kWideWarps * 32, 1)
segmax_fwd_wide_kernel(const float* __restrict__ P, int64_t ldp, const int32_t* __restrict__ indptr,
                       const int32_t* __restrict__ indices, int32_t N,
                       float* __restrict__ neigh, int64_t ldn, int32_t* __restrict__ argmax, int64_t ldarg,
                       int64_t nodes_per_cta, const FoldConst fc) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const unsigned full = 0xffffffffu;
  const int slab4 = blockIdx.y * 32 * VEC;       // gridDim.y column slabs of 128*VEC floats each
  const char* __restrict__ Pl = reinterpret_cast<const char*>(reinterpret_cast<const float4*>(P) + slab4 + lane);
  const uint32_t ld_bytes = (uint32_t)(ldp << 2);   // host guarantees N * ldp * 4 < 2^32: one IMAD.WIDE per row address
  auto row_of = [&](int32_t u) {
    return reinterpret_cast<const float4*>(Pl + (uint64_t)(uint32_t)u * ld_bytes);
  };

  // sweep: chunk blockIdx.x + s * gridDim.x at step s (one step when nodes_per_cta = ceil(N / grid))
  for (int64_t v_begin = (int64_t)blockIdx.x * nodes_per_cta; v_begin < N; v_begin += (int64_t)gridDim.x * nodes_per_cta) {
  const int64_t v_end = v_begin + nodes_per_cta < N ? v_begin + nodes_per_cta : N;
  for (int64_t v = v_begin + warp; v < v_end; v += kWideWarps) {
    const int32_t beg = indptr[v], end = indptr[v + 1];
    float4 best[VEC];
    int4 arg[VEC];
#pragma unroll
    for (int c = 0; c < VEC; ++c) {
      best[c] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      arg[c] = make_int4(-1, -1, -1, -1);
    }
    for (int32_t base = beg; base < end; base += 32) {
      const int32_t my_idx = (base + lane < end) ? indices[base + lane] : 0;
      const int32_t last = min(32, end - base) - 1;
      // whole groups of four; a short last group repeats the row's last neighbour, which can never win
      // again (strictly-greater test), so the loop body has no per-neighbour predicate at all
      for (int32_t j = 0; j <= last; j += 4) {
        const int32_t u0 = __shfl_sync(full, my_idx, j), u1 = __shfl_sync(full, my_idx, min(j + 1, last));
        const int32_t u2 = __shfl_sync(full, my_idx, min(j + 2, last)), u3 = __shfl_sync(full, my_idx, min(j + 3, last));
        const float4* p0 = row_of(u0);
        const float4* p1 = row_of(u1);
        const float4* p2 = row_of(u2);
        const float4* p3 = row_of(u3);
        float4 r0[VEC], r1[VEC], r2[VEC], r3[VEC];
#pragma unroll
        for (int c = 0; c < VEC; ++c) { r0[c] = ldg_nc(p0 + 32 * c); r1[c] = ldg_nc(p1 + 32 * c); }
#pragma unroll
        for (int c = 0; c < VEC; ++c) { r2[c] = ldg_nc(p2 + 32 * c); r3[c] = ldg_nc(p3 + 32 * c); }
#pragma unroll
        for (int c = 0; c < VEC; ++c) {
          if (WRITE_ARG) {
            fold_max(best[c], arg[c], r0[c], u0, fc);
            fold_max(best[c], arg[c], r1[c], u1, fc);
            fold_max(best[c], arg[c], r2[c], u2, fc);
            fold_max(best[c], arg[c], r3[c], u3, fc);
          } else {
            fold_val(best[c], r0[c]); fold_val(best[c], r1[c]); fold_val(best[c], r2[c]); fold_val(best[c], r3[c]);
          }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < VEC; ++c) {
      float4 o = best[c];
      if (end == beg) o = make_float4(0.f, 0.f, 0.f, 0.f);     // no in-edges: DGL fills 0
      stg_na(reinterpret_cast<float4*>(neigh + v * ldn) + slab4 + lane + 32 * c, o);
      if (WRITE_ARG) stg_na(reinterpret_cast<int4*>(argmax + v * ldarg) + slab4 + lane + 32 * c, arg[c]);
    }
  }
  }
}

template <int VEC, int kWideWarps>
static int launch_fwd_wide(const float* P, int64_t ldp, const int32_t* indptr, const int32_t* indices, int32_t N,
                           float* neigh, int64_t ldn, int32_t* argmax, int64_t ldarg, cudaStream_t st, int slabs = 1) {
  // one CTA per SM, contiguous id ranges of (almost) equal length
  int64_t grid = sm_count();
  static const int chunk_req = getenv("GTS_SEGMAX_CHUNK") ? atoi(getenv("GTS_SEGMAX_CHUNK")) : 0;
  const int64_t per_cta = chunk_req > 0 ? chunk_req : ceil_div<int64_t>(N, grid);
  grid = std::min<int64_t>(grid, ceil_div<int64_t>(N, per_cta));
  if (argmax)
    segmax_fwd_wide_kernel<VEC, true, kWideWarps><<<dim3((unsigned)grid, slabs), kWideWarps * 32, 0, st>>>(P, ldp, indptr, indices, N, neigh, ldn, argmax, ldarg, per_cta, fold_const());
  else
    segmax_fwd_wide_kernel<VEC, false, kWideWarps><<<dim3((unsigned)grid, slabs), kWideWarps * 32, 0, st>>>(P, ldp, indptr, indices, N, neigh, ldn, nullptr, 0, per_cta, fold_const());
  GTS_LAUNCH_CHECK();
  return GTS_OK;
}

// Second form of the wide kernel: software-pipelined.  What the measurements said about the first one
// (profiles/r01_segmax_variants.md): it is neither purely issue-bound nor DRAM-bound - a warp loads four rows, waits
// for the slowest of them (almost every group holds a +-z neighbour row that misses L1) and only then folds, so time
// goes with the number of resident warps; and every row is fetched from DRAM about twice because the CTAs that share
// it (one z-slab of ids apart) touch it a whole kernel duration apart.  This form changes four things:
//   * stages of R rows are double-buffered (ping-pong A / C): the loads of the next stage are in flight while the
//     current one is folded, so a warp never drains its memory pipeline inside a row.  R = 1 needs 16 row registers
//     (32 resident warps in 64 registers), R = 2 needs 32 (24 warps in 80);
//   * the row's first neighbour INITIALISES (best, arg) instead of being folded into (-inf, -1) (inputs are finite:
//     P is a ReLU output; a row of -inf/NaN would differ from the fold form), and tails are exact: the head takes
//     1..2R rows so that whole double-stages remain - no padded repeats of the last neighbour;
//   * the next node's (beg, end) and neighbour ids are fetched while the current node is processed, which takes
//     the indptr -> indices -> row dependent-latency chain off the per-node critical path;
//   * the two selects share ONE opaque zero register (fold_one_p) instead of two float constants and an int;
//   * the id space is swept by all CTAs side by side in short pieces (see the iterator below) instead of one long
//     contiguous range per CTA, which keeps the shared rows in L2.
// Rows with more than 32 in-edges take the plain loop at the end (never on supervoxel RAGs: max degree ~30).
// Tried and dropped: prefetch.global.L2 of the next node's rows (+1.4 GB of L2 requests for lines that mostly hit L1
// anyway: 92 -> 111 us).
constexpr int kMaxPieceSteps = 512;

template <int VEC, bool WRITE_ARG, int kWarps, int FV, int R>
__global__ void __launch_bounds__(kWarps * 32, 1)
segmax_fwd_pipe_kernel(const float* __restrict__ P, int64_t ldp, const int32_t* __restrict__ indptr,
                       const int32_t* __restrict__ indices, int32_t N,
                       float* __restrict__ neigh, int64_t ldn, int32_t* __restrict__ argmax, int64_t ldarg,
                       int32_t steps, const FoldConst fc, uint32_t* __restrict__ pos_bits, int64_t ld_bits) {
  constexpr int G = 2 * R;      // rows per loop iteration (two stages)
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const unsigned full = 0xffffffffu;
  // piece table.  The id space is cut into steps * gridDim.x nearly equal pieces; at sweep step s the CTA owns piece
  // s * gridDim.x + blockIdx.x and its warps walk the concatenation of its pieces kWarps apart (equal node counts
  // per warp and per CTA, +-1).  steps = 1 is one contiguous range per CTA; with short pieces the rows a CTA shares
  // with its neighbours in id space are touched by everybody at about the same time (L2 hits), while neighbours
  // within a piece still hit L1.
  __shared__ int32_t s_lo[kMaxPieceSteps + 1], s_len[kMaxPieceSteps + 1];
  {
    const int64_t n_pieces = (int64_t)steps * gridDim.x;
    for (int32_t t = threadIdx.x; t < steps; t += kWarps * 32) {
      const int64_t k = (int64_t)t * gridDim.x + blockIdx.x;
      const int32_t lo = (int32_t)(k * N / n_pieces);
      s_lo[t] = lo;
      s_len[t] = (int32_t)((k + 1) * N / n_pieces) - lo;
    }
    if (threadIdx.x == 0) { s_lo[steps] = N; s_len[steps] = 0x7fffffff; }     // sentinel: the iterator parks on N
    __syncthreads();
  }
  int32_t it_s = 0, it_off = warp;
  auto next_node = [&]() {
    int32_t len = s_len[it_s];
    while (it_off >= len) { it_off -= len; len = s_len[++it_s]; }
    const int32_t r = s_lo[it_s] + it_off;
    if (it_s < steps) it_off += kWarps; else it_off = 0;
    return r;             // == N once exhausted
  };

  const int slab4 = blockIdx.y * 32 * VEC;       // gridDim.y column slabs of 128 * VEC floats each
  const char* __restrict__ Pl = reinterpret_cast<const char*>(reinterpret_cast<const float4*>(P) + slab4 + lane);
  const uint32_t ld_bytes = (uint32_t)(ldp << 2);   // host guarantees N * ldp * 4 < 2^32
  struct Row { float4 c[VEC]; };
  auto row_ptr = [&](int32_t u) {
    uint64_t a;      // base + u * ld_bytes
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(a) : "r"(u), "r"(ld_bytes), "l"(Pl));
    return reinterpret_cast<const float4*>(a);
  };
  auto load_row = [&](Row& r, int32_t u) {
    const float4* p = row_ptr(u);
#pragma unroll
    for (int c = 0; c < VEC; ++c) r.c[c] = ldg_nc(p + 32 * c);
  };
  // predicated form (registers keep their contents when !on): keeps the stage loop free of branches
  auto load_row_if = [&](Row& r, int32_t u, int32_t on) {
    const float4* p = row_ptr(u);
#pragma unroll
    for (int c = 0; c < VEC; ++c)
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %5, 0;\n\t"
                   "@p ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];\n\t}"
                   : "+f"(r.c[c].x), "+f"(r.c[c].y), "+f"(r.c[c].z), "+f"(r.c[c].w) : "l"(p + 32 * c), "r"(on));
  };
#define FOLD_ROW(ROW, U)                                                                                         \
  _Pragma("unroll") for (int c_ = 0; c_ < VEC; ++c_) {                                                           \
    if (WRITE_ARG) fold_max_p<FV>(best[c_], arg[c_], (ROW).c[c_], (U), fc); else fold_val(best[c_], (ROW).c[c_]); \
  }

  int32_t v = next_node();
  if (v >= N) return;
  // software pipeline over nodes: the next node's (beg, end) are requested when the current node starts and its
  // neighbour ids once the head rows are folded (the offsets have arrived by then)
  int32_t beg = indptr[v], end = indptr[v + 1];
  int32_t my_idx = (beg + lane < end) ? indices[beg + lane] : 0;

  while (v < N) {
    const int32_t deg = end - beg;
    const int32_t v1 = next_node();
    int32_t nbeg = 0, nend = 0, n_idx = 0;
    if (v1 < N) { nbeg = indptr[v1]; nend = indptr[v1 + 1]; }

    float4 best[VEC];
    int4 arg[VEC];
    if (deg > 0 && deg <= 32) {
      // head: 1..G rows so that a multiple of G remains; the first one initialises (best, arg)
      const int32_t head = ((deg - 1) & (G - 1)) + 1;
      Row A[R], C[R];
      int32_t ua[R], uc[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
#pragma unroll
        for (int c = 0; c < VEC; ++c) A[r].c[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        ua[r] = 0; uc[r] = 0;
      }
      const int32_t u0 = __shfl_sync(full, my_idx, 0);
      {
        Row F;
        load_row(F, u0);
#pragma unroll
        for (int c = 0; c < VEC; ++c) { best[c] = F.c[c]; arg[c] = make_int4(u0, u0, u0, u0); }
      }
      // head rows 1 .. head-1 (at most G-1 = 2R-1) are staged in C[0..R-1] then A[0..R-2]
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (head >= 2 + r) { uc[r] = __shfl_sync(full, my_idx, 1 + r); load_row(C[r], uc[r]); }
#pragma unroll
      for (int r = 0; r < R - 1; ++r)
        if (head >= 2 + R + r) { ua[r] = __shfl_sync(full, my_idx, 1 + R + r); load_row(A[r], ua[r]); }
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (head >= 2 + r) FOLD_ROW(C[r], uc[r]);
#pragma unroll
      for (int r = 0; r < R - 1; ++r)
        if (head >= 2 + R + r) FOLD_ROW(A[r], ua[r]);
      int32_t j = head;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        ua[r] = __shfl_sync(full, my_idx, j + r);      // lane index wraps mod 32
        load_row_if(A[r], ua[r], j < deg);
      }
      if (nbeg + lane < nend) n_idx = indices[nbeg + lane];
      // whole double-stages from here: A holds neighbours j .. j+R-1 on entry
      while (j < deg) {
#pragma unroll
        for (int r = 0; r < R; ++r) { uc[r] = __shfl_sync(full, my_idx, j + R + r); load_row(C[r], uc[r]); }
#pragma unroll
        for (int r = 0; r < R; ++r) FOLD_ROW(A[r], ua[r]);
        j += G;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          ua[r] = __shfl_sync(full, my_idx, j + r);
          load_row_if(A[r], ua[r], j < deg);
        }
#pragma unroll
        for (int r = 0; r < R; ++r) FOLD_ROW(C[r], uc[r]);
      }
    } else {
#pragma unroll
      for (int c = 0; c < VEC; ++c) {
        best[c] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        arg[c] = make_int4(-1, -1, -1, -1);
      }
      if (nbeg + lane < nend) n_idx = indices[nbeg + lane];
      for (int32_t e = beg; e < end; ++e) {      // deg > 32: plain loop, ids by warp-uniform loads
        const int32_t u = indices[e];
        Row T;
        load_row(T, u);
#pragma unroll
        for (int c = 0; c < VEC; ++c) {
          if (WRITE_ARG) fold_max(best[c], arg[c], T.c[c], u, fc); else fold_val(best[c], T.c[c]);
        }
      }
      if (deg == 0) {                            // no in-edges: DGL fills 0
#pragma unroll
        for (int c = 0; c < VEC; ++c) best[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int c = 0; c < VEC; ++c) {
      stg_na(reinterpret_cast<float4*>(neigh + (int64_t)v * ldn) + slab4 + lane + 32 * c, best[c]);
      if (WRITE_ARG) stg_na(reinterpret_cast<int4*>(argmax + (int64_t)v * ldarg) + slab4 + lane + 32 * c, arg[c]);
    }
    if (pos_bits) {
      // (neigh > 0) as bits in the layout of GTS_ACT_MASK_BITS: lanes 8q .. 8q+7 hold the 32 columns of word 4c + q
      // (column 4 * (lane & 7) + comp <-> bit 8 * comp + (lane & 7)): byte q of each component's ballot
      const int q = lane >> 3;
      const uint32_t sel = (uint32_t)q | ((4u + (uint32_t)q) << 4);
#pragma unroll
      for (int c = 0; c < VEC; ++c) {
        const uint32_t b0 = __ballot_sync(full, best[c].x > 0.f), b1 = __ballot_sync(full, best[c].y > 0.f);
        const uint32_t b2 = __ballot_sync(full, best[c].z > 0.f), b3 = __ballot_sync(full, best[c].w > 0.f);
        const uint32_t w = __byte_perm(__byte_perm(b0, b1, sel), __byte_perm(b2, b3, sel), 0x5410);
        if ((lane & 7) == 0) pos_bits[(int64_t)v * ld_bits + (slab4 >> 3) + 4 * c + q] = w;
      }
    }
    v = v1;
    beg = nbeg; end = nend; my_idx = n_idx;
  }
#undef FOLD_ROW
}

template <int VEC, int kWarps, int FV, int R>
static int launch_fwd_pipe(const float* P, int64_t ldp, const int32_t* indptr, const int32_t* indices, int32_t N,
                           float* neigh, int64_t ldn, int32_t* argmax, int64_t ldarg, cudaStream_t st, int piece_req,
                           int slabs = 1, uint32_t* pos_bits = nullptr, int64_t ld_bits = 0) {
  // piece_req = target nodes per piece (0: one contiguous range per CTA)
  const int64_t grid = std::min<int64_t>(sm_count(), ceil_div<int64_t>(N, kWarps));
  int64_t steps = 1;
  if (piece_req > 0) steps = std::min<int64_t>(kMaxPieceSteps, std::max<int64_t>(1, (N + grid * piece_req / 2) / (grid * piece_req)));
  static const bool max_l1 = [] {     // the gathers live on L1 hits: ask for the largest L1 carve-out (A/B: GTS_SEGMAX_MAXL1=0)
    const bool on = !(getenv("GTS_SEGMAX_MAXL1") && atoi(getenv("GTS_SEGMAX_MAXL1")) == 0);
    if (on) {
      cudaFuncSetAttribute(segmax_fwd_pipe_kernel<VEC, true, kWarps, FV, R>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1);
      cudaFuncSetAttribute(segmax_fwd_pipe_kernel<VEC, false, kWarps, FV, R>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1);
    }
    return on;
  }();
  (void)max_l1;
  if (argmax)
    segmax_fwd_pipe_kernel<VEC, true, kWarps, FV, R><<<dim3((unsigned)grid, slabs), kWarps * 32, 0, st>>>(P, ldp, indptr, indices, N, neigh, ldn, argmax, ldarg, (int32_t)steps, fold_const(), pos_bits, ld_bits);
  else
    segmax_fwd_pipe_kernel<VEC, false, kWarps, FV, R><<<dim3((unsigned)grid, slabs), kWarps * 32, 0, st>>>(P, ldp, indptr, indices, N, neigh, ldn, nullptr, 0, (int32_t)steps, fold_const(), pos_bits, ld_bits);
  GTS_LAUNCH_CHECK();
  return GTS_OK;
}

// Scalar fallback: any D / alignment.  One warp per node, lanes stride over columns.
__global__ void segmax_fwd_scalar_kernel(const float* __restrict__ P, int64_t ldp,
                                         const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                         int32_t N, int32_t D, float* __restrict__ neigh, int64_t ldn,
                                         int32_t* __restrict__ argmax, int64_t ldarg) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t v = warp_global; v < N; v += n_warps) {
    const int32_t beg = indptr[v], end = indptr[v + 1];
    for (int k = lane; k < D; k += 32) {
      float best = -INFINITY;
      int32_t a = -1;
      for (int32_t p = beg; p < end; ++p) {
        const int32_t u = indices[p];
        const float x = P[(int64_t)u * ldp + k];
        if (x > best) { best = x; a = u; }
      }
      neigh[v * ldn + k] = (a < 0) ? 0.f : best;
      if (argmax) argmax[v * ldarg + k] = a;
    }
  }
}

// K2b (atomic form): one thread per float4 of dNeigh.
__global__ void segmax_bwd_vec_kernel(const float* __restrict__ dN, int64_t ldd, const int32_t* __restrict__ arg,
                                      int64_t ldarg, int64_t total4, int32_t D4, float* __restrict__ dP, int64_t lddp) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = i / D4;
    const int c = (int)(i - v * D4);
    const float4 g = ldg_nc_na(reinterpret_cast<const float4*>(dN + v * ldd) + c);
    const int4 a = *(reinterpret_cast<const int4*>(arg + v * ldarg) + c);
    const int k = c * 4;
    if (a.x >= 0 && g.x != 0.f) atomicAdd(dP + (int64_t)a.x * lddp + k + 0, g.x);
    if (a.y >= 0 && g.y != 0.f) atomicAdd(dP + (int64_t)a.y * lddp + k + 1, g.y);
    if (a.z >= 0 && g.z != 0.f) atomicAdd(dP + (int64_t)a.z * lddp + k + 2, g.z);
    if (a.w >= 0 && g.w != 0.f) atomicAdd(dP + (int64_t)a.w * lddp + k + 3, g.w);
  }
}

__global__ void segmax_bwd_scalar_kernel(const float* __restrict__ dN, int64_t ldd, const int32_t* __restrict__ arg,
                                         int64_t ldarg, int64_t N, int32_t D, float* __restrict__ dP, int64_t lddp) {
  const int64_t total = N * D;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = i / D;
    const int k = (int)(i - v * D);
    const int32_t a = arg[v * ldarg + k];
    const float g = dN[v * ldd + k];
    if (a >= 0 && g != 0.f) atomicAdd(dP + (int64_t)a * lddp + k, g);
  }
}

// K2b (deterministic form): one warp per SOURCE node, transposed gather over
// the out-edge CSC; sums in CSC (edge id) order.
__global__ void __launch_bounds__(kSegThreads)
segmax_bwd_det_kernel(const float* __restrict__ dN, int64_t ldd, const int32_t* __restrict__ arg, int64_t ldarg,
                      const int32_t* __restrict__ cptr, const int32_t* __restrict__ cidx, int32_t N, int32_t D,
                      float* __restrict__ dP, int64_t lddp) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t u = warp_global; u < N; u += n_warps) {
    const int32_t beg = cptr[u], end = cptr[u + 1];
    for (int k = lane; k < D; k += 32) {
      float acc = 0.f;
      for (int32_t p = beg; p < end; ++p) {
        const int32_t v = cidx[p];
        if (arg[(int64_t)v * ldarg + k] == (int32_t)u) acc += dN[(int64_t)v * ldd + k];
      }
      dP[u * lddp + k] = acc;
    }
  }
}

// K2b, deterministic AND vectorised: the transposed gather in the shape of the forward kernel.  One warp per SOURCE
// node u on a contiguous id range per SM, lanes own float4 column chunks; for every out-neighbour v (CSC row of u, in
// edge-id order) the warp reads arg[v, :] (int4 per lane) and adds dN[v, k] where arg[v, k] == u.  The dN chunk is
// only fetched by lanes with a match in their four columns (~1/4 of them: each (v, k) has exactly one winner among
// ~15 neighbours).  No zero-fill, no atomics: every dP element is written exactly once, sums run in CSC order, so the
// gradients are bitwise reproducible.  The adds are predicated FADDs (FMA pipe), the compares ISETPs (ALU pipe).
template <int VEC>
__global__ void __launch_bounds__(1024, 1)
segmax_bwd_gather_wide_kernel(const float* __restrict__ dN, int64_t ldd, const int32_t* __restrict__ arg, int64_t ldarg,
                              const int32_t* __restrict__ cptr, const int32_t* __restrict__ cidx, int32_t N,
                              float* __restrict__ dP, int64_t lddp, int64_t nodes_per_cta) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const unsigned full = 0xffffffffu;
  const int64_t u_begin = (int64_t)blockIdx.x * nodes_per_cta;
  const int64_t u_end = u_begin + nodes_per_cta < N ? u_begin + nodes_per_cta : N;
  const char* __restrict__ argl = reinterpret_cast<const char*>(reinterpret_cast<const int4*>(arg) + lane);
  const char* __restrict__ dNl = reinterpret_cast<const char*>(reinterpret_cast<const float4*>(dN) + lane);
  const uint32_t lda_bytes = (uint32_t)(ldarg << 2), ldd_bytes = (uint32_t)(ldd << 2);   // host: N * ld * 4 < 2^32

  for (int64_t u64 = u_begin + warp; u64 < u_end; u64 += 32) {
    const int32_t u = (int32_t)u64;
    const int32_t beg = cptr[u], end = cptr[u + 1];
    float4 acc[VEC];
#pragma unroll
    for (int c = 0; c < VEC; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int32_t base = beg; base < end; base += 32) {
      const int32_t my_idx = (base + lane < end) ? cidx[base + lane] : 0;
      const int32_t cnt = min(32, end - base);
      for (int32_t j = 0; j < cnt; j += 2) {
        const bool two = j + 1 < cnt;
        const int32_t v0 = __shfl_sync(full, my_idx, j), v1 = __shfl_sync(full, my_idx, two ? j + 1 : j);
        const int4* a0p = reinterpret_cast<const int4*>(argl + (uint64_t)(uint32_t)v0 * lda_bytes);
        const int4* a1p = reinterpret_cast<const int4*>(argl + (uint64_t)(uint32_t)v1 * lda_bytes);
        int4 a0[VEC], a1[VEC];
#pragma unroll
        for (int c = 0; c < VEC; ++c) { a0[c] = __ldg(a0p + 32 * c); a1[c] = __ldg(a1p + 32 * c); }
        const int32_t m1 = two ? u : -2;                  // a repeated last neighbour must not be counted twice
        float4 g0[VEC], g1[VEC];
#pragma unroll
        for (int c = 0; c < VEC; ++c) {
          g0[c] = make_float4(0.f, 0.f, 0.f, 0.f);
          g1[c] = g0[c];
          if (a0[c].x == u || a0[c].y == u || a0[c].z == u || a0[c].w == u)
            g0[c] = ldg_nc(reinterpret_cast<const float4*>(dNl + (uint64_t)(uint32_t)v0 * ldd_bytes) + 32 * c);
          if (a1[c].x == m1 || a1[c].y == m1 || a1[c].z == m1 || a1[c].w == m1)
            g1[c] = ldg_nc(reinterpret_cast<const float4*>(dNl + (uint64_t)(uint32_t)v1 * ldd_bytes) + 32 * c);
        }
#pragma unroll
        for (int c = 0; c < VEC; ++c) {
          if (a0[c].x == u) acc[c].x += g0[c].x;
          if (a0[c].y == u) acc[c].y += g0[c].y;
          if (a0[c].z == u) acc[c].z += g0[c].z;
          if (a0[c].w == u) acc[c].w += g0[c].w;
          if (a1[c].x == m1) acc[c].x += g1[c].x;
          if (a1[c].y == m1) acc[c].y += g1[c].y;
          if (a1[c].z == m1) acc[c].z += g1[c].z;
          if (a1[c].w == m1) acc[c].w += g1[c].w;
        }
      }
    }
#pragma unroll
    for (int c = 0; c < VEC; ++c) stg_na(reinterpret_cast<float4*>(dP + u64 * lddp) + lane + 32 * c, acc[c]);
  }
}

// Sum-type aggregators (mean / gcn) — same traversal, additive fold.
__global__ void __launch_bounds__(kSegThreads)
segsum_fwd_kernel(const float* __restrict__ P, int64_t ldp, const int32_t* __restrict__ indptr,
                  const int32_t* __restrict__ indices, int32_t N, int32_t D, int mode,
                  float* __restrict__ out, int64_t ldo) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t v = warp_global; v < N; v += n_warps) {
    const int32_t beg = indptr[v], end = indptr[v + 1];
    const int32_t deg = end - beg;
    for (int k = lane; k < D; k += 32) {
      float acc = 0.f;
      for (int32_t p = beg; p < end; ++p) acc += P[(int64_t)indices[p] * ldp + k];
      if (mode == 1) acc = deg > 0 ? acc / (float)deg : 0.f;
      else if (mode == 2) acc = (acc + P[v * ldp + k]) / (float)(deg + 1);
      out[v * ldo + k] = acc;
    }
  }
}


__global__ void __launch_bounds__(kSegThreads)
segsum_bwd_kernel(const float* __restrict__ dO, int64_t ldd, const int32_t* __restrict__ cptr,
                  const int32_t* __restrict__ cidx, const int32_t* __restrict__ in_ptr, int32_t N, int32_t D, int mode,
                  float* __restrict__ dP, int64_t lddp) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t u = warp_global; u < N; u += n_warps) {
    const int32_t beg = cptr[u], end = cptr[u + 1];
    for (int k = lane; k < D; k += 32) {
      float acc = 0.f;
      for (int32_t p = beg; p < end; ++p) {
        const int32_t v = cidx[p];
        const int32_t deg = in_ptr[v + 1] - in_ptr[v];
        const float s = mode == 0 ? 1.f : (mode == 1 ? 1.f / (float)deg : 1.f / (float)(deg + 1));
        acc = fmaf(s, dO[(int64_t)v * ldd + k], acc);
      }
      if (mode == 2) {
        const int32_t deg = in_ptr[u + 1] - in_ptr[u];
        acc += dO[u * ldd + k] / (float)(deg + 1);
      }
      dP[u * lddp + k] = acc;
    }
  }
}

__global__ void mask_pos_kernel(const float* __restrict__ g, const float* __restrict__ ref, int64_t n, float* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = ref[i] > 0.f ? g[i] : 0.f;
}

static inline int seg_grid(int64_t warps_needed) {
  const int wpb = kSegThreads / 32;
  int64_t blocks = ceil_div<int64_t>(warps_needed, wpb);
  const int64_t cap = (int64_t)sm_count() * 32;   // grid-stride beyond 32 resident-sized waves
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

template <int LPN, int VEC>
static int launch_fwd_vec(const float* P, int64_t ldp, const int32_t* indptr, const int32_t* indices, int32_t N,
                          int32_t D4, float* neigh, int64_t ldn, int32_t* argmax, int64_t ldarg, cudaStream_t st) {
  // one CTA per SM (768 threads for rows up to 2 chunks per lane, 512 above), each on a contiguous id range
  constexpr int THREADS = VEC <= 2 ? 768 : (VEC <= 4 ? 384 : 256);
  constexpr int NPW = 32 / LPN;
  const int64_t per_pass = (int64_t)(THREADS / 32) * NPW;            // nodes one CTA touches per sweep
  int64_t grid = sm_count();
  int64_t per_cta = ceil_div<int64_t>(N, grid);
  per_cta = ceil_div<int64_t>(per_cta, per_pass) * per_pass;         // whole sweeps
  grid = ceil_div<int64_t>(N, per_cta);
  if (argmax)
    segmax_fwd_vec_kernel<LPN, VEC, true, THREADS><<<(int)grid, THREADS, 0, st>>>(P, ldp, indptr, indices, N, D4, neigh, ldn, argmax, ldarg, per_cta, fold_const());
  else
    segmax_fwd_vec_kernel<LPN, VEC, false, THREADS><<<(int)grid, THREADS, 0, st>>>(P, ldp, indptr, indices, N, D4, neigh, ldn, nullptr, 0, per_cta, fold_const());
  GTS_LAUNCH_CHECK();
  return GTS_OK;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Sum-type aggregators, wide form (D == 128 * VEC, aligned rows): the traversal of the forward seg-max kernel with an
// additive fold (8 FADD/FFMA per neighbour and lane, on the FMA pipe).  One kernel for both directions:
//   forward  (row = destination v over the in-edge CSR):  out[v] = sum_u X[u]            (mode 0)
//                                                                  / deg(v)              (mode 1, mean)
//                                                         (sum_u X[u] + X[v]) / (deg+1)  (mode 2, gcn)
//   backward (row = source u over the out-edge CSC):      out[u] = sum_v s(v) X[v] (+ s(u) X[u] for gcn),
//            s(v) = 1, 1/deg_in(v), 1/(deg_in(v)+1); deg_in comes from the in-edge indptr `wptr`.
template <int VEC, bool BACKWARD>
__global__ void __launch_bounds__(1024, 1)
segsum_wide_kernel(const float* __restrict__ X, int64_t ldx, const int32_t* __restrict__ ptr_,
                   const int32_t* __restrict__ idx, const int32_t* __restrict__ wptr, int32_t N, int mode,
                   float* __restrict__ out, int64_t ldo, int64_t nodes_per_cta) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const unsigned full = 0xffffffffu;
  const int64_t r_begin = (int64_t)blockIdx.x * nodes_per_cta;
  const int64_t r_end = r_begin + nodes_per_cta < N ? r_begin + nodes_per_cta : N;
  const char* __restrict__ Xl = reinterpret_cast<const char*>(reinterpret_cast<const float4*>(X) + lane);
  const uint32_t ld_bytes = (uint32_t)(ldx << 2);
  auto scale_of = [&](int32_t v) -> float {
    if (mode == 0) return 1.f;
    const int32_t d = wptr[v + 1] - wptr[v];
    return mode == 1 ? (d > 0 ? 1.f / (float)d : 0.f) : 1.f / (float)(d + 1);
  };
  for (int64_t r = r_begin + warp; r < r_end; r += 32) {
    const int32_t beg = ptr_[r], end = ptr_[r + 1];
    float4 acc[VEC];
#pragma unroll
    for (int c = 0; c < VEC; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int32_t base = beg; base < end; base += 32) {
      int32_t my_idx = 0;
      float my_w = 0.f;                                   // lanes past the row end carry weight 0
      if (base + lane < end) {
        my_idx = idx[base + lane];
        my_w = BACKWARD ? scale_of(my_idx) : 1.f;
      }
      const int32_t cnt = min(32, end - base);
      for (int32_t j = 0; j < cnt; j += 4) {              // a short last group re-reads lane `cnt` onwards: index 0, weight 0
        int32_t u[4];
        float w[4];
        float4 x[4][VEC];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          u[q] = __shfl_sync(full, my_idx, (j + q) & 31);
          w[q] = __shfl_sync(full, my_w, (j + q) & 31);
          const float4* p = reinterpret_cast<const float4*>(Xl + (uint64_t)(uint32_t)u[q] * ld_bytes);
#pragma unroll
          for (int c = 0; c < VEC; ++c) x[q][c] = ldg_nc(p + 32 * c);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (j + q < cnt) {
#pragma unroll
            for (int c = 0; c < VEC; ++c) {
              acc[c].x = fmaf(w[q], x[q][c].x, acc[c].x); acc[c].y = fmaf(w[q], x[q][c].y, acc[c].y);
              acc[c].z = fmaf(w[q], x[q][c].z, acc[c].z); acc[c].w = fmaf(w[q], x[q][c].w, acc[c].w);
            }
          }
        }
      }
    }
    const int32_t deg = end - beg;
    float self_w = 0.f;
    if (!BACKWARD) {
      if (mode == 2) self_w = 1.f;
    } else if (mode == 2) {
      self_w = scale_of((int32_t)r);
    }
#pragma unroll
    for (int c = 0; c < VEC; ++c) {
      float4 o = acc[c];
      if (self_w != 0.f) {
        const float4 sx = ldg_nc(reinterpret_cast<const float4*>(Xl + (uint64_t)(uint32_t)r * ld_bytes) + 32 * c);
        o.x = fmaf(self_w, sx.x, o.x); o.y = fmaf(self_w, sx.y, o.y); o.z = fmaf(self_w, sx.z, o.z); o.w = fmaf(self_w, sx.w, o.w);
      }
      if (!BACKWARD && mode != 0) {                       // the reference divides: (sum [+ self]) / n
        const float n = mode == 1 ? (float)deg : (float)(deg + 1);
        if (mode == 1 && deg == 0) o = make_float4(0.f, 0.f, 0.f, 0.f);
        else { o.x /= n; o.y /= n; o.z /= n; o.w /= n; }
      }
      stg_na(reinterpret_cast<float4*>(out + r * ldo) + lane + 32 * c, o);
    }
  }
}

template <bool BACKWARD>
static bool launch_segsum_wide(const float* X, int64_t ldx, const int32_t* ptr_, const int32_t* idx, const int32_t* wptr,
                               int32_t N, int32_t D, int mode, float* out, int64_t ldo, cudaStream_t st) {
  const bool ok = (D == 128 || D == 256) && ldx % 4 == 0 && ldo % 4 == 0 && aligned16(X) && aligned16(out) &&
                  (int64_t)N * ldx * 4 < ((int64_t)1 << 32);
  if (!ok) return false;
  int64_t grid = sm_count();
  const int64_t per_cta = ceil_div<int64_t>(N, grid);
  grid = ceil_div<int64_t>(N, per_cta);
  if (D == 256) segsum_wide_kernel<2, BACKWARD><<<(unsigned)grid, 1024, 0, st>>>(X, ldx, ptr_, idx, wptr, N, mode, out, ldo, per_cta);
  else segsum_wide_kernel<1, BACKWARD><<<(unsigned)grid, 1024, 0, st>>>(X, ldx, ptr_, idx, wptr, N, mode, out, ldo, per_cta);
  return true;
}


}  // namespace gts

using namespace gts;

extern "C" {

int gts_segmax_fwd(const float* P, int64_t ldp, const int32_t* indptr, const int32_t* indices,
                   int32_t n_nodes, int32_t D, float* neigh, int64_t ldn,
                   int32_t* argmax, int64_t ldarg, gts_stream_t stream) {
  GTS_CHECK_ARG(n_nodes >= 0 && D >= 0, "gts_segmax_fwd: negative size");
  if (n_nodes == 0 || D == 0) return GTS_OK;
  GTS_CHECK_ARG(P && indptr && neigh, "gts_segmax_fwd: null pointer");
  GTS_CHECK_ARG(ldp >= D && ldn >= D && (!argmax || ldarg >= D), "gts_segmax_fwd: leading dimension < D");
  cudaStream_t st = as_stream(stream);
  const bool vec_ok = (D % 4 == 0) && (ldp % 4 == 0) && (ldn % 4 == 0) && (!argmax || ldarg % 4 == 0) &&
                      aligned16(P) && aligned16(neigh) && (!argmax || aligned16(argmax)) && D <= 1024;
  static const bool no_wide = getenv("GTS_SEGMAX_GENERIC") != nullptr;   // A/B switch for profiling
  static const int wide_warps = getenv("GTS_SEGMAX_WARPS") ? atoi(getenv("GTS_SEGMAX_WARPS")) : 32;
  const bool wide_ok = !no_wide && vec_ok && (int64_t)n_nodes * ldp * 4 < ((int64_t)1 << 32);
  if (wide_ok && D == 128) return launch_fwd_wide<1, 32>(P, ldp, indptr, indices, n_nodes, neigh, ldn, argmax, ldarg, st);
  // D == 256 (the hidden layers): the software-pipelined form (segmax_fwd_pipe_kernel), 32 warps, one row per stage,
  // sweep pieces of ~64 nodes - the fastest of the measured matrix (profiles/r01_segmax_variants.md).  A/B switches:
  // GTS_SEGMAX_PIPE=<warps> (0 = the grouped form), GTS_SEGMAX_STAGE=<rows per stage>, GTS_SEGMAX_FOLD=<0|3>,
  // GTS_SEGMAX_CHUNK=<nodes per sweep piece; 0 = one contiguous range per CTA>, GTS_SEGMAX_SLABS=2, GTS_SEGMAX_MAXL1=0.
  static const int pipe_warps = getenv("GTS_SEGMAX_PIPE") ? atoi(getenv("GTS_SEGMAX_PIPE")) : 32;
  static const int seg_chunk = getenv("GTS_SEGMAX_CHUNK") ? atoi(getenv("GTS_SEGMAX_CHUNK")) : (pipe_warps > 0 ? 64 : 0);
  static const int pipe_fold = getenv("GTS_SEGMAX_FOLD") ? atoi(getenv("GTS_SEGMAX_FOLD")) : 3;
  static const int pipe_stage = getenv("GTS_SEGMAX_STAGE") ? atoi(getenv("GTS_SEGMAX_STAGE")) : 1;
  if (wide_ok && D == 256 && pipe_warps > 0) {
#define GTS_PIPE(W, F, S)                                                                            \
    if (pipe_warps == W && pipe_fold == F && pipe_stage == S)                                         \
      return launch_fwd_pipe<2, W, F, S>(P, ldp, indptr, indices, n_nodes, neigh, ldn, argmax, ldarg, st, seg_chunk);
    static const int pipe_slabs = getenv("GTS_SEGMAX_SLABS") ? atoi(getenv("GTS_SEGMAX_SLABS")) : 1;
    if (pipe_slabs == 2) {     // two 128-column slabs, one CTA each: half the L1 working set per SM
      if (pipe_stage == 1) return launch_fwd_pipe<1, 32, 3, 1>(P, ldp, indptr, indices, n_nodes, neigh, ldn, argmax, ldarg, st, seg_chunk, 2);
      if (pipe_stage == 2) return launch_fwd_pipe<1, 32, 3, 2>(P, ldp, indptr, indices, n_nodes, neigh, ldn, argmax, ldarg, st, seg_chunk, 2);
      return launch_fwd_pipe<1, 32, 3, 4>(P, ldp, indptr, indices, n_nodes, neigh, ldn, argmax, ldarg, st, seg_chunk, 2);
    }
    GTS_PIPE(20, 3, 2) GTS_PIPE(24, 3, 2) GTS_PIPE(24, 0, 2)
    GTS_PIPE(24, 3, 1) GTS_PIPE(28, 3, 1) GTS_PIPE(32, 3, 1) GTS_PIPE(32, 0, 1)
#undef GTS_PIPE
    GTS_CHECK_ARG(false, "gts_segmax_fwd: unsupported GTS_SEGMAX_PIPE / _FOLD / _STAGE combination");
  }
  if (wide_ok && D == 256) {
    if (wide_warps == 48) return launch_fwd_wide<1, 24>(P, ldp, indptr, indices, n_nodes, neigh, ldn, argmax, ldarg, st, 2);
    if (wide_warps == 64) return launch_fwd_wide<1, 32>(P, ldp, indptr, indices, n_nodes, neigh, ldn, argmax, ldarg, st, 2);
    if (wide_warps == 24) return launch_fwd_wide<2, 24>(P, ldp, indptr, indices, n_nodes, neigh, ldn, argmax, ldarg, st);
    if (wide_warps == 16) return launch_fwd_wide<2, 16>(P, ldp, indptr, indices, n_nodes, neigh, ldn, argmax, ldarg, st);
    return launch_fwd_wide<2, 32>(P, ldp, indptr, indices, n_nodes, neigh, ldn, argmax, ldarg, st);
  }
  if (vec_ok) {
    const int D4 = D / 4;
    if (D4 <= 4)   return launch_fwd_vec<4, 1>(P, ldp, indptr, indices, n_nodes, D4, neigh, ldn, argmax, ldarg, st);
    if (D4 <= 8)   return launch_fwd_vec<8, 1>(P, ldp, indptr, indices, n_nodes, D4, neigh, ldn, argmax, ldarg, st);
    if (D4 <= 16)  return launch_fwd_vec<16, 1>(P, ldp, indptr, indices, n_nodes, D4, neigh, ldn, argmax, ldarg, st);
    if (D4 <= 32)  return launch_fwd_vec<32, 1>(P, ldp, indptr, indices, n_nodes, D4, neigh, ldn, argmax, ldarg, st);
    if (D4 <= 64)  return launch_fwd_vec<32, 2>(P, ldp, indptr, indices, n_nodes, D4, neigh, ldn, argmax, ldarg, st);
    if (D4 <= 128) return launch_fwd_vec<32, 4>(P, ldp, indptr, indices, n_nodes, D4, neigh, ldn, argmax, ldarg, st);
    return launch_fwd_vec<32, 8>(P, ldp, indptr, indices, n_nodes, D4, neigh, ldn, argmax, ldarg, st);
  }
  segmax_fwd_scalar_kernel<<<seg_grid(n_nodes), kSegThreads, 0, st>>>(P, ldp, indptr, indices, n_nodes, D, neigh, ldn, argmax, ldarg);
  GTS_LAUNCH_CHECK();
  return GTS_OK;
}

int gts_segmax_fwd_bits_supported(int32_t n_nodes, int32_t D, int64_t ldp) {
  // the default software-pipelined D = 256 kernel (no A/B environment overrides of its configuration)
  static const bool plain = !getenv("GTS_SEGMAX_GENERIC") && !getenv("GTS_SEGMAX_PIPE") && !getenv("GTS_SEGMAX_FOLD") &&
                            !getenv("GTS_SEGMAX_STAGE") && !getenv("GTS_SEGMAX_SLABS");
  return plain && D == 256 && ldp % 4 == 0 && n_nodes > 0 && (int64_t)n_nodes * ldp * 4 < ((int64_t)1 << 32);
}

int gts_segmax_fwd_bits(const float* P, int64_t ldp, const int32_t* indptr, const int32_t* indices,
                        int32_t n_nodes, int32_t D, float* neigh, int64_t ldn,
                        int32_t* argmax, int64_t ldarg, uint32_t* pos_bits, int64_t ld_bits, gts_stream_t stream) {
  GTS_CHECK_ARG(n_nodes >= 0 && D >= 0, "gts_segmax_fwd_bits: negative size");
  if (n_nodes == 0 || D == 0) return GTS_OK;
  GTS_CHECK_ARG(P && indptr && neigh && pos_bits, "gts_segmax_fwd_bits: null pointer");
  GTS_CHECK_ARG(ldp >= D && ldn >= D && (!argmax || ldarg >= D) && ld_bits >= D / 32, "gts_segmax_fwd_bits: leading dimension too small");
  const bool vec_ok = (ldn % 4 == 0) && (!argmax || ldarg % 4 == 0) && aligned16(P) && aligned16(neigh) && (!argmax || aligned16(argmax));
  if (!gts_segmax_fwd_bits_supported(n_nodes, D, ldp) || !vec_ok) {
    set_error("gts_segmax_fwd_bits: D = %d / this layout is not supported (gts_segmax_fwd_bits_supported)", D);
    return GTS_ERR_UNSUPPORTED;
  }
  return launch_fwd_pipe<2, 32, 3, 1>(P, ldp, indptr, indices, n_nodes, neigh, ldn, argmax, ldarg, as_stream(stream), 64, 1,
                                      pos_bits, ld_bits);
}

int gts_segmax_bwd_add(const float* dNeigh, int64_t ldd, const int32_t* argmax, int64_t ldarg,
                       int32_t n_nodes, int32_t D, float* dP, int64_t lddp, gts_stream_t stream) {
  GTS_CHECK_ARG(n_nodes >= 0 && D >= 0, "gts_segmax_bwd_add: negative size");
  if (n_nodes == 0 || D == 0) return GTS_OK;
  GTS_CHECK_ARG(dNeigh && argmax && dP, "gts_segmax_bwd_add: null pointer");
  GTS_CHECK_ARG(lddp >= D, "gts_segmax_bwd_add: lddp < D");
  cudaStream_t st = as_stream(stream);
  const bool vec_ok = (D % 4 == 0) && (ldd % 4 == 0) && (ldarg % 4 == 0) && aligned16(dNeigh) && aligned16(argmax);
  const int threads = 256;
  if (vec_ok) {
    const int64_t total4 = (int64_t)n_nodes * (D / 4);
    int64_t blocks = ceil_div<int64_t>(total4, threads);
    const int64_t cap = (int64_t)sm_count() * 64;
    if (blocks > cap) blocks = cap;
    segmax_bwd_vec_kernel<<<(int)blocks, threads, 0, st>>>(dNeigh, ldd, argmax, ldarg, total4, D / 4, dP, lddp);
  } else {
    const int64_t total = (int64_t)n_nodes * D;
    int64_t blocks = ceil_div<int64_t>(total, threads);
    const int64_t cap = (int64_t)sm_count() * 64;
    if (blocks > cap) blocks = cap;
    segmax_bwd_scalar_kernel<<<(int)blocks, threads, 0, st>>>(dNeigh, ldd, argmax, ldarg, n_nodes, D, dP, lddp);
  }
  GTS_LAUNCH_CHECK();
  return GTS_OK;
}

int gts_segmax_bwd(const float* dNeigh, int64_t ldd, const int32_t* argmax, int64_t ldarg,
                   int32_t n_nodes, int32_t D, float* dP, int64_t lddp, int32_t n_src_rows,
                   gts_stream_t stream) {
  GTS_CHECK_ARG(n_nodes >= 0 && D >= 0 && n_src_rows >= 0, "gts_segmax_bwd: negative size");
  if (n_src_rows == 0 || D == 0) return GTS_OK;
  GTS_CHECK_ARG(dP != nullptr, "gts_segmax_bwd: dP is null");
  GTS_CHECK_ARG(lddp >= D, "gts_segmax_bwd: lddp < D");
  cudaStream_t st = as_stream(stream);
  if (lddp == D) {
    GTS_CUDA(cudaMemsetAsync(dP, 0, sizeof(float) * (size_t)n_src_rows * D, st));
  } else {
    GTS_CUDA(cudaMemset2DAsync(dP, sizeof(float) * lddp, 0, sizeof(float) * D, n_src_rows, st));
  }
  if (n_nodes == 0) return GTS_OK;
  GTS_CHECK_ARG(dNeigh && argmax, "gts_segmax_bwd: null pointer");
  return gts_segmax_bwd_add(dNeigh, ldd, argmax, ldarg, n_nodes, D, dP, lddp, stream);
}

int gts_segmax_bwd_det(const float* dNeigh, int64_t ldd, const int32_t* argmax, int64_t ldarg,
                       const int32_t* csc_indptr, const int32_t* csc_indices,
                       int32_t n_nodes, int32_t D, float* dP, int64_t lddp, gts_stream_t stream) {
  GTS_CHECK_ARG(n_nodes >= 0 && D >= 0, "gts_segmax_bwd_det: negative size");
  if (n_nodes == 0 || D == 0) return GTS_OK;
  GTS_CHECK_ARG(dNeigh && argmax && csc_indptr && dP, "gts_segmax_bwd_det: null pointer");
  const bool wide_ok = (D == 128 || D == 256) && ldd % 4 == 0 && ldarg % 4 == 0 && lddp % 4 == 0 && aligned16(dNeigh) &&
                       aligned16(argmax) && aligned16(dP) && (int64_t)n_nodes * ldd * 4 < ((int64_t)1 << 32) &&
                       (int64_t)n_nodes * ldarg * 4 < ((int64_t)1 << 32);
  static const bool no_wide = getenv("GTS_SEGMAX_GENERIC") != nullptr;
  if (wide_ok && !no_wide) {
    int64_t grid = sm_count();
    const int64_t per_cta = ceil_div<int64_t>(n_nodes, grid);
    grid = ceil_div<int64_t>(n_nodes, per_cta);
    if (D == 256)
      segmax_bwd_gather_wide_kernel<2><<<(unsigned)grid, 1024, 0, as_stream(stream)>>>(dNeigh, ldd, argmax, ldarg, csc_indptr,
                                                                                     csc_indices, n_nodes, dP, lddp, per_cta);
    else
      segmax_bwd_gather_wide_kernel<1><<<(unsigned)grid, 1024, 0, as_stream(stream)>>>(dNeigh, ldd, argmax, ldarg, csc_indptr,
                                                                                     csc_indices, n_nodes, dP, lddp, per_cta);
    GTS_LAUNCH_CHECK();
    return GTS_OK;
  }
  segmax_bwd_det_kernel<<<seg_grid(n_nodes), kSegThreads, 0, as_stream(stream)>>>(dNeigh, ldd, argmax, ldarg, csc_indptr,
                                                                                   csc_indices, n_nodes, D, dP, lddp);
  GTS_LAUNCH_CHECK();
  return GTS_OK;
}

int gts_segsum_fwd(const float* P, int64_t ldp, const int32_t* indptr, const int32_t* indices,
                   int32_t n_nodes, int32_t D, int32_t mode, float* out, int64_t ldo, gts_stream_t stream) {
  GTS_CHECK_ARG(n_nodes >= 0 && D >= 0, "gts_segsum_fwd: negative size");
  GTS_CHECK_ARG(mode >= 0 && mode <= 2, "gts_segsum_fwd: mode must be 0 (sum), 1 (mean) or 2 (gcn)");
  if (n_nodes == 0 || D == 0) return GTS_OK;
  GTS_CHECK_ARG(P && indptr && out, "gts_segsum_fwd: null pointer");
  if (launch_segsum_wide<false>(P, ldp, indptr, indices, indptr, n_nodes, D, mode, out, ldo, as_stream(stream))) {
    GTS_LAUNCH_CHECK();
    return GTS_OK;
  }
  segsum_fwd_kernel<<<seg_grid(n_nodes), kSegThreads, 0, as_stream(stream)>>>(P, ldp, indptr, indices, n_nodes, D, mode, out, ldo);
  GTS_LAUNCH_CHECK();
  return GTS_OK;
}

int gts_segsum_bwd(const float* dOut, int64_t ldd, const int32_t* csc_indptr, const int32_t* csc_indices,
                   const int32_t* in_indptr, int32_t n_nodes, int32_t D, int32_t mode,
                   float* dP, int64_t lddp, gts_stream_t stream) {
  GTS_CHECK_ARG(n_nodes >= 0 && D >= 0, "gts_segsum_bwd: negative size");
  GTS_CHECK_ARG(mode >= 0 && mode <= 2, "gts_segsum_bwd: mode must be 0, 1 or 2");
  if (n_nodes == 0 || D == 0) return GTS_OK;
  GTS_CHECK_ARG(dOut && csc_indptr && in_indptr && dP, "gts_segsum_bwd: null pointer");
  if (launch_segsum_wide<true>(dOut, ldd, csc_indptr, csc_indices, in_indptr, n_nodes, D, mode, dP, lddp, as_stream(stream))) {
    GTS_LAUNCH_CHECK();
    return GTS_OK;
  }
  segsum_bwd_kernel<<<seg_grid(n_nodes), kSegThreads, 0, as_stream(stream)>>>(dOut, ldd, csc_indptr, csc_indices, in_indptr,
                                                                               n_nodes, D, mode, dP, lddp);
  GTS_LAUNCH_CHECK();
  return GTS_OK;
}

int gts_mask_pos(const float* grad, const float* ref, int64_t n, float* out, gts_stream_t stream) {
  GTS_CHECK_ARG(n >= 0, "gts_mask_pos: negative size");
  if (n == 0) return GTS_OK;
  GTS_CHECK_ARG(grad && ref && out, "gts_mask_pos: null pointer");
  int64_t b = ceil_div<int64_t>(n, 256);
  const int64_t cap = (int64_t)sm_count() * 16;
  if (b > cap) b = cap;
  mask_pos_kernel<<<(int)b, 256, 0, as_stream(stream)>>>(grad, ref, n, out);
  GTS_LAUNCH_CHECK();
  return GTS_OK;
}

}  // extern "C"

// project.cu — K7: node -> voxel reprojection (integer gathers over the int16
// supervoxel map).  Replaces, on the device:
//   project_nodes_to_img         data_processing/graph_io.py:21-24
//   uncrop_to_brats_size         data_processing/image_processing.py:21-25
//   swap_labels_to_brats         scripts/preprocess_dataset.py:159-169
//   torch.max(logits,1)          scripts/generate_gnn_predictions.py:66
//   save_voxel_logits gather     scripts/generate_gnn_predictions.py:55-62
// HBM-bound byte work: algorithmic bytes for gts_project_labels =
// 2*X*Y*Z (map read) + 2*VX*VY*VZ (volume write) + 4*N.
#include "common.cuh"

namespace gts {

__global__ void argmax_rows_kernel(const float* __restrict__ logits, int64_t ld, int32_t N, int32_t C,
                                   int32_t* __restrict__ cls) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    float best = logits[i * ld];
    int32_t a = 0;
    for (int c = 1; c < C; ++c) {
      const float x = logits[i * ld + c];
      if (x > best) { best = x; a = c; }      // strictly greater: first maximum wins (torch.max)
    }
    cls[i] = a;
  }
}

__global__ void project_nodes_kernel(const int16_t* __restrict__ svs, int64_t n_vox,
                                     const int64_t* __restrict__ node_labels, int32_t N,
                                     int64_t* __restrict__ out, int32_t* __restrict__ err) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_vox; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t s = svs[i];
    int64_t v = 0;
    if (s >= 0 && s < N) v = node_labels[s];
    else if (s != -1) *err = 1;
    out[i] = v;
  }
}

// determine_tumor_crop (data_processing/image_processing.py:8-17) without materialising the voxel predictions:
// a voxel is "tumour" when its supervoxel's class is non-zero (background -1 -> healthy).  The crop the
// reference returns is np.ix_ of the planes that contain a voxel of binary_dilation(mask) (3-D cross, one step,
// border 0) — and a plane contains such a voxel iff it or one of its two neighbour planes contains a tumour
// voxel.  So one pass marks plane occupancy (one warp-level vote, then at most one atomicOr per warp and axis
// plane), the +-1 dilation of three short vectors is host-side arithmetic on 536 flags.
__global__ void __launch_bounds__(256)
plane_occupancy_kernel(const int16_t* __restrict__ svs, int32_t X, int32_t Y, int32_t Z,
                       const int32_t* __restrict__ node_cls, int32_t N,
                       int32_t* __restrict__ occ_x, int32_t* __restrict__ occ_y, int32_t* __restrict__ occ_z,
                       int32_t* __restrict__ err) {
  const int64_t total = (int64_t)X * Y * Z;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t s = svs[i];
    bool tumour = false;
    if (s >= 0 && s < N) tumour = node_cls[s] != 0;
    else if (s != -1) *err = 1;
    if (tumour) {
      const int32_t z = (int32_t)(i % Z);
      const int64_t t = i / Z;
      // idempotent flags: plain stores race benignly (every writer stores 1)
      occ_z[z] = 1;
      occ_y[(int32_t)(t % Y)] = 1;
      occ_x[(int32_t)(t / Y)] = 1;
    }
  }
}

// One thread per 8 consecutive output voxels (one 16-byte store).  The flat
// output index is decomposed once, then walked along z with carries.
__global__ void __launch_bounds__(256)
project_labels_kernel(const int16_t* __restrict__ svs, int32_t X, int32_t Y, int32_t Z,
                      const int32_t* __restrict__ inv_x, const int32_t* __restrict__ inv_y,
                      const int32_t* __restrict__ inv_z, const int32_t* __restrict__ node_cls, int32_t N,
                      const int16_t* __restrict__ lut, int32_t n_lut, int16_t* __restrict__ vol,
                      int32_t VX, int32_t VY, int32_t VZ, int32_t* __restrict__ err) {
  const int64_t total = (int64_t)VX * VY * VZ;
  const int64_t n_groups = (total + 7) / 8;
  for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < n_groups; g += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i0 = g * 8;
    int32_t z = (int32_t)(i0 % VZ);
    int64_t t = i0 / VZ;
    int32_t y = (int32_t)(t % VY);
    int32_t x = (int32_t)(t / VY);
    int16_t o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int16_t val = 0;
      if (i0 + j < total) {
        const int32_t cx = inv_x[x], cy = inv_y[y], cz = inv_z[z];
        if ((cx | cy | cz) >= 0) {
          const int32_t s = svs[((int64_t)cx * Y + cy) * Z + cz];
          if (s >= 0 && s < N) {
            const int32_t c = node_cls[s];
            if (c >= 0 && c < n_lut) val = lut[c];
            else *err = 1;
          } else if (s != -1) {
            *err = 1;
          }
        }
      }
      o[j] = val;
      if (++z == VZ) { z = 0; if (++y == VY) { y = 0; ++x; } }
    }
    if (i0 + 8 <= total) {
      int4 pk;
      pk.x = (uint16_t)o[0] | ((uint32_t)(uint16_t)o[1] << 16);
      pk.y = (uint16_t)o[2] | ((uint32_t)(uint16_t)o[3] << 16);
      pk.z = (uint16_t)o[4] | ((uint32_t)(uint16_t)o[5] << 16);
      pk.w = (uint16_t)o[6] | ((uint32_t)(uint16_t)o[7] << 16);
      stg_na(reinterpret_cast<int4*>(vol + i0), pk);
    } else {
      for (int j = 0; j < 8 && i0 + j < total; ++j) vol[i0 + j] = o[j];
    }
  }
}

__global__ void project_logits_kernel(const int16_t* __restrict__ svs, int64_t n_vox, const float* __restrict__ node_logits,
                                      int64_t ld, int32_t N, int32_t C, const float* __restrict__ bg,
                                      float* __restrict__ out, int32_t* __restrict__ err) {
  const int64_t total = n_vox * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t vx = i / C;
    const int c = (int)(i - vx * C);
    const int32_t s = svs[vx];
    float v;
    if (s >= 0 && s < N) v = node_logits[(int64_t)s * ld + c];
    else { v = bg[c]; if (s != -1) *err = 1; }
    out[i] = v;
  }
}

static inline int pj_grid(int64_t n, int threads) {
  int64_t b = ceil_div<int64_t>(n, threads);
  const int64_t cap = (int64_t)sm_count() * 32;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace gts

using namespace gts;

extern "C" {

int gts_argmax_rows(const float* logits, int64_t ld, int32_t n_nodes, int32_t n_classes,
                    int32_t* cls, gts_stream_t stream) {
  GTS_CHECK_ARG(n_nodes >= 0 && n_classes >= 1, "gts_argmax_rows: bad size");
  if (n_nodes == 0) return GTS_OK;
  GTS_CHECK_ARG(logits && cls, "gts_argmax_rows: null pointer");
  argmax_rows_kernel<<<pj_grid(n_nodes, 256), 256, 0, as_stream(stream)>>>(logits, ld, n_nodes, n_classes, cls);
  GTS_LAUNCH_CHECK();
  return GTS_OK;
}

int gts_project_nodes(const int16_t* svs, int64_t n_vox, const int64_t* node_labels,
                      int32_t n_nodes, int64_t* out, int32_t* err_flag, gts_stream_t stream) {
  GTS_CHECK_ARG(n_vox >= 0 && n_nodes >= 0, "gts_project_nodes: negative size");
  if (n_vox == 0) return GTS_OK;
  GTS_CHECK_ARG(svs && out && err_flag && (node_labels || n_nodes == 0), "gts_project_nodes: null pointer");
  project_nodes_kernel<<<pj_grid(n_vox, 256), 256, 0, as_stream(stream)>>>(svs, n_vox, node_labels, n_nodes, out, err_flag);
  GTS_LAUNCH_CHECK();
  return GTS_OK;
}

int gts_project_labels(const int16_t* svs, int32_t X, int32_t Y, int32_t Z,
                       const int32_t* inv_x, const int32_t* inv_y, const int32_t* inv_z,
                       const int32_t* node_cls, int32_t n_nodes,
                       const int16_t* lut, int32_t n_lut,
                       int16_t* vol, int32_t VX, int32_t VY, int32_t VZ,
                       int32_t* err_flag, gts_stream_t stream) {
  GTS_CHECK_ARG(X >= 0 && Y >= 0 && Z >= 0 && VX >= 0 && VY >= 0 && VZ >= 0 && n_nodes >= 0 && n_lut >= 0,
                "gts_project_labels: negative size");
  const int64_t total = (int64_t)VX * VY * VZ;
  if (total == 0) return GTS_OK;
  GTS_CHECK_ARG(vol && inv_x && inv_y && inv_z && err_flag, "gts_project_labels: null pointer");
  GTS_CHECK_ARG((reinterpret_cast<uintptr_t>(vol) & 15u) == 0, "gts_project_labels: vol must be 16-byte aligned");
  GTS_CHECK_ARG((int64_t)X * Y * Z == 0 || (svs && node_cls && lut), "gts_project_labels: null pointer");
  project_labels_kernel<<<pj_grid((total + 7) / 8, 256), 256, 0, as_stream(stream)>>>(
      svs, X, Y, Z, inv_x, inv_y, inv_z, node_cls, n_nodes, lut, n_lut, vol, VX, VY, VZ, err_flag);
  GTS_LAUNCH_CHECK();
  return GTS_OK;
}

int gts_project_logits(const int16_t* svs, int64_t n_vox, const float* node_logits, int64_t ld,
                       int32_t n_nodes, int32_t n_classes, const float* bg_row, float* out,
                       int32_t* err_flag, gts_stream_t stream) {
  GTS_CHECK_ARG(n_vox >= 0 && n_nodes >= 0 && n_classes >= 1, "gts_project_logits: bad size");
  if (n_vox == 0) return GTS_OK;
  GTS_CHECK_ARG(svs && bg_row && out && err_flag && (node_logits || n_nodes == 0), "gts_project_logits: null pointer");
  project_logits_kernel<<<pj_grid(n_vox * n_classes, 256), 256, 0, as_stream(stream)>>>(svs, n_vox, node_logits, ld, n_nodes,
                                                                                      n_classes, bg_row, out, err_flag);
  GTS_LAUNCH_CHECK();
  return GTS_OK;
}

int gts_tumor_plane_occupancy(const int16_t* svs, int32_t X, int32_t Y, int32_t Z,
                              const int32_t* node_cls, int32_t n_nodes,
                              int32_t* occ_x, int32_t* occ_y, int32_t* occ_z,
                              int32_t* err_flag, gts_stream_t stream) {
  GTS_CHECK_ARG(X >= 0 && Y >= 0 && Z >= 0 && n_nodes >= 0, "gts_tumor_plane_occupancy: negative size");
  GTS_CHECK_ARG(occ_x && occ_y && occ_z && err_flag, "gts_tumor_plane_occupancy: null output");
  cudaStream_t st = as_stream(stream);
  if (X > 0) GTS_CUDA(cudaMemsetAsync(occ_x, 0, sizeof(int32_t) * X, st));
  if (Y > 0) GTS_CUDA(cudaMemsetAsync(occ_y, 0, sizeof(int32_t) * Y, st));
  if (Z > 0) GTS_CUDA(cudaMemsetAsync(occ_z, 0, sizeof(int32_t) * Z, st));
  const int64_t total = (int64_t)X * Y * Z;
  if (total == 0) return GTS_OK;
  GTS_CHECK_ARG(svs && (node_cls || n_nodes == 0), "gts_tumor_plane_occupancy: null input");
  int64_t blocks = ceil_div<int64_t>(total, 256 * 8);
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  plane_occupancy_kernel<<<(int)blocks, 256, 0, st>>>(svs, X, Y, Z, node_cls, n_nodes, occ_x, occ_y, occ_z, err_flag);
  GTS_LAUNCH_CHECK();
  return GTS_OK;
}

}  // extern "C"

// Scene TSDF -> triangle mesh on the GPU (SURVEY.md section 8 f row 3).
//
// Replaces the CPU tail of the reference's mesh export: skimage.measure.marching_cubes(tsdf_vol, level=0) followed by the
// nearest-voxel semantic / instance lookup (utils.py:231-247; called from SaveScene.save_scene_eval / vis_incremental,
// utils.py:362-388, on the dense scene volumes GRUFusion.save_mesh builds, models/gru_fusion.py:217-257).
//
// Three kernels over the dense [dx,dy,dz] volume (default value 1 = unobserved, exactly what the reference hands to skimage):
//   classify : per voxel, the cube index of the cell it anchors -> triangle count; per (voxel, axis) "the surface crosses the
//              grid edge leaving this voxel along +axis" -> one UNIQUE mesh vertex per crossed grid edge (no duplicates, so no
//              welding pass: vertex ids come from a stable compaction of the edge flags);
//   vertices : position by linear interpolation of the zero crossing, normal = normalised central-difference gradient
//              interpolated the same way (pointing towards increasing TSDF, i.e. into free space), semantic / instance label
//              of the nearest voxel (round-half-even, clipped: np.round + np.clip of the reference);
//   faces    : per active cell (ascending raster order) the table's triangles, written at scan offsets -> deterministic order.
// The case table is DERIVED (tools/gen_mc_table.py), not copied: face-consistent disambiguation, hole-free.
#include "common.cuh"
#include "mc_table.cuh"

namespace {

__device__ __forceinline__ float vol_at(const float* __restrict__ v, int dy, int dz, int x, int y, int z) {
  return v[((size_t)x * dy + y) * dz + z];
}

__global__ void __launch_bounds__(256)
mc_classify_kernel(const float* __restrict__ vol, int dx, int dy, int dz, float level, uint8_t* __restrict__ edge_flags,
                   uint8_t* __restrict__ cell_ntri) {
  const long long n = (long long)dx * dy * dz;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int z = (int)(i % dz), y = (int)((i / dz) % dy), x = (int)(i / ((long long)dz * dy));
  const bool in0 = vol[i] < level;
  const bool hx = x + 1 < dx, hy = y + 1 < dy, hz = z + 1 < dz;
  bool in[8];
  in[0] = in0;
#pragma unroll
  for (int c = 1; c < 8; ++c) {
    const int cx = c & 1, cy = (c >> 1) & 1, cz = c >> 2;
    const bool ok = (!cx || hx) && (!cy || hy) && (!cz || hz);
    in[c] = ok ? vol_at(vol, dy, dz, x + cx, y + cy, z + cz) < level : false;
  }
  edge_flags[3 * i + 0] = hx && (in[1] != in0);
  edge_flags[3 * i + 1] = hy && (in[2] != in0);
  edge_flags[3 * i + 2] = hz && (in[4] != in0);
  int idx = 0;
  if (hx && hy && hz) {
#pragma unroll
    for (int c = 0; c < 8; ++c) idx |= (in[c] ? 1 : 0) << c;
  }
  cell_ntri[i] = kMcCount[idx];
}

__device__ __forceinline__ float grad_axis(const float* __restrict__ v, int dx, int dy, int dz, int x, int y, int z, int axis) {
  // np.gradient: central differences inside, one-sided at the borders
  int p[3] = {x, y, z}, q[3] = {x, y, z};
  const int d[3] = {dx, dy, dz};
  float scale = 0.5f;
  if (p[axis] + 1 < d[axis]) p[axis] += 1; else scale = 1.f;
  if (q[axis] > 0) q[axis] -= 1; else scale = 1.f;
  if (d[axis] == 1) return 0.f;
  return (vol_at(v, dy, dz, p[0], p[1], p[2]) - vol_at(v, dy, dz, q[0], q[1], q[2])) * scale;
}

__global__ void __launch_bounds__(256)
mc_vertices_kernel(const float* __restrict__ vol, int dx, int dy, int dz, float level, const int* __restrict__ edge_index,
                   int n_verts, float* __restrict__ verts, float* __restrict__ normals, const int* __restrict__ sem_vol,
                   const int* __restrict__ inst_vol, int* __restrict__ sem_out, int* __restrict__ inst_out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_verts) return;
  const int e = edge_index[t];
  const int axis = e % 3;
  const long long i = e / 3;
  const int z = (int)(i % dz), y = (int)((i / dz) % dy), x = (int)(i / ((long long)dz * dy));
  int x1 = x, y1 = y, z1 = z;
  if (axis == 0) ++x1; else if (axis == 1) ++y1; else ++z1;
  const float v0 = vol[i], v1 = vol_at(vol, dy, dz, x1, y1, z1);
  const float tt = __fdiv_rn(__fsub_rn(level, v0), __fsub_rn(v1, v0));
  float p[3] = {(float)x, (float)y, (float)z};
  p[axis] = __fadd_rn(p[axis], tt);
  verts[3 * t + 0] = p[0]; verts[3 * t + 1] = p[1]; verts[3 * t + 2] = p[2];
  if (normals) {
    float g[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float g0 = grad_axis(vol, dx, dy, dz, x, y, z, a), g1 = grad_axis(vol, dx, dy, dz, x1, y1, z1, a);
      g[a] = __fadd_rn(g0, __fmul_rn(tt, __fsub_rn(g1, g0)));
    }
    const float len = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(g[0], g[0]), __fmul_rn(g[1], g[1])), __fmul_rn(g[2], g[2])));
    const float inv = len > 0.f ? __fdiv_rn(1.f, len) : 0.f;
    normals[3 * t + 0] = __fmul_rn(g[0], inv); normals[3 * t + 1] = __fmul_rn(g[1], inv); normals[3 * t + 2] = __fmul_rn(g[2], inv);
  }
  if (sem_out || inst_out) {
    const int rx = min(max((int)rintf(p[0]), 0), dx - 1), ry = min(max((int)rintf(p[1]), 0), dy - 1),
              rz = min(max((int)rintf(p[2]), 0), dz - 1);
    const size_t r = ((size_t)rx * dy + ry) * dz + rz;
    if (sem_out) sem_out[t] = sem_vol[r];
    if (inst_out) inst_out[t] = inst_vol[r];
  }
}

__global__ void __launch_bounds__(256)
mc_faces_kernel(const float* __restrict__ vol, int dx, int dy, int dz, float level, const int* __restrict__ cell_index,
                const int* __restrict__ tri_offset, int n_cells, const int* __restrict__ edge_pos, int* __restrict__ faces) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_cells) return;
  const long long i = cell_index[t];
  const int z = (int)(i % dz), y = (int)((i / dz) % dy), x = (int)(i / ((long long)dz * dy));
  int idx = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c)
    idx |= (vol_at(vol, dy, dz, x + (c & 1), y + ((c >> 1) & 1), z + (c >> 2)) < level ? 1 : 0) << c;
  const int nt = kMcCount[idx];
  int* out = faces + 3 * (size_t)tri_offset[t];
  for (int k = 0; k < 3 * nt; ++k) {
    const int e = kMcTris[idx][k];
    const int axis = e >> 2, j = e & 3;
    int o[3] = {0, 0, 0};
    o[(axis + 1) % 3] = j & 1;
    o[(axis + 2) % 3] = j >> 1;
    const long long owner = ((long long)(x + o[0]) * dy + (y + o[1])) * dz + (z + o[2]);
    out[k] = edge_pos[3 * owner + axis];
  }
}

}  // namespace

extern "C" {

// vol f32 [dx,dy,dz]; edge_flags uint8 [3*n] (index = 3 * voxel + axis); cell_ntri uint8 [n]
int ep_mc_classify(const float* vol, int dx, int dy, int dz, float level, uint8_t* edge_flags, uint8_t* cell_ntri,
                   cudaStream_t stream) {
  if (dx < 1 || dy < 1 || dz < 1) return EP_ERR_ARG;
  const long long n = (long long)dx * dy * dz;
  if (3 * n > 0x7fffffffLL) return EP_ERR_UNSUPPORTED;
  mc_classify_kernel<<<ep_div_up(n, 256), 256, 0, stream>>>(vol, dx, dy, dz, level, edge_flags, cell_ntri);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

// edge_index int32 [n_verts]: flagged (voxel, axis) ids ascending; verts / normals f32 [n_verts,3] (normals optional);
// sem_vol / inst_vol int32 [dx,dy,dz] with sem_out / inst_out int32 [n_verts] (optional)
int ep_mc_vertices(const float* vol, int dx, int dy, int dz, float level, const int32_t* edge_index, int64_t n_verts, float* verts,
                   float* normals, const int32_t* sem_vol, const int32_t* inst_vol, int32_t* sem_out, int32_t* inst_out,
                   cudaStream_t stream) {
  if (n_verts <= 0 || (sem_out && !sem_vol) || (inst_out && !inst_vol)) return EP_ERR_ARG;
  mc_vertices_kernel<<<ep_div_up(n_verts, 256), 256, 0, stream>>>(vol, dx, dy, dz, level, edge_index, (int)n_verts, verts, normals,
                                                                  sem_vol, inst_vol, sem_out, inst_out);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

// cell_index int32 [n_cells]: cells with >= 1 triangle, ascending; tri_offset int32 [n_cells]: exclusive scan of their triangle
// counts; edge_pos int32 [3*n]: vertex id of every flagged (voxel, axis); faces int32 [n_tris,3]
int ep_mc_faces(const float* vol, int dx, int dy, int dz, float level, const int32_t* cell_index, const int32_t* tri_offset,
                int64_t n_cells, const int32_t* edge_pos, int32_t* faces, cudaStream_t stream) {
  if (n_cells <= 0) return EP_ERR_ARG;
  mc_faces_kernel<<<ep_div_up(n_cells, 256), 256, 0, stream>>>(vol, dx, dy, dz, level, cell_index, tri_offset, (int)n_cells, edge_pos,
                                                               faces);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

}  // extern "C"

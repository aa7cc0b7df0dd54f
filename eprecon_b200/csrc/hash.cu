// Coordinate hashing, hash tables and kernel maps for the sparse 3-D ops.
//
// Replaces the torchsparse v2.0.0 primitives the reference calls through ops/torchsparse_utils.py:15-105
// and inside every spnn.Conv3d (models/modules.py:19,35,50,56,64,90,181): sphash / sphashquery
// (hash_cuda.cu + cuckoo hashmap), spdownsample, the per-conv neighbour search, calc_ti_weights; and the
// indice-pair generation of spconv.SubMConv3d (models/modules.py:252,444).
//
// Design: one open-addressing table (64-bit key -> row id, load factor <= 0.5) per coordinate set, and
// OUTPUT-MAJOR neighbour tables nbr[M_out, K] (-1 = miss) built once per (coordinate set, kernel
// geometry) and shared by every conv on that set -- the conv kernels are then atomics-free gathers.
// Keys are the reference's own 60-bit FNV hash so voxel identity (and torch.unique ordering) match it.
#include "common.cuh"

namespace {

__device__ __forceinline__ void table_insert(uint64_t* __restrict__ tk, int* __restrict__ tv, uint32_t mask,
                                             uint64_t key, int val) {
  uint32_t slot = (uint32_t)ep_mix64(key) & mask;
  while (true) {
    unsigned long long prev = atomicCAS((unsigned long long*)&tk[slot], (unsigned long long)EP_HASH_EMPTY,
                                        (unsigned long long)key);
    if (prev == EP_HASH_EMPTY) { atomicMin(&tv[slot], val); return; }   // tv is pre-set to INT_MAX: a racing duplicate's smaller row survives
    if (prev == key) { atomicMin(&tv[slot], val); return; }  // duplicate key: first (lowest) row wins
    slot = (slot + 1) & mask;
  }
}

__device__ __forceinline__ int table_find(const uint64_t* __restrict__ tk, const int* __restrict__ tv, uint32_t mask,
                                          uint64_t key) {
  uint32_t slot = (uint32_t)ep_mix64(key) & mask;
  while (true) {
    uint64_t k = tk[slot];
    if (k == key) return tv[slot];
    if (k == EP_HASH_EMPTY) return -1;
    slot = (slot + 1) & mask;
  }
}

__global__ void table_clear_kernel(uint64_t* tk, int* tv, int cap) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cap) { tk[i] = EP_HASH_EMPTY; tv[i] = 0x7fffffff; }
}

__global__ void table_build_kernel(const uint64_t* __restrict__ keys, int m, uint64_t* tk, int* tv, uint32_t mask) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) table_insert(tk, tv, mask, keys[i], i);
}

// keys from int coords; batch_first selects (b,x,y,z) vs (x,y,z,b) column order
__global__ void coord_keys_kernel(const int4* __restrict__ coords, int m, int batch_first, uint64_t* __restrict__ keys) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) {
    int4 c = coords[i];
    keys[i] = batch_first ? ep_sphash(c.y, c.z, c.w, c.x) : ep_sphash(c.x, c.y, c.z, c.w);
  }
}

// initial_voxelize front end (ops/torchsparse_utils.py:16-19): new = [C.xyz / vres, b]; key = sphash(floor(new))
// cb > 0: compact Morton keys (common.cuh) -- a point outside the compact range raises *violation and the caller redoes the
// keys in a wider form.
__global__ void point_keys_kernel(const float4* __restrict__ pts, int n, float vres, int spatial,
                                  float4* __restrict__ pts_scaled, uint64_t* __restrict__ keys, int cb, int bb,
                                  int* __restrict__ violation) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    float4 p = pts[i];
    float4 q = make_float4(__fdiv_rn(p.x, vres), __fdiv_rn(p.y, vres), __fdiv_rn(p.z, vres), p.w);
    if (pts_scaled) pts_scaled[i] = q;
    const int x = (int)floorf(q.x), y = (int)floorf(q.y), z = (int)floorf(q.z), b = (int)floorf(q.w);
    // spatial != 0: Morton key -> voxels come out in Z-order (gather locality); identical grouping, only the internal
    // row order differs from the reference's ascending-hash order
    if (cb > 0) {
      const bool ok = ep_morton_compact_ok(x, y, z, b, cb, bb);
      if (!ok) *violation = 1;
      keys[i] = ok ? ep_morton_key_compact(x, y, z, b, cb) : 0ull;
    } else {
      keys[i] = spatial ? ep_morton_key(x, y, z, b) : ep_sphash(x, y, z, b);
    }
  }
}

// voxel coords of each segment = floor(scaled point) of its first member (all members agree)
__global__ void segment_coords_kernel(const float4* __restrict__ pts_scaled, const int* __restrict__ perm,
                                      const int* __restrict__ seg_start, int m, int4* __restrict__ vox) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < m) {
    float4 q = pts_scaled[perm[seg_start[s]]];
    vox[s] = make_int4((int)floorf(q.x), (int)floorf(q.y), (int)floorf(q.z), (int)floorf(q.w));
  }
}

// Generic output-major neighbour table: nbr[j*K + k] = row of (coord_j + off_k) in the table, or -1.
// Optional bounds (spconv spatial_shape) reject neighbours outside [0, shape).
__global__ void kmap_kernel(const int4* __restrict__ out_coords, int m, int batch_first, const int* __restrict__ offsets,
                            int K, const uint64_t* __restrict__ tk, const int* __restrict__ tv, uint32_t mask,
                            int sx, int sy, int sz, int* __restrict__ nbr) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)m * K) return;
  int j = (int)(t / K), k = (int)(t - (long long)j * K);
  int4 c = out_coords[j];
  int x, y, z, b;
  if (batch_first) { b = c.x; x = c.y; y = c.z; z = c.w; } else { x = c.x; y = c.y; z = c.z; b = c.w; }
  x += offsets[3 * k]; y += offsets[3 * k + 1]; z += offsets[3 * k + 2];
  int r = -1;
  bool ok = (sx <= 0) || (x >= 0 && x < sx && y >= 0 && y < sy && z >= 0 && z < sz);
  if (ok) r = table_find(tk, tv, mask, ep_sphash(x, y, z, b));
  nbr[t] = r;
}

// inverse of a k==s strided map as a one-hot neighbour table: up[i, k] = j for the unique (coarse j, offset k) that
// reaches fine row i (table pre-filled with -1) -- the transposed conv then runs as a plain gather-GEMM
__global__ void kmap_inverse_kernel(const int* __restrict__ nbr, int m_out, int K, int* __restrict__ up) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)m_out * K) return;
  int i = nbr[t];
  if (i >= 0) up[(size_t)i * K + (int)(t % K)] = (int)(t / K);
}

// torchsparse spdownsample for k == s (nn/functional/downsample.py): trunc(c / (s*ts)) * (s*ts), then a
// packed (b,x,y,z) key whose ascending order equals torch.unique(dim=0) on [b,x,y,z] rows.
__global__ void down_keys_kernel(const int4* __restrict__ coords, int m, int step, uint64_t* __restrict__ keys, int cb) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) {
    int4 c = coords[i];
    // torch.div(int, float).trunc() * step : exact in fp32 for |c| < 2^24
    int x = (int)truncf(__fdiv_rn((float)c.x, (float)step)) * step;
    int y = (int)truncf(__fdiv_rn((float)c.y, (float)step)) * step;
    int z = (int)truncf(__fdiv_rn((float)c.z, (float)step)) * step;
    // (the reference sorts the coarse sites by (b,x,y,z); the order is internal, Z-order keeps conv tiles compact)
    keys[i] = cb > 0 ? ep_morton_key_compact(x, y, z, c.w, cb) : ep_morton_key(x, y, z, c.w);
  }
}

__global__ void unpack_down_keys_kernel(const uint64_t* __restrict__ keys_sorted, const int* __restrict__ seg_start,
                                        int m, int4* __restrict__ coords, int cb) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < m) {
    int x, y, z, b;
    if (cb > 0) ep_morton_unkey_compact(keys_sorted[seg_start[s]], cb, x, y, z, b);
    else ep_morton_unkey(keys_sorted[seg_start[s]], x, y, z, b);
    coords[s] = make_int4(x, y, z, b);
  }
}

// voxel_to_point front end (ops/torchsparse_utils.py:71-86 + calc_ti_weights): 8 corner rows + trilinear
// weights, zeroed on misses and renormalised by (sum + 1e-8).  pts are the *scaled* float coords.
__global__ void devox_prepare_kernel(const float4* __restrict__ pts, int n, int s, const uint64_t* __restrict__ tk,
                                     const int* __restrict__ tv, uint32_t mask, int* __restrict__ idx,
                                     float* __restrict__ wts) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = pts[i];
  const float fs = (float)s;
  int bx, by, bz;
  float xf, yf, zf;
  if (s == 1) {
    xf = floorf(p.x); yf = floorf(p.y); zf = floorf(p.z);
    bx = (int)xf; by = (int)yf; bz = (int)zf;
  } else {
    float qx = floorf(__fdiv_rn(p.x, fs)), qy = floorf(__fdiv_rn(p.y, fs)), qz = floorf(__fdiv_rn(p.z, fs));
    bx = (int)qx * s; by = (int)qy * s; bz = (int)qz * s;
    xf = __fmul_rn(qx, fs); yf = __fmul_rn(qy, fs); zf = __fmul_rn(qz, fs);
  }
  const int b = (int)p.w;  // .int() truncation of the batch column
  const float xc = __fadd_rn(xf, fs), yc = __fadd_rn(yf, fs), zc = __fadd_rn(zf, fs);
  const float ax[2] = {__fsub_rn(xc, p.x), __fsub_rn(p.x, xf)};
  const float ay[2] = {__fsub_rn(yc, p.y), __fsub_rn(p.y, yf)};
  const float az[2] = {__fsub_rn(zc, p.z), __fsub_rn(p.z, zf)};
  const float inv_s3 = fs * fs * fs;
  float w[8], sum = 0.f;
  int id[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int dx = k >> 2, dy = (k >> 1) & 1, dz = k & 1;  // x-outer, z-inner
    id[k] = table_find(tk, tv, mask, ep_sphash(bx + dx * s, by + dy * s, bz + dz * s, b));
    float wk = __fmul_rn(__fmul_rn(ax[dx], ay[dy]), az[dz]);
    if (s != 1) wk = __fdiv_rn(wk, inv_s3);
    if (id[k] < 0) wk = 0.f;
    w[k] = wk;
    sum = __fadd_rn(sum, wk);
  }
  const float den = __fadd_rn(sum, 1e-8f);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    idx[(size_t)i * 8 + k] = id[k];
    wts[(size_t)i * 8 + k] = __fdiv_rn(w[k], den);
  }
}

// point_to_voxel front end (ops/torchsparse_utils.py:44-50): row of floor(C/s)*s in the voxel table, or -1
__global__ void point_query_kernel(const float4* __restrict__ pts, int n, int s, const uint64_t* __restrict__ tk,
                                   const int* __restrict__ tv, uint32_t mask, int* __restrict__ idx,
                                   uint64_t* __restrict__ idx_as_key, int m_sentinel) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = pts[i];
  int bx, by, bz;
  if (s == 1) { bx = (int)floorf(p.x); by = (int)floorf(p.y); bz = (int)floorf(p.z); }
  else {
    const float fs = (float)s;
    bx = (int)floorf(__fdiv_rn(p.x, fs)) * s; by = (int)floorf(__fdiv_rn(p.y, fs)) * s; bz = (int)floorf(__fdiv_rn(p.z, fs)) * s;
  }
  int r = table_find(tk, tv, mask, ep_sphash(bx, by, bz, (int)p.w));
  idx[i] = r;
  if (idx_as_key) idx_as_key[i] = r < 0 ? (uint64_t)m_sentinel : (uint64_t)r;
}

}  // namespace

extern "C" {

// capacity must be a power of two >= 2*m
int ep_hash_build(const uint64_t* keys, int64_t m, uint64_t* table_keys, int32_t* table_vals, int64_t capacity,
                  cudaStream_t stream) {
  if (m < 0 || capacity < 2 || (capacity & (capacity - 1)) != 0 || capacity < 2 * m) return EP_ERR_ARG;
  table_clear_kernel<<<ep_div_up(capacity, 256), 256, 0, stream>>>(table_keys, table_vals, (int)capacity);
  if (m > 0)
    table_build_kernel<<<ep_div_up(m, 256), 256, 0, stream>>>(keys, (int)m, table_keys, table_vals,
                                                              (uint32_t)(capacity - 1));
  EP_CHECK_LAUNCH();
  return EP_OK;
}

int ep_coord_keys(const int32_t* coords, int64_t m, int batch_first, uint64_t* keys, cudaStream_t stream) {
  if (m <= 0) return EP_ERR_ARG;
  coord_keys_kernel<<<ep_div_up(m, 256), 256, 0, stream>>>((const int4*)coords, (int)m, batch_first, keys);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

int ep_point_keys(const float* pts, int64_t n, float vres, int spatial, float* pts_scaled, uint64_t* keys,
                  cudaStream_t stream) {
  if (n <= 0) return EP_ERR_ARG;
  point_keys_kernel<<<ep_div_up(n, 256), 256, 0, stream>>>((const float4*)pts, (int)n, vres, spatial,
                                                           (float4*)pts_scaled, keys, 0, 0, nullptr);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

// Compact Z-order keys (3 * coord_bits + batch_bits bits, see common.cuh): same grouping and row order as the spatial keys of
// ep_point_keys for voxel coordinates in [-2^(coord_bits-1), 2^(coord_bits-1)) and batch ids below 2^batch_bits; any point
// outside sets *violation (int32, zeroed by the caller) and the keys must be redone in a wider form.
int ep_point_keys_compact(const float* pts, int64_t n, float vres, int coord_bits, int batch_bits, float* pts_scaled,
                          uint64_t* keys, int32_t* violation, cudaStream_t stream) {
  if (n <= 0 || coord_bits < 1 || coord_bits > 16 || batch_bits < 0 || 3 * coord_bits + batch_bits > 63 || !violation) return EP_ERR_ARG;
  point_keys_kernel<<<ep_div_up(n, 256), 256, 0, stream>>>((const float4*)pts, (int)n, vres, 1, (float4*)pts_scaled, keys,
                                                           coord_bits, batch_bits, violation);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

int ep_segment_coords(const float* pts_scaled, const int32_t* perm, const int32_t* seg_start, int64_t m,
                      int32_t* vox_coords, cudaStream_t stream) {
  if (m <= 0) return EP_ERR_ARG;
  segment_coords_kernel<<<ep_div_up(m, 256), 256, 0, stream>>>((const float4*)pts_scaled, perm, seg_start, (int)m,
                                                               (int4*)vox_coords);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

int ep_kmap_build(const int32_t* out_coords, int64_t m_out, int batch_first, const int32_t* offsets, int K,
                  const uint64_t* table_keys, const int32_t* table_vals, int64_t capacity, int sx, int sy, int sz,
                  int32_t* nbr, cudaStream_t stream) {
  if (m_out <= 0 || K < 1) return EP_ERR_ARG;
  long long total = (long long)m_out * K;
  kmap_kernel<<<ep_div_up(total, 256), 256, 0, stream>>>((const int4*)out_coords, (int)m_out, batch_first, offsets, K,
                                                        table_keys, table_vals, (uint32_t)(capacity - 1), sx, sy, sz,
                                                        nbr);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

// up (int32 [m_in, K]) must be pre-filled with -1 by the caller
int ep_kmap_inverse(const int32_t* nbr, int64_t m_out, int K, int32_t* inv, cudaStream_t stream) {
  if (m_out <= 0) return EP_ERR_ARG;
  long long total = (long long)m_out * K;
  kmap_inverse_kernel<<<ep_div_up(total, 256), 256, 0, stream>>>(nbr, (int)m_out, K, inv);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

int ep_down_keys(const int32_t* coords, int64_t m, int step, uint64_t* keys, cudaStream_t stream) {
  if (m <= 0 || step < 1) return EP_ERR_ARG;
  down_keys_kernel<<<ep_div_up(m, 256), 256, 0, stream>>>((const int4*)coords, (int)m, step, keys, 0);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

// coarse-site keys / coordinates in the compact form; the fine coordinates must already be inside the compact range (coarse
// sites are trunc(c / step) * step, never farther from the origin than c)
int ep_down_keys_compact(const int32_t* coords, int64_t m, int step, int coord_bits, uint64_t* keys, cudaStream_t stream) {
  if (m <= 0 || step < 1 || coord_bits < 1 || coord_bits > 16) return EP_ERR_ARG;
  down_keys_kernel<<<ep_div_up(m, 256), 256, 0, stream>>>((const int4*)coords, (int)m, step, keys, coord_bits);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

int ep_down_unpack_compact(const uint64_t* keys_sorted, const int32_t* seg_start, int64_t m, int coord_bits, int32_t* coords,
                           cudaStream_t stream) {
  if (m <= 0 || coord_bits < 1 || coord_bits > 16) return EP_ERR_ARG;
  unpack_down_keys_kernel<<<ep_div_up(m, 256), 256, 0, stream>>>(keys_sorted, seg_start, (int)m, (int4*)coords, coord_bits);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

int ep_down_unpack(const uint64_t* keys_sorted, const int32_t* seg_start, int64_t m, int32_t* coords,
                   cudaStream_t stream) {
  if (m <= 0) return EP_ERR_ARG;
  unpack_down_keys_kernel<<<ep_div_up(m, 256), 256, 0, stream>>>(keys_sorted, seg_start, (int)m, (int4*)coords, 0);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

int ep_devox_prepare(const float* pts_scaled, int64_t n, int stride, const uint64_t* table_keys,
                     const int32_t* table_vals, int64_t capacity, int32_t* idx, float* weights, cudaStream_t stream) {
  if (n <= 0 || stride < 1) return EP_ERR_ARG;
  devox_prepare_kernel<<<ep_div_up(n, 128), 128, 0, stream>>>((const float4*)pts_scaled, (int)n, stride, table_keys,
                                                              table_vals, (uint32_t)(capacity - 1), idx, weights);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

int ep_point_query(const float* pts_scaled, int64_t n, int stride, const uint64_t* table_keys,
                   const int32_t* table_vals, int64_t capacity, int32_t* idx, uint64_t* idx_as_key, int m_sentinel,
                   cudaStream_t stream) {
  if (n <= 0 || stride < 1) return EP_ERR_ARG;
  point_query_kernel<<<ep_div_up(n, 256), 256, 0, stream>>>((const float4*)pts_scaled, (int)n, stride, table_keys,
                                                            table_vals, (uint32_t)(capacity - 1), idx, idx_as_key,
                                                            m_sentinel);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

}  // extern "C"

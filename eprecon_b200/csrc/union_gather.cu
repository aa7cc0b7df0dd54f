// GRU-fusion sparse union: merge the current fragment's sites with the cropped global state WITHOUT the
// reference's two dense [d,d,d,C] feature volumes (models/gru_fusion.py:67-96,321-322; utils.py:176-180,
// 9.7 / 38.9 / 169.9 MB each at the three levels).  Only two int32 row-index volumes [d^3] are touched;
// the union comes out in raster (x,y,z) order, identical to torch.nonzero on the dense boolean volume.
#include "common.cuh"

namespace {

// row test: mode 0 -> any(x != 0) (feature volumes), mode 1 -> any(|x| < 1) (TSDF direct substitute)
__device__ __forceinline__ bool row_active(const float* __restrict__ f, int c, int mode) {
  for (int k = 0; k < c; ++k) {
    float v = f[k];
    if (mode == 0 ? (v != 0.f) : (fabsf(v) < 1.f)) return true;
  }
  return false;
}

// coords: int32 [n, cw] with xyz in columns [c0, c0+3); site = xyz / div - offset; rows outside [0,d)^3 are
// flagged invalid (valid[i] = 0) and skipped.  vol[lin] = i for active in-range rows, -(i+2) for inactive ones.
__global__ void __launch_bounds__(256)
scatter_rows_kernel(const int* __restrict__ coords, int cw, int c0, int n, int div, int ox, int oy, int oz, int dx,
                    int dy, int dz, const float* __restrict__ feat, int ld, int c, int mode, int* __restrict__ vol,
                    uint8_t* __restrict__ valid) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int* p = coords + (size_t)i * cw + c0;
  // floor division (torch.div(..., rounding_mode='floor')) for div > 1
  auto fdiv = [](int a, int b) { int q = a / b; return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q; };
  int x = fdiv(p[0], div) - ox, y = fdiv(p[1], div) - oy, z = fdiv(p[2], div) - oz;
  bool in = x >= 0 && x < dx && y >= 0 && y < dy && z >= 0 && z < dz;
  if (valid) valid[i] = in;
  // active rows store their id (>= 0); inactive in-range rows store -(id+2) so their values can still be gathered
  if (in) vol[((size_t)x * dy + y) * dz + z] = row_active(feat + (size_t)i * ld, c, mode) ? i : -(i + 2);
}

__global__ void __launch_bounds__(256)
union_flags_kernel(const int* __restrict__ va, const int* __restrict__ vb, int n, uint8_t* __restrict__ flags) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = (va[i] >= 0) || (vb && vb[i] >= 0);
}

// site list (linear ids, ascending) -> coords (b, x*scale, y*scale, z*scale) and the two row ids
__global__ void __launch_bounds__(256)
union_sites_kernel(const int* __restrict__ sites, int u, int dy, int dz, int batch, int scale,
                   const int* __restrict__ va, const int* __restrict__ vb, int4* __restrict__ out_coords,
                   int* __restrict__ row_a, int* __restrict__ row_b) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= u) return;
  int lin = sites[i];
  int z = lin % dz, y = (lin / dz) % dy, x = lin / (dz * dy);
  out_coords[i] = make_int4(batch, x * scale, y * scale, z * scale);
  auto decode = [](int v) { return v >= 0 ? v : (v <= -2 ? -(v + 2) : -1); };
  row_a[i] = decode(va[lin]);
  if (row_b) row_b[i] = vb ? decode(vb[lin]) : -1;
}

__global__ void fill_i32_kernel(int* p, int n, int v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace

extern "C" {

int ep_fill_i32(int32_t* p, int64_t n, int32_t v, cudaStream_t stream) {
  if (n <= 0) return EP_ERR_ARG;
  fill_i32_kernel<<<ep_div_up(n, 256), 256, 0, stream>>>(p, (int)n, v);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

int ep_scatter_rows_to_volume(const int32_t* coords, int coord_width, int coord_col0, int64_t n, int div, int ox,
                              int oy, int oz, int dx, int dy, int dz, const float* feat, int ld, int c, int mode,
                              int32_t* vol, uint8_t* valid, cudaStream_t stream) {
  if (n <= 0 || div < 1) return EP_ERR_ARG;
  scatter_rows_kernel<<<ep_div_up(n, 256), 256, 0, stream>>>(coords, coord_width, coord_col0, (int)n, div, ox, oy, oz,
                                                             dx, dy, dz, feat, ld, c, mode, vol, valid);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

int ep_union_flags(const int32_t* vol_a, const int32_t* vol_b, int64_t n, uint8_t* flags, cudaStream_t stream) {
  if (n <= 0) return EP_ERR_ARG;
  union_flags_kernel<<<ep_div_up(n, 256), 256, 0, stream>>>(vol_a, vol_b, (int)n, flags);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

int ep_union_sites(const int32_t* sites, int64_t u, int dy, int dz, int batch, int scale, const int32_t* vol_a,
                   const int32_t* vol_b, int32_t* out_coords, int32_t* row_a, int32_t* row_b, cudaStream_t stream) {
  if (u <= 0) return EP_ERR_ARG;
  union_sites_kernel<<<ep_div_up(u, 256), 256, 0, stream>>>(sites, (int)u, dy, dz, batch, scale, vol_a, vol_b,
                                                            (int4*)out_coords, row_a, row_b);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

}  // extern "C"

// ---- panoptic level alignment (models/neucon_network.py:516-544): keep a coarse voxel only if a finer-level voxel
// survives inside it.  The reference does this with an O(N*M) broadcast compare of coordinate rows; here the finer set
// marks its parents in a byte volume and the coarse set looks itself up.
namespace {
__global__ void mark_parents_kernel(const int4* __restrict__ coords, int n, int step, int dx, int dy, int dz, int bs,
                                    uint8_t* __restrict__ vol) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int4 c = coords[i];  // (b, x, y, z) in finest-voxel units; parent cell index = floor(coord / step)
  int x = c.y / step, y = c.z / step, z = c.w / step;
  if (c.x >= 0 && c.x < bs && x >= 0 && x < dx && y >= 0 && y < dy && z >= 0 && z < dz)
    vol[(((size_t)c.x * dx + x) * dy + y) * dz + z] = 1;
}
__global__ void lookup_marks_kernel(const int4* __restrict__ coords, int n, int step, int dx, int dy, int dz, int bs,
                                    const uint8_t* __restrict__ vol, uint8_t* __restrict__ flags) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int4 c = coords[i];
  int x = c.y / step, y = c.z / step, z = c.w / step;
  bool in = c.x >= 0 && c.x < bs && x >= 0 && x < dx && y >= 0 && y < dy && z >= 0 && z < dz;
  flags[i] = in ? vol[(((size_t)c.x * dx + x) * dy + y) * dz + z] : 0;
}
}  // namespace

extern "C" {
// vol: uint8 [bs, dx, dy, dz], pre-zeroed.  coords are non-negative (fragment-local) voxel indices.
int ep_mark_parents(const int32_t* coords, int64_t n, int step, int dx, int dy, int dz, int bs, uint8_t* vol,
                    cudaStream_t stream) {
  if (n <= 0 || step < 1) return EP_ERR_ARG;
  mark_parents_kernel<<<ep_div_up(n, 256), 256, 0, stream>>>((const int4*)coords, (int)n, step, dx, dy, dz, bs, vol);
  EP_CHECK_LAUNCH();
  return EP_OK;
}
int ep_lookup_marks(const int32_t* coords, int64_t n, int step, int dx, int dy, int dz, int bs, const uint8_t* vol,
                    uint8_t* flags, cudaStream_t stream) {
  if (n <= 0 || step < 1) return EP_ERR_ARG;
  lookup_marks_kernel<<<ep_div_up(n, 256), 256, 0, stream>>>((const int4*)coords, (int)n, step, dx, dy, dz, bs, vol, flags);
  EP_CHECK_LAUNCH();
  return EP_OK;
}
}  // extern "C"

// ---- global-volume merge on the holder rank (BASELINE configs[3]; rule from GRUFusion(direct_substitute=True),
// models/gru_fusion.py:93-94,198-204: a fragment REPLACES every global voxel inside its bounding volume).  Fragments are
// applied in order, so a row of fragment f survives iff no LATER fragment's box contains it: one pass over all rows at
// once instead of the reference's per-fragment mask + cat (O(F * rows) tensor traffic, F dependent steps).
namespace {
constexpr int kMaxMergeFragments = 1024;
__global__ void __launch_bounds__(256)
merge_substitute_flags_kernel(const int4* __restrict__ rows, int n, const int* __restrict__ frag_start,
                              const int* __restrict__ boxes, int F, uint8_t* __restrict__ flags) {
  __shared__ int s_start[kMaxMergeFragments + 1];
  __shared__ int s_box[kMaxMergeFragments * 6];
  for (int i = threadIdx.x; i <= F; i += blockDim.x) s_start[i] = frag_start[i];
  for (int i = threadIdx.x; i < 6 * F; i += blockDim.x) s_box[i] = boxes[i];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int lo = 0, hi = F;            // fragment of row i: last f with s_start[f] <= i
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (s_start[mid] <= i) lo = mid; else hi = mid;
  }
  const int4 r = rows[i];
  bool keep = true;
  for (int f = lo + 1; f < F && keep; ++f) {
    if (s_start[f + 1] == s_start[f]) continue;      // an empty fragment substitutes nothing (the reference skips it)
    const int* b = s_box + 6 * f;
    keep = !(r.x >= b[0] && r.x < b[3] && r.y >= b[1] && r.y < b[4] && r.z >= b[2] && r.z < b[5]);
  }
  flags[i] = keep;
}
}  // namespace

extern "C" {
// rows int32 [n,4] = (x, y, z, tsdf bits) of F fragments back to back; frag_start int32 [F+1]; boxes int32 [F,6]
// (lo xyz inclusive, hi xyz exclusive, global voxel units).  flags[i] = 1 iff row i survives the in-order substitution.
int ep_merge_substitute_flags(const int32_t* rows, int64_t n, const int32_t* frag_start, const int32_t* boxes, int n_fragments,
                              uint8_t* flags, cudaStream_t stream) {
  if (n <= 0 || n_fragments < 1 || n_fragments > kMaxMergeFragments) return EP_ERR_ARG;
  merge_substitute_flags_kernel<<<ep_div_up(n, 256), 256, 0, stream>>>((const int4*)rows, (int)n, frag_start, boxes, n_fragments,
                                                                      flags);
  EP_CHECK_LAUNCH();
  return EP_OK;
}
}  // extern "C"

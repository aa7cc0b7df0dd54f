// Multi-view back-projection for sm_100a: voxel-major gather of channels-last 2D features.
//
// Replaces (reference, zhen6618/EPRecon):
//   models/occupancy_initialization.py:189-261  Back_Project.forward         (mode MEAN)
//   models/occupancy_initialization.py:79-128   init-stage project+variance   (mode MEANVAR)
//   ops/back_project.py:5-80                     legacy back_project           (mode MEAN + zbar)
//
// Arithmetic spec.  Everything that decides an index or a mask
// (world point, projection, perspective divide, grid normalisation, visibility, sample position)
// is evaluated with explicitly NON-FUSED IEEE fp32 ops in a fixed order so that the CPU oracle
// (oracle/restate.py, numpy fp32) reproduces masks and counts bit for bit.  Only the bilinear
// accumulation of feature values uses FMA (float outputs carry a 1e-3 relative tolerance).
//
// Data layout: features are consumed channels-last [V, bs, H, W, C] so one bilinear tap is one
// contiguous C*4-byte read; a warp covers 32/(C/4) consecutive voxels, lanes split the channel
// quads of their voxel, and each lane keeps a float4 accumulator across views.  The feature maps
// (3.5-16.6 MB per level) are L2-resident; HBM traffic is the first touch of the maps plus the
// coords / count / feature streams.
#include "common.cuh"

namespace {

constexpr int kCountThreads = 256;
constexpr int kMaxViews = 32;

struct ProjOut { float gx, gy, z; bool vis; };

// One view of the projection, op for op as the reference evaluates it:
//   im_p = P @ [x,y,z,1]; u = X/Z; v = Y/Z; gx = 2*u/(W-1) - 1; vis = |gx|<=1 & |gy|<=1 & Z>0
// with the 4-term dot product summed left to right.
__device__ __forceinline__ ProjOut project_one(const float* __restrict__ P, float wx, float wy, float wz,
                                               float wm1, float hm1) {
  float X = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(P[0], wx), __fmul_rn(P[1], wy)), __fmul_rn(P[2], wz)), P[3]);
  float Y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(P[4], wx), __fmul_rn(P[5], wy)), __fmul_rn(P[6], wz)), P[7]);
  float Z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(P[8], wx), __fmul_rn(P[9], wy)), __fmul_rn(P[10], wz)), P[11]);
  float u = __fdiv_rn(X, Z);
  float v = __fdiv_rn(Y, Z);
  ProjOut o;
  o.gx = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, u), wm1), 1.f);
  o.gy = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, v), hm1), 1.f);
  o.z = Z;
  o.vis = (fabsf(o.gx) <= 1.f) && (fabsf(o.gy) <= 1.f) && (Z > 0.f);  // NaN -> false
  return o;
}

__device__ __forceinline__ void world_point(int x, int y, int z, float vs, const float* __restrict__ org,
                                            float& wx, float& wy, float& wz) {
  wx = __fadd_rn(__fmul_rn((float)x, vs), org[0]);
  wy = __fadd_rn(__fmul_rn((float)y, vs), org[1]);
  wz = __fadd_rn(__fmul_rn((float)z, vs), org[2]);
}

// ---------------------------------------------------------------------------------------------
// pass 1: visibility bitmask + view count per candidate voxel, per-block survivor counts
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCountThreads)
bp_count_kernel(const int4* __restrict__ coords, int n, const float* __restrict__ origin, float vs,
                const float* __restrict__ kr, int V, int bs, float wm1, float hm1, int min_views,
                float* __restrict__ count, uint32_t* __restrict__ vismask, int* __restrict__ block_count,
                int* __restrict__ valid_per_batch) {
  extern __shared__ float s_kr[];  // [V*bs*16]
  __shared__ int s_scan[33];
  for (int i = threadIdx.x; i < V * bs * 16; i += blockDim.x) s_kr[i] = kr[i];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int keep = 0, b = 0;
  if (i < n) {
    int4 c = coords[i];
    b = c.x;
    float wx, wy, wz;
    world_point(c.y, c.z, c.w, vs, origin + 3 * b, wx, wy, wz);
    uint32_t m = 0;
    for (int v = 0; v < V; ++v) {
      ProjOut p = project_one(s_kr + (v * bs + b) * 16, wx, wy, wz, wm1, hm1);
      m |= (p.vis ? 1u : 0u) << v;
    }
    int cnt = __popc(m);
    count[i] = (float)cnt;
    vismask[i] = m;
    keep = cnt >= min_views;
  }
  int total;
  ep_block_excl_scan(keep, s_scan, &total);
  if (threadIdx.x == 0) block_count[blockIdx.x] = total;
  if (bs == 1) {
    if (threadIdx.x == 0 && total) atomicAdd(valid_per_batch, total);
  } else {
    // batched fragments: one atomic per (warp, batch entry) instead of one per surviving voxel
    const unsigned act = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const unsigned peers = __match_any_sync(act, b);
      if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(valid_per_batch + b, __popc(peers));
    }
  }
}

// exclusive scan of per-block counts (single CTA, loops); also writes the grand total
__global__ void __launch_bounds__(1024) scan_partials_kernel(const int* __restrict__ in, int* __restrict__ out,
                                                               int nblk, int* __restrict__ total_out) {
  __shared__ int s_scan[33];
  int carry = 0;
  for (int base = 0; base < nblk; base += blockDim.x) {
    int i = base + threadIdx.x;
    int v = i < nblk ? in[i] : 0;
    int tot;
    int ex = ep_block_excl_scan(v, s_scan, &tot);
    if (i < nblk) out[i] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0) *total_out = carry;
}

// pass 2: stable compaction of survivors (coords, visibility mask, source row)
__global__ void __launch_bounds__(kCountThreads)
bp_compact_kernel(const int4* __restrict__ coords, const uint32_t* __restrict__ vismask, int n, int min_views,
                  const int* __restrict__ block_offset, int4* __restrict__ out_coords,
                  uint32_t* __restrict__ out_vis, int* __restrict__ out_src) {
  __shared__ int s_scan[33];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t m = 0;
  int keep = 0;
  if (i < n) {
    m = vismask[i];
    keep = __popc(m) >= min_views;
  }
  int total;
  int ex = ep_block_excl_scan(keep, s_scan, &total);
  if (keep) {
    int pos = block_offset[blockIdx.x] + ex;
    out_coords[pos] = coords[i];
    out_vis[pos] = m;
    if (out_src) out_src[pos] = i;
  }
}

// ---------------------------------------------------------------------------------------------
// pass 3: the gather.  MODE 0 = masked mean over views, 1 = masked two-pass variance.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void fma4(float4& a, float w, const float4 v) {
  a.x = fmaf(w, v.x, a.x); a.y = fmaf(w, v.y, a.y); a.z = fmaf(w, v.z, a.z); a.w = fmaf(w, v.w, a.w);
}

// bilinear sample (align_corners=True, zero padding) of one channel quad; ix,iy already un-normalised
__device__ __forceinline__ float4 sample_quad(const float* __restrict__ fmap, int H, int W, int C, int cq,
                                              float ix, float iy) {
  float fx0 = floorf(ix), fy0 = floorf(iy);
  int x0 = (int)fx0, y0 = (int)fy0, x1 = x0 + 1, y1 = y0 + 1;
  float fx1 = fx0 + 1.f, fy1 = fy0 + 1.f;
  float w_nw = (fx1 - ix) * (fy1 - iy), w_ne = (ix - fx0) * (fy1 - iy);
  float w_sw = (fx1 - ix) * (iy - fy0), w_se = (ix - fx0) * (iy - fy0);
  bool xin0 = x0 >= 0 && x0 < W, xin1 = x1 >= 0 && x1 < W, yin0 = y0 >= 0 && y0 < H, yin1 = y1 >= 0 && y1 < H;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 zero = a;
  const float* r0 = fmap + ((size_t)y0 * W) * C + cq * 4;
  const float* r1 = fmap + ((size_t)y1 * W) * C + cq * 4;
  float4 v_nw = (yin0 && xin0) ? ep_ldg4(r0 + (size_t)x0 * C) : zero;
  float4 v_ne = (yin0 && xin1) ? ep_ldg4(r0 + (size_t)x1 * C) : zero;
  float4 v_sw = (yin1 && xin0) ? ep_ldg4(r1 + (size_t)x0 * C) : zero;
  float4 v_se = (yin1 && xin1) ? ep_ldg4(r1 + (size_t)x1 * C) : zero;
  fma4(a, w_nw, v_nw); fma4(a, w_ne, v_ne); fma4(a, w_sw, v_sw); fma4(a, w_se, v_se);
  return a;
}

template <int C, int MODE>
__global__ void __launch_bounds__(256)
bp_gather_kernel(const int4* __restrict__ out_coords, const uint32_t* __restrict__ out_vis, int m,
                 const float* __restrict__ feats /*[V,bs,H,W,C]*/, int V, int bs, int H, int W,
                 const float* __restrict__ origin, float vs, const float* __restrict__ kr,
                 float* __restrict__ out, int ldo, float* __restrict__ zbar) {
  constexpr int G = C / 4;             // lanes per voxel
  constexpr int VPW = 32 / G;          // voxels per warp pass
  constexpr int WARPS = 8;
  extern __shared__ float s_kr[];      // [V*bs*16]
  __shared__ float2 s_grid[WARPS][VPW][kMaxViews];
  __shared__ float s_z[WARPS][VPW][kMaxViews];
  for (int i = threadIdx.x; i < V * bs * 16; i += blockDim.x) s_kr[i] = kr[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / G, cq = lane % G;
  const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
  const size_t plane = (size_t)H * W * C;
  const int ngroups = (m + VPW - 1) / VPW;
  for (int grp = blockIdx.x * WARPS + warp; grp < ngroups; grp += gridDim.x * WARPS) {
    const int base = grp * VPW;
    // phase 1: (voxel, view) pairs spread over the lanes -> un-normalised sample positions
    for (int p = lane; p < VPW * V; p += 32) {
      int s = p / V, v = p - s * V, o = base + s;
      if (o < m) {
        uint32_t msk = out_vis[o];
        if ((msk >> v) & 1u) {
          int4 c = out_coords[o];
          float wx, wy, wz;
          world_point(c.y, c.z, c.w, vs, origin + 3 * c.x, wx, wy, wz);
          ProjOut pr = project_one(s_kr + (v * bs + c.x) * 16, wx, wy, wz, wm1, hm1);
          // ATen grid_sampler unnormalize, align_corners=True: ((g + 1) / 2) * (size - 1)
          float ix = __fmul_rn(__fdiv_rn(__fadd_rn(pr.gx, 1.f), 2.f), wm1);
          float iy = __fmul_rn(__fdiv_rn(__fadd_rn(pr.gy, 1.f), 2.f), hm1);
          s_grid[warp][s][v] = make_float2(ix, iy);
          s_z[warp][s][v] = pr.z;
        }
      }
    }
    __syncwarp();
    const int o = base + sub;
    if (sub < VPW && o < m) {
      const uint32_t msk = out_vis[o];
      const int b = out_coords[o].x;
      const float cntf = (float)max(__popc(msk), 1);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      float zsum = 0.f;
#pragma unroll 3
      for (int v = 0; v < V; ++v) {
        if ((msk >> v) & 1u) {
          float2 g = s_grid[warp][sub][v];
          float4 s = sample_quad(feats + (size_t)(v * bs + b) * plane, H, W, C, cq, g.x, g.y);
          acc.x += s.x; acc.y += s.y; acc.z += s.z; acc.w += s.w;
          zsum += s_z[warp][sub][v];
        }
      }
      float4 mean = make_float4(__fdiv_rn(acc.x, cntf), __fdiv_rn(acc.y, cntf), __fdiv_rn(acc.z, cntf),
                                __fdiv_rn(acc.w, cntf));
      float4 res = mean;
      if (MODE == 1) {
        float4 var = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 3
        for (int v = 0; v < V; ++v) {
          if ((msk >> v) & 1u) {
            float2 g = s_grid[warp][sub][v];
            float4 s = sample_quad(feats + (size_t)(v * bs + b) * plane, H, W, C, cq, g.x, g.y);
            float dx = s.x - mean.x, dy = s.y - mean.y, dz = s.z - mean.z, dw = s.w - mean.w;
            var.x = fmaf(dx, dx, var.x); var.y = fmaf(dy, dy, var.y);
            var.z = fmaf(dz, dz, var.z); var.w = fmaf(dw, dw, var.w);
          }
        }
        res = make_float4(__fdiv_rn(var.x, cntf), __fdiv_rn(var.y, cntf), __fdiv_rn(var.z, cntf),
                          __fdiv_rn(var.w, cntf));
      }
      *reinterpret_cast<float4*>(out + (size_t)o * ldo + cq * 4) = res;
      if (zbar && cq == 0) zbar[o] = __fdiv_rn(zsum, cntf);
    }
    __syncwarp();
  }
}

// generic-channel fallback: one voxel per warp, lanes loop over channel quads (any C % 4 == 0)
template <int MODE>
__global__ void __launch_bounds__(256)
bp_gather_generic_kernel(const int4* __restrict__ out_coords, const uint32_t* __restrict__ out_vis, int m,
                         const float* __restrict__ feats, int C, int V, int bs, int H, int W,
                         const float* __restrict__ origin, float vs, const float* __restrict__ kr,
                         float* __restrict__ out, int ldo, float* __restrict__ zbar) {
  constexpr int WARPS = 8;
  extern __shared__ float s_kr[];
  __shared__ float2 s_grid[WARPS][kMaxViews];
  __shared__ float s_z[WARPS][kMaxViews];
  for (int i = threadIdx.x; i < V * bs * 16; i += blockDim.x) s_kr[i] = kr[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
  const size_t plane = (size_t)H * W * C;
  for (int o = blockIdx.x * WARPS + warp; o < m; o += gridDim.x * WARPS) {
    const uint32_t msk = out_vis[o];
    const int4 c = out_coords[o];
    if (lane < V && ((msk >> lane) & 1u)) {
      float wx, wy, wz;
      world_point(c.y, c.z, c.w, vs, origin + 3 * c.x, wx, wy, wz);
      ProjOut pr = project_one(s_kr + (lane * bs + c.x) * 16, wx, wy, wz, wm1, hm1);
      s_grid[warp][lane] = make_float2(__fmul_rn(__fdiv_rn(__fadd_rn(pr.gx, 1.f), 2.f), wm1),
                                       __fmul_rn(__fdiv_rn(__fadd_rn(pr.gy, 1.f), 2.f), hm1));
      s_z[warp][lane] = pr.z;
    }
    __syncwarp();
    const float cntf = (float)max(__popc(msk), 1);
    for (int cq = lane; cq < C / 4; cq += 32) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int v = 0; v < V; ++v)
        if ((msk >> v) & 1u) {
          float2 g = s_grid[warp][v];
          float4 s = sample_quad(feats + (size_t)(v * bs + c.x) * plane, H, W, C, cq, g.x, g.y);
          acc.x += s.x; acc.y += s.y; acc.z += s.z; acc.w += s.w;
        }
      float4 mean = make_float4(__fdiv_rn(acc.x, cntf), __fdiv_rn(acc.y, cntf), __fdiv_rn(acc.z, cntf),
                                __fdiv_rn(acc.w, cntf));
      float4 res = mean;
      if (MODE == 1) {
        float4 var = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int v = 0; v < V; ++v)
          if ((msk >> v) & 1u) {
            float2 g = s_grid[warp][v];
            float4 s = sample_quad(feats + (size_t)(v * bs + c.x) * plane, H, W, C, cq, g.x, g.y);
            float dx = s.x - mean.x, dy = s.y - mean.y, dz = s.z - mean.z, dw = s.w - mean.w;
            var.x = fmaf(dx, dx, var.x); var.y = fmaf(dy, dy, var.y);
            var.z = fmaf(dz, dz, var.z); var.w = fmaf(dw, dw, var.w);
          }
        res = make_float4(__fdiv_rn(var.x, cntf), __fdiv_rn(var.y, cntf), __fdiv_rn(var.z, cntf),
                          __fdiv_rn(var.w, cntf));
      }
      *reinterpret_cast<float4*>(out + (size_t)o * ldo + cq * 4) = res;
    }
    if (zbar && lane == 0) {
      float zsum = 0.f;
      for (int v = 0; v < V; ++v)
        if ((msk >> v) & 1u) zsum += s_z[warp][v];
      zbar[o] = __fdiv_rn(zsum, cntf);
    }
    __syncwarp();
  }
}

// im_grid [V,M,2] + mask [V,M] (uint8) for callers that want the reference's full 5-tuple
__global__ void bp_grid_kernel(const int4* __restrict__ out_coords, const uint32_t* __restrict__ out_vis, int m,
                               int V, int bs, float wm1, float hm1, const float* __restrict__ origin, float vs,
                               const float* __restrict__ kr, float* __restrict__ im_grid,
                               uint8_t* __restrict__ mask) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)V * m) return;
  int v = (int)(t / m), o = (int)(t - (long long)v * m);
  int4 c = out_coords[o];
  float wx, wy, wz;
  world_point(c.y, c.z, c.w, vs, origin + 3 * c.x, wx, wy, wz);
  ProjOut p = project_one(kr + (v * bs + c.x) * 16, wx, wy, wz, wm1, hm1);
  im_grid[2 * t] = p.gx;
  im_grid[2 * t + 1] = p.gy;
  mask[t] = (out_vis[o] >> v) & 1u;
}

// NCHW -> NHWC for a stack of images: in [n_img, C, HW] -> out [n_img, HW, C]; 32x32 smem tiles
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                           int C, int HW) {
  __shared__ float tile[32][33];
  const int img = blockIdx.z;
  const int hw0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const float* src = in + (size_t)img * C * HW;
  float* dst = out + (size_t)img * C * HW;
#pragma unroll
  for (int r = 0; r < 32; r += 8) {
    int c = c0 + ty + r, hw = hw0 + tx;
    if (c < C && hw < HW) tile[ty + r][tx] = src[(size_t)c * HW + hw];
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 32; r += 8) {
    int hw = hw0 + ty + r, c = c0 + tx;
    if (c < C && hw < HW) dst[(size_t)hw * C + c] = tile[tx][ty + r];
  }
}

// Input side: the reference hands the path a LIST of per-view NCHW maps (models/neuralrecon.py:53-54,
// models/backbone.py:59-77) and stacks them (torch.stack, neucon_network.py:364).  This packs the V separately allocated maps
// [bs, C, H, W] straight into the channels-last buffer [V, bs, H, W, C] the gather consumes: one launch instead of a stack
// copy + a transpose, and each element moves once.
struct ViewPtrs { const float* p[32]; };
__global__ void __launch_bounds__(256) pack_views_nhwc_kernel(ViewPtrs views, float* __restrict__ out, int bs, int C, int HW) {
  __shared__ float tile[32][33];
  const int img = blockIdx.z;                 // v * bs + b
  const int v = img / bs, b = img - v * bs;
  const int hw0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* src = views.p[v] + (size_t)b * C * HW;
  float* dst = out + (size_t)img * C * HW;
#pragma unroll
  for (int r = 0; r < 32; r += 8) {
    int c = c0 + ty + r, hw = hw0 + tx;
    if (c < C && hw < HW) tile[ty + r][tx] = src[(size_t)c * HW + hw];
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 32; r += 8) {
    int hw = hw0 + ty + r, c = c0 + tx;
    if (c < C && hw < HW) dst[(size_t)hw * C + c] = tile[tx][ty + r];
  }
}

// ---------------------------------------------------------------------------------------------
// Single-pass back-projection: projection + visibility + view count + stable compaction (decoupled
// look-back over tiles) + bilinear gather in ONE kernel.  Sample positions computed in the visibility
// phase stay in shared memory for the gather, nothing is recomputed and no intermediate touches HBM.
// ---------------------------------------------------------------------------------------------
constexpr int FT = 256;  // threads per CTA
constexpr int TV = 64;   // candidate voxels per CTA: 4 threads share a voxel's views in the projection phase, and the
                         // gather phase still has 8 warps for <= 64 survivors (latency hiding needs warps, not voxels)
constexpr unsigned long long TS_AGG = 1ULL << 32, TS_PREFIX = 2ULL << 32;

template <int C, int MODE>
__global__ void __launch_bounds__(FT, (C <= 40 ? 5 : 3))
bp_fused_kernel(const int4* __restrict__ coords, int n, const float* __restrict__ origin, float vs,
                const float* __restrict__ kr, int V, int bs, int H, int W, const float* __restrict__ feats,
                int min_views, float* __restrict__ count, int4* __restrict__ out_coords,
                uint32_t* __restrict__ out_vis, int* __restrict__ out_src, float* __restrict__ out, int ldo,
                float* __restrict__ zbar, unsigned long long* __restrict__ tile_state, int* __restrict__ ticket,
                int* __restrict__ totals /*[bs+1]*/, int ntiles) {
  constexpr int G = C / 4, VPW = 32 / G;
  extern __shared__ __align__(16) float s_dyn[];   // [V*bs*16] KRt | float2 s_pos[V][TV] | float s_z[V][TV] | float4 s_out[TV][C/4]
  float* s_kr = s_dyn;
  float2* s_pos = reinterpret_cast<float2*>(s_dyn + ((V * bs * 16 + 3) & ~3));
  float* s_z = reinterpret_cast<float*>(s_pos + V * TV);
  float4* s_out = reinterpret_cast<float4*>(s_z + V * TV);   // [TV][C/4] gathered rows of the tile's survivors
  __shared__ float s_zres[TV];
  __shared__ uint32_t s_mask[TV];
  __shared__ int4 s_coord[TV];
  __shared__ int s_list[TV];
  __shared__ int s_rank[TV];
  __shared__ int s_scan[33];
  __shared__ int s_tile, s_base;
  const int t = threadIdx.x;
  if (t == 0) s_tile = atomicAdd(ticket, 1);   // tiles are numbered in scheduling order -> look-back always progresses
  for (int i = t; i < V * bs * 16; i += FT) s_kr[i] = kr[i];
  if (t < TV) s_mask[t] = 0;
  __syncthreads();
  const int tile = s_tile;
  const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
  // (A lane-per-(voxel, view) variant with a division-free rejection band and a __ballot_sync mask was measured in round 2:
  //  1457 us vs 950 us on the batched probe -- the band only pays when whole warps skip the divisions, and with ~27 % of
  //  the (voxel, view) pairs visible almost every warp still runs them, while the three dependent passes per warp expose the
  //  coordinate-load latency three times.  Kept: the exact power-of-two scaling of the sample position.)
  // ---- phase A: thread (vox = t % TV, part = t / TV) projects views part, part+4, ... of its voxel
  {
    const int vox = t & (TV - 1), part = t / TV;
    const int i = tile * TV + vox;
    if (i < n) {
      const int4 c = coords[i];
      if (part == 0) s_coord[vox] = c;
      float wx, wy, wz;
      world_point(c.y, c.z, c.w, vs, origin + 3 * c.x, wx, wy, wz);
      uint32_t m = 0;
      for (int v = part; v < V; v += FT / TV) {
        ProjOut p = project_one(s_kr + (v * bs + c.x) * 16, wx, wy, wz, wm1, hm1);
        if (p.vis) {
          m |= 1u << v;
          // (g + 1) / 2 is an exact scaling by a power of two: multiplying by 0.5 is bit-identical to the division
          s_pos[v * TV + vox] = make_float2(__fmul_rn(__fmul_rn(__fadd_rn(p.gx, 1.f), 0.5f), wm1),
                                            __fmul_rn(__fmul_rn(__fadd_rn(p.gy, 1.f), 0.5f), hm1));
          if (zbar) s_z[v * TV + vox] = p.z;
        }
      }
      if (m) atomicOr(&s_mask[vox], m);
    }
  }
  __syncthreads();
  // ---- phase B1: survivors of this tile (block scan); the tile's aggregate is published at once so that later tiles
  //      never wait on this one's gather
  int keep = 0;
  uint32_t m = 0;
  int4 c = make_int4(0, 0, 0, 0);
  const int i = tile * TV + t;
  if (t < TV && i < n) {
    m = s_mask[t];
    c = s_coord[t];
    const int cnt = __popc(m);
    count[i] = (float)cnt;
    keep = cnt >= min_views;
  }
  int total;
  const int rank = ep_block_excl_scan(keep, s_scan, &total);
  if (t == 0) {
    volatile unsigned long long* ts = tile_state;
    __threadfence();
    ts[tile] = (tile == 0 ? TS_PREFIX : TS_AGG) | (unsigned long long)(unsigned)total;
  }
  if (bs > 1 && t < TV) {   // warps 0-1 hold the tile's voxels: one atomic per (warp, batch entry), not one per voxel
    const unsigned act = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const unsigned peers = __match_any_sync(act, c.x);
      if ((t & 31) == __ffs(peers) - 1) atomicAdd(totals + c.x, __popc(peers));
    }
  }
  if (keep) s_list[rank] = t;
  if (t < TV) s_rank[t] = keep ? rank : -1;
  __syncthreads();
  // ---- phase C: gather into shared memory (lanes split the channel quads of VPW voxels per warp pass).  The output
  //      row of a survivor needs the tile's global offset, which is looked up AFTER the gather: by then the predecessors
  //      have published their counts and the look-back returns without spinning (the first version looked back first
  //      and parked 7 of 8 warps at a barrier: 36 % of all stall samples, profiles/r01_bp_fused_ncu_summary.txt).
  const int lane = t & 31, warp = t >> 5;
  const int sub = lane / G, cq = lane % G;
  const size_t plane = (size_t)H * W * C;
  const int ngroups = (total + VPW - 1) / VPW;
  for (int g = warp; g < ngroups; g += FT / 32) {
    const int k = g * VPW + sub;
    if (sub < VPW && k < total) {
      const int tv = s_list[k];
      const uint32_t msk = s_mask[tv];
      const int b = bs == 1 ? 0 : s_coord[tv].x;
      const float cntf = (float)max(__popc(msk), 1);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      float zsum = 0.f;
      uint32_t rem = msk;
      while (rem) {                         // two visible views per trip: 8 independent 16-byte taps in flight
        const int v0 = __ffs(rem) - 1;
        rem &= rem - 1;
        const int v1 = rem ? __ffs(rem) - 1 : -1;
        if (v1 >= 0) rem &= rem - 1;
        const float2 g0 = s_pos[v0 * TV + tv];
        const float2 g1 = v1 >= 0 ? s_pos[v1 * TV + tv] : g0;
        float4 s0 = sample_quad(feats + (size_t)(v0 * bs + b) * plane, H, W, C, cq, g0.x, g0.y);
        float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (v1 >= 0) s1 = sample_quad(feats + (size_t)(v1 * bs + b) * plane, H, W, C, cq, g1.x, g1.y);
        acc.x += s0.x; acc.y += s0.y; acc.z += s0.z; acc.w += s0.w;
        acc.x += s1.x; acc.y += s1.y; acc.z += s1.z; acc.w += s1.w;
        if (zbar) { zsum += s_z[v0 * TV + tv]; if (v1 >= 0) zsum += s_z[v1 * TV + tv]; }
      }
      float4 mean = make_float4(__fdiv_rn(acc.x, cntf), __fdiv_rn(acc.y, cntf), __fdiv_rn(acc.z, cntf),
                                __fdiv_rn(acc.w, cntf));
      float4 res = mean;
      if (MODE == 1) {
        float4 var = make_float4(0.f, 0.f, 0.f, 0.f);
        rem = msk;
        while (rem) {
          const int v0 = __ffs(rem) - 1;
          rem &= rem - 1;
          const float2 g0 = s_pos[v0 * TV + tv];
          float4 s0 = sample_quad(feats + (size_t)(v0 * bs + b) * plane, H, W, C, cq, g0.x, g0.y);
          float dx = s0.x - mean.x, dy = s0.y - mean.y, dz = s0.z - mean.z, dw = s0.w - mean.w;
          var.x = fmaf(dx, dx, var.x); var.y = fmaf(dy, dy, var.y);
          var.z = fmaf(dz, dz, var.z); var.w = fmaf(dw, dw, var.w);
        }
        res = make_float4(__fdiv_rn(var.x, cntf), __fdiv_rn(var.y, cntf), __fdiv_rn(var.z, cntf),
                          __fdiv_rn(var.w, cntf));
      }
      s_out[k * G + cq] = res;
      if (zbar && cq == 0) s_zres[k] = __fdiv_rn(zsum, cntf);
    }
  }
  // ---- phase B2: global offset of the tile (warp-parallel decoupled look-back over the predecessors' tile states)
  if (t < 32) {
    volatile unsigned long long* ts = tile_state;
    int prefix = 0;
    int p = tile - 1 - t;
    bool done = tile == 0;
    while (!done) {
      unsigned long long st = TS_PREFIX;            // tiles before 0: an empty prefix
      if (p >= 0) { do { st = ts[p]; } while ((st >> 32) == 0); }
      const unsigned has_prefix = __ballot_sync(0xffffffffu, (st >> 32) == 2);
      const int first = has_prefix ? __ffs(has_prefix) - 1 : 32;
      int contrib = (t <= first) ? (int)(st & 0xffffffffULL) : 0;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, d);
      prefix += contrib;
      done = has_prefix != 0;
      p -= 32;
    }
    if (t == 0) {
      if (tile > 0) { __threadfence(); ts[tile] = TS_PREFIX | (unsigned long long)(unsigned)(prefix + total); }
      s_base = prefix;
      if (tile == ntiles - 1) totals[bs] = prefix + total;
      if (bs == 1 && total) atomicAdd(totals, total);
    }
  }
  __syncthreads();
  // ---- stores: compacted coords / visibility / source row, then the gathered rows
  const int base = s_base;
  if (t < TV) {
    const int r = s_rank[t];
    if (r >= 0) {
      out_coords[base + r] = s_coord[t];
      out_vis[base + r] = s_mask[t];
      if (out_src) out_src[base + r] = tile * TV + t;
    }
  }
  // rows base .. base+total-1 are contiguous in `out` (up to the row pitch): all 256 threads copy quads, coalesced
  for (int e = t; e < total * G; e += FT) {
    const int k = e / G, q = e - k * G;
    *reinterpret_cast<float4*>(out + (size_t)(base + k) * ldo + q * 4) = s_out[e];
  }
  if (zbar)
    for (int k = t; k < total; k += FT) zbar[base + k] = s_zres[k];
}

template <int MODE>
int launch_gather(const int4* oc, const uint32_t* ov, int m, const float* feats, int C, int V, int bs, int H, int W,
                  const float* origin, float vs, const float* kr, float* out, int ldo, float* zbar,
                  cudaStream_t st) {
  const size_t smem = (size_t)V * bs * 16 * sizeof(float);
  auto grid_for = [&](int vpw) {
    long long groups = ((long long)m + vpw - 1) / vpw;
    long long blocks = (groups + 7) / 8;
    long long cap = (long long)EP_NUM_SMS * 8;
    return (int)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
  };
  switch (C) {
    case 24: bp_gather_kernel<24, MODE><<<grid_for(5), 256, smem, st>>>(oc, ov, m, feats, V, bs, H, W, origin, vs, kr, out, ldo, zbar); break;
    case 32: bp_gather_kernel<32, MODE><<<grid_for(4), 256, smem, st>>>(oc, ov, m, feats, V, bs, H, W, origin, vs, kr, out, ldo, zbar); break;
    case 40: bp_gather_kernel<40, MODE><<<grid_for(3), 256, smem, st>>>(oc, ov, m, feats, V, bs, H, W, origin, vs, kr, out, ldo, zbar); break;
    case 80: bp_gather_kernel<80, MODE><<<grid_for(1), 256, smem, st>>>(oc, ov, m, feats, V, bs, H, W, origin, vs, kr, out, ldo, zbar); break;
    default: bp_gather_generic_kernel<MODE><<<grid_for(1), 256, smem, st>>>(oc, ov, m, feats, C, V, bs, H, W, origin, vs, kr, out, ldo, zbar); break;
  }
  EP_CHECK_LAUNCH();
  return EP_OK;
}

}  // namespace

extern "C" {

size_t ep_backproject_workspace_bytes(int64_t n) {
  size_t nblk = (size_t)ep_div_up(n > 0 ? n : 1, kCountThreads);
  return 2 * nblk * sizeof(int) + 256;
}

int ep_backproject_count(const int32_t* coords, int64_t n, const float* origin, float voxel_size,
                         const float* krcam, int n_views, int bs, int feat_h, int feat_w, int min_views,
                         float* count, uint32_t* vismask, int32_t* valid_per_batch, int32_t* n_valid_total,
                         void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (n <= 0 || n_views < 1 || n_views > kMaxViews || bs < 1 || n > 0x7fffffffLL) return EP_ERR_ARG;
  if (workspace_bytes < ep_backproject_workspace_bytes(n)) return EP_ERR_WORKSPACE;
  const int nblk = ep_div_up(n, kCountThreads);
  int* block_count = (int*)workspace;
  int* block_offset = block_count + nblk;
  const size_t smem = (size_t)n_views * bs * 16 * sizeof(float);
  if (smem > 48 * 1024) return EP_ERR_UNSUPPORTED;
  cudaMemsetAsync(valid_per_batch, 0, sizeof(int) * bs, stream);
  bp_count_kernel<<<nblk, kCountThreads, smem, stream>>>((const int4*)coords, (int)n, origin, voxel_size, krcam,
                                                        n_views, bs, (float)(feat_w - 1), (float)(feat_h - 1),
                                                        min_views, count, vismask, block_count, valid_per_batch);
  EP_CHECK_LAUNCH();
  scan_partials_kernel<<<1, 1024, 0, stream>>>(block_count, block_offset, nblk, n_valid_total);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

int ep_backproject_compact(const int32_t* coords, const uint32_t* vismask, int64_t n, int min_views,
                           int32_t* out_coords, uint32_t* out_vis, int32_t* out_src, const void* workspace,
                           cudaStream_t stream) {
  if (n <= 0) return EP_ERR_ARG;
  const int nblk = ep_div_up(n, kCountThreads);
  const int* block_offset = (const int*)workspace + nblk;
  bp_compact_kernel<<<nblk, kCountThreads, 0, stream>>>((const int4*)coords, vismask, (int)n, min_views, block_offset,
                                                        (int4*)out_coords, out_vis, out_src);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

int ep_backproject_gather(const int32_t* out_coords, const uint32_t* out_vis, int64_t m, const float* feats_nhwc,
                          int channels, int n_views, int bs, int feat_h, int feat_w, const float* origin,
                          float voxel_size, const float* krcam, int mode, float* out, int ld_out, float* zbar,
                          cudaStream_t stream) {
  if (m <= 0 || channels % 4 != 0 || ld_out % 4 != 0 || n_views > kMaxViews) return EP_ERR_ARG;
  if ((size_t)n_views * bs * 16 * sizeof(float) > 40 * 1024) return EP_ERR_UNSUPPORTED;
  if (mode == 0)
    return launch_gather<0>((const int4*)out_coords, out_vis, (int)m, feats_nhwc, channels, n_views, bs, feat_h,
                            feat_w, origin, voxel_size, krcam, out, ld_out, zbar, stream);
  if (mode == 1)
    return launch_gather<1>((const int4*)out_coords, out_vis, (int)m, feats_nhwc, channels, n_views, bs, feat_h,
                            feat_w, origin, voxel_size, krcam, out, ld_out, zbar, stream);
  return EP_ERR_ARG;
}

size_t ep_backproject_fused_workspace_bytes(int64_t n) {
  return (size_t)ep_div_up(n > 0 ? n : 1, TV) * sizeof(unsigned long long) + 256;
}

// One launch: count[n], compacted out_coords / out_vis / out_src / out rows (capacity n each) and totals[bs+1]
// (survivors per batch entry, grand total).  channels must be one of 24 / 32 / 40 / 80 (the path's pyramids + the
// fused init map); other widths use the three-pass entry points above.
int ep_backproject_fused(const int32_t* coords, int64_t n, const float* origin, float voxel_size, const float* krcam,
                         int n_views, int bs, int feat_h, int feat_w, const float* feats_nhwc, int channels,
                         int min_views, int mode, float* count, int32_t* out_coords, uint32_t* out_vis,
                         int32_t* out_src, float* out, int ld_out, float* zbar, int32_t* totals, void* workspace,
                         size_t workspace_bytes, cudaStream_t stream) {
  if (n <= 0 || n_views < 1 || n_views > kMaxViews || bs < 1 || n > 0x7fffffffLL || ld_out % 4 != 0) return EP_ERR_ARG;
  if (mode != 0 && mode != 1) return EP_ERR_ARG;
  if (workspace_bytes < ep_backproject_fused_workspace_bytes(n)) return EP_ERR_WORKSPACE;
  const int ntiles = ep_div_up(n, TV);
  unsigned long long* tile_state = (unsigned long long*)workspace;
  int* ticket = (int*)(tile_state + ntiles);
  cudaMemsetAsync(workspace, 0, (size_t)ntiles * sizeof(unsigned long long) + sizeof(int), stream);
  cudaMemsetAsync(totals, 0, sizeof(int) * (bs + 1), stream);
  const size_t smem = (size_t)(((n_views * bs * 16 + 3) & ~3) + n_views * TV * 2 + n_views * TV + TV * channels) * sizeof(float);
  if (smem > 160 * 1024) return EP_ERR_UNSUPPORTED;
#define EP_BP_FUSED(CC, MM)                                                                                           \
  do {                                                                                                                \
    if (smem > 48 * 1024 &&                                                                                           \
        cudaFuncSetAttribute(bp_fused_kernel<CC, MM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024) !=   \
            cudaSuccess)                                                                                              \
      return EP_ERR_CUDA;                                                                                             \
    bp_fused_kernel<CC, MM><<<ntiles, FT, smem, stream>>>((const int4*)coords, (int)n, origin, voxel_size, krcam,     \
                                                          n_views, bs, feat_h, feat_w, feats_nhwc, min_views, count,  \
                                                          (int4*)out_coords, out_vis, out_src, out, ld_out, zbar,     \
                                                          tile_state, ticket, totals, ntiles);                        \
  } while (0)
  if (mode == 0) {
    switch (channels) {
      case 24: EP_BP_FUSED(24, 0); break;
      case 32: EP_BP_FUSED(32, 0); break;
      case 40: EP_BP_FUSED(40, 0); break;
      case 80: EP_BP_FUSED(80, 0); break;
      default: return EP_ERR_UNSUPPORTED;
    }
  } else {
    switch (channels) {
      case 24: EP_BP_FUSED(24, 1); break;
      case 32: EP_BP_FUSED(32, 1); break;
      case 40: EP_BP_FUSED(40, 1); break;
      case 80: EP_BP_FUSED(80, 1); break;
      default: return EP_ERR_UNSUPPORTED;
    }
  }
#undef EP_BP_FUSED
  EP_CHECK_LAUNCH();
  return EP_OK;
}

int ep_backproject_grid(const int32_t* out_coords, const uint32_t* out_vis, int64_t m, int n_views, int bs,
                        int feat_h, int feat_w, const float* origin, float voxel_size, const float* krcam,
                        float* im_grid, uint8_t* mask, cudaStream_t stream) {
  if (m <= 0) return EP_ERR_ARG;
  long long total = (long long)n_views * m;
  bp_grid_kernel<<<ep_div_up(total, 256), 256, 0, stream>>>((const int4*)out_coords, out_vis, (int)m, n_views, bs,
                                                           (float)(feat_w - 1), (float)(feat_h - 1), origin,
                                                           voxel_size, krcam, im_grid, mask);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

// views: host array of n_views device pointers, each a contiguous NCHW map [bs, channels, hw]; out [n_views, bs, hw, channels]
int ep_pack_views_nhwc(const float* const* views, int n_views, int bs, int channels, int hw, float* out, cudaStream_t stream) {
  if (!views || n_views < 1 || n_views > 32 || bs < 1 || channels < 1 || hw < 1) return EP_ERR_ARG;
  ViewPtrs vp;
  for (int v = 0; v < 32; ++v) vp.p[v] = v < n_views ? views[v] : nullptr;
  dim3 grid(ep_div_up(hw, 32), ep_div_up(channels, 32), n_views * bs);
  pack_views_nhwc_kernel<<<grid, 256, 0, stream>>>(vp, out, bs, channels, hw);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

int ep_nchw_to_nhwc(const float* in, float* out, int n_img, int channels, int hw, cudaStream_t stream) {
  if (n_img <= 0 || channels <= 0 || hw <= 0) return EP_ERR_ARG;
  dim3 grid(ep_div_up(hw, 32), ep_div_up(channels, 32), n_img);
  nchw_to_nhwc_kernel<<<grid, 256, 0, stream>>>(in, out, channels, hw);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

}  // extern "C"

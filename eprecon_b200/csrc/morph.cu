// Occupancy-initialisation pruning: sigmoid threshold -> 2x max-pool -> erode(3) -> dilate(3) x2 -> ordered
// compaction, fused into one single-CTA kernel over the coarse grid held in shared memory.
//
// Replaces models/neucon_network.py:264,298-318 (boolean scatter, F.max_pool3d, two F.conv3d-based morphology
// helpers :216-228, torch.nonzero): 7 library launches + 1 host sync on a 24^3 volume.
#include "common.cuh"

namespace {

// fine[src[r]] = sigmoid(logit[r]) > thr   (fine volume pre-zeroed)
__global__ void __launch_bounds__(256)
scatter_selected_kernel(const float* __restrict__ logit, int ld, const int* __restrict__ src, int n, float thr,
                        uint8_t* __restrict__ fine) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  float s = 1.f / (1.f + expf(-logit[(size_t)r * ld]));
  fine[src[r]] = s > thr;
}

__global__ void __launch_bounds__(1024)
init_prune_kernel(const uint8_t* __restrict__ fine, int bs, int D /*coarse dim*/, int out_scale,
                  int4* __restrict__ out_coords, int* __restrict__ out_count /*[bs+1]: per batch, total*/) {
  extern __shared__ uint8_t sm[];
  __shared__ int s_scan[33];
  const int n = D * D * D, F = 2 * D;
  uint8_t* A = sm;
  uint8_t* B = sm + n;
  int written = 0;
  for (int b = 0; b < bs; ++b) {
    const uint8_t* fv = fine + (size_t)b * F * F * F;
    // max-pool 2x2x2
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      int x = i / (D * D), y = (i / D) % D, z = i % D;
      uint8_t v = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        v |= fv[((size_t)(2 * x + (k >> 2)) * F + (2 * y + ((k >> 1) & 1))) * F + (2 * z + (k & 1))];
      A[i] = v;
    }
    __syncthreads();
    // erode: all 27 cells of the zero-padded neighbourhood set
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      int x = i / (D * D), y = (i / D) % D, z = i % D;
      bool all = x > 0 && x < D - 1 && y > 0 && y < D - 1 && z > 0 && z < D - 1;
      if (all)
        for (int dx = -1; dx <= 1 && all; ++dx)
          for (int dy = -1; dy <= 1 && all; ++dy)
            for (int dz = -1; dz <= 1; ++dz)
              if (!A[((x + dx) * D + (y + dy)) * D + (z + dz)]) { all = false; break; }
      B[i] = all;
    }
    __syncthreads();
    // dilate twice (B -> A -> B)
    for (int pass = 0; pass < 2; ++pass) {
      uint8_t* src = pass == 0 ? B : A;
      uint8_t* dst = pass == 0 ? A : B;
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int x = i / (D * D), y = (i / D) % D, z = i % D;
        bool any = false;
        for (int dx = -1; dx <= 1 && !any; ++dx)
          for (int dy = -1; dy <= 1 && !any; ++dy)
            for (int dz = -1; dz <= 1; ++dz) {
              int xx = x + dx, yy = y + dy, zz = z + dz;
              if (xx >= 0 && xx < D && yy >= 0 && yy < D && zz >= 0 && zz < D && src[(xx * D + yy) * D + zz]) { any = true; break; }
            }
        dst[i] = any;
      }
      __syncthreads();
    }
    // ordered compaction in raster order (== torch.nonzero)
    int batch_count = 0;
    for (int base = 0; base < n; base += blockDim.x) {
      int i = base + threadIdx.x;
      int f = (i < n) ? B[i] : 0;
      int tot;
      int ex = ep_block_excl_scan(f, s_scan, &tot);
      if (f) {
        int x = i / (D * D), y = (i / D) % D, z = i % D;
        out_coords[written + batch_count + ex] = make_int4(b, x * out_scale, y * out_scale, z * out_scale);
      }
      batch_count += tot;
    }
    if (threadIdx.x == 0) out_count[b] = batch_count;
    written += batch_count;
    __syncthreads();
  }
  if (threadIdx.x == 0) out_count[bs] = written;
}

}  // namespace

extern "C" {

int ep_scatter_selected(const float* logit, int ld, const int32_t* src, int64_t n, float thr, uint8_t* fine,
                        cudaStream_t stream) {
  if (n <= 0) return EP_ERR_ARG;
  scatter_selected_kernel<<<ep_div_up(n, 256), 256, 0, stream>>>(logit, ld, src, (int)n, thr, fine);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

// fine: uint8 [bs, 2D, 2D, 2D]; out_coords int32 [<= bs*D^3, 4] = (b, x*s, y*s, z*s); out_count int32[bs+1]
int ep_init_prune(const uint8_t* fine, int bs, int coarse_dim, int out_scale, int32_t* out_coords, int32_t* out_count,
                  cudaStream_t stream) {
  if (bs < 1 || coarse_dim < 1) return EP_ERR_ARG;
  const size_t smem = 2 * (size_t)coarse_dim * coarse_dim * coarse_dim;
  if (smem > 200 * 1024) return EP_ERR_UNSUPPORTED;
  if (cudaFuncSetAttribute(init_prune_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
    return EP_ERR_CUDA;
  init_prune_kernel<<<1, 1024, smem, stream>>>(fine, bs, coarse_dim, out_scale, (int4*)out_coords, out_count);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

}  // extern "C"

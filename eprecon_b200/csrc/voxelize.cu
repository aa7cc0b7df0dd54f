// Point <-> voxel feature transfer.
//
// Replaces torchsparse v2.0.0 spvoxelize (atomicAdd mean pooling) and spdevoxelize (8-tap weighted gather)
// as called from ops/torchsparse_utils.py:27,58,86,98.  The mean pooling here is a segmented sum over the
// CSR built by ep_sort_segments (points of a voxel in ascending point order) -> deterministic, no float atomics.
#include "common.cuh"

namespace {

// out[s, :] = sum_{p in segment s} feat[perm[p], :] / count(s)     (each term divided first, as the reference does)
__global__ void __launch_bounds__(256)
segment_mean_kernel(const float* __restrict__ feat, int ld_in, int c, const int* __restrict__ perm,
                    const int* __restrict__ seg_start, const int* __restrict__ seg_end, int m,
                    float* __restrict__ out, int ld_out) {
  const int cq_n = (c + 3) / 4;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)m * cq_n) return;
  const int s = (int)(t / cq_n), cq = (int)(t - (long long)s * cq_n);
  const int a = seg_start[s], b = seg_end[s];
  const float cnt = (float)(b - a);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int p = a; p < b; ++p) {
    const float4 v = *reinterpret_cast<const float4*>(feat + (size_t)perm[p] * ld_in + cq * 4);
    acc.x += __fdiv_rn(v.x, cnt); acc.y += __fdiv_rn(v.y, cnt); acc.z += __fdiv_rn(v.z, cnt); acc.w += __fdiv_rn(v.w, cnt);
  }
  *reinterpret_cast<float4*>(out + (size_t)s * ld_out + cq * 4) = acc;
}

// out[i, :] = sum_k w[i,k] * feat[idx[i,k], :]  (+ add[i, :])
__global__ void __launch_bounds__(256)
devoxelize_kernel(const float* __restrict__ feat, int ld_in, int c, const int* __restrict__ idx,
                  const float* __restrict__ wts, int n, const float* __restrict__ add, int ld_add,
                  float* __restrict__ out, int ld_out) {
  const int cq_n = (c + 3) / 4;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * cq_n) return;
  const int i = (int)(t / cq_n), cq = (int)(t - (long long)i * cq_n);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int r = idx[(size_t)i * 8 + k];
    if (r >= 0) {
      const float w = wts[(size_t)i * 8 + k];
      const float4 v = *reinterpret_cast<const float4*>(feat + (size_t)r * ld_in + cq * 4);
      acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y); acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
    }
  }
  if (add) {
    const float4 v = *reinterpret_cast<const float4*>(add + (size_t)i * ld_add + cq * 4);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  *reinterpret_cast<float4*>(out + (size_t)i * ld_out + cq * 4) = acc;
}

}  // namespace

extern "C" {

// all leading dimensions must be multiples of 4 and cover c rounded up to 4 (padding columns are carried along)
int ep_segment_mean(const float* feat, int ld_in, int c, const int32_t* perm, const int32_t* seg_start,
                    const int32_t* seg_end, int64_t m, float* out, int ld_out, cudaStream_t stream) {
  if (m <= 0 || c < 1 || ld_in % 4 || ld_out % 4) return EP_ERR_ARG;
  long long total = (long long)m * ((c + 3) / 4);
  segment_mean_kernel<<<ep_div_up(total, 256), 256, 0, stream>>>(feat, ld_in, c, perm, seg_start, seg_end, (int)m, out, ld_out);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

int ep_devoxelize(const float* feat, int ld_in, int c, const int32_t* idx, const float* weights, int64_t n,
                  const float* add, int ld_add, float* out, int ld_out, cudaStream_t stream) {
  if (n <= 0 || c < 1 || ld_in % 4 || ld_out % 4 || (add && ld_add % 4)) return EP_ERR_ARG;
  long long total = (long long)n * ((c + 3) / 4);
  devoxelize_kernel<<<ep_div_up(total, 256), 256, 0, stream>>>(feat, ld_in, c, idx, weights, (int)n, add, ld_add, out, ld_out);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

}  // extern "C"

// Row-wise / element-wise kernels around the sparse convs: BatchNorm apply (+residual, +ReLU), LayerNorm
// variants, ConvGRU gates, row gathers / concats, coordinate transforms, x8 upsampling.
//
// Reference call sites: spnn.BatchNorm / ReLU / residual add (models/modules.py:23-24,54,60,65,71-73),
// LayerNorm+ReLU (modules.py:288-309,447-450,476-480; occupancy_initialization.py:143-168),
// ConvGRU gate math (modules.py:214-221), torch.cat feature concats (neucon_network.py:378,404),
// world->aligned-camera coords (neucon_network.py:387-398; gru_fusion.py:331-337), NeuConNet.upsample (:193-214).
#include "common.cuh"

namespace {

// y = act( a*sa + ta  [+ b*sb + tb] ),  per-column scale/shift optional (NULL = identity)
__global__ void __launch_bounds__(256)
affine_act_kernel(const float* __restrict__ a, int ld_a, const float* __restrict__ ssa /*[2,c] or null*/,
                  const float* __restrict__ b, int ld_b, const float* __restrict__ ssb, int relu, long long m, int c,
                  float* __restrict__ out, int ld_out) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m * c) return;
  const long long i = t / c;
  const int col = (int)(t - i * c);
  float v = a[i * ld_a + col];
  if (ssa) v = fmaf(v, ssa[col], ssa[c + col]);
  if (b) {
    float w = b[i * ld_b + col];
    if (ssb) w = fmaf(w, ssb[col], ssb[c + col]);
    v += w;
  }
  if (relu) v = fmaxf(v, 0.f);
  out[i * ld_out + col] = v;
}

// one warp per row: v = x; relu_before -> v = max(v,0); res -> v += res; LN(v)*g + b; relu_after
template <int MAXQ>  // max columns per lane
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, int ld_x, const float* __restrict__ res, int ld_res, int relu_before,
                 const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int relu_after, int m,
                 int c, float* __restrict__ out, int ld_out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= m) return;
  float v[MAXQ];
  float s = 0.f;
#pragma unroll
  for (int q = 0; q < MAXQ; ++q) {
    const int col = lane + 32 * q;
    float t = 0.f;
    if (col < c) {
      t = x[(size_t)warp * ld_x + col];
      if (relu_before) t = fmaxf(t, 0.f);
      if (res) t += res[(size_t)warp * ld_res + col];
      s += t;
    }
    v[q] = t;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  const float mean = s / (float)c;
  float q2 = 0.f;
#pragma unroll
  for (int q = 0; q < MAXQ; ++q) {
    const int col = lane + 32 * q;
    if (col < c) { float d = v[q] - mean; q2 = fmaf(d, d, q2); }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) q2 += __shfl_xor_sync(0xffffffffu, q2, d);
  const float inv = rsqrtf(q2 / (float)c + eps);
#pragma unroll
  for (int q = 0; q < MAXQ; ++q) {
    const int col = lane + 32 * q;
    if (col < c) {
      float y = (v[q] - mean) * inv;
      y = fmaf(y, gamma ? gamma[col] : 1.f, beta ? beta[col] : 0.f);
      if (relu_after) y = fmaxf(y, 0.f);
      out[(size_t)warp * ld_out + col] = y;
    }
  }
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// out[:, 0:c] = sigmoid(r_pre) * h ; out[:, c:2c] = x     (ConvGRU: x.F = cat([r*h, x]))
__global__ void __launch_bounds__(256)
gru_rh_kernel(const float* __restrict__ r_pre, int ld_r, const float* __restrict__ h, int ld_h,
              const float* __restrict__ x, int ld_x, long long m, int c, float* __restrict__ out, int ld_out) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m * c) return;
  const long long i = t / c;
  const int col = (int)(t - i * c);
  out[i * ld_out + col] = sigmoidf_(r_pre[i * ld_r + col]) * h[i * ld_h + col];
  out[i * ld_out + c + col] = x[i * ld_x + col];
}

// h' = (1 - z) * h + z * tanh(q_pre),  z = sigmoid(z_pre)
__global__ void __launch_bounds__(256)
gru_out_kernel(const float* __restrict__ z_pre, int ld_z, const float* __restrict__ q_pre, int ld_q,
               const float* __restrict__ h, int ld_h, long long m, int c, float* __restrict__ out, int ld_out) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m * c) return;
  const long long i = t / c;
  const int col = (int)(t - i * c);
  const float z = sigmoidf_(z_pre[i * ld_z + col]);
  const float hv = h[i * ld_h + col];
  out[i * ld_out + col] = (1.f - z) * hv + z * tanhf(q_pre[i * ld_q + col]);
}

// out[i, 0:c] = src[(index ? index[i] >> shift : i), 0:c]; rows with index < 0 get `fill`
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ src, int ld_src, const int* __restrict__ index, int shift, float fill,
                   long long m, int c, float* __restrict__ out, int ld_out) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m * c) return;
  const long long i = t / c;
  const int col = (int)(t - i * c);
  long long r = i;
  if (index) { int v = index[i]; r = v < 0 ? -1 : (v >> shift); }
  out[i * ld_out + col] = r < 0 ? fill : src[r * ld_src + col];
}

__global__ void __launch_bounds__(256)
gather_rows_i32x4_kernel(const int4* __restrict__ src, const int* __restrict__ index, int m, int4* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) out[i] = src[index[i]];
}

// (b,x,y,z) int voxel index -> aligned-camera float point (x,y,z,b):
//   w = idx*vs + origin[b];  r = ((R0*wx + R1*wy) + R2*wz) + R3      (non-fused, left to right)
__global__ void __launch_bounds__(256)
aligned_coords_kernel(const int4* __restrict__ coords, int n, const float* __restrict__ origin, float vs,
                      const float* __restrict__ w2ac /*[bs,4,4]*/, int zero_batch, float4* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int4 c = coords[i];
  const float* o = origin + 3 * c.x;
  const float* R = w2ac + 16 * c.x;
  float wx = __fadd_rn(__fmul_rn((float)c.y, vs), o[0]);
  float wy = __fadd_rn(__fmul_rn((float)c.z, vs), o[1]);
  float wz = __fadd_rn(__fmul_rn((float)c.w, vs), o[2]);
  float4 r;
  r.x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[0], wx), __fmul_rn(R[1], wy)), __fmul_rn(R[2], wz)), R[3]);
  r.y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[4], wx), __fmul_rn(R[5], wy)), __fmul_rn(R[6], wz)), R[7]);
  r.z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[8], wx), __fmul_rn(R[9], wy)), __fmul_rn(R[10], wz)), R[11]);
  r.w = zero_batch ? 0.f : (float)c.x;
  out[i] = r;
}

// children of each voxel in the reference's order: self, +x, +y, +z, +xy, +xz, +yz, +xyz
__global__ void __launch_bounds__(256)
upsample8_kernel(const int4* __restrict__ coords, int n, int interval, int4* __restrict__ out) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * 8) return;
  int4 c = coords[t >> 3];
  const int k = t & 7;
  const int dx = (k == 1 || k == 4 || k == 5 || k == 7), dy = (k == 2 || k == 4 || k == 6 || k == 7),
            dz = (k == 3 || k == 5 || k == 6 || k == 7);
  out[t] = make_int4(c.x, c.y + dx * interval, c.z + dy * interval, c.w + dz * interval);
}

// flags[i] = x[i*ld] > thr   (mode 0)   |   sigmoid(x[i*ld]) > thr   (mode 1)
__global__ void __launch_bounds__(256)
threshold_flags_kernel(const float* __restrict__ x, int ld, int n, float thr, int mode, uint8_t* __restrict__ flags) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = x[(size_t)i * ld];
  if (mode == 1) v = sigmoidf_(v);
  flags[i] = v > thr;
}

}  // namespace

extern "C" {

int ep_affine_act(const float* a, int ld_a, const float* ss_a, const float* b, int ld_b, const float* ss_b, int relu,
                  int64_t m, int c, float* out, int ld_out, cudaStream_t stream) {
  if (m <= 0 || c < 1) return EP_ERR_ARG;
  affine_act_kernel<<<ep_div_up(m * c, 256), 256, 0, stream>>>(a, ld_a, ss_a, b, ld_b, ss_b, relu, m, c, out, ld_out);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

int ep_layernorm(const float* x, int ld_x, const float* res, int ld_res, int relu_before, const float* gamma,
                 const float* beta, float eps, int relu_after, int64_t m, int c, float* out, int ld_out,
                 cudaStream_t stream) {
  if (m <= 0 || c < 1 || c > 1024) return EP_ERR_ARG;
  const int blocks = ep_div_up(m * 32, 256);
  if (c <= 32) layernorm_kernel<1><<<blocks, 256, 0, stream>>>(x, ld_x, res, ld_res, relu_before, gamma, beta, eps, relu_after, (int)m, c, out, ld_out);
  else if (c <= 64) layernorm_kernel<2><<<blocks, 256, 0, stream>>>(x, ld_x, res, ld_res, relu_before, gamma, beta, eps, relu_after, (int)m, c, out, ld_out);
  else if (c <= 128) layernorm_kernel<4><<<blocks, 256, 0, stream>>>(x, ld_x, res, ld_res, relu_before, gamma, beta, eps, relu_after, (int)m, c, out, ld_out);
  else if (c <= 512) layernorm_kernel<16><<<blocks, 256, 0, stream>>>(x, ld_x, res, ld_res, relu_before, gamma, beta, eps, relu_after, (int)m, c, out, ld_out);
  else layernorm_kernel<32><<<blocks, 256, 0, stream>>>(x, ld_x, res, ld_res, relu_before, gamma, beta, eps, relu_after, (int)m, c, out, ld_out);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

int ep_gru_rh(const float* r_pre, int ld_r, const float* h, int ld_h, const float* x, int ld_x, int64_t m, int c,
              float* out, int ld_out, cudaStream_t stream) {
  if (m <= 0 || c < 1) return EP_ERR_ARG;
  gru_rh_kernel<<<ep_div_up(m * c, 256), 256, 0, stream>>>(r_pre, ld_r, h, ld_h, x, ld_x, m, c, out, ld_out);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

int ep_gru_out(const float* z_pre, int ld_z, const float* q_pre, int ld_q, const float* h, int ld_h, int64_t m, int c,
               float* out, int ld_out, cudaStream_t stream) {
  if (m <= 0 || c < 1) return EP_ERR_ARG;
  gru_out_kernel<<<ep_div_up(m * c, 256), 256, 0, stream>>>(z_pre, ld_z, q_pre, ld_q, h, ld_h, m, c, out, ld_out);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

int ep_gather_rows(const float* src, int ld_src, const int32_t* index, int shift, float fill, int64_t m, int c,
                   float* out, int ld_out, cudaStream_t stream) {
  if (m <= 0 || c < 1) return EP_ERR_ARG;
  gather_rows_kernel<<<ep_div_up(m * c, 256), 256, 0, stream>>>(src, ld_src, index, shift, fill, m, c, out, ld_out);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

int ep_gather_coords(const int32_t* src, const int32_t* index, int64_t m, int32_t* out, cudaStream_t stream) {
  if (m <= 0) return EP_ERR_ARG;
  gather_rows_i32x4_kernel<<<ep_div_up(m, 256), 256, 0, stream>>>((const int4*)src, index, (int)m, (int4*)out);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

int ep_aligned_coords(const int32_t* coords, int64_t n, const float* origin, float voxel_size, const float* w2ac,
                      int zero_batch, float* out, cudaStream_t stream) {
  if (n <= 0) return EP_ERR_ARG;
  aligned_coords_kernel<<<ep_div_up(n, 256), 256, 0, stream>>>((const int4*)coords, (int)n, origin, voxel_size, w2ac,
                                                               zero_batch, (float4*)out);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

int ep_upsample8(const int32_t* coords, int64_t n, int interval, int32_t* out, cudaStream_t stream) {
  if (n <= 0) return EP_ERR_ARG;
  upsample8_kernel<<<ep_div_up(n * 8, 256), 256, 0, stream>>>((const int4*)coords, (int)n, interval, (int4*)out);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

int ep_threshold_flags(const float* x, int ld, int64_t n, float thr, int mode, uint8_t* flags, cudaStream_t stream) {
  if (n <= 0) return EP_ERR_ARG;
  threshold_flags_kernel<<<ep_div_up(n, 256), 256, 0, stream>>>(x, ld, (int)n, thr, mode, flags);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

}  // extern "C"

// ---- train-mode BatchNorm2d (+ residual-before, + ReLU) for the dense 2-D fusion (models/modules.py:313-399,
// models/occupancy_initialization.py:41-58): statistics over (N, H, W) of an NCHW map.  Two launches per layer instead of
// ATen's var_mean + rsqrt + mul + sub + addcmul + relu chain (8 launches inside the captured graph).
namespace {
constexpr int BN2D_CHUNKS = 16;   // partial sums per channel: fixed count, fixed order -> deterministic

// x (+ relu_pre?  y = relu(x) + res : x).  part[c][chunk] = (sum, sumsq) in double over the chunk's elements
__global__ void __launch_bounds__(256)
bn2d_stats_kernel(const float* __restrict__ x, const float* __restrict__ res, int relu_pre, int n_img, int c, int hw,
                  double2* __restrict__ part) {
  const int ch = blockIdx.x, chunk = blockIdx.y;
  const long long total = (long long)n_img * hw;
  const long long per = (total + BN2D_CHUNKS - 1) / BN2D_CHUNKS;
  const long long beg = chunk * per, end = min(total, beg + per);
  double s = 0.0, q = 0.0;
  for (long long e = beg + threadIdx.x; e < end; e += blockDim.x) {
    const long long img = e / hw, p = e - img * hw;
    const size_t off = ((size_t)img * c + ch) * hw + p;
    float v = x[off];
    if (relu_pre) v = fmaxf(v, 0.f);
    if (res) v += res[off];
    s += v;
    q += (double)v * v;
  }
  __shared__ double ss[256], sq[256];
  ss[threadIdx.x] = s; sq[threadIdx.x] = q;
  __syncthreads();
  for (int d = 128; d > 0; d >>= 1) {
    if ((int)threadIdx.x < d) { ss[threadIdx.x] += ss[threadIdx.x + d]; sq[threadIdx.x] += sq[threadIdx.x + d]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) part[ch * BN2D_CHUNKS + chunk] = make_double2(ss[0], sq[0]);
}

// one CTA = one (image, channel) plane chunk: the channel's scale / shift are reduced ONCE per CTA from the 16 partials
// (first version: every element re-read the 16 double2 partials -- 20 us per layer, slower than the ATen chain it replaced)
__global__ void __launch_bounds__(256)
bn2d_apply_kernel(const float* __restrict__ x, const float* __restrict__ res, int relu_pre, int n_img, int c, int hw,
                  const double2* __restrict__ part, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                  int relu_post, float* __restrict__ out) {
  __shared__ float s_ss[2];
  const int plane = blockIdx.y;                 // img * c + ch
  const int ch = plane % c;
  if (threadIdx.x == 0) {
    double s = 0.0, q = 0.0;
#pragma unroll
    for (int k = 0; k < BN2D_CHUNKS; ++k) { const double2 p = part[ch * BN2D_CHUNKS + k]; s += p.x; q += p.y; }
    const double cnt = (double)n_img * hw;
    const double mean = s / cnt;
    const double var = fmax(q / cnt - mean * mean, 0.0);
    const float scale = gamma[ch] * rsqrtf((float)var + eps);
    s_ss[0] = scale;
    s_ss[1] = beta[ch] - (float)mean * scale;
  }
  __syncthreads();
  const float scale = s_ss[0], shift = s_ss[1];
  const size_t base = (size_t)plane * hw;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += gridDim.x * blockDim.x) {
    float v = x[base + p];
    if (relu_pre) v = fmaxf(v, 0.f);
    if (res) v += res[base + p];
    v = fmaf(v, scale, shift);
    if (relu_post) v = fmaxf(v, 0.f);
    out[base + p] = v;
  }
}
}  // namespace

extern "C" {
size_t ep_bn2d_workspace_bytes(int c) { return (size_t)c * BN2D_CHUNKS * sizeof(double2); }

// out = [relu]( BN_batchstats( res ? relu?(x) + res : relu?(x) ) ); x, res, out: NCHW f32 [n_img, c, hw] contiguous.
int ep_bn2d_train(const float* x, const float* res, int relu_pre, int n_img, int c, int hw, const float* gamma, const float* beta,
                  float eps, int relu_post, float* out, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (n_img < 1 || c < 1 || hw < 1 || !gamma || !beta) return EP_ERR_ARG;
  if (workspace_bytes < ep_bn2d_workspace_bytes(c) || ((uintptr_t)workspace & 15)) return EP_ERR_WORKSPACE;
  double2* part = reinterpret_cast<double2*>(workspace);
  bn2d_stats_kernel<<<dim3(c, BN2D_CHUNKS), 256, 0, stream>>>(x, res, relu_pre, n_img, c, hw, part);
  const int chunks = hw >= 4096 ? 4 : 1;        // a few CTAs per (image, channel) plane on the large maps
  bn2d_apply_kernel<<<dim3(chunks, n_img * c), 256, 0, stream>>>(x, res, relu_pre, n_img, c, hw, part, gamma, beta, eps, relu_post, out);
  EP_CHECK_LAUNCH();
  return EP_OK;
}
}  // extern "C"

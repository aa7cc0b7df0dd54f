// Sparse 3-D convolution as an output-stationary gather-GEMM (fp32 accumulate, no atomics).
//
// Replaces torchsparse v2.0.0 `spnn.Conv3d` forward (gather -> cuBLAS GEMM -> scatter-add per kernel
// offset; reference call sites models/modules.py:19,35,50,56,64,90,181), spconv `SubMConv3d`
// (models/modules.py:252,444) and, with K == 1 and no neighbour table, the dense per-row nn.Linear layers of the path
// (models/modules.py:127-136,187,279-284) whose weights are too wide for the tensor-core row kernel (csrc/linear_mma.cu).
//
//   out[j, :] = bias + sum_k  W[k]^T . in[nbr[j, k], :]        (rows with nbr < 0 contribute nothing)
//
// One CTA owns a tile of TM output rows x TN output channels and walks the K kernel offsets, gathering
// the A tile through the neighbour table; offsets with no neighbour in the whole tile are skipped.  The
// epilogue can emit per-CTA column sums / sums of squares so batch-statistics BatchNorm (the reference
// runs BN in train mode at inference, main.py:357) needs no extra pass over the output.
// Deterministic: fixed summation order, per-CTA partials reduced in order by bn_finalize.
#include <cstdlib>

#include "common.cuh"

namespace {

constexpr int TM = 64;       // output rows per CTA
constexpr int KC = 16;       // input channels per smem stage
constexpr int THREADS = 256; // 16 (cols) x 16 (rows) thread grid, 4 rows x TNM cols per thread

template <int TN>
__global__ void __launch_bounds__(THREADS)
spconv_kernel(const float* __restrict__ in, int ld_in, int cin, const int* __restrict__ nbr, int K,
              const float* __restrict__ W, int ldw /*padded Cout*/, int cout, const float* __restrict__ bias,
              float* __restrict__ out, int ld_out, int m_out, float* __restrict__ bn_partial) {
  constexpr int TNM = TN / 16;  // columns per thread
  __shared__ __align__(16) float As[2][KC][TM + 4];
  __shared__ __align__(16) float Bs[2][KC][TN];
  __shared__ int s_nbr[TM];
  __shared__ int s_any;

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int row0 = blockIdx.x * TM;
  const int col0 = blockIdx.y * TN;

  // A-tile loader: thread -> (row r_a, channel quad q_a); B-tile loader: thread -> (kk_b, 4*j cols)
  const int r_a = tid >> 2, q_a = tid & 3;

  float acc[4][TNM];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TNM; ++j) acc[i][j] = 0.f;

  const int nchunk = (cin + KC - 1) / KC;

  for (int k = 0; k < K; ++k) {
    // neighbour rows of this tile for offset k
    __syncthreads();
    if (tid == 0) s_any = 0;
    __syncthreads();
    if (tid < TM) {
      int j = row0 + tid;
      int r = -1;
      if (j < m_out) r = nbr ? nbr[(size_t)j * K + k] : j;
      s_nbr[tid] = r;
      if (r >= 0) s_any = 1;
    }
    __syncthreads();
    if (!s_any) continue;
    const int my_src = s_nbr[r_a];
    const float* a_row = my_src >= 0 ? in + (size_t)my_src * ld_in : nullptr;
    const float* w_k = W + (size_t)k * cin * ldw;

    float4 a_reg = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 b_reg[(KC * TN / 4 + THREADS - 1) / THREADS];
    constexpr int B_ITERS = (KC * TN / 4 + THREADS - 1) / THREADS;

    auto load_tiles = [&](int chunk) {
      const int c = chunk * KC + q_a * 4;
      a_reg = (a_row && c < cin) ? *reinterpret_cast<const float4*>(a_row + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int it = 0; it < B_ITERS; ++it) {
        int e = tid + it * THREADS;           // float4 index inside the [KC][TN] tile
        int kk = e / (TN / 4), cq = e % (TN / 4);
        int cc = chunk * KC + kk, col = col0 + cq * 4;
        b_reg[it] = (e < KC * TN / 4 && cc < cin && col < ldw)
                        ? __ldg(reinterpret_cast<const float4*>(w_k + (size_t)cc * ldw + col))
                        : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto store_tiles = [&](int buf) {
      As[buf][q_a * 4 + 0][r_a] = a_reg.x;
      As[buf][q_a * 4 + 1][r_a] = a_reg.y;
      As[buf][q_a * 4 + 2][r_a] = a_reg.z;
      As[buf][q_a * 4 + 3][r_a] = a_reg.w;
#pragma unroll
      for (int it = 0; it < B_ITERS; ++it) {
        int e = tid + it * THREADS;
        if (e < KC * TN / 4) {
          int kk = e / (TN / 4), cq = e % (TN / 4);
          *reinterpret_cast<float4*>(&Bs[buf][kk][cq * 4]) = b_reg[it];
        }
      }
    };

    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    for (int chunk = 0; chunk < nchunk; ++chunk) {
      const int buf = chunk & 1;
      if (chunk + 1 < nchunk) load_tiles(chunk + 1);
#pragma unroll
      for (int kk = 0; kk < KC; ++kk) {
        const float4 a4 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
        const float a[4] = {a4.x, a4.y, a4.z, a4.w};
        float b[TNM];
#pragma unroll
        for (int j = 0; j < TNM; ++j) b[j] = Bs[buf][kk][tx + 16 * j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < TNM; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      if (chunk + 1 < nchunk) {
        store_tiles(buf ^ 1);
        __syncthreads();
      }
    }
  }

  // epilogue: bias, store, optional BN partial statistics (column sum / sum of squares over this tile)
  float csum[TNM], csq[TNM];
#pragma unroll
  for (int j = 0; j < TNM; ++j) { csum[j] = 0.f; csq[j] = 0.f; }
#pragma unroll
  for (int j = 0; j < TNM; ++j) {
    const int col = col0 + tx + 16 * j;
    const float bv = (bias && col < cout) ? bias[col] : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = row0 + ty * 4 + i;
      if (row < m_out && col < cout) {
        const float v = acc[i][j] + bv;
        out[(size_t)row * ld_out + col] = v;
        csum[j] += v;
        csq[j] = fmaf(v, v, csq[j]);
      }
    }
  }
  if (bn_partial) {
    __syncthreads();
    float* red = &Bs[0][0][0];  // reuse: [16 ty][TN] sums then [16][TN] squares (2*16*TN floats == sizeof(Bs))
#pragma unroll
    for (int j = 0; j < TNM; ++j) {
      red[ty * TN + tx + 16 * j] = csum[j];
      red[16 * TN + ty * TN + tx + 16 * j] = csq[j];
    }
    __syncthreads();
    if (tid < TN) {
      float s = 0.f, q = 0.f;
#pragma unroll
      for (int r = 0; r < 16; ++r) { s += red[r * TN + tid]; q += red[16 * TN + r * TN + tid]; }
      const int col = col0 + tid;
      if (col < cout) {
        bn_partial[((size_t)blockIdx.x * 2 + 0) * cout + col] = s;
        bn_partial[((size_t)blockIdx.x * 2 + 1) * cout + col] = q;
      }
    }
  }
}

// column statistics of an existing [m, c] matrix (for BatchNorm1d on tensors no conv of ours produced)
__global__ void __launch_bounds__(256) colstats_kernel(const float* __restrict__ x, int ld, int m, int c,
                                                       float* __restrict__ bn_partial) {
  // CTA b covers rows [b*64, b*64+64); thread t loops over columns
  const int row0 = blockIdx.x * TM;
  for (int col = threadIdx.x; col < c; col += blockDim.x) {
    float s = 0.f, q = 0.f;
    for (int r = 0; r < TM; ++r) {
      int row = row0 + r;
      if (row < m) { float v = x[(size_t)row * ld + col]; s += v; q = fmaf(v, v, q); }
    }
    bn_partial[((size_t)blockIdx.x * 2 + 0) * c + col] = s;
    bn_partial[((size_t)blockIdx.x * 2 + 1) * c + col] = q;
  }
}

// reduce per-CTA partials in CTA order (double accumulation) -> scale/shift of train-mode BatchNorm:
//   y = (x - mean) * rsqrt(var_biased + eps) * gamma + beta  ==  x * scale + shift
__global__ void __launch_bounds__(1024)
bn_finalize_kernel(const float* __restrict__ bn_partial, int nblk, int c, int m, float eps,
                   const float* __restrict__ gamma, const float* __restrict__ beta,
                   float* __restrict__ scale_shift /*[2,c]*/, float* __restrict__ mean_var /*[2,c] or null*/) {
  // 32 columns x 32 tile-lanes per CTA: lane ty sums tiles ty, ty+32, ... (double), then a fixed-order combine
  __shared__ double s_s[32][33], s_q[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + tx;
  double s = 0.0, q = 0.0;
  if (col < c)
#pragma unroll 4   // same sequential accumulation order, 8 independent loads in flight instead of 2 (latency-bound at 3000 tiles)
    for (int b = ty; b < nblk; b += 32) {
      s += (double)bn_partial[((size_t)b * 2 + 0) * c + col];
      q += (double)bn_partial[((size_t)b * 2 + 1) * c + col];
    }
  s_s[ty][tx] = s;
  s_q[ty][tx] = q;
  __syncthreads();
  if (ty != 0 || col >= c) return;
#pragma unroll
  for (int r = 1; r < 32; ++r) { s += s_s[r][tx]; q += s_q[r][tx]; }
  const double mean = s / m;
  double var = q / m - mean * mean;
  if (var < 0.0) var = 0.0;
  const float inv = (float)(1.0 / sqrt(var + (double)eps));
  const float g = gamma ? gamma[col] : 1.f, bt = beta ? beta[col] : 0.f;
  const float sc = inv * g;
  scale_shift[col] = sc;
  scale_shift[c + col] = bt - (float)mean * sc;
  if (mean_var) { mean_var[col] = (float)mean; mean_var[c + col] = (float)var; }
}


}  // namespace

int ep_internal_linear_mma(const float* in, int ld_in, int cin, const float* W, int ldw, int cout, const float* bias, float* out,
                           int ld_out, int64_t m, float* bn_partial, cudaStream_t stream);   // csrc/linear_mma.cu

extern "C" {

int ep_spconv_num_row_tiles(int64_t m_out) { return ep_div_up(m_out, TM); }

// in [*, ld_in] fp32, nbr int32 [m_out, K] or NULL (identity, K must be 1), W [K, cin, ldw] (ldw = cout padded to 4),
// out [m_out, ld_out].  bn_partial: NULL or float[num_row_tiles, 2, cout].
int ep_spconv_fwd(const float* in, int ld_in, int cin, const int32_t* nbr, int K, const float* W, int ldw, int cout,
                  const float* bias, float* out, int ld_out, int64_t m_out, float* bn_partial, cudaStream_t stream) {
  if (m_out <= 0 || cin < 1 || cout < 1 || K < 1 || ld_in % 4 != 0 || ldw % 4 != 0 || ldw < cout) return EP_ERR_ARG;
  if (!nbr && K != 1) return EP_ERR_ARG;
  // dense linears run on the 3xTF32 row kernel (csrc/linear_mma.cu) unless their weights are too wide for it or
  // EPRECON_LINEAR=tile asks for the fp32 FFMA tile kernel below
  static const bool knob_tile = [] { const char* v = getenv("EPRECON_LINEAR"); return v && v[0] == 't'; }();
  if (!nbr && !knob_tile && in != out) {
    const int st = ep_internal_linear_mma(in, ld_in, cin, W, ldw, cout, bias, out, ld_out, m_out, bn_partial, stream);
    if (st != EP_ERR_UNSUPPORTED) {
      if (st != EP_OK) return st;
      EP_CHECK_LAUNCH();
      return EP_OK;
    }
  }
  const int rt = ep_div_up(m_out, TM);
  if (cout <= 32) {
    dim3 grid(rt, ep_div_up(cout, 32));
    spconv_kernel<32><<<grid, THREADS, 0, stream>>>(in, ld_in, cin, nbr, K, W, ldw, cout, bias, out, ld_out, (int)m_out, bn_partial);
  } else if (cout <= 64) {
    dim3 grid(rt, ep_div_up(cout, 64));
    spconv_kernel<64><<<grid, THREADS, 0, stream>>>(in, ld_in, cin, nbr, K, W, ldw, cout, bias, out, ld_out, (int)m_out, bn_partial);
  } else {
    dim3 grid(rt, ep_div_up(cout, 128));
    spconv_kernel<128><<<grid, THREADS, 0, stream>>>(in, ld_in, cin, nbr, K, W, ldw, cout, bias, out, ld_out, (int)m_out, bn_partial);
  }
  EP_CHECK_LAUNCH();
  return EP_OK;
}

int ep_colstats(const float* x, int ld, int64_t m, int c, float* bn_partial, cudaStream_t stream) {
  if (m <= 0 || c < 1) return EP_ERR_ARG;
  colstats_kernel<<<ep_div_up(m, TM), 256, 0, stream>>>(x, ld, (int)m, c, bn_partial);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

int ep_bn_finalize(const float* bn_partial, int num_row_tiles, int c, int64_t m, float eps, const float* gamma,
                   const float* beta, float* scale_shift, float* mean_var, cudaStream_t stream) {
  if (num_row_tiles < 1 || c < 1 || m < 1) return EP_ERR_ARG;
  bn_finalize_kernel<<<ep_div_up(c, 32), 1024, 0, stream>>>(bn_partial, num_row_tiles, c, (int)m, eps, gamma, beta,
                                                           scale_shift, mean_var);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

}  // extern "C"

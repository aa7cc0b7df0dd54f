// Dense per-row linear  out[r, :cout] = bias + in[r, :cin] W  (ep_spconv_fwd with K == 1 and no neighbour table: every nn.Linear
// of the path -- models/modules.py:127-136,187,279-284 -- 57 launches and ~1.5 ms per bench fragment on the FFMA tile kernel).
//
// The op moves 4 (cin + cout) bytes per row against 2 cin cout flops: with cout = 24...96 it is bound by how fast rows can be
// pulled through the SM, not by arithmetic.  The FFMA kernels (csrc/spconv.cu) need a shared-memory operand read per 4-8 FMAs and
// top out at 1.3-1.9 TB/s of row traffic.  Here the rows never touch shared memory:
//   * every lane loads 16-byte pieces of its two rows of a 16-row block straight into the A-fragment registers of
//     mma.sync.m16n8k8 (tf32): the 16 input channels of a k-block are assigned to the two k8 MMAs so that a lane's float4 is
//     exactly its four A elements (the weights are staged in shared memory in the matching permuted fragment order) -> fully
//     used sectors, no transposition, all loads of a k-block in flight at once;
//   * fp32 accuracy through the 3xTF32 split done in registers (a = hi + lo, both tf32; lo*hi + hi*lo + hi*hi, small terms
//     first): ~5e-7 relative, the same class as the half-pair tensor-core convs;
//   * a warp owns 64 rows = one BatchNorm statistics tile: column sums / sums of squares are reduced with 3 shuffles in a fixed
//     order (deterministic), outputs leave as 8-byte stores that fill whole sectors.
// This is a legacy warp-level MMA on purpose: the tensor pipe is idle on this op either way, a tcgen05 version would need the
// activations re-staged (or pre-split) in shared memory -- the traffic this kernel exists to avoid.
#include <cstdlib>

#include "common.cuh"

namespace {

constexpr int LM_THREADS = 128;              // 4 warps x 64 rows
constexpr int LM_ROWS = 256;
constexpr int LM_MB = 4;                     // 16-row blocks per warp

__device__ __forceinline__ uint32_t tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// NNT: 8-column tiles per pass (accumulators: LM_MB * NNT * 4 registers)
template <int NNT, bool ALIGNED8>
__global__ void __launch_bounds__(LM_THREADS)
linear_mma_kernel(const float* __restrict__ in, int ld_in, int cin, const float* __restrict__ W, int ldw, int cout,
                  const float* __restrict__ bias, float* __restrict__ out, int ld_out, int m, float* __restrict__ bn_partial,
                  int nkb /* 16-channel blocks */, int nnt /* 8-column tiles, all passes */) {
  // weight fragments: [kb][nt][half][hi|lo][lane] float2 = (b0, b1) of the MMA that consumes channels {4t + 2 half, 4t + 2 half + 1}
  extern __shared__ __align__(16) float2 lm_w[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
#pragma unroll 4
  for (int e = tid; e < nkb * nnt * 2 * 32; e += LM_THREADS) {
    const int ln = e & 31, half = (e >> 5) & 1, nt = (e >> 6) % nnt, kb = (e >> 6) / nnt;
    const int k0 = kb * 16 + 4 * (ln & 3) + 2 * half, n = nt * 8 + (ln >> 2);
    const float w0 = (k0 < cin && n < ldw) ? __ldg(W + (size_t)k0 * ldw + n) : 0.f;
    const float w1 = (k0 + 1 < cin && n < ldw) ? __ldg(W + (size_t)(k0 + 1) * ldw + n) : 0.f;
    const uint32_t h0 = tf32_rna(w0), h1 = tf32_rna(w1);
    const size_t base = ((size_t)((kb * nnt + nt) * 2 + half) * 2) * 32 + ln;
    lm_w[base] = make_float2(__uint_as_float(h0), __uint_as_float(h1));
    lm_w[base + 32] = make_float2(__uint_as_float(tf32_rna(w0 - __uint_as_float(h0))), __uint_as_float(tf32_rna(w1 - __uint_as_float(h1))));
  }
  __syncthreads();
  const int cin4 = (cin + 3) & ~3;
  const int nchunk = (m + LM_ROWS - 1) / LM_ROWS;
  for (int chunk = blockIdx.x; chunk < nchunk; chunk += gridDim.x) {
    const int row0 = chunk * LM_ROWS + warp * 64;
    if (row0 >= m) continue;
    for (int nt0 = 0; nt0 < nnt; nt0 += NNT) {
      float acc[LM_MB][NNT][4];
#pragma unroll
      for (int mb = 0; mb < LM_MB; ++mb)
#pragma unroll
        for (int j = 0; j < NNT; ++j)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[mb][j][q] = 0.f;
      for (int kb = 0; kb < nkb; ++kb) {
        const int ch = kb * 16 + 4 * t;
        float4 raw[LM_MB][2];
#pragma unroll
        for (int mb = 0; mb < LM_MB; ++mb)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int r = row0 + mb * 16 + g + 8 * h;
            raw[mb][h] = (r < m && ch < cin4) ? __ldg(reinterpret_cast<const float4*>(in + (size_t)r * ld_in + ch))
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t ahi[LM_MB][4], alo[LM_MB][4];
#pragma unroll
          for (int mb = 0; mb < LM_MB; ++mb) {
            // a0 = (row g, slot t), a1 = (row g + 8, slot t), a2 = (row g, slot t + 4), a3 = (row g + 8, slot t + 4)
            const float v[4] = {half ? raw[mb][0].z : raw[mb][0].x, half ? raw[mb][1].z : raw[mb][1].x,
                                half ? raw[mb][0].w : raw[mb][0].y, half ? raw[mb][1].w : raw[mb][1].y};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              ahi[mb][q] = tf32_rna(v[q]);
              alo[mb][q] = tf32_rna(v[q] - __uint_as_float(ahi[mb][q]));
            }
          }
#pragma unroll
          for (int j = 0; j < NNT; ++j) {
            if (nt0 + j < nnt) {
              const size_t base = ((size_t)((kb * nnt + nt0 + j) * 2 + half) * 2) * 32 + lane;
              const float2 bh = lm_w[base], bl = lm_w[base + 32];
              const uint32_t bh0 = __float_as_uint(bh.x), bh1 = __float_as_uint(bh.y);
              const uint32_t bl0 = __float_as_uint(bl.x), bl1 = __float_as_uint(bl.y);
#pragma unroll
              for (int mb = 0; mb < LM_MB; ++mb) {
                mma_tf32(acc[mb][j], alo[mb], bh0, bh1);
                mma_tf32(acc[mb][j], ahi[mb], bl0, bl1);
                mma_tf32(acc[mb][j], ahi[mb], bh0, bh1);
              }
            }
          }
        }
      }
      // epilogue: bias, stores (c0, c1 = row g, columns 2t, 2t + 1; c2, c3 = row g + 8), per-tile column statistics
#pragma unroll
      for (int j = 0; j < NNT; ++j) {
        if (nt0 + j >= nnt) continue;                       // warp-uniform
        const int col = (nt0 + j) * 8 + 2 * t;
        const bool one = col < cout, two = col + 1 < cout;  // per lane: the shuffles below stay warp-wide
        const float b0 = (bias && one) ? __ldg(bias + col) : 0.f, b1 = (bias && two) ? __ldg(bias + col + 1) : 0.f;
        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
        for (int mb = 0; mb < LM_MB; ++mb)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int r = row0 + mb * 16 + g + 8 * h;
            if (r < m && one) {
              const float v0 = acc[mb][j][2 * h] + b0, v1 = acc[mb][j][2 * h + 1] + b1;
              float* dst = out + (size_t)r * ld_out + col;
              if (ALIGNED8 && two) {
                *reinterpret_cast<float2*>(dst) = make_float2(v0, v1);
              } else {
                dst[0] = v0;
                if (two) dst[1] = v1;
              }
              s0 += v0; q0 = fmaf(v0, v0, q0);
              s1 += v1; q1 = fmaf(v1, v1, q1);
            }
          }
        if (bn_partial) {
#pragma unroll
          for (int d = 4; d < 32; d <<= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, d); q0 += __shfl_xor_sync(0xffffffffu, q0, d);
            s1 += __shfl_xor_sync(0xffffffffu, s1, d); q1 += __shfl_xor_sync(0xffffffffu, q1, d);
          }
          if (g == 0 && one) {
            const size_t tile = (size_t)(row0 >> 6);
            bn_partial[(tile * 2 + 0) * cout + col] = s0;
            bn_partial[(tile * 2 + 1) * cout + col] = q0;
            if (two) {
              bn_partial[(tile * 2 + 0) * cout + col + 1] = s1;
              bn_partial[(tile * 2 + 1) * cout + col + 1] = q1;
            }
          }
        }
      }
    }
  }
}

template <int NNT>
int launch(bool aligned8, int grid, size_t smem, cudaStream_t stream, const float* in, int ld_in, int cin, const float* W, int ldw,
           int cout, const float* bias, float* out, int ld_out, int m, float* bn_partial, int nkb, int nnt) {
  static const cudaError_t attr = [] {
    cudaError_t a = cudaFuncSetAttribute(linear_mma_kernel<NNT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
    cudaError_t b = cudaFuncSetAttribute(linear_mma_kernel<NNT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
    return a != cudaSuccess ? a : b;
  }();
  if (attr != cudaSuccess) return EP_ERR_CUDA;
  if (aligned8)
    linear_mma_kernel<NNT, true><<<grid, LM_THREADS, smem, stream>>>(in, ld_in, cin, W, ldw, cout, bias, out, ld_out, m, bn_partial, nkb, nnt);
  else
    linear_mma_kernel<NNT, false><<<grid, LM_THREADS, smem, stream>>>(in, ld_in, cin, W, ldw, cout, bias, out, ld_out, m, bn_partial, nkb, nnt);
  return EP_OK;
}

}  // namespace

// EP_ERR_UNSUPPORTED: the weight fragments do not fit in shared memory (the caller keeps the FFMA tile kernel)
int ep_internal_linear_mma(const float* in, int ld_in, int cin, const float* W, int ldw, int cout, const float* bias, float* out,
                           int ld_out, int64_t m, float* bn_partial, cudaStream_t stream) {
  const int nkb = ep_div_up(cin, 16), nnt = ep_div_up(cout, 8);
  const size_t smem = (size_t)nkb * nnt * 2 * 2 * 32 * sizeof(float2);
  // > 48 KB of weight fragments (192-wide layers): the per-CTA staging and the column passes eat the advantage -- measured on
  // par with / behind the FFMA tile kernel (profiles/r02_probe_linear_mma.json), which keeps those
  if (smem > 48 * 1024 || ((uintptr_t)in & 15) || ld_in % 4) return EP_ERR_UNSUPPORTED;
  const int npass_cols = nnt <= 4 ? nnt : (nnt % 4 == 0 ? 4 : (nnt % 3 == 0 ? 3 : 4));
  static const int sms = [] {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 148;
    return n;
  }();
  // persistent CTAs, c per SM: the c whose last round of 256-row chunks is fullest
  const int nchunk = ep_div_up(m, LM_ROWS);
  int per_sm = (int)(200 * 1024 / (smem + 2048));
  per_sm = per_sm < 1 ? 1 : per_sm > 4 ? 4 : per_sm;
  int grid = nchunk;
  if (nchunk > sms) {
    double best = -1.0;
    for (int c = 1; c <= per_sm; ++c) {
      const int gsz = sms * c;
      const double eff = (double)nchunk / ((double)ep_div_up(nchunk, gsz) * gsz) + 1e-3 * c;
      if (eff > best) { best = eff; grid = gsz; }
    }
    if (grid > nchunk) grid = nchunk;
  }
  const bool aligned8 = (((uintptr_t)out & 7) == 0) && (ld_out % 2 == 0);
  switch (npass_cols) {
    case 1: return launch<1>(aligned8, grid, smem, stream, in, ld_in, cin, W, ldw, cout, bias, out, ld_out, (int)m, bn_partial, nkb, nnt);
    case 2: return launch<2>(aligned8, grid, smem, stream, in, ld_in, cin, W, ldw, cout, bias, out, ld_out, (int)m, bn_partial, nkb, nnt);
    case 3: return launch<3>(aligned8, grid, smem, stream, in, ld_in, cin, W, ldw, cout, bias, out, ld_out, (int)m, bn_partial, nkb, nnt);
    default: return launch<4>(aligned8, grid, smem, stream, in, ld_in, cin, W, ldw, cout, bias, out, ld_out, (int)m, bn_partial, nkb, nnt);
  }
}

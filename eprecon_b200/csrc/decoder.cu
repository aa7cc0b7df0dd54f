// Panoptic decoder body, native (SURVEY.md section 8 f row 1): MultiScaleMaskedTransformerDecoder.forward
// (models/mask3dformer.py:337-445; layers :70-200; position encoding models/voxel_position_encoding.py:42-146) as ONE C call.
//
// Round 1 kept everything but the masked cross-attention core on cuBLAS/ATen: 419 launches issued from Python, 2 ms of
// kernels inside an 8.5 ms host span per fragment.  Here the whole forward is ~50 launches issued from C++:
//   dec_kv_kernel        per layer: K = (rows + level_embed + fourier(xyz)) Wk^T + bk, V = (rows + level_embed) Wv^T + bv over
//                        the level's N_l voxel rows (one thread per row, weights in shared memory) -- replaces the embedding
//                        add, the position encoding and two 105 k x 48 x 48 GEMMs;
//   ep_masked_attention  (csrc/attention.cu) score -> mask -> online softmax -> weighted sum, one pass over the keys;
//   dec_rows_a_kernel    per query row: cross-attention out-projection + residual + LayerNorm, then the self-attention q / k / v;
//   dec_selfattn_kernel  80 x 80 softmax attention, 8 heads, one CTA;
//   dec_rows_b_kernel    per query row: self-attention out-projection + LN, FFN + LN, decoder norm, class logits, mask-embed
//                        MLP, and the NEXT layer's cross-attention query;
//   dec_masks_kernel     per voxel of the next memory level: mask logit of its nearest level-2 voxel for all 80 queries ->
//                        blocked flags (logit < 0) + "query has an unblocked key" flags; the full [Q, N2] mask logits are only
//                        written where the caller needs them (final output, optional aux outputs).
// A query whose keys are all blocked attends everywhere (mask3dformer.py:392): the attention kernel takes the per-query flag
// instead of a rewritten mask.  fp32 FMA throughout; every reduction in a fixed order.
#include <cmath>
#include <cstring>

#include "common.cuh"
#include "eprecon_b200.h"

namespace {

constexpr int DE = 48;        // hidden_dim == mask_dim
constexpr int DH = 8;         // heads
constexpr int DHD = 6;        // channels per head
constexpr int DQ = 80;        // queries
constexpr int DFF = 192;      // dim_feedforward == mask-embed hidden
constexpr int DCLS = 21;      // num_classes + 1
constexpr int DR = 4;         // query rows per CTA in the row kernels
constexpr int DMAX = 192;

struct Lin { const float* w; const float* b; };
struct Nrm { const float* g; const float* b; };
struct LayerP {
  const float* ca_in_w; const float* ca_in_b; Lin ca_out; Nrm ca_norm;
  const float* sa_in_w; const float* sa_in_b; Lin sa_out; Nrm sa_norm;
  Lin ff1, ff2; Nrm ff_norm;
};
struct DecP {
  const float* query_feat; const float* query_embed; const float* level_embed; const float* gauss_b;
  Nrm dec_norm; Lin cls; Lin me[3];
  LayerP layer[6];
};

// ------------------------------------------------------------------------------------------------ K / V projections
__global__ void __launch_bounds__(128)
dec_kv_kernel(const float* __restrict__ rows, int ld_rows, const int64_t* __restrict__ xyz /*[n,3]*/, int n, const float* __restrict__ level_embed,
              const float* __restrict__ gauss_b /*[3][24]*/, float ex, float ey, float ez, const float* __restrict__ in_w /*[144][48]*/,
              const float* __restrict__ in_b, float* __restrict__ k_out, float* __restrict__ v_out) {
  __shared__ float s_wk[DE][DE + 4], s_wv[DE][DE + 4];   // transposed: [i][o]
  __shared__ float s_b[2][DE], s_le[DE], s_g[3][DE / 2];
  for (int e = threadIdx.x; e < DE * DE; e += blockDim.x) {
    const int o = e / DE, i = e - o * DE;
    s_wk[i][o] = in_w[(size_t)(DE + o) * DE + i];
    s_wv[i][o] = in_w[(size_t)(2 * DE + o) * DE + i];
  }
  for (int e = threadIdx.x; e < DE; e += blockDim.x) { s_b[0][e] = in_b[DE + e]; s_b[1][e] = in_b[2 * DE + e]; s_le[e] = level_embed[e]; }
  for (int e = threadIdx.x; e < 3 * (DE / 2); e += blockDim.x) s_g[e / (DE / 2)][e % (DE / 2)] = gauss_b[e];
  __syncthreads();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  float s[DE], key[DE];
#pragma unroll
  for (int c = 0; c < DE; c += 4) {
    const float4 f = __ldg(reinterpret_cast<const float4*>(rows + (size_t)r * ld_rows + c));
    s[c] = f.x + s_le[c]; s[c + 1] = f.y + s_le[c + 1]; s[c + 2] = f.z + s_le[c + 2]; s[c + 3] = f.w + s_le[c + 3];
  }
  // PositionEmbeddingCoordsSine.rows: x / extent * 2 pi @ gauss_B -> [sin | cos]
  const float two_pi = 6.283185307179586f;
  const float px = ((float)xyz[(size_t)r * 3 + 0] / ex) * two_pi, py = ((float)xyz[(size_t)r * 3 + 1] / ey) * two_pi,
              pz = ((float)xyz[(size_t)r * 3 + 2] / ez) * two_pi;
#pragma unroll
  for (int j = 0; j < DE / 2; ++j) {
    const float p = px * s_g[0][j] + py * s_g[1][j] + pz * s_g[2][j];
    key[j] = s[j] + sinf(p);
    key[DE / 2 + j] = s[DE / 2 + j] + cosf(p);
  }
  for (int o = 0; o < DE; o += 4) {
    float4 ak = make_float4(s_b[0][o], s_b[0][o + 1], s_b[0][o + 2], s_b[0][o + 3]);
    float4 av = make_float4(s_b[1][o], s_b[1][o + 1], s_b[1][o + 2], s_b[1][o + 3]);
#pragma unroll
    for (int i = 0; i < DE; ++i) {
      const float4 wk = *reinterpret_cast<const float4*>(&s_wk[i][o]);
      const float4 wv = *reinterpret_cast<const float4*>(&s_wv[i][o]);
      ak.x = fmaf(key[i], wk.x, ak.x); ak.y = fmaf(key[i], wk.y, ak.y); ak.z = fmaf(key[i], wk.z, ak.z); ak.w = fmaf(key[i], wk.w, ak.w);
      av.x = fmaf(s[i], wv.x, av.x); av.y = fmaf(s[i], wv.y, av.y); av.z = fmaf(s[i], wv.z, av.z); av.w = fmaf(s[i], wv.w, av.w);
    }
    *reinterpret_cast<float4*>(k_out + (size_t)r * DE + o) = ak;
    *reinterpret_cast<float4*>(v_out + (size_t)r * DE + o) = av;
  }
}

// ------------------------------------------------------------------------------------------------ row helpers (DR rows per CTA)
// out[r][o] = b[o] + sum_i in[r][i] W[o][i]  (+ res[r][o]) (relu), for the CTA's DR rows; W is nn.Linear's [cout][cin]
__device__ __forceinline__ void row_linear(const float (*in)[DMAX], int cin, const float* __restrict__ W, const float* __restrict__ b,
                                           int cout, float (*out)[DMAX], bool relu, const float (*res)[DMAX]) {
  for (int o = threadIdx.x; o < cout; o += blockDim.x) {
    float acc[DR];
    const float b0 = b ? b[o] : 0.f;
#pragma unroll
    for (int r = 0; r < DR; ++r) acc[r] = b0;
    const float4* w4 = reinterpret_cast<const float4*>(W + (size_t)o * cin);
    for (int i = 0; i < cin; i += 4) {
      const float4 w = __ldg(w4 + (i >> 2));
#pragma unroll
      for (int r = 0; r < DR; ++r) {
        acc[r] = fmaf(in[r][i], w.x, acc[r]);
        acc[r] = fmaf(in[r][i + 1], w.y, acc[r]);
        acc[r] = fmaf(in[r][i + 2], w.z, acc[r]);
        acc[r] = fmaf(in[r][i + 3], w.w, acc[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < DR; ++r) {
      float v = acc[r] + (res ? res[r][o] : 0.f);
      out[r][o] = relu ? fmaxf(v, 0.f) : v;
    }
  }
  __syncthreads();
}

// LayerNorm over the first c entries of each of the CTA's rows (warp r owns row r; biased variance, as nn.LayerNorm)
__device__ __forceinline__ void row_layernorm(float (*buf)[DMAX], int c, const Nrm& n, float eps, float (*out)[DMAX]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < DR) {
    float s = 0.f;
    for (int i = lane; i < c; i += 32) s += buf[warp][i];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    const float mean = s / (float)c;
    float q = 0.f;
    for (int i = lane; i < c; i += 32) { const float d0 = buf[warp][i] - mean; q = fmaf(d0, d0, q); }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) q += __shfl_xor_sync(0xffffffffu, q, d);
    const float inv = rsqrtf(q / (float)c + eps);
    for (int i = lane; i < c; i += 32) out[warp][i] = fmaf((buf[warp][i] - mean) * inv, n.g[i], n.b[i]);
  }
  __syncthreads();
}

__device__ __forceinline__ void rows_load(const float* __restrict__ src, int row0, int nrows, int c, float (*dst)[DMAX]) {
  for (int e = threadIdx.x; e < DR * c; e += blockDim.x) {
    const int r = e / c, i = e - r * c;
    dst[r][i] = (row0 + r < nrows) ? src[(size_t)(row0 + r) * c + i] : 0.f;
  }
  __syncthreads();
}
__device__ __forceinline__ void rows_store(const float (*src)[DMAX], int row0, int nrows, int c, float* __restrict__ dst) {
  for (int e = threadIdx.x; e < DR * c; e += blockDim.x) {
    const int r = e / c, i = e - r * c;
    if (row0 + r < nrows) dst[(size_t)(row0 + r) * c + i] = src[r][i];
  }
}

// cross-attention tail + self-attention projections for DR query rows
__global__ void __launch_bounds__(128)
dec_rows_a_kernel(LayerP L, const float* __restrict__ out_in, const float* __restrict__ att, const float* __restrict__ qpos, int nq,
                  float* __restrict__ out1, float* __restrict__ qs, float* __restrict__ ks, float* __restrict__ vs) {
  __shared__ float a[DR][DMAX], b[DR][DMAX], c[DR][DMAX];
  const int row0 = blockIdx.x * DR;
  rows_load(att, row0, nq, DE, a);
  rows_load(out_in, row0, nq, DE, b);
  row_linear(a, DE, L.ca_out.w, L.ca_out.b, DE, c, false, b);            // out + out_proj(att)
  row_layernorm(c, DE, L.ca_norm, 1e-5f, b);                             // out1
  rows_store(b, row0, nq, DE, out1);
  rows_load(qpos, row0, nq, DE, a);
  for (int e = threadIdx.x; e < DR * DE; e += blockDim.x) a[e / DE][e % DE] += b[e / DE][e % DE];   // qk = out1 + query_pos
  __syncthreads();
  row_linear(a, DE, L.sa_in_w, L.sa_in_b, DE, c, false, nullptr);
  rows_store(c, row0, nq, DE, qs);
  __syncthreads();
  row_linear(a, DE, L.sa_in_w + DE * DE, L.sa_in_b + DE, DE, c, false, nullptr);
  rows_store(c, row0, nq, DE, ks);
  __syncthreads();
  row_linear(b, DE, L.sa_in_w + 2 * DE * DE, L.sa_in_b + 2 * DE, DE, c, false, nullptr);
  rows_store(c, row0, nq, DE, vs);
}

// softmax(q k^T / sqrt(6)) v for 8 heads x 80 queries in one CTA: thread = (head, query)
__global__ void __launch_bounds__(DH * 96)
dec_selfattn_kernel(const float* __restrict__ qs, const float* __restrict__ ks, const float* __restrict__ vs, int nq, float* __restrict__ sa) {
  __shared__ float s_k[96][DE], s_v[96][DE];
  for (int e = threadIdx.x; e < nq * DE; e += blockDim.x) { s_k[e / DE][e % DE] = ks[e]; s_v[e / DE][e % DE] = vs[e]; }
  __syncthreads();
  const int h = threadIdx.x / 96, qi = threadIdx.x % 96;
  if (qi >= nq) return;
  const float scale = 0.408248290463863f;     // 1 / sqrt(6)
  float q[DHD];
#pragma unroll
  for (int d = 0; d < DHD; ++d) q[d] = qs[(size_t)qi * DE + h * DHD + d] * scale;
  float m = -INFINITY;
  for (int j = 0; j < nq; ++j) {
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < DHD; ++d) s = fmaf(q[d], s_k[j][h * DHD + d], s);
    m = fmaxf(m, s);
  }
  float l = 0.f, acc[DHD];
#pragma unroll
  for (int d = 0; d < DHD; ++d) acc[d] = 0.f;
  for (int j = 0; j < nq; ++j) {
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < DHD; ++d) s = fmaf(q[d], s_k[j][h * DHD + d], s);
    const float p = expf(s - m);
    l += p;
#pragma unroll
    for (int d = 0; d < DHD; ++d) acc[d] = fmaf(p, s_v[j][h * DHD + d], acc[d]);
  }
#pragma unroll
  for (int d = 0; d < DHD; ++d) sa[(size_t)qi * DE + h * DHD + d] = acc[d] / l;
}

// self-attention tail, FFN, prediction heads and the next layer's cross-attention query for DR query rows.
// heads_only: `out1` already holds the layer output (the initial prediction on query_feat).
__global__ void __launch_bounds__(128)
dec_rows_b_kernel(LayerP L, int heads_only, const float* __restrict__ out1, const float* __restrict__ sa, const float* __restrict__ qpos,
                  int nq, Nrm dec_norm, Lin cls, Lin me0, Lin me1, Lin me2, const float* __restrict__ next_in_w,
                  const float* __restrict__ next_in_b, float* __restrict__ out3, float* __restrict__ logits, float* __restrict__ membed,
                  float* __restrict__ q_next, int* __restrict__ has_unblocked) {
  __shared__ float a[DR][DMAX], b[DR][DMAX], c[DR][DMAX];
  const int row0 = blockIdx.x * DR;
  if (blockIdx.x == 0 && has_unblocked)
    for (int e = threadIdx.x; e < nq; e += blockDim.x) has_unblocked[e] = 0;      // consumed by the masks kernel that follows
  rows_load(out1, row0, nq, DE, b);
  if (!heads_only) {
    rows_load(sa, row0, nq, DE, a);
    row_linear(a, DE, L.sa_out.w, L.sa_out.b, DE, c, false, b);         // out1 + out_proj(self-attention)
    row_layernorm(c, DE, L.sa_norm, 1e-5f, b);                          // out2
    row_linear(b, DE, L.ff1.w, L.ff1.b, DFF, a, true, nullptr);         // relu(linear1)
    row_linear(a, DFF, L.ff2.w, L.ff2.b, DE, c, false, b);              // out2 + linear2
    row_layernorm(c, DE, L.ff_norm, 1e-5f, b);                          // out3
  }
  rows_store(b, row0, nq, DE, out3);
  row_layernorm(b, DE, dec_norm, 1e-5f, c);                             // decoder_norm(output)
  row_linear(c, DE, cls.w, cls.b, DCLS, a, false, nullptr);
  rows_store(a, row0, nq, DCLS, logits);
  __syncthreads();
  row_linear(c, DE, me0.w, me0.b, DFF, a, true, nullptr);
  row_linear(a, DFF, me1.w, me1.b, DFF, c, true, nullptr);
  row_linear(c, DFF, me2.w, me2.b, DE, a, false, nullptr);
  rows_store(a, row0, nq, DE, membed);
  if (q_next) {
    __syncthreads();
    rows_load(qpos, row0, nq, DE, a);
    for (int e = threadIdx.x; e < DR * DE; e += blockDim.x) a[e / DE][e % DE] += b[e / DE][e % DE];
    __syncthreads();
    row_linear(a, DE, next_in_w, next_in_b, DE, c, false, nullptr);
    rows_store(c, row0, nq, DE, q_next);
  }
}

// mask logits of the voxels a memory level attends through: s[q][i] = membed[q] . mask_rows[index ? index[i] : i]
// -> blocked[q][i] = s < 0 (optional), has_unblocked[q] |= !blocked, masks[q][i] = s (optional)
__global__ void __launch_bounds__(128)
dec_masks_kernel(const float* __restrict__ membed /*[nq][48]*/, int nq, const float* __restrict__ mask_rows, int ld_mask,
                 const int64_t* __restrict__ index, int n, uint8_t* __restrict__ blocked, int* __restrict__ has_unblocked,
                 float* __restrict__ masks) {
  __shared__ float s_me[96][DE];
  for (int e = threadIdx.x; e < nq * DE; e += blockDim.x) s_me[e / DE][e % DE] = membed[e];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool on = i < n;
  float row[DE];
  if (on) {
    const size_t src = index ? (size_t)index[i] : (size_t)i;
#pragma unroll
    for (int c = 0; c < DE; c += 4) {
      const float4 f = __ldg(reinterpret_cast<const float4*>(mask_rows + src * ld_mask + c));
      row[c] = f.x; row[c + 1] = f.y; row[c + 2] = f.z; row[c + 3] = f.w;
    }
  }
  for (int q = 0; q < nq; ++q) {
    float s = 0.f;
    if (on) {
#pragma unroll
      for (int c = 0; c < DE; ++c) s = fmaf(s_me[q][c], row[c], s);
      if (masks) masks[(size_t)q * n + i] = s;
      if (blocked) blocked[(size_t)q * n + i] = s < 0.f;
    }
    if (has_unblocked) {
      const unsigned any = __ballot_sync(0xffffffffu, on && !(s < 0.f));
      if ((threadIdx.x & 31) == 0 && any) atomicOr(&has_unblocked[q], 1);
    }
  }
}

struct PReader {
  const int64_t* p;
  const float* f() { return (const float*)(uintptr_t)*p++; }
  Lin lin() { Lin l; l.w = f(); l.b = f(); return l; }
  Nrm nrm() { Nrm n; n.g = f(); n.b = f(); return n; }
};

}  // namespace

extern "C" {

size_t ep_exec_decoder_workspace_bytes(int64_t n0, int64_t n1, int64_t n2) {
  const int64_t nmax = n0 > n1 ? (n0 > n2 ? n0 : n2) : (n1 > n2 ? n1 : n2);
  // K, V of the current layer + blocked flags + the small per-query buffers + the attention kernel's partials (+ 256-byte
  // rounding of each of the 14 pieces)
  return (size_t)(2 * nmax * DE * sizeof(float)) + (size_t)DQ * nmax + 9 * DQ * DE * sizeof(float) + DQ * sizeof(int) +
         ep_masked_attention_workspace_bytes(nmax, DH) + 16 * 256;
}

// desc (int64 pointers, order of eprecon_b200/executor.py::_build_decoder): query_feat, query_embed, level_embed, gauss_B,
// decoder_norm (g, b), class_embed (w, b), mask_embed 3 x (w, b), then per layer: ca in_proj (w, b), ca out_proj (w, b), ca norm
// (g, b), sa in_proj (w, b), sa out_proj (w, b), sa norm (g, b), ffn linear1 (w, b), linear2 (w, b), ffn norm (g, b); END marker.
// rows[l] f32 [n_l, ld_rows_l] (48 channels), xyz[l] int64 [n_l,3], mask_rows f32 [n2, ld_mask], index0 / index1 int64 (nearest
// level-2 row of every level-0 / level-1 voxel).  Outputs: pred_logits f32 [7][80][21] (initial prediction + 6 layers; the last is
// the decoder's output), pred_masks f32 [80][n2] (last layer), aux_masks f32 [6][80][n2] or NULL.
int ep_exec_decoder(const int64_t* desc, const float* rows0, int ld0, const float* rows1, int ld1, const float* rows2, int ld2,
                    const int64_t* xyz0, const int64_t* xyz1, const int64_t* xyz2, int64_t n0, int64_t n1, int64_t n2,
                    const float* mask_rows, int ld_mask, const int64_t* index0, const int64_t* index1, float ex, float ey,
                    float ez, float* pred_logits, float* pred_masks, float* aux_masks, void* workspace, size_t workspace_bytes,
                    cudaStream_t stream) {
  if (!desc || n0 <= 0 || n1 <= 0 || n2 <= 0 || n0 > 0x7fffffff / DQ || n1 > 0x7fffffff / DQ || n2 > 0x7fffffff / DQ) return EP_ERR_ARG;
  if (ld0 % 4 || ld1 % 4 || ld2 % 4 || ld_mask % 4 || ld0 < DE || ld1 < DE || ld2 < DE || ld_mask < DE) return EP_ERR_ARG;
  if (workspace_bytes < ep_exec_decoder_workspace_bytes(n0, n1, n2) || ((uintptr_t)workspace & 255)) return EP_ERR_WORKSPACE;
  PReader r{desc};
  DecP P;
  P.query_feat = r.f(); P.query_embed = r.f(); P.level_embed = r.f(); P.gauss_b = r.f();
  P.dec_norm = r.nrm(); P.cls = r.lin();
  for (int i = 0; i < 3; ++i) P.me[i] = r.lin();
  for (int j = 0; j < 6; ++j) {
    LayerP& L = P.layer[j];
    L.ca_in_w = r.f(); L.ca_in_b = r.f(); L.ca_out = r.lin(); L.ca_norm = r.nrm();
    L.sa_in_w = r.f(); L.sa_in_b = r.f(); L.sa_out = r.lin(); L.sa_norm = r.nrm();
    L.ff1 = r.lin(); L.ff2 = r.lin(); L.ff_norm = r.nrm();
  }
  if (*r.p != 0x44454344ll) return EP_ERR_ARG;
  const float* rows[3] = {rows0, rows1, rows2};
  const int ld[3] = {ld0, ld1, ld2};
  const int64_t* xyz[3] = {xyz0, xyz1, xyz2};
  const int64_t nl[3] = {n0, n1, n2};
  const int64_t* index[3] = {index0, index1, nullptr};
  const int64_t nmax = n0 > n1 ? (n0 > n2 ? n0 : n2) : (n1 > n2 ? n1 : n2);
  // workspace carve-up
  char* w = (char*)workspace;
  auto take = [&](size_t bytes) { char* p = w; w += (bytes + 255) & ~(size_t)255; return p; };
  float* kbuf = (float*)take((size_t)nmax * DE * sizeof(float));
  float* vbuf = (float*)take((size_t)nmax * DE * sizeof(float));
  uint8_t* blocked = (uint8_t*)take((size_t)DQ * nmax);
  float* out_a = (float*)take(DQ * DE * sizeof(float));
  float* out_b = (float*)take(DQ * DE * sizeof(float));
  float* qbuf = (float*)take(DQ * DE * sizeof(float));
  float* att = (float*)take(DQ * DE * sizeof(float));
  float* qs = (float*)take(DQ * DE * sizeof(float));
  float* ks = (float*)take(DQ * DE * sizeof(float));
  float* vs = (float*)take(DQ * DE * sizeof(float));
  float* sa = (float*)take(DQ * DE * sizeof(float));
  float* membed = (float*)take(DQ * DE * sizeof(float));
  int* has_unb = (int*)take(DQ * sizeof(int));
  const size_t att_ws_bytes = ep_masked_attention_workspace_bytes(nmax, DH);
  // the attention partials live after the fixed buffers (the workspace query reserves DQ * nmax + slack; attention needs far less)
  void* att_ws = take(att_ws_bytes);
  if ((size_t)(w - (char*)workspace) > workspace_bytes) return EP_ERR_WORKSPACE;
  const int row_ctas = ep_div_up(DQ, DR);
  const float scale = 0.408248290463863f;

  // initial prediction on query_feat + the first layer's query
  dec_rows_b_kernel<<<row_ctas, 128, 0, stream>>>(P.layer[0], 1, P.query_feat, nullptr, P.query_embed, DQ, P.dec_norm, P.cls, P.me[0],
                                                  P.me[1], P.me[2], P.layer[0].ca_in_w, P.layer[0].ca_in_b, out_a, pred_logits, membed,
                                                  qbuf, has_unb);
  float* cur = out_a;
  float* nxt = out_b;
  for (int j = 0; j <= 6; ++j) {
    // masks for the memory level layer j attends to (j == 6: only the final full masks)
    const int l = j % 3;
    const bool last = j == 6;
    if (!last) {
      // prediction j (0 = on query_feat, j = after layer j-1) also is aux output j: its full [Q, n2] mask logits are written
      // only on request; when the level IS level 2 (j = 2, 5) the same pass produces them
      float* aux_j = aux_masks ? aux_masks + (size_t)j * DQ * n2 : nullptr;
      dec_masks_kernel<<<ep_div_up(nl[l], 128), 128, 0, stream>>>(membed, DQ, mask_rows, ld_mask, index[l], (int)nl[l], blocked, has_unb,
                                                                  l == 2 ? aux_j : nullptr);
      if (aux_j && l != 2)
        dec_masks_kernel<<<ep_div_up(n2, 128), 128, 0, stream>>>(membed, DQ, mask_rows, ld_mask, nullptr, (int)n2, nullptr, nullptr, aux_j);
    } else {
      dec_masks_kernel<<<ep_div_up(n2, 128), 128, 0, stream>>>(membed, DQ, mask_rows, ld_mask, nullptr, (int)n2, nullptr, nullptr, pred_masks);
      break;
    }
    const LayerP& L = P.layer[j];
    dec_kv_kernel<<<ep_div_up(nl[l], 128), 128, 0, stream>>>(rows[l], ld[l], xyz[l], (int)nl[l], P.level_embed + l * DE, P.gauss_b, ex, ey, ez,
                                                            L.ca_in_w, L.ca_in_b, kbuf, vbuf);
    const int st = ep_masked_attention_flagged(qbuf, kbuf, vbuf, DE, blocked, has_unb, nl[l], DQ, DH, DHD, scale, att, att_ws, att_ws_bytes, stream);
    if (st != EP_OK) return st;
    dec_rows_a_kernel<<<row_ctas, 128, 0, stream>>>(L, cur, att, P.query_embed, DQ, nxt, qs, ks, vs);
    dec_selfattn_kernel<<<1, DH * 96, 0, stream>>>(qs, ks, vs, DQ, sa);
    const bool more = j + 1 < 6;
    dec_rows_b_kernel<<<row_ctas, 128, 0, stream>>>(L, 0, nxt, sa, P.query_embed, DQ, P.dec_norm, P.cls, P.me[0], P.me[1], P.me[2],
                                                    more ? P.layer[j + 1].ca_in_w : nullptr, more ? P.layer[j + 1].ca_in_b : nullptr, cur,
                                                    pred_logits + (size_t)(j + 1) * DQ * DCLS, membed, more ? qbuf : nullptr, has_unb);
    // `cur` now holds the layer output again (rows_b wrote out3 into it); nxt is scratch
  }
  EP_CHECK_LAUNCH();
  return EP_OK;
}

}  // extern "C"

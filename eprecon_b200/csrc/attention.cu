// Masked cross-attention of the panoptic decoder (models/mask3dformer.py:70-130 CrossAttentionLayer over
// nn.MultiheadAttention, called at :392-397): Q = 80 queries x 8 heads x 6 channels against N <= ~120 k voxel keys with
// a per-(query, key) boolean mask shared by the heads.
//
// The library route (two batched GEMMs with an inner / outer dimension of 6, a [heads, Q, N] score tensor written,
// masked, soft-maxed and read back) costs ~1.5 ms per level-2 layer, 6.2 of the decoder's 10 ms
// (profiles/r01_profile_panoptic_v1.txt).  Here one pass over the keys does score -> mask -> online softmax -> weighted
// sum: every CTA owns a contiguous chunk of keys, stages K / V rows and the mask in shared memory tile by tile, and
// each thread owns one (head, query) pair with its running maximum, normaliser and 6 accumulators in registers (all
// threads of a warp share the head, so K / V reads are shared-memory broadcasts).  Per-chunk partials are merged in
// fixed chunk order by a second tiny kernel: deterministic, no atomics.  HBM-bound: 2 x 192 B of K / V + Q bytes of
// mask per key, read once.
#include "common.cuh"
#include "eprecon_b200.h"

namespace {

constexpr int AD = 6;        // channels per head
constexpr int AQ = 96;       // query slots per head (3 warps); n_queries <= AQ
constexpr int ATK = 64;      // keys per shared-memory tile
constexpr int AKB = 8;       // keys per online-softmax block (one rescale per block)

struct Partial { float m, l, acc[AD]; };   // 32 bytes

__global__ void __launch_bounds__(8 * AQ)
masked_attention_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v, int ld_kv,
                        const uint8_t* __restrict__ blocked, const int* __restrict__ row_unblocked, int n_keys, int n_queries,
                        int n_heads, float scale, int keys_per_cta, Partial* __restrict__ part) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int E = n_heads * AD;
  float* s_k = reinterpret_cast<float*>(smem);                    // [ATK][E]
  float* s_v = s_k + ATK * E;                                     // [ATK][E]
  uint8_t* s_m = reinterpret_cast<uint8_t*>(s_v + ATK * E);       // [ATK][AQ]
  const int t = threadIdx.x, nthreads = blockDim.x;
  const int h = t / AQ, qi = t - h * AQ;
  const bool active = qi < n_queries;
  float qr[AD];
#pragma unroll
  for (int d = 0; d < AD; ++d) qr[d] = active ? q[(size_t)qi * E + h * AD + d] * scale : 0.f;
  float m = -INFINITY, l = 0.f, acc[AD];
#pragma unroll
  for (int d = 0; d < AD; ++d) acc[d] = 0.f;
  const int k0 = blockIdx.x * keys_per_cta;
  const int k1 = min(n_keys, k0 + keys_per_cta);
  const int e4 = E / 4;                                           // float4 items per key row (E % 4 == 0)
  for (int base = k0; base < k1; base += ATK) {
    const int nt = min(ATK, k1 - base);
    __syncthreads();                                              // previous tile fully consumed
    for (int e = t; e < nt * e4; e += nthreads) {
      const int j = e / e4, c = e - j * e4;
      reinterpret_cast<float4*>(s_k)[j * e4 + c] = __ldg(reinterpret_cast<const float4*>(k + (size_t)(base + j) * ld_kv) + c);
      reinterpret_cast<float4*>(s_v)[j * e4 + c] = __ldg(reinterpret_cast<const float4*>(v + (size_t)(base + j) * ld_kv) + c);
    }
    for (int e = t; e < n_queries * ATK; e += nthreads) {
      const int qq = e / ATK, j = e - qq * ATK;
      // row_unblocked[q] == 0: every key of query q is blocked -> the query attends everywhere (mask3dformer.py:392)
      const bool use_mask = blocked && (!row_unblocked || row_unblocked[qq] != 0);
      s_m[j * AQ + qq] = (j < nt && use_mask) ? blocked[(size_t)qq * n_keys + base + j] : (uint8_t)(j >= nt);
    }
    __syncthreads();
    if (active) {
      for (int j0 = 0; j0 < nt; j0 += AKB) {
        float s[AKB];
        float mb = -INFINITY;
#pragma unroll
        for (int u = 0; u < AKB; ++u) {
          const int j = j0 + u;
          s[u] = -INFINITY;
          if (j < nt && !s_m[j * AQ + qi]) {
            const float* kr = s_k + j * E + h * AD;
            float d0 = 0.f;
#pragma unroll
            for (int d = 0; d < AD; ++d) d0 = fmaf(qr[d], kr[d], d0);
            s[u] = d0;
          }
          mb = fmaxf(mb, s[u]);
        }
        if (mb == -INFINITY) continue;                            // every key of the block is masked for this query
        if (mb > m) {
          const float corr = expf(m - mb);                        // m = -inf on the first hit -> 0
          l *= corr;
#pragma unroll
          for (int d = 0; d < AD; ++d) acc[d] *= corr;
          m = mb;
        }
#pragma unroll
        for (int u = 0; u < AKB; ++u) {
          if (s[u] != -INFINITY) {
            const float p = expf(s[u] - m);
            const float* vr = s_v + (j0 + u) * E + h * AD;
            l += p;
#pragma unroll
            for (int d = 0; d < AD; ++d) acc[d] = fmaf(p, vr[d], acc[d]);
          }
        }
      }
    }
  }
  Partial* o = part + ((size_t)blockIdx.x * n_heads + h) * AQ + qi;
  o->m = m;
  o->l = l;
#pragma unroll
  for (int d = 0; d < AD; ++d) o->acc[d] = acc[d];
}

// out[q, h*6 + d] = sum_c acc_c * exp(m_c - M) / sum_c l_c * exp(m_c - M), chunks in ascending order
__global__ void __launch_bounds__(8 * AQ)
attention_combine_kernel(const Partial* __restrict__ part, int n_chunks, int n_queries, int n_heads, float* __restrict__ out) {
  const int t = threadIdx.x;
  const int h = t / AQ, qi = t - h * AQ;
  if (qi >= n_queries) return;
  const size_t stride = (size_t)n_heads * AQ;
  const Partial* p = part + (size_t)h * AQ + qi;
  float M = -INFINITY;
  for (int c = 0; c < n_chunks; ++c) M = fmaxf(M, p[c * stride].m);
  float L = 0.f, acc[AD];
#pragma unroll
  for (int d = 0; d < AD; ++d) acc[d] = 0.f;
  if (M != -INFINITY) {
    for (int c = 0; c < n_chunks; ++c) {
      const Partial x = p[c * stride];
      if (x.m == -INFINITY) continue;
      const float w = expf(x.m - M);
      L = fmaf(x.l, w, L);
#pragma unroll
      for (int d = 0; d < AD; ++d) acc[d] = fmaf(x.acc[d], w, acc[d]);
    }
  }
  const int E = n_heads * AD;
#pragma unroll
  for (int d = 0; d < AD; ++d) out[(size_t)qi * E + h * AD + d] = L > 0.f ? acc[d] / L : 0.f;   // no visible key -> 0 (torch: NaN)
}

inline int attn_chunks(int64_t n_keys) {
  const int tiles = ep_div_up(n_keys, ATK);
  return tiles < EP_NUM_SMS ? tiles : EP_NUM_SMS;
}

}  // namespace

extern "C" {

size_t ep_masked_attention_workspace_bytes(int64_t n_keys, int n_heads) {
  return (size_t)attn_chunks(n_keys > 0 ? n_keys : 1) * (size_t)n_heads * AQ * sizeof(Partial);
}

// q [n_queries, n_heads*6] (projected, unscaled), k / v [n_keys, ld_kv] (projected; ld_kv % 4 == 0, 16-byte aligned),
// blocked uint8 [n_queries, n_keys] (non-zero = may not attend; NULL = no mask), out [n_queries, n_heads*6].
int ep_masked_attention(const float* q, const float* k, const float* v, int ld_kv, const uint8_t* blocked, int64_t n_keys,
                        int n_queries, int n_heads, int head_dim, float scale, float* out, void* workspace,
                        size_t workspace_bytes, cudaStream_t stream) {
  return ep_masked_attention_flagged(q, k, v, ld_kv, blocked, nullptr, n_keys, n_queries, n_heads, head_dim, scale, out, workspace,
                                     workspace_bytes, stream);
}

// row_unblocked (optional, int32 [n_queries]): 0 = every key of that query is blocked in `blocked`, so the query ignores the
// mask (what the reference does by rewriting the mask, mask3dformer.py:392); saves a pass over the [n_queries, n_keys] flags.
int ep_masked_attention_flagged(const float* q, const float* k, const float* v, int ld_kv, const uint8_t* blocked,
                                const int32_t* row_unblocked, int64_t n_keys, int n_queries, int n_heads, int head_dim, float scale,
                                float* out, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (n_keys <= 0 || n_keys > 0x7fffffffLL || n_queries < 1 || n_heads < 1 || ld_kv % 4 != 0 || ld_kv < n_heads * head_dim)
    return EP_ERR_ARG;
  if (head_dim != AD || n_queries > AQ || n_heads > 8 || (n_heads * AD) % 4 != 0) return EP_ERR_UNSUPPORTED;
  if (workspace_bytes < ep_masked_attention_workspace_bytes(n_keys, n_heads)) return EP_ERR_WORKSPACE;
  const int chunks = attn_chunks(n_keys);
  int keys_per_cta = ep_div_up(n_keys, chunks);
  keys_per_cta = ep_div_up(keys_per_cta, ATK) * ATK;
  const int grid = ep_div_up(n_keys, keys_per_cta);
  const int E = n_heads * AD;
  const size_t smem = (size_t)2 * ATK * E * sizeof(float) + (size_t)ATK * AQ;
  Partial* part = (Partial*)workspace;
  masked_attention_kernel<<<grid, n_heads * AQ, smem, stream>>>(q, k, v, ld_kv, blocked, row_unblocked, (int)n_keys, n_queries, n_heads,
                                                              scale, keys_per_cta, part);
  EP_CHECK_LAUNCH();
  attention_combine_kernel<<<1, n_heads * AQ, 0, stream>>>(part, grid, n_queries, n_heads, out);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

}  // extern "C"

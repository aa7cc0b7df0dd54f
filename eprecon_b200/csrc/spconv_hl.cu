// Sparse 3-D convolution, operand path rebuilt around the TMA engine ("hl" = half-pair operands).
//
//     out[j, :] = bias + sum_k W[k]^T . in[nbr[j, k], :]           (same contract as ep_spconv_fwd / ep_spconv_tc_fwd)
//
// What changed against csrc/spconv_tc.cu (round-1 kernel: 8 producer warps LDG -> cvt -> STS an fp32 operand that was
// split into tf32 hi/lo planes on the fly; L1TEX-bound at 81 %, tensor pipe 8.6 %, profiles/r01_spconv_tc_v7_*):
//   * activations reach this kernel PRE-SPLIT: every fp32 value x is stored by its producer as a pair of halfs
//     (h, l) with h = fp16(x), l = fp16((x - h) * 2^11) -- 22 significant bits, the same budget as the 3xTF32 split,
//     in 4 bytes per value instead of 8.  Rows are laid out in 32-channel slabs of 128 bytes [h0..h31 | l0..l31], which
//     is exactly one SWIZZLE_128B row of a K-major UMMA operand.
//   * no thread touches an operand: one warp issues 32 `cp.async.bulk.tensor.2d.tile::gather4` per pipeline stage (4
//     neighbour rows each, row indices straight from the staged neighbour table; a missing neighbour is an
//     out-of-bounds row, which the TMA unit zero-fills) and one tiled TMA load for the matching weight slab; both land
//     in shared memory in the canonical swizzled layout and complete on the stage's mbarrier (complete_tx).
//   * one thread issues tcgen05.mma.kind::f16 (M = 128, N = padded cout, K = 16): per stage and 16-channel step the
//     products h.h (rotating over three TMEM accumulators: the tensor core truncates when it chains an accumulator, see
//     spconv_tc.cu), l.h and h.l (a fourth accumulator, scaled by 2^-11 in the epilogue).
//   * 4 epilogue warps: tcgen05.ld, sum the accumulators in fp32, bias, store, per-CTA BatchNorm partial sums.
// SASS evidence: UTMALDG (gather4 + tile), UTCHMMA, UTCBAR, LDTM; no LDG/STS on the operand path.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"

extern "C" int ep_bn_finalize(const float* bn_partial, int num_row_tiles, int c, int64_t m, float eps, const float* gamma,
                              const float* beta, float* scale_shift, float* mean_var, cudaStream_t stream);

namespace {

using namespace eptc;

constexpr int HTM = 128;                   // output rows per CTA (UMMA M)
constexpr int H_THREADS = 192;             // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-5: epilogue
constexpr int A_STAGE = HTM * 128;         // bytes of one gathered A slab (128 rows x [32 h | 32 l] halfs)
constexpr int NBS = 132;                   // ints per offset in the staged neighbour table (16-byte aligned rows)
constexpr int MAX_STAGES = 8;

__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
                                            // leading byte offset 0: one swizzle atom along K (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;         // stride byte offset: one 8-row x 128-byte swizzle atom
  d |= (uint64_t)1 << 46;                   // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                   // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tma_gather4(const CUtensorMap* map, uint64_t* bar, void* smem_dst, int col, int r0, int r1,
                                            int r2, int r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3),
        "r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ void tma_tile2d(const CUtensorMap* map, uint64_t* bar, void* smem_dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}

// mbarrier wait that cannot hang the device: a barrier that does not flip within ~2 s is a protocol bug -> trap (the
// launch fails with an error instead of wedging the GPU).
__device__ __forceinline__ void mbar_wait_b(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  const long long t0 = clock64();
  uint32_t spins = 0;
  for (;;) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    if ((++spins & 1023u) == 0 && clock64() - t0 > 4000000000LL) __trap();
  }
}

__device__ __forceinline__ long long gtime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return (long long)t;
}

__global__ void __launch_bounds__(H_THREADS)
spconv_hl_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                 const int* __restrict__ nbr, int K, int nslab, int npad, int nt, int tmem_cols, int nstage, int cout,
                 int neg_row /* row index used for "no neighbour": any out-of-bounds row */,
                 const float* __restrict__ bias, float* __restrict__ out, int ld_out, int m_out,
                 float* __restrict__ bn_partial, int bn_rows, int splits, float* __restrict__ partial) {
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);   // SWIZZLE_128B atoms: 1024-byte aligned
  const int b_stage = nt * 128;
  const int stage_bytes = A_STAGE + b_stage;
  int* s_nbr = reinterpret_cast<int*>(smem + (size_t)nstage * stage_bytes);       // [K][NBS]
  __shared__ uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES];
  __shared__ uint64_t all_done;
  __shared__ uint32_t tmem_base_s;
  __shared__ int s_mask, s_used;
  __shared__ float s_red[8][128];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * HTM;
  const int col0 = blockIdx.y * nt;
  const int kb = (int)(((long long)blockIdx.z * K) / splits), ke = (int)(((long long)(blockIdx.z + 1) * K) / splits);

  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int s = 0; s < nstage; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&all_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_mask = 0;
    s_used = 0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // neighbour table of the tile, transposed to [k][row]; rows past m_out read as "no neighbour"
  {
    constexpr int PU = 6;
    const int total = HTM * K;
    const long long base = (long long)row0 * K, lim = (long long)m_out * K;
    unsigned mine = 0;
    for (int e0 = 0; e0 < total; e0 += H_THREADS * PU) {
      int v[PU];
#pragma unroll
      for (int u = 0; u < PU; ++u) {
        const int e = e0 + u * H_THREADS + tid;
        v[u] = -1;
        if (e < total && base + e < lim) v[u] = nbr ? __ldg(nbr + base + e) : row0 + e;   // no table: K == 1, identity
      }
#pragma unroll
      for (int u = 0; u < PU; ++u) {
        const int e = e0 + u * H_THREADS + tid;
        if (e < total) {
          const int r = e / K, kq = e - r * K;
          s_nbr[kq * NBS + r] = v[u] >= 0 ? v[u] : neg_row;
          if (v[u] >= 0 && kq >= kb && kq < ke) mine |= 1u << kq;
        }
      }
    }
    mine = __reduce_or_sync(0xffffffffu, mine);
    if (lane == 0 && mine) atomicOr(&s_mask, (int)mine);
  }
  __syncthreads();
  const unsigned kmask = (unsigned)s_mask;
  const uint32_t tmem_d = tmem_base_s;
  const int T = __popc(kmask) * nslab;      // pipeline stages of this CTA

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer: 32 gather4 + 1 weight tile per stage
    int stage = 0, round = 0;
    unsigned rem = kmask;
    while (rem) {
      const int k = __ffs(rem) - 1;
      rem &= rem - 1;
      const int4 rows = *reinterpret_cast<const int4*>(&s_nbr[k * NBS + 4 * lane]);
      for (int c = 0; c < nslab; ++c) {
        if (round > 0) mbar_wait_b(&empty_bar[stage], (round - 1) & 1);
        uint8_t* sa = smem + (size_t)stage * stage_bytes;
        if (lane == 0) {
          mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(A_STAGE + b_stage));
          tma_tile2d(&tm_b, &full_bar[stage], sa + A_STAGE, 0, (k * nslab + c) * npad + col0);
        }
        __syncwarp();
        tma_gather4(&tm_a, &full_bar[stage], sa + lane * 512, c * 64, rows.x, rows.y, rows.z, rows.w);
        if (++stage == nstage) { stage = 0; ++round; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------- MMA issuer (one thread)
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(nt >> 3) << 17) | ((uint32_t)(HTM >> 4) << 24);   // f16 x f16 -> f32
      uint32_t used = 0;
      int stage = 0, round = 0;
      for (int t = 0; t < T; ++t) {
        mbar_wait_b(&full_bar[stage], round & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes), sb = sa + A_STAGE;
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const uint64_t dah = umma_desc_sw128(sa + kk * 32), dal = umma_desc_sw128(sa + 64 + kk * 32);
          const uint64_t dbh = umma_desc_sw128(sb + kk * 32), dbl = umma_desc_sw128(sb + 64 + kk * 32);
          const int am = (t * 2 + kk) % 3;
          umma_f16(tmem_d + am * nt, dah, dbh, idesc, (used >> am) & 1u);
          used |= 1u << am;
          umma_f16(tmem_d + 3 * nt, dal, dbh, idesc, (used >> 3) & 1u);
          used |= 1u << 3;
          umma_f16(tmem_d + 3 * nt, dah, dbl, idesc, 1u);
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == nstage) { stage = 0; ++round; }
      }
      umma_commit(&all_done);
      s_used = (int)used;
    }
  }
  __syncthreads();
  if (warp >= 2) {
    // ------------------------------------------------------------- epilogue (warps 2..5 own TMEM lanes 32 * (warp & 3))
    if (T > 0) mbar_wait_b(&all_done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t used = (uint32_t)s_used;
    const int q = warp & 3;
    const int row = row0 + q * 32 + lane;
    for (int cb = 0; cb < nt; cb += 16) {
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = 0.f;
      if (T > 0) {
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          if ((used >> a) & 1u) {
            float t16[16];
            tmem_ld16(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * nt + cb), t16);
            const float sc = a == 3 ? (1.f / 2048.f) : 1.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaf(t16[i], sc, v[i]);
          }
        }
      }
      if (splits > 1) {   // raw partial sums; bias / statistics are applied by the split-K reduce
        if (row < m_out) {
          float4* dst = reinterpret_cast<float4*>(partial + ((size_t)blockIdx.z * m_out + row) * npad + col0 + cb);
#pragma unroll
          for (int i = 0; i < 4; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
        continue;
      }
      float s[16], sq[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int col = col0 + cb + i;
        const float val = v[i] + ((bias && col < cout) ? bias[col] : 0.f);
        const bool ok = row < m_out && col < cout;
        if (ok) out[(size_t)row * ld_out + col] = val;
        s[i] = ok ? val : 0.f;
        sq[i] = ok ? val * val : 0.f;
      }
      if (bn_partial) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
#pragma unroll
          for (int d = 16; d > 0; d >>= 1) {
            s[i] += __shfl_xor_sync(0xffffffffu, s[i], d);
            sq[i] += __shfl_xor_sync(0xffffffffu, sq[i], d);
          }
        }
        if (lane == 0) {
#pragma unroll
          for (int i = 0; i < 16; ++i) { s_red[q][cb + i] = s[i]; s_red[4 + q][cb + i] = sq[i]; }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (bn_partial && splits == 1 && tid < nt) {
    const int col = col0 + tid;
    if (col < cout) {
      const float s = (s_red[0][tid] + s_red[1][tid]) + (s_red[2][tid] + s_red[3][tid]);
      const float sq = (s_red[4][tid] + s_red[5][tid]) + (s_red[6][tid] + s_red[7][tid]);
      const int r0 = 2 * blockIdx.x;   // bn_partial is sized for 64-row tiles: fill entry 2*bx, zero 2*bx+1
      bn_partial[((size_t)r0 * 2 + 0) * cout + col] = s;
      bn_partial[((size_t)r0 * 2 + 1) * cout + col] = sq;
      if (r0 + 1 < bn_rows) {
        bn_partial[((size_t)(r0 + 1) * 2 + 0) * cout + col] = 0.f;
        bn_partial[((size_t)(r0 + 1) * 2 + 1) * cout + col] = 0.f;
      }
    }
  }
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols));
  }
}

// fp32 rows [m, ld] (c valid columns) -> half-pair rows [m][nslab][32 h | 32 l].  One thread = 8 channels of one row.
__global__ void __launch_bounds__(256)
hl_split_kernel(const float* __restrict__ src, int ld, int c, long long m, int nslab, uint4* __restrict__ dst,
                int* __restrict__ overflow) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m * nslab * 4) return;
  const int oct = (int)(t & 3);
  const long long rs = t >> 2;
  const int slab = (int)(rs % nslab);
  const long long row = rs / nslab;
  const int ch0 = slab * 32 + oct * 8;
  float v[8];
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const int ch = ch0 + 4 * g;
    float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ch < c) f = __ldg(reinterpret_cast<const float4*>(src + row * ld + ch));   // ld % 4 == 0: whole quads are in-row
    v[4 * g + 0] = f.x;
    v[4 * g + 1] = ch + 1 < c ? f.y : 0.f;
    v[4 * g + 2] = ch + 2 < c ? f.z : 0.f;
    v[4 * g + 3] = ch + 3 < c ? f.w : 0.f;
  }
  __half h[8], l[8];
  bool big = false;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    h[i] = __float2half_rn(v[i]);
    l[i] = __float2half_rn((v[i] - __half2float(h[i])) * 2048.f);
    big |= fabsf(v[i]) > 65000.f;
  }
  if (big && overflow) atomicOr(overflow, 1);
  uint4* d = dst + (row * nslab + slab) * 8;     // 8 x 16 bytes per slab row
  d[oct] = *reinterpret_cast<const uint4*>(h);
  d[4 + oct] = *reinterpret_cast<const uint4*>(l);
}

// ------------------------------------------------------------------------------------------------------------------
// cp.async producers.  Measured on B200 (profiles/r02_probe_hl_v1_tma_gather4_timing.json): the TMA unit retires a
// tile::gather4 of 128-byte rows in ~62 cycles (1.8-2.3 TB/s chip-wide, 8 bytes / clock / SM), slower than the register
// producers it replaced.  The half-pair layout makes the cheaper alternative possible: a 16-byte cp.async (LDGSTS) lands a
// gathered chunk directly in its swizzled slot of the UMMA operand -- no register staging, no conversion, no STS -- 8
// lanes per 128-byte row, so a warp instruction moves 4 whole rows (4 L1TEX wavefronts for 512 bytes).
constexpr int CP_THREADS = 256;            // warps 0-3: cp.async producers, warp 4 lane 0: MMA issuer, all 8 warps: epilogue
constexpr int CP_PRODUCERS = 128;
// Completion: every producer thread issues `cp.async.mbarrier.arrive.noinc` on the stage's full-barrier right after its
// copies -- the barrier receives that arrival when the thread's copies have landed -- and moves on to the next stage at once
// (the same protocol as CUTLASS's sm100 cp.async/UMMA mainloop).  The first version retired stages with cp.async.wait_group +
// fence.proxy.async + arrive in the producer: the proxy fence drains ALL of the thread's outstanding copies, i.e. one stage
// in flight per thread, and the kernel ran at 1 / (memory latency) per CTA (profiles/r02_probe_hl_v2_cpasync_fence_timing.json:
// 1 CTA per SM with an 8-deep ring was 1.75x SLOWER than 2 CTAs with 4 stages).  `.ca` instead of `.cg` made no difference.
template <bool CA>
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  if (CA) asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
  else asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// TICKET = false (default) compiles the in-kernel split-K reduce / BatchNorm finalisation out: with them the kernel needs 112
// registers (2 CTAs / SM), without 79 (3 CTAs / SM, the configuration the producers were tuned for; forcing 80 through
// __launch_bounds__ on the ticket variant spills).  tests/test_capi_cpu.py pins the register count of the default variant.
template <bool CA, bool TICKET>
__global__ void __launch_bounds__(CP_THREADS)
spconv_hl_cp_kernel(const uint8_t* __restrict__ in_hl, const __grid_constant__ CUtensorMap tm_b,
                    const int* __restrict__ nbr, int K, int nslab, int npad, int nt, int tmem_cols, int nstage, int cout,
                    const float* __restrict__ bias, float* __restrict__ out, int ld_out, int m_out,
                    float* __restrict__ bn_partial, int bn_rows, int splits, float* __restrict__ partial,
                    int* __restrict__ counters_arg /* zeroed tickets: [0] finished tiles, [1 + tile] finished splits; or null */,
                    const float* __restrict__ gamma, const float* __restrict__ beta, float bn_eps, float* __restrict__ ss_out,
                    long long* __restrict__ dbg /* optional timeline of one CTA: [stage][6] clock64 stamps */) {
  int* const counters = TICKET ? counters_arg : nullptr;
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const int b_stage = nt * 128;
  const int stage_bytes = A_STAGE + b_stage;
  int* s_nbr = reinterpret_cast<int*>(smem + (size_t)nstage * stage_bytes);       // [K][NBS]
  __shared__ uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES];
  __shared__ uint64_t all_done;
  __shared__ uint32_t tmem_base_s;
  __shared__ int s_mask, s_used;
  __shared__ float s_red[8][128];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * HTM;
  const int col0 = blockIdx.y * nt;
  const int kb = (int)(((long long)blockIdx.z * K) / splits), ke = (int)(((long long)(blockIdx.z + 1) * K) / splits);
  if (dbg && tid == 0 && blockIdx.x == gridDim.x / 2 && blockIdx.y == 0 && blockIdx.z == 0) { dbg[6 * 64 + 2] = clock64(); dbg[394] = gtime_ns(); }

  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    // full: every producer thread arrives once per stage + the expect_tx arrive that covers the weight tile's TMA bytes
    for (int s = 0; s < nstage; ++s) { mbar_init(&full_bar[s], CP_PRODUCERS + 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&all_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_mask = 0;
    s_used = 0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    constexpr int PU = 14;      // 128 x 27 / 256 threads: every load of the tile's neighbour slab is in flight at once
    const int total = HTM * K;
    const long long base = (long long)row0 * K, lim = (long long)m_out * K;
    unsigned mine = 0;
    for (int e0 = 0; e0 < total; e0 += CP_THREADS * PU) {
      int v[PU];
#pragma unroll
      for (int u = 0; u < PU; ++u) {
        const int e = e0 + u * CP_THREADS + tid;
        v[u] = -1;
        if (e < total && base + e < lim) v[u] = nbr ? __ldg(nbr + base + e) : row0 + e;
      }
#pragma unroll
      for (int u = 0; u < PU; ++u) {
        const int e = e0 + u * CP_THREADS + tid;
        if (e < total) {
          const int r = e / K, kq = e - r * K;
          s_nbr[kq * NBS + r] = v[u];
          if (v[u] >= 0 && kq >= kb && kq < ke) mine |= 1u << kq;
        }
      }
    }
    mine = __reduce_or_sync(0xffffffffu, mine);
    if (lane == 0 && mine) atomicOr(&s_mask, (int)mine);
  }
  __syncthreads();
  const unsigned kmask = (unsigned)s_mask;
  const uint32_t tmem_d = tmem_base_s;
  const int T = __popc(kmask) * nslab;

  if (warp < 4) {
    // ------------------------------------------------------------- producers: thread = chunk (tid & 7) of rows (tid >> 3) + 16 i
    const int chunk = tid & 7, rbase = tid >> 3;
    const uint32_t dst_thread = (uint32_t)(rbase * 128 + ((chunk ^ (rbase & 7)) << 4));   // SWIZZLE_128B: chunk ^ (row % 8)
    const size_t pitch = (size_t)nslab * 128;
    const uint8_t* src_thread = in_hl + chunk * 16;
    unsigned rem = kmask;
    int k = 0, c = 0;
    const uint8_t* rowp[8];
    auto set_k = [&]() {
      k = __ffs(rem) - 1;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = s_nbr[k * NBS + rbase + 16 * i];
        rowp[i] = r >= 0 ? src_thread + (size_t)r * pitch : nullptr;
      }
    };
    if (T > 0) set_k();
    int stage = 0, round = 0;
    const bool rec = dbg && tid == 0 && blockIdx.x == gridDim.x / 2 && blockIdx.y == 0 && blockIdx.z == 0;
    for (int t = 0; t < T; ++t) {
      if (rec && t < 64) dbg[6 * t + 0] = clock64();
      if (round > 0) mbar_wait_b(&empty_bar[stage], (round - 1) & 1);
      if (rec && t < 64) dbg[6 * t + 1] = clock64();
      const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
      if (tid == 0) {
        mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)b_stage);
        tma_tile2d(&tm_b, &full_bar[stage], smem + (size_t)stage * stage_bytes + A_STAGE, 0, (k * nslab + c) * npad + col0);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint8_t* p = rowp[i];
        cp_async16<CA>(sa + dst_thread + i * (16 * 128), p ? p + c * 128 : in_hl, p ? 16u : 0u);   // missing neighbour: zero fill
      }
      cp_async_arrive_noinc(&full_bar[stage]);    // arrives once this thread's copies (of this and earlier stages) have landed
      if (rec && t < 64) dbg[6 * t + 2] = clock64();
      if (++c == nslab) {
        c = 0;
        rem &= rem - 1;
        if (rem) set_k();
      }
      if (++stage == nstage) { stage = 0; ++round; }
    }
  } else if (warp == 4) {
    // ------------------------------------------------------------- MMA issuer: warp 4 stays converged, one elected lane issues
    // (a divergent `if (lane == 0)` makes the compiler wrap every UTCHMMA in an ELECT / BRA.U.ANY loop; the first version also
    //  ran a proxy fence = MEMBAR.ALL.CTA per stage here: ~600 cycles to issue 6 MMAs, the bottleneck of the whole kernel,
    //  profiles/r02_timeline_hl_v3.json)
    const uint32_t idesc = (1u << 4) | ((uint32_t)(nt >> 3) << 17) | ((uint32_t)(HTM >> 4) << 24);   // f16 x f16 -> f32
    // descriptors differ only in their 14-bit start-address field: base + stage * (stage_bytes >> 4) + {0,2,4,6} (32-byte k steps)
    const uint64_t da0 = umma_desc_sw128(smem_u32(smem)), db0 = umma_desc_sw128(smem_u32(smem) + A_STAGE);
    const uint32_t dstep = (uint32_t)stage_bytes >> 4;
    uint32_t elected;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(elected));
    uint32_t used = 0;
    int stage = 0, round = 0;
    const bool rec = dbg && elected && blockIdx.x == gridDim.x / 2 && blockIdx.y == 0 && blockIdx.z == 0;
    for (int t = 0; t < T; ++t) {
      if (rec && t < 64) dbg[6 * t + 3] = clock64();
      mbar_wait_b(&full_bar[stage], round & 1);
      if (rec && t < 64) dbg[6 * t + 4] = clock64();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elected) {
        const uint64_t da = da0 + (uint64_t)(stage * dstep), db = db0 + (uint64_t)(stage * dstep);
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const uint64_t dah = da + 2 * kk, dal = da + 4 + 2 * kk, dbh = db + 2 * kk, dbl = db + 4 + 2 * kk;
          const int am = (t * 2 + kk) % 3;
          umma_f16(tmem_d + am * nt, dah, dbh, idesc, (used >> am) & 1u);
          used |= 1u << am;
          umma_f16(tmem_d + 3 * nt, dal, dbh, idesc, (used >> 3) & 1u);
          used |= 1u << 3;
          umma_f16(tmem_d + 3 * nt, dah, dbl, idesc, 1u);
        }
        umma_commit(&empty_bar[stage]);
      }
      __syncwarp();
      if (rec && t < 64) dbg[6 * t + 5] = clock64();
      if (++stage == nstage) { stage = 0; ++round; }
    }
    if (elected) {
      umma_commit(&all_done);
      s_used = (int)used;
    }
  }
  __syncthreads();
  if (dbg && tid == 0 && blockIdx.x == gridDim.x / 2 && blockIdx.y == 0 && blockIdx.z == 0) dbg[6 * 64] = clock64();
  // ------------------------------------------------------------- epilogue: warp w owns TMEM lanes 32 (w & 3) and one half of the columns
  if (T > 0) mbar_wait_b(&all_done, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t used = (uint32_t)s_used;
  const int q = warp & 3, half = warp >> 2;
  const int row = row0 + q * 32 + lane;
  const int ncol_half = (nt / 16 + 1) / 2 * 16;
  const int cbeg = half == 0 ? 0 : ncol_half, cend = half == 0 ? min(ncol_half, nt) : nt;
  auto load_acc = [&](int cb, float (&v)[16]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = 0.f;
    if (T > 0) {
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        if ((used >> a) & 1u) {
          float t16[16];
          tmem_ld16(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * nt + cb), t16);
          const float sc = a == 3 ? (1.f / 2048.f) : 1.f;
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = fmaf(t16[i], sc, v[i]);
        }
      }
    }
  };
  // split-K: every split writes its raw partial sums; with `counters` the LAST split to finish a tile (ticket) reduces them
  // in fixed z order right here (deterministic: the order does not depend on which CTA happens to be last), without
  // `counters` a separate reduce kernel does (ep_internal_splitk_reduce).
  bool do_final = true, from_partial = false;
  if (splits > 1) {
    for (int cb = cbeg; cb < cend; cb += 16) {
      float v[16];
      load_acc(cb, v);
      if (row < m_out) {
        float4* dst = reinterpret_cast<float4*>(partial + ((size_t)blockIdx.z * m_out + row) * npad + col0 + cb);
#pragma unroll
        for (int i = 0; i < 4; ++i) __stcg(dst + i, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
      }
    }
    do_final = false;
    if (counters) {
      __threadfence();
      __syncthreads();
      if (tid == 0) {
        int* ctr = counters + 1 + blockIdx.y * gridDim.x + blockIdx.x;
        const int old = atomicAdd(ctr, 1);
        s_mask = old == splits - 1;
        if (old == splits - 1) *ctr = 0;          // self-resetting: the buffer is all zero again when the launch retires
      }
      __syncthreads();
      do_final = s_mask != 0;
      from_partial = true;
      if (do_final) __threadfence();
      if (dbg && do_final && tid == 0 && blockIdx.x == gridDim.x / 2 && blockIdx.y == 0) { dbg[390] = gtime_ns(); dbg[389] = blockIdx.z; }
    }
  }
  if (do_final && from_partial) {
    // ---- last split of the tile: reduce the `splits` partial planes, coalesced -- a warp walks rows, its lanes the row's
    //      float4 column quads (one 512-byte row per load instruction; the first version mapped a lane to a ROW: 32 L1TEX
    //      wavefronts per load, 50-120 us per tile).  Fixed order: z ascending per element, rows ascending per warp, warps
    //      0..7 for the column statistics -> deterministic regardless of which CTA won the ticket.
    const int nq = nt >> 2;
    const size_t zstride = (size_t)m_out * npad;
    float cs[4] = {0.f, 0.f, 0.f, 0.f}, cq[4] = {0.f, 0.f, 0.f, 0.f};
    float bv[4] = {0.f, 0.f, 0.f, 0.f};
    if (lane < nq && bias) {
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = (col0 + 4 * lane + j < cout) ? bias[col0 + 4 * lane + j] : 0.f;
    }
    for (int r = warp; r < HTM; r += 2 * (CP_THREADS / 32)) {
      // two rows per trip: 2 x 8 independent 16-byte loads in flight per lane
      const int rowa = row0 + r, rowb = rowa + CP_THREADS / 32;
      if (rowa >= m_out) break;
      const bool hasb = rowb < m_out && r + CP_THREADS / 32 < HTM;
      if (lane < nq) {
        const float* srca = partial + (size_t)rowa * npad + col0 + 4 * lane;
        const float* srcb = partial + (size_t)rowb * npad + col0 + 4 * lane;
        float4 acca = make_float4(0.f, 0.f, 0.f, 0.f), accb = acca;
        for (int z0 = 0; z0 < splits; z0 += 8) {
          float4 ta[8], tb[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const bool on = z0 + u < splits;
            ta[u] = on ? __ldcg(reinterpret_cast<const float4*>(srca + (size_t)(z0 + u) * zstride)) : make_float4(0.f, 0.f, 0.f, 0.f);
            tb[u] = (on && hasb) ? __ldcg(reinterpret_cast<const float4*>(srcb + (size_t)(z0 + u) * zstride)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            if (z0 + u < splits) {
              acca.x += ta[u].x; acca.y += ta[u].y; acca.z += ta[u].z; acca.w += ta[u].w;
              accb.x += tb[u].x; accb.y += tb[u].y; accb.z += tb[u].z; accb.w += tb[u].w;
            }
          }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (h == 1 && !hasb) break;
          const float4 acc = h == 0 ? acca : accb;
          const int row_h = h == 0 ? rowa : rowb;
          const float vals[4] = {acc.x + bv[0], acc.y + bv[1], acc.z + bv[2], acc.w + bv[3]};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int col = col0 + 4 * lane + j;
            if (col < cout) {
              out[(size_t)row_h * ld_out + col] = vals[j];
              cs[j] += vals[j];
              cq[j] += vals[j] * vals[j];
            }
          }
        }
      }
    }
    if (bn_partial) {
      float* red = reinterpret_cast<float*>(smem);     // [8 warps][2][128]; the operand ring is idle by now
      if (lane < nq) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          red[(warp * 2 + 0) * 128 + 4 * lane + j] = cs[j];
          red[(warp * 2 + 1) * 128 + 4 * lane + j] = cq[j];
        }
      }
      __syncthreads();
      if (tid < nt && col0 + tid < cout) {
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int w = 0; w < CP_THREADS / 32; ++w) { s1 += red[(w * 2 + 0) * 128 + tid]; s2 += red[(w * 2 + 1) * 128 + tid]; }
        const int col = col0 + tid;
        const int r0 = 2 * blockIdx.x;   // bn_partial is sized for 64-row tiles: fill entry 2*bx, zero 2*bx+1
        __stcg(&bn_partial[((size_t)r0 * 2 + 0) * cout + col], s1);
        __stcg(&bn_partial[((size_t)r0 * 2 + 1) * cout + col], s2);
        if (r0 + 1 < bn_rows) {
          __stcg(&bn_partial[((size_t)(r0 + 1) * 2 + 0) * cout + col], 0.f);
          __stcg(&bn_partial[((size_t)(r0 + 1) * 2 + 1) * cout + col], 0.f);
        }
      }
    }
  } else if (do_final) {
    for (int cb = cbeg; cb < cend; cb += 16) {
      float v[16];
      load_acc(cb, v);
      float s[16], sq[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int col = col0 + cb + i;
        const float val = v[i] + ((bias && col < cout) ? bias[col] : 0.f);
        const bool ok = row < m_out && col < cout;
        if (ok) out[(size_t)row * ld_out + col] = val;
        s[i] = ok ? val : 0.f;
        sq[i] = ok ? val * val : 0.f;
      }
      if (bn_partial) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
#pragma unroll
          for (int d = 16; d > 0; d >>= 1) {
            s[i] += __shfl_xor_sync(0xffffffffu, s[i], d);
            sq[i] += __shfl_xor_sync(0xffffffffu, sq[i], d);
          }
        }
        if (lane == 0) {
#pragma unroll
          for (int i = 0; i < 16; ++i) { s_red[q][cb + i] = s[i]; s_red[4 + q][cb + i] = sq[i]; }
        }
      }
    }
  }
  if (dbg && do_final && from_partial && tid == 0 && blockIdx.x == gridDim.x / 2 && blockIdx.y == 0) dbg[391] = gtime_ns();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (bn_partial && do_final && !from_partial && tid < nt) {
    const int col = col0 + tid;
    if (col < cout) {
      const float s = (s_red[0][tid] + s_red[1][tid]) + (s_red[2][tid] + s_red[3][tid]);
      const float sq = (s_red[4][tid] + s_red[5][tid]) + (s_red[6][tid] + s_red[7][tid]);
      const int r0 = 2 * blockIdx.x;   // bn_partial is sized for 64-row tiles: fill entry 2*bx, zero 2*bx+1
      __stcg(&bn_partial[((size_t)r0 * 2 + 0) * cout + col], s);
      __stcg(&bn_partial[((size_t)r0 * 2 + 1) * cout + col], sq);
      if (r0 + 1 < bn_rows) {
        __stcg(&bn_partial[((size_t)(r0 + 1) * 2 + 0) * cout + col], 0.f);
        __stcg(&bn_partial[((size_t)(r0 + 1) * 2 + 1) * cout + col], 0.f);
      }
    }
  }
  // fused BatchNorm finalisation: the last tile to finish (ticket) turns the per-tile partial sums into the per-column
  // scale / shift, summing in exactly the order of bn_finalize_kernel (csrc/spconv.cu) -- bit-identical to the separate launch.
  if (ss_out && bn_partial && counters && do_final) {
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      const int old = atomicAdd(counters, 1);
      const int total_tiles = gridDim.x * gridDim.y;
      s_mask = old == total_tiles - 1;
      if (old == total_tiles - 1) *counters = 0;
    }
    __syncthreads();
    if (dbg && s_mask && tid == 0) dbg[392] = gtime_ns();
    if (s_mask) {
      __threadfence();
      double* s_s = reinterpret_cast<double*>(smem);          // [32][33]; the operand ring is idle by now
      double* s_q = s_s + 32 * 33;
      const int tx = tid & 31, ty = tid >> 5;
      for (int c0 = 0; c0 < cout; c0 += 32) {
        const int col = c0 + tx;
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          const int r = ty + 8 * rr;
          double sd = 0.0, qd = 0.0;
          if (col < cout)
#pragma unroll 4
            for (int b2 = r; b2 < bn_rows; b2 += 32) {
              sd += (double)__ldcg(&bn_partial[((size_t)b2 * 2 + 0) * cout + col]);
              qd += (double)__ldcg(&bn_partial[((size_t)b2 * 2 + 1) * cout + col]);
            }
          s_s[r * 33 + tx] = sd;
          s_q[r * 33 + tx] = qd;
        }
        __syncthreads();
        if (ty == 0 && col < cout) {
          double sd = s_s[tx], qd = s_q[tx];
#pragma unroll
          for (int r = 1; r < 32; ++r) { sd += s_s[r * 33 + tx]; qd += s_q[r * 33 + tx]; }
          const double mean = sd / m_out;
          double var = qd / m_out - mean * mean;
          if (var < 0.0) var = 0.0;
          const float inv = (float)(1.0 / sqrt(var + (double)bn_eps));
          const float g = gamma ? gamma[col] : 1.f, bt = beta ? beta[col] : 0.f;
          const float sc = inv * g;
          ss_out[col] = sc;
          ss_out[cout + col] = bt - (float)mean * sc;
        }
        __syncthreads();
      }
      if (dbg && tid == 0) dbg[393] = gtime_ns();
    }
  }
  if (dbg && tid == 0 && blockIdx.x == gridDim.x / 2 && blockIdx.y == 0 && blockIdx.z == 0) { dbg[6 * 64 + 1] = clock64(); dbg[395] = gtime_ns(); }
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols));
  }
}

__device__ __forceinline__ void hl_store8(uint4* __restrict__ dst_slab_row /* 8 x uint4 */, int oct, const float (&v)[8],
                                          int* __restrict__ overflow) {
  __half h[8], l[8];
  bool big = false;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    h[i] = __float2half_rn(v[i]);
    l[i] = __float2half_rn((v[i] - __half2float(h[i])) * 2048.f);
    big |= fabsf(v[i]) > 65000.f;
  }
  if (big && overflow) atomicOr(overflow, 1);
  dst_slab_row[oct] = *reinterpret_cast<const uint4*>(h);
  dst_slab_row[4 + oct] = *reinterpret_cast<const uint4*>(l);
}

// BatchNorm apply (+ residual, + ReLU) that ALSO writes the half-pair copy the next sparse conv gathers (fuses
// ep_affine_act with ep_hl_split_rows):  y = act(a*sa + ta [+ b*sb + tb]);  out (fp32, may alias a) and out_hl.
__global__ void __launch_bounds__(256)
hl_affine_act_kernel(const float* __restrict__ a, int ld_a, const float* __restrict__ ssa, const float* __restrict__ b, int ld_b,
                     const float* __restrict__ ssb, int relu, long long m, int c, int nslab, float* __restrict__ out, int ld_out,
                     uint4* __restrict__ out_hl, int* __restrict__ overflow) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m * nslab * 4) return;
  const int oct = (int)(t & 3);
  const long long rs = t >> 2;
  const int slab = (int)(rs % nslab);
  const long long row = rs / nslab;
  const int ch0 = slab * 32 + oct * 8;
  float v[8];
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const int ch = ch0 + 4 * g;
    float x[4] = {0.f, 0.f, 0.f, 0.f};
    if (ch < c) {
      const float4 f = __ldg(reinterpret_cast<const float4*>(a + row * ld_a + ch));
      x[0] = f.x; x[1] = f.y; x[2] = f.z; x[3] = f.w;
      float y[4] = {0.f, 0.f, 0.f, 0.f};
      if (b) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(b + row * ld_b + ch));
        y[0] = q.x; y[1] = q.y; y[2] = q.z; y[3] = q.w;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = ch + j;
        float r = 0.f;
        if (col < c) {
          r = x[j];
          if (ssa) r = fmaf(r, ssa[col], ssa[c + col]);
          if (b) {
            float w = y[j];
            if (ssb) w = fmaf(w, ssb[col], ssb[c + col]);
            r += w;
          }
          if (relu) r = fmaxf(r, 0.f);
        }
        x[j] = r;
      }
      *reinterpret_cast<float4*>(out + row * ld_out + ch) = make_float4(x[0], x[1], x[2], x[3]);
    }
    v[4 * g + 0] = x[0]; v[4 * g + 1] = x[1]; v[4 * g + 2] = x[2]; v[4 * g + 3] = x[3];
  }
  hl_store8(out_hl + (row * nslab + slab) * 8, oct, v, overflow);
}

// voxelisation (segmented mean of point rows, as ep_segment_mean) that also writes the half-pair copy of its output
__global__ void __launch_bounds__(256)
hl_segment_mean_kernel(const float* __restrict__ feat, int ld_in, int c, const int* __restrict__ perm,
                       const int* __restrict__ seg_start, const int* __restrict__ seg_end, long long m, int nslab,
                       float* __restrict__ out, int ld_out, uint4* __restrict__ out_hl, int* __restrict__ overflow) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m * nslab * 4) return;
  const int oct = (int)(t & 3);
  const long long rs = t >> 2;
  const int slab = (int)(rs % nslab);
  const long long s = rs / nslab;
  const int ch0 = slab * 32 + oct * 8;
  const int a = seg_start[s], b = seg_end[s];
  const float cnt = (float)(b - a);
  float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const bool g0 = ch0 < c, g1 = ch0 + 4 < c;      // whole quads: the input's padding columns (ld % 4 == 0) are carried along
  if (g0) {
    for (int p = a; p < b; ++p) {
      const float* src = feat + (size_t)perm[p] * ld_in + ch0;
      const float4 x = *reinterpret_cast<const float4*>(src);
      v[0] += __fdiv_rn(x.x, cnt); v[1] += __fdiv_rn(x.y, cnt); v[2] += __fdiv_rn(x.z, cnt); v[3] += __fdiv_rn(x.w, cnt);
      if (g1) {
        const float4 y = *reinterpret_cast<const float4*>(src + 4);
        v[4] += __fdiv_rn(y.x, cnt); v[5] += __fdiv_rn(y.y, cnt); v[6] += __fdiv_rn(y.z, cnt); v[7] += __fdiv_rn(y.w, cnt);
      }
    }
    *reinterpret_cast<float4*>(out + (size_t)s * ld_out + ch0) = make_float4(v[0], v[1], v[2], v[3]);
    if (g1) *reinterpret_cast<float4*>(out + (size_t)s * ld_out + ch0 + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (ch0 + j >= c) v[j] = 0.f;               // the half-pair copy is zero beyond the valid channels
  hl_store8(out_hl + (s * nslab + slab) * 8, oct, v, overflow);
}

// debugging aid: one warp gathers 128 rows of slab `slab` with 32 gather4 and dumps the raw 16 KB of shared memory
__global__ void __launch_bounds__(32)
hl_probe_gather4_kernel(const __grid_constant__ CUtensorMap tm_a, const int* __restrict__ rows128, int slab,
                        uint4* __restrict__ out, int* __restrict__ status) {
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  const int lane = threadIdx.x;
  if (lane == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = lane; i < A_STAGE / 16; i += 32) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0xdeadbeefu, 0xdeadbeefu, 0xdeadbeefu, 0xdeadbeefu);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (lane == 0) mbar_arrive_expect_tx(&bar, A_STAGE);
  __syncwarp();
  const int4 r = reinterpret_cast<const int4*>(rows128)[lane];
  tma_gather4(&tm_a, &bar, smem + lane * 512, slab * 64, r.x, r.y, r.z, r.w);
  // bounded wait: report a timeout instead of trapping
  const uint32_t addr = smem_u32(&bar);
  const long long t0 = clock64();
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(0u) : "memory");
    if (!done && clock64() - t0 > 400000000LL) break;
  }
  if (lane == 0) *status = done ? 1 : -1;
  __syncwarp();
  for (int i = lane; i < A_STAGE / 16; i += 32) out[i] = reinterpret_cast<const uint4*>(smem)[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}

// 2-D map over rows of `row_halfs` halfs (pitch = row_halfs * 2 bytes), box = 64 halfs x box_rows, SWIZZLE_128B
bool make_map(CUtensorMap* map, const void* base, uint64_t row_halfs, uint64_t rows, uint32_t box_rows) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  const cuuint64_t dims[2] = {row_halfs, rows};
  const cuuint64_t strides[1] = {row_halfs * 2};
  const cuuint32_t box[2] = {64, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

long long* g_hl_timeline = nullptr;   // device buffer [6 * 64 + 3] set by ep_hl_set_timeline (debug)
int g_hl_debug = 0;   // last failure site of ep_spconv_hl_fwd (1 map A, 2 map B, 3 smem attribute, 4 launch); ep_hl_debug_code()

inline int pow2_cols(int n) {
  int c = 32;
  while (c < n) c <<= 1;
  return c;
}

// split-K over the kernel offsets for sub-wave problems.  nslab > 0 additionally keeps >= ~8 pipeline stages per CTA: a CTA
// pays ~6 us of prologue + epilogue and writes a full fp32 partial tile, so 27 splits of 4 stages each (1-tile problems) spent
// more time on partial sums than on the convolution; nslab == 0 gives the upper bound the workspace query needs.
int hl_splits(int64_t m_out, int npad, int K, int nslab = 0) {
  const int nt = npad > 128 ? 128 : npad;
  const long long ctas = (long long)ep_div_up(m_out, HTM) * (npad / nt);
  if (K < 2 || ctas * 2 > EP_NUM_SMS) return 1;
  long long s = EP_NUM_SMS / ctas;
  if (s > K) s = K;
  if (nslab > 0) {
    const long long by_stages = ((long long)K * nslab) / 8;
    if (s > by_stages) s = by_stages;
  }
  return s < 2 ? 1 : (int)s;
}

}  // namespace

extern "C" {

int ep_hl_slabs(int c) { return (c + 31) / 32; }

// src f32 [m, ld_src] (c valid columns, ld_src % 4 == 0) -> dst half pairs [m][ep_hl_slabs(c)][64] (128 bytes per slab).
// overflow (optional): set to 1 when a value is outside the half range (|x| > 65000).
int ep_hl_split_rows(const float* src, int ld_src, int c, int64_t m, uint16_t* dst, int32_t* overflow, cudaStream_t stream) {
  if (m <= 0 || c < 1 || ld_src % 4 != 0 || ((c + 3) / 4 * 4) > ld_src) return EP_ERR_ARG;
  if (((uintptr_t)src & 15) || ((uintptr_t)dst & 15)) return EP_ERR_ARG;
  const int nslab = (c + 31) / 32;
  const long long total = (long long)m * nslab * 4;
  hl_split_kernel<<<ep_div_up(total, 256), 256, 0, stream>>>(src, ld_src, c, (long long)m, nslab, reinterpret_cast<uint4*>(dst),
                                                             overflow);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

// ep_affine_act + half-pair copy of the result for the next sparse conv (out may alias a; out_hl [m][ep_hl_slabs(c)][64])
int ep_hl_affine_act(const float* a, int ld_a, const float* ss_a, const float* b, int ld_b, const float* ss_b, int relu, int64_t m,
                     int c, float* out, int ld_out, uint16_t* out_hl, int32_t* overflow, cudaStream_t stream) {
  if (m <= 0 || c < 1 || ld_a % 4 || ld_out % 4 || (b && ld_b % 4) || !out_hl) return EP_ERR_ARG;
  if (((uintptr_t)a & 15) || ((uintptr_t)out & 15) || ((uintptr_t)out_hl & 15) || (b && ((uintptr_t)b & 15))) return EP_ERR_ARG;
  const int nslab = (c + 31) / 32;
  const long long total = (long long)m * nslab * 4;
  hl_affine_act_kernel<<<ep_div_up(total, 256), 256, 0, stream>>>(a, ld_a, ss_a, b, ld_b, ss_b, relu, (long long)m, c, nslab, out, ld_out,
                                                                  reinterpret_cast<uint4*>(out_hl), overflow);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

// ep_segment_mean + half-pair copy of the result
int ep_hl_segment_mean(const float* feat, int ld_in, int c, const int32_t* perm, const int32_t* seg_start, const int32_t* seg_end,
                       int64_t m, float* out, int ld_out, uint16_t* out_hl, int32_t* overflow, cudaStream_t stream) {
  if (m <= 0 || c < 1 || ld_in % 4 || ld_out % 4 || !out_hl) return EP_ERR_ARG;
  if (((uintptr_t)feat & 15) || ((uintptr_t)out & 15) || ((uintptr_t)out_hl & 15)) return EP_ERR_ARG;
  const int nslab = (c + 31) / 32;
  const long long total = (long long)m * nslab * 4;
  hl_segment_mean_kernel<<<ep_div_up(total, 256), 256, 0, stream>>>(feat, ld_in, c, perm, seg_start, seg_end, (long long)m, nslab, out,
                                                                    ld_out, reinterpret_cast<uint4*>(out_hl), overflow);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

size_t ep_spconv_hl_workspace_bytes(int64_t m_out, int npad, int K) {
  const int s = hl_splits(m_out, npad, K);
  return s > 1 ? (size_t)s * (size_t)m_out * (size_t)npad * sizeof(float) : 0;
}

// in_hl: half-pair rows of the m_in input rows (ep_hl_split_rows); w_hl: half-pair weights [K][nslab][npad][64]
// (w_hl[k][s][n][j] = fp16(W[k][32 s + j][n]), [..][32 + j] = fp16 of the remainder * 2^11; zero padded), npad = cout
// rounded up to a multiple of 16 (of 128 when larger).  neg_row_mode 0: a missing neighbour is gathered as row m_in (first
// out-of-bounds row), 1: as row -1.  Other arguments as ep_spconv_tc_fwd.
// Fused variant.  counters: device int32 [counters_len], all zero on entry and all zero again when the launch retires
// (tickets; one buffer per stream), or NULL.  With counters, split-K partial sums are reduced by the last split of each tile
// inside the conv kernel (no second launch).  ss_out (optional, float [2, cout]) receives the train-mode BatchNorm scale /
// shift of the output columns (gamma / beta may be NULL = 1 / 0): fused into the conv kernel (last tile) for small outputs,
// a separate ep_bn_finalize launch otherwise -- bit-identical either way.  bn_partial must be given with ss_out.
int ep_spconv_hl_fused_fwd(const uint16_t* in_hl, int64_t m_in, int cin, const int32_t* nbr, int K, const uint16_t* w_hl, int npad,
                           int cout, const float* bias, float* out, int ld_out, int64_t m_out, float* bn_partial, void* workspace,
                           size_t workspace_bytes, int neg_row_mode, int32_t* counters, int counters_len, const float* gamma,
                           const float* beta, float eps, float* ss_out, cudaStream_t stream) {
  if (m_out <= 0 || m_in <= 0 || m_in > 0x7ffffff0LL || cin < 1 || cout < 1 || K < 1 || npad % 16 != 0 || npad < cout)
    return EP_ERR_ARG;
  if (!nbr && K != 1) return EP_ERR_ARG;
  if (K > 27) return EP_ERR_UNSUPPORTED;
  if (((uintptr_t)in_hl & 15) || ((uintptr_t)w_hl & 15)) return EP_ERR_ARG;
  if (ss_out && !bn_partial) return EP_ERR_ARG;
  int nt = npad;
  if (npad > 128) {
    if (npad % 128 != 0) return EP_ERR_ARG;
    nt = 128;
  }
  const int nslab = (cin + 31) / 32;
  const int tmem_cols = pow2_cols(4 * nt);
  if (tmem_cols > 512) return EP_ERR_UNSUPPORTED;
  // operand producer: cp.async (default) or the TMA tile::gather4 path (EPRECON_HL_PRODUCER=tma; measured slower, kept as evidence)
  static const bool use_tma = [] { const char* v = getenv("EPRECON_HL_PRODUCER"); return v && v[0] == 't'; }();
  CUtensorMap tm_a, tm_b;
  if (use_tma && !make_map(&tm_a, in_hl, (uint64_t)nslab * 64, (uint64_t)m_in, 1)) { g_hl_debug = 1; return EP_ERR_CUDA; }
  if (!make_map(&tm_b, w_hl, 64, (uint64_t)K * nslab * npad, (uint32_t)nt)) { g_hl_debug = 2; return EP_ERR_CUDA; }
  const int stage_bytes = A_STAGE + nt * 128;
  // TMEM decides how many CTAs share an SM (512 columns): give each the deepest ring its share of shared memory allows
  static const int knob_ctas = [] { const char* v = getenv("EPRECON_HL_CTAS"); return v ? atoi(v) : 0; }();
  // CTAs per SM: as many as TMEM allows, up to 3 -- the LSU's cp.async rate (~17 cycles per 512-byte gather instruction per SM)
  // bounds the main loop, and only OTHER CTAs' main loops can fill the prologue / epilogue bubbles of a tile (2 -> 3 CTAs:
  // -20 % on the level-2 convs even though the ring shrinks from 4 to 2 stages; profiles/r02_probe_hl_v4_*)
  int ctas_per_sm = 512 / tmem_cols >= 3 ? 3 : (512 / tmem_cols >= 2 ? 2 : 1);
  if (knob_ctas >= 1 && knob_ctas <= 3 && 512 / tmem_cols >= knob_ctas) ctas_per_sm = knob_ctas;
  // shared memory per SM: 228 KB, per CTA at most 227 KB incl. ~4.4 KB of static barriers / reduction scratch + 1 KB reserved
  const int budget = (ctas_per_sm >= 4 ? 50 : ctas_per_sm == 3 ? 68 : ctas_per_sm == 2 ? 104 : 216) * 1024 - K * NBS * 4 - 1024;
  int nstage = budget / stage_bytes;
  if (nstage > MAX_STAGES) nstage = MAX_STAGES;
  if (nstage < 2) return EP_ERR_UNSUPPORTED;
  const size_t smem = (size_t)nstage * stage_bytes + (size_t)K * NBS * sizeof(int) + 1024;
  static const cudaError_t attr = [] {
    cudaError_t e1 = cudaFuncSetAttribute(spconv_hl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024);
    cudaError_t e2 = cudaFuncSetAttribute(spconv_hl_cp_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024);
    cudaError_t e3 = cudaFuncSetAttribute(spconv_hl_cp_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024);
    cudaError_t e4 = cudaFuncSetAttribute(spconv_hl_cp_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024);
    cudaError_t e5 = cudaFuncSetAttribute(spconv_hl_cp_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024);
    return e1 != cudaSuccess ? e1 : e2 != cudaSuccess ? e2 : e3 != cudaSuccess ? e3 : e4 != cudaSuccess ? e4 : e5;
  }();
  if (attr != cudaSuccess) { g_hl_debug = 3; return EP_ERR_CUDA; }
  const int splits = hl_splits(m_out, npad, K, nslab);
  if (splits > 1 && workspace_bytes < (size_t)splits * (size_t)m_out * (size_t)npad * sizeof(float)) return EP_ERR_WORKSPACE;
  float* partial = splits > 1 ? (float*)workspace : nullptr;
  dim3 grid(ep_div_up(m_out, HTM), npad / nt, splits);
  const int bn_rows = ep_div_up(m_out, 64);
  const int tiles = (int)(grid.x * grid.y);
  // The ticket path (last split of a tile reduces the partial sums in-kernel; last tile finalises the BatchNorm) is OFF by
  // default: measured on B200 it is 2-4x SLOWER than the two small extra launches on exactly the problems it was meant for
  // (87-640 rows: 185 vs 42 us) -- one CTA re-reading 27 split planes with a lane-per-row pattern (32 L1TEX wavefronts per
  // load) and a single-CTA finalisation cannot compete with the wide, coalesced reduce / finalise kernels.
  // EPRECON_HL_FUSE=1 enables it (profiles/r02_probe_small_conv_fused_vs_unfused.txt).
  static const bool knob_fuse = [] { const char* v = getenv("EPRECON_HL_FUSE"); return v && v[0] == '1'; }();
  int32_t* ctr = (knob_fuse && !use_tma && counters && counters_len >= 1 + tiles) ? counters : nullptr;
  // the finalisation runs on ONE CTA when fused: worth it while the partial-sum table is small (<= 1024 64-row tiles)
  const bool fuse_bn = ss_out && ctr && bn_rows <= 1024;
  if (use_tma) {
    const int neg_row = neg_row_mode == 1 ? -1 : (int)m_in;
    spconv_hl_kernel<<<grid, H_THREADS, smem, stream>>>(tm_a, tm_b, nbr, K, nslab, npad, nt, tmem_cols, nstage, cout, neg_row, bias,
                                                        out, ld_out, (int)m_out, bn_partial, bn_rows, splits, partial);
  } else {
    static const bool knob_ca = [] { const char* v = getenv("EPRECON_HL_CA"); return v && v[0] == '1'; }();
    auto kern = knob_ca ? (ctr ? spconv_hl_cp_kernel<true, true> : spconv_hl_cp_kernel<true, false>)
                        : (ctr ? spconv_hl_cp_kernel<false, true> : spconv_hl_cp_kernel<false, false>);
    kern<<<grid, CP_THREADS, smem, stream>>>(reinterpret_cast<const uint8_t*>(in_hl), tm_b, nbr, K, nslab, npad, nt, tmem_cols, nstage,
                                             cout, bias, out, ld_out, (int)m_out, bn_partial, bn_rows, splits, partial, ctr, gamma,
                                             beta, eps, fuse_bn ? ss_out : nullptr, g_hl_timeline);
  }
  if (splits > 1 && !ctr) {
    const int st = ep_internal_splitk_reduce(partial, splits, (int)m_out, npad, cout, bias, out, ld_out, bn_partial, stream);
    if (st != EP_OK) return st;
  }
  if (cudaGetLastError() != cudaSuccess) { g_hl_debug = 4; return EP_ERR_CUDA; }
  if (ss_out && !fuse_bn) return ep_bn_finalize(bn_partial, bn_rows, cout, m_out, eps, gamma, beta, ss_out, nullptr, stream);
  return EP_OK;
}

// kernels ep_spconv_hl_fused_fwd launches for these arguments (1-3): conv [+ split-K reduce] [+ BatchNorm finalisation]
int ep_spconv_hl_launches(int64_t m_out, int cin, int npad, int K, int have_counters, int want_ss) {
  static const bool knob_fuse = [] { const char* v = getenv("EPRECON_HL_FUSE"); return v && v[0] == '1'; }();
  if (!knob_fuse) have_counters = 0;
  const int nt = npad > 128 ? 128 : npad;
  const int splits = hl_splits(m_out, npad, K, (cin + 31) / 32);
  const int bn_rows = ep_div_up(m_out, 64);
  const int tiles = ep_div_up(m_out, HTM) * (npad / nt);
  (void)tiles;
  return 1 + ((splits > 1 && !have_counters) ? 1 : 0) + ((want_ss && !(have_counters && bn_rows <= 1024)) ? 1 : 0);
}

int ep_spconv_hl_fwd(const uint16_t* in_hl, int64_t m_in, int cin, const int32_t* nbr, int K, const uint16_t* w_hl, int npad,
                     int cout, const float* bias, float* out, int ld_out, int64_t m_out, float* bn_partial, void* workspace,
                     size_t workspace_bytes, int neg_row_mode, cudaStream_t stream) {
  return ep_spconv_hl_fused_fwd(in_hl, m_in, cin, nbr, K, w_hl, npad, cout, bias, out, ld_out, m_out, bn_partial, workspace,
                                workspace_bytes, neg_row_mode, nullptr, 0, nullptr, nullptr, 0.f, nullptr, stream);
}

int ep_hl_debug_code(void) { return g_hl_debug; }

// debug: device int64 [6 * 64 + 3] that receives the clock64 timeline of the middle CTA of every following launch (NULL = off):
// per stage t: producer {wait-empty begin, end, copies issued}, MMA thread {wait-full begin, end, MMAs issued + committed};
// then [384] main loop done, [385] epilogue done, [386] kernel entry
int ep_hl_set_timeline(void* dev_buffer) { g_hl_timeline = (long long*)dev_buffer; return EP_OK; }

// Debug entry: gathers the 128 rows `rows128` (device int32[128]; any value, out-of-range rows must read as zeros) of slab
// `slab` of in_hl [m_in][nslab][64] with 32 gather4 and returns the raw shared-memory image (16 KB) + status (1 ok, -1 timeout).
int ep_hl_probe_gather4(const uint16_t* in_hl, int64_t m_in, int nslab, const int32_t* rows128, int slab, void* out16k,
                        int32_t* status, cudaStream_t stream) {
  CUtensorMap tm_a;
  if (!make_map(&tm_a, in_hl, (uint64_t)nslab * 64, (uint64_t)m_in, 1)) return EP_ERR_CUDA;
  hl_probe_gather4_kernel<<<1, 32, A_STAGE + 1024, stream>>>(tm_a, rows128, slab, reinterpret_cast<uint4*>(out16k), status);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

}  // extern "C"

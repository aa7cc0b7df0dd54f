// Sparse 3-D convolution on the 5th-gen tensor cores: gather -> tcgen05.mma (kind::tf32, fp32 accumulate in
// TMEM) -> fused epilogue.  Same contract as ep_spconv_fwd (csrc/spconv.cu):
//     out[j, :] = bias + sum_k W[k]^T . in[nbr[j, k], :]
//
// One CTA owns 128 output rows x NT output channels (NT <= 128, multiple of 16).  For every kernel offset with a
// neighbour in the tile and every 16-channel slab of Cin, 8 producer warps gather the A slab (rows through the
// neighbour table, 16-byte loads) and copy the matching weight slab into shared memory in the UMMA canonical
// K-major / no-swizzle layout (8x16B core matrices); a dedicated issuer thread waits on the slab's full-mbarrier,
// issues the MMAs and commits them to the slab's empty-mbarrier (4-stage ring, producers run ahead).  Accumulators
// never leave TMEM until the epilogue (tcgen05.ld), which adds the bias, stores the rows and emits per-CTA column
// sums / sums of squares for batch-statistics BatchNorm.
//
// Precision: kind::tf32 reads the top 19 bits of each fp32 operand.  PREC == 1 rounds both operands to tf32
// (rel. error 2^-11 per product); PREC == 3 is the error-compensated split  a = a_hi + a_lo, b = b_hi + b_lo,
// a.b ~= a_hi.b_hi + a_lo.b_hi + a_hi.b_lo  (dropped term 2^-22), i.e. fp32-grade results at 3 MMAs per slab --
// this keeps the 1e-3 end-to-end parity budget of the north star through ~40 stacked layers.
#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int TMR = 128;      // output rows per CTA (UMMA M)
constexpr int KCT = 16;       // input channels per stage (2 MMAs of K = 8)
constexpr int NSTAGE = 3;
constexpr int TC_THREADS = 256;
constexpr int kMaxDynSmem = 200 * 1024;   // dynamic part; static barriers + reduction scratch (~5 KB) come on top (227 KB per CTA on sm_100)
constexpr int NBS = TMR + 1;  // row stride of the staged neighbour table (odd -> conflict-free when lanes run over offsets)
constexpr int APL = TMR + 2;  // float4 slots per K-chunk plane of the A tile: 128 rows + 32 B pad, so the 4 chunks of a row
                              // land in different banks (lanes run along a row's channels for coalesced gathers)

using namespace eptc;

// Warp-specialised pipeline: warps 0..7 are PRODUCERS (gather A rows through the neighbour table, split to tf32
// hi/lo, copy the weight slab; run up to NSTAGE slabs ahead of the tensor core), warp 8 lane 0 is the MMA ISSUER
// (waits full[s], issues tcgen05.mma, commits to empty[s]).  After the last slab warps 0..7 run the epilogue.
//
// Accumulators (PREC == 3): the hi.hi products round-robin over 3 TMEM accumulators and the two cross terms go to a
// 4th one; they are summed in fp32 registers in the epilogue.  The tensor core adds into its accumulator with
// truncation, so the error grows with the number of MMAs chained on one accumulator: keeping the (tiny) cross terms
// off the main chain and splitting it three ways brings the result back to fp32-FMA grade.
template <int PREC>
__global__ void __launch_bounds__(TC_THREADS + 32, 3)
spconv_tc_kernel(const float* __restrict__ in, int ld_in, int cin4 /*ceil4(cin)*/, const int* __restrict__ nbr, int K,
                 const float* __restrict__ w_hi, const float* __restrict__ w_lo, int nq /*4 * ceil(cin/16)*/,
                 int npad /*total padded cout*/, int nt /*columns of this launch's tile*/, int tmem_cols, int cout,
                 const float* __restrict__ bias, float* __restrict__ out, int ld_out, int m_out,
                 float* __restrict__ bn_partial, int bn_rows, int splits, float* __restrict__ partial) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int NACC = PREC == 3 ? 4 : 1;
  // split-K over the kernel offsets: blockIdx.z owns offsets [kb, ke) and (when splits > 1) writes raw partial sums
  const int kb = (int)(((long long)blockIdx.z * K) / splits), ke = (int)(((long long)(blockIdx.z + 1) * K) / splits);
  const int a_bytes = (KCT / 4) * APL * 16;
  const int b_bytes = (KCT / 4) * nt * 16;
  const int stage_bytes = (PREC == 3 ? 2 : 1) * (a_bytes + b_bytes);
  int* s_nbr = reinterpret_cast<int*>(smem_raw + (size_t)NSTAGE * stage_bytes);  // [K][128] neighbour rows of the tile
  __shared__ uint64_t full_bar[NSTAGE], empty_bar[NSTAGE];
  __shared__ uint64_t all_done;
  __shared__ uint32_t tmem_base_s;
  __shared__ int s_klist[32];
  __shared__ int s_nk;
  __shared__ float s_red[8][128];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * TMR;
  const int col0 = blockIdx.y * nt;
  const bool producer = warp < 8;

  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full_bar[s], 8); mbar_init(&empty_bar[s], 1); }
    mbar_init(&all_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid == 0) s_nk = 0;   // reused as the bit mask of kernel offsets that have a neighbour in this tile
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // stage the tile's neighbour table transposed to [k][row].  The tile's slab nbr[row0*K .. (row0+128)*K) is contiguous:
  // every thread issues PU independent coalesced loads before the first store (the first version walked one row per warp
  // with a load -> store dependency per iteration and showed up as 19 % of the stall samples of a one-wave launch,
  // profiles/r01_spconv_tc_v5_ncu_summary.txt) and collects the mask of offsets that have a neighbour in the tile.
  {
    constexpr int NT = TC_THREADS + 32, PU = 6;
    const int total = TMR * K;
    const long long base = (long long)row0 * K, lim = (long long)m_out * K;
    unsigned mine = 0;
    for (int e0 = 0; e0 < total; e0 += NT * PU) {
      int v[PU];
#pragma unroll
      for (int u = 0; u < PU; ++u) {
        const int e = e0 + u * NT + tid;
        v[u] = -1;
        if (e < total && base + e < lim) v[u] = nbr ? __ldg(nbr + base + e) : row0 + e;   // no table: K == 1, identity
      }
#pragma unroll
      for (int u = 0; u < PU; ++u) {
        const int e = e0 + u * NT + tid;
        if (e < total) {
          const int r = e / K, kq = e - r * K;
          s_nbr[kq * NBS + r] = v[u];
          if (v[u] >= 0 && kq >= kb && kq < ke) mine |= 1u << kq;
        }
      }
    }
    mine = __reduce_or_sync(0xffffffffu, mine);
    if (lane == 0 && mine) atomicOr(&s_nk, (int)mine);
  }
  __syncthreads();
  const unsigned kmask = (unsigned)s_nk;   // every thread walks the set bits in ascending order: the active offsets
  const uint32_t tmem_d = tmem_base_s;
  const int nk = __popc(kmask);
  const int nchunk = (cin4 + KCT - 1) / KCT;
  const int T = nk * nchunk;  // pipeline slabs of this CTA

  if (producer) {
    // ---------------------------------------------------------------- producers
    // lanes run along the channels of a row: 4 consecutive lanes read one row's 64 contiguous bytes (1 L1 wavefront
    // per row instead of one per 16-byte item); this thread owns items (rowA, jqA) and (rowA + 64, jqA)
    const int rowA = tid >> 2, jqA = tid & 3;
    const int nb_items = (KCT / 4) * nt;                         // float4 items of the weight slab
    float4 a_reg[2], bh_reg[2], bl_reg[2];
    // loop-invariant parts of this thread's two weight-slab items: K-chunk of the item and its element offset in a slab
    unsigned b_rel[2];
    bool b_on[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int e = tid + i * TC_THREADS;
      b_on[i] = e < nb_items;
      const int jq = b_on[i] ? e / nt : 0;
      b_rel[i] = b_on[i] ? (unsigned)(jq * npad + (e - jq * nt)) * 4u : 0u;
      bh_reg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      bl_reg[i] = bh_reg[i];
    }
    const int a_item = jqA * APL + rowA;          // float4 slot of this thread's first A item inside a plane set
    // load cursor: `rem` = offsets still to load (lowest set bit = current offset k), `c` = channel slab within k.
    // Everything that only depends on k (the two gathered rows' addresses, the offset's weight block) is computed once
    // per offset, not once per slab; the weights are addressed with 32-bit element offsets (K*nq*npad*4 < 2^31).
    unsigned rem = kmask;
    int c = 0;
    unsigned w_k = 0;
    const float* arow[2] = {nullptr, nullptr};
    auto set_k = [&]() {
      const int k = __ffs(rem) - 1;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int src = s_nbr[k * NBS + rowA + 64 * i];
        arow[i] = src >= 0 ? in + (size_t)src * ld_in + jqA * 4 : nullptr;
      }
      w_k = (unsigned)(k * nq * npad + col0) * 4u;
    };
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    auto issue_loads = [&]() {
      const bool col_ok = c * KCT + jqA * 4 < cin4;
#pragma unroll
      for (int i = 0; i < 2; ++i)
        a_reg[i] = (arow[i] && col_ok) ? __ldg(reinterpret_cast<const float4*>(arow[i] + c * KCT)) : zero4;
      // the weight planes are zero-padded to whole slabs (nq is a multiple of KCT/4): no bounds check, and a 32-bit
      // unsigned element offset from the (warp-uniform) plane base
      const unsigned w_s = w_k + (unsigned)(c * (KCT / 4) * npad) * 4u;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        if (b_on[i]) {
          bh_reg[i] = __ldg(reinterpret_cast<const float4*>(w_hi + (w_s + b_rel[i])));
          if (PREC == 3) bl_reg[i] = __ldg(reinterpret_cast<const float4*>(w_lo + (w_s + b_rel[i])));
        }
      }
      if (++c == nchunk) {
        c = 0;
        rem &= rem - 1;
        if (rem) set_k();
      }
    };
    if (T > 0) { set_k(); issue_loads(); }
    int stage = 0, round = 0;                      // slab t lives in stage t % NSTAGE, round = t / NSTAGE
    for (int t = 0; t < T; ++t) {
      if (round > 0) mbar_wait(&empty_bar[stage], (round - 1) & 1);
      uint8_t* sbase = smem_raw + (size_t)stage * stage_bytes;
      float4* a_hi = reinterpret_cast<float4*>(sbase) + a_item;
      float4* a_lo = reinterpret_cast<float4*>(sbase + a_bytes) + a_item;
      float4* b_hi = reinterpret_cast<float4*>(sbase + (PREC == 3 ? 2 : 1) * a_bytes) + tid;
      float4* b_lo = reinterpret_cast<float4*>(sbase + (PREC == 3 ? 2 : 1) * a_bytes + b_bytes) + tid;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float4 v = a_reg[i];
        const float4 h = make_float4(tf32_rn_finite(v.x), tf32_rn_finite(v.y), tf32_rn_finite(v.z), tf32_rn_finite(v.w));
        a_hi[64 * i] = h;
        if (PREC == 3) a_lo[64 * i] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        if (b_on[i]) {
          b_hi[i * TC_THREADS] = bh_reg[i];
          if (PREC == 3) b_lo[i * TC_THREADS] = bl_reg[i];
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy STS -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[stage]);
      // The next slab's loads are issued AFTER the fence (it drains the warp's outstanding loads, so loads issued before
      // it cannot overlap anything) straight into the registers just stored: no second register set, no copies.
      if (t + 1 < T) issue_loads();
      if (++stage == NSTAGE) { stage = 0; ++round; }
    }
  } else if (lane == 0) {
    // ---------------------------------------------------------------- MMA issuer (one thread)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(nt >> 3) << 17) | ((uint32_t)(TMR >> 4) << 24);
    uint32_t used = 0;  // bit a: accumulator a already written
    for (int t = 0; t < T; ++t) {
      const int stage = t % NSTAGE;
      mbar_wait(&full_bar[stage], (t / NSTAGE) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sbase = smem_u32(smem_raw + (size_t)stage * stage_bytes);
      const uint32_t sa_hi = sbase, sa_lo = sbase + a_bytes;
      const uint32_t sb_hi = sbase + (PREC == 3 ? 2 : 1) * a_bytes, sb_lo = sb_hi + b_bytes;
#pragma unroll
      for (int kk = 0; kk < KCT / 8; ++kk) {
        const uint32_t a_off = kk * 2 * APL * 16, b_off = kk * 2 * nt * 16;
        const uint64_t dah = umma_desc(sa_hi + a_off, APL * 16, 128), dbh = umma_desc(sb_hi + b_off, nt * 16, 128);
        const int am = (PREC == 3) ? ((t * (KCT / 8) + kk) % 3) : 0;
        umma_tf32(tmem_d + am * nt, dah, dbh, idesc, (used >> am) & 1u);
        used |= 1u << am;
        if (PREC == 3) {
          const uint64_t dal = umma_desc(sa_lo + a_off, APL * 16, 128), dbl = umma_desc(sb_lo + b_off, nt * 16, 128);
          umma_tf32(tmem_d + 3 * nt, dal, dbh, idesc, (used >> 3) & 1u);
          used |= 1u << 3;
          umma_tf32(tmem_d + 3 * nt, dah, dbl, idesc, 1u);
        }
      }
      umma_commit(&empty_bar[stage]);  // arrives once the MMAs above have finished reading this stage
    }
    umma_commit(&all_done);
    // tell the epilogue which accumulators hold data
    s_klist[31] = (int)used;
  }
  __syncthreads();
  if (producer) {
    // ---------------------------------------------------------------- epilogue (warps 0..7)
    if (T > 0) mbar_wait(&all_done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t used = (uint32_t)s_klist[31];
    const int q = warp & 3, half = warp >> 2;
    const int row = row0 + q * 32 + lane;
    const int ncol_half = (nt / 16 + 1) / 2 * 16;
    const int cbeg = half == 0 ? 0 : ncol_half, cend = half == 0 ? min(ncol_half, nt) : nt;
    for (int cb = cbeg; cb < cend; cb += 16) {
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = 0.f;
#pragma unroll
      for (int a = 0; a < NACC; ++a) {
        if (T > 0 && ((used >> a) & 1u)) {
          float t16[16];
          tmem_ld16(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * nt + cb), t16);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += t16[i];
        }
      }
      if (splits > 1) {  // raw partial sums; bias / statistics are applied by splitk_reduce_kernel
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
      // bn_partial is sized for 64-row tiles: this CTA fills entry 2*bx and zeroes 2*bx+1
      const int r0 = 2 * blockIdx.x;
      bn_partial[((size_t)r0 * 2 + 0) * cout + col] = s;
      bn_partial[((size_t)r0 * 2 + 1) * cout + col] = sq;
      if (r0 + 1 < bn_rows) {
        bn_partial[((size_t)(r0 + 1) * 2 + 0) * cout + col] = 0.f;
        bn_partial[((size_t)(r0 + 1) * 2 + 1) * cout + col] = 0.f;
      }
    }
  }
  if (warp == 8) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols));
  }
}

// out[row, col] = bias + sum_z partial[z][row][col] (fixed z order), plus the per-64-row-tile BN statistics.
// One CTA per (64-row tile, 128-column chunk); 256 threads = 8 row lanes x 32 float4 column lanes.  For every split z a
// thread issues the loads of its 8 rows back to back and only then accumulates, so 8 (16 with the unroll) independent
// 16-byte loads are in flight per thread and a launch needs `splits` dependent round trips.  (v1 walked 16 rows x splits
// scalar loads per thread: 34 us; v2 nested the split loop inside the row loop -- 8 x splits round trips, 50-60 us on
// the 2-11 CTA grids of the coarse levels, longer than the convolution it finished,
// profiles/r01_launches_v7_one_fragment_summary.txt.)  Same summation order as before: bit-identical results.
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ partial, int splits, int m_out, int npad, int cout,
                     const float* __restrict__ bias, float* __restrict__ out, int ld_out, float* __restrict__ bn_partial) {
  __shared__ float s_s[8][128], s_q[8][128];
  const int cq = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int row0 = blockIdx.x * 64;
  const int c0 = blockIdx.y * 128;
  const size_t plane = (size_t)m_out * npad;
  const int col = c0 + cq * 4;
  float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
  if (col < npad && col < cout) {
    float4 v[8];
    const float* p[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      const int row = row0 + ry + 8 * i;
      p[i] = row < m_out ? partial + (size_t)row * npad + col : nullptr;
    }
#pragma unroll 2
    for (int z = 0; z < splits; ++z) {
      float4 t[8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        t[i] = p[i] ? __ldg(reinterpret_cast<const float4*>(p[i] + (size_t)z * plane)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 8; ++i) { v[i].x += t[i].x; v[i].y += t[i].y; v[i].z += t[i].z; v[i].w += t[i].w; }
    }
    float bv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) bv[j] = (bias && col + j < cout) ? bias[col + j] : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = row0 + ry + 8 * i;
      if (row < m_out) {
        const float r[4] = {v[i].x + bv[0], v[i].y + bv[1], v[i].z + bv[2], v[i].w + bv[3]};
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (col + j < cout) {
            out[(size_t)row * ld_out + col + j] = r[j];
            s[j] += r[j];
            q[j] = fmaf(r[j], r[j], q[j]);
          }
      }
    }
  }
  if (bn_partial) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { s_s[ry][cq * 4 + j] = s[j]; s_q[ry][cq * 4 + j] = q[j]; }
    __syncthreads();
    if (threadIdx.x < 128 && c0 + (int)threadIdx.x < cout) {
      const int t = threadIdx.x;
      float a = 0.f, b = 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) { a += s_s[r][t]; b += s_q[r][t]; }   // fixed order: deterministic
      bn_partial[((size_t)blockIdx.x * 2 + 0) * cout + c0 + t] = a;
      bn_partial[((size_t)blockIdx.x * 2 + 1) * cout + c0 + t] = b;
    }
  }
}

inline int pow2_cols(int n) {
  int c = 32;
  while (c < n) c <<= 1;
  return c;
}

}  // namespace

// shared with csrc/spconv_hl.cu: fixed-order reduction of split-K partial sums (+ bias, + BatchNorm partial statistics)
int ep_internal_splitk_reduce(const float* partial, int splits, int m_out, int npad, int cout, const float* bias, float* out,
                              int ld_out, float* bn_partial, cudaStream_t stream) {
  splitk_reduce_kernel<<<dim3(ep_div_up(m_out, 64), ep_div_up(cout, 128)), 256, 0, stream>>>(partial, splits, m_out, npad, cout, bias,
                                                                                            out, ld_out, bn_partial);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

extern "C" {

// Weights must be pre-arranged as float[K][nq][npad][4] (w[k][q][n][i] = W[k][4q+i][n], zero padded; nq = ceil(cin/4)
// rounded up to a multiple of 4 (whole 16-channel slabs); npad = cout rounded up to a multiple of 16, and of 128 when
// larger than 128).  w_lo is only read when prec == 3.
// bn_partial: NULL or float[ep_spconv_num_row_tiles(m_out), 2, cout].
// split-K factor for small problems: spread the K kernel offsets over up to ~one wave of CTAs
// EPRECON_TC_SPLIT_CTAS=<n> (experiment knob): aim for ~n CTAs per launch instead of one per SM, rounding the factor up
static int tc_split_target() {
  static const int target = [] {
    const char* s = getenv("EPRECON_TC_SPLIT_CTAS");
    const int v = s ? atoi(s) : 0;
    return v > 0 ? v : 0;
  }();
  return target;
}

static int tc_splits(int64_t m_out, int npad, int K) {
  const int nt = npad > 128 ? 128 : npad;
  const long long ctas = (long long)ep_div_up(m_out, TMR) * (npad / nt);
  const int target = tc_split_target();
  if (target > 0) {
    if (K < 2 || ctas >= target) return 1;
    long long s = (target + ctas - 1) / ctas;
    if (s > K) s = K;
    return s < 2 ? 1 : (int)s;
  }
  if (K < 2 || ctas * 2 > EP_NUM_SMS) return 1;
  long long s = EP_NUM_SMS / ctas;
  if (s > K) s = K;
  return s < 2 ? 1 : (int)s;
}

size_t ep_spconv_tc_workspace_bytes(int64_t m_out, int npad, int K) {
  const int s = tc_splits(m_out, npad, K);
  return s > 1 ? (size_t)s * (size_t)m_out * (size_t)npad * sizeof(float) : 0;
}

int ep_spconv_tc_fwd(const float* in, int ld_in, int cin, const int32_t* nbr, int K, const float* w_hi,
                     const float* w_lo, int npad, int cout, const float* bias, float* out, int ld_out, int64_t m_out,
                     float* bn_partial, int prec, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (m_out <= 0 || cin < 1 || cout < 1 || K < 1 || ld_in % 4 != 0 || npad % 16 != 0 || npad < cout) return EP_ERR_ARG;
  if (!nbr && K != 1) return EP_ERR_ARG;
  if (prec != 1 && prec != 3) return EP_ERR_ARG;
  if (prec == 3 && !w_lo) return EP_ERR_ARG;
  const int cin4 = (cin + 3) / 4 * 4;
  if (cin4 > ld_in) return EP_ERR_ARG;
  int nt = npad;
  if (npad > 128) {
    if (npad % 128 != 0) return EP_ERR_ARG;
    nt = 128;
  }
  if (K > 27) return EP_ERR_UNSUPPORTED;
  const int tmem_cols = pow2_cols((prec == 3 ? 4 : 1) * nt);
  if (tmem_cols > 512) return EP_ERR_UNSUPPORTED;
  const size_t stage = (size_t)(prec == 3 ? 2 : 1) * ((KCT / 4) * APL * 16 + (KCT / 4) * nt * 16);
  const size_t smem = stage * NSTAGE + (size_t)K * NBS * sizeof(int);
  cudaError_t e;
  // always the same (maximum) value: concurrent host threads (one per CUDA stream) may not race on a per-launch size
  if (prec == 3) e = cudaFuncSetAttribute(spconv_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
  else e = cudaFuncSetAttribute(spconv_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
  if (e != cudaSuccess) return EP_ERR_CUDA;
  const int splits = tc_splits(m_out, npad, K);
  if (splits > 1 && workspace_bytes < ep_spconv_tc_workspace_bytes(m_out, npad, K)) return EP_ERR_WORKSPACE;
  float* partial = splits > 1 ? (float*)workspace : nullptr;
  dim3 grid(ep_div_up(m_out, TMR), npad / nt, splits);
  const int bn_rows = ep_div_up(m_out, 64);
  if (prec == 3)
    spconv_tc_kernel<3><<<grid, TC_THREADS + 32, smem, stream>>>(in, ld_in, cin4, nbr, K, w_hi, w_lo, ((cin + 15) / 16) * 4, npad, nt,
                                                           tmem_cols, cout, bias, out, ld_out, (int)m_out, bn_partial, bn_rows, splits, partial);
  else
    spconv_tc_kernel<1><<<grid, TC_THREADS + 32, smem, stream>>>(in, ld_in, cin4, nbr, K, w_hi, w_lo, ((cin + 15) / 16) * 4, npad, nt,
                                                           tmem_cols, cout, bias, out, ld_out, (int)m_out, bn_partial, bn_rows, splits, partial);
  if (splits > 1)
    splitk_reduce_kernel<<<dim3(ep_div_up(m_out, 64), ep_div_up(cout, 128)), 256, 0, stream>>>(partial, splits, (int)m_out, npad, cout,
                                                                                              bias, out, ld_out, bn_partial);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

}  // extern "C"

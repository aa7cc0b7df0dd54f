// Internal helpers shared by the eprecon_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define EP_OK 0
#define EP_ERR_ARG (-1)
#define EP_ERR_WORKSPACE (-2)
#define EP_ERR_CUDA (-3)
#define EP_ERR_UNSUPPORTED (-4)

#define EP_NUM_SMS 148

#define EP_CHECK_LAUNCH()                                   \
  do {                                                      \
    cudaError_t e__ = cudaGetLastError();                   \
    if (e__ != cudaSuccess) return EP_ERR_CUDA;             \
  } while (0)

static inline int ep_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// grid size for a grid-stride kernel: whole waves of resident CTAs, capped by the work
static inline int ep_grid(long long work_items, int threads, int ctas_per_sm) {
  long long need = (work_items + threads - 1) / threads;
  long long cap = (long long)EP_NUM_SMS * ctas_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

__device__ __forceinline__ float4 ep_ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ int ep_warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += t;
  }
  return v;
}

// Block-wide exclusive scan of one int per thread (blockDim.x <= 1024, multiple of 32).
// Returns the exclusive prefix; *total receives the block sum (valid for all threads).
__device__ __forceinline__ int ep_block_excl_scan(int v, int* smem_warp /*>=33 ints*/, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  int incl = ep_warp_incl_scan(v, lane);
  if (lane == 31) smem_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = lane < nwarp ? smem_warp[lane] : 0;
    int wi = ep_warp_incl_scan(w, lane);
    smem_warp[lane] = wi - w;
    if (lane == 31) smem_warp[32] = wi;
  }
  __syncthreads();
  int out = incl - v + smem_warp[warp];
  *total = smem_warp[32];
  __syncthreads();
  return out;
}

// torchsparse v2.0.0 coordinate hash (hash_cuda.cu): FNV-1a over (x,y,z,b) as uint32, folded to 60 bits.
__host__ __device__ __forceinline__ uint64_t ep_sphash(int x, int y, int z, int b) {
  uint64_t h = 14695981039346656037ULL;
  h ^= (uint32_t)x; h *= 1099511628211ULL;
  h ^= (uint32_t)y; h *= 1099511628211ULL;
  h ^= (uint32_t)z; h *= 1099511628211ULL;
  h ^= (uint32_t)b; h *= 1099511628211ULL;
  return (h >> 60) ^ (h & 0x0FFFFFFFFFFFFFFFULL);
}

// 64-bit mix for open-addressing slots (splitmix64 finaliser)
__host__ __device__ __forceinline__ uint64_t ep_mix64(uint64_t k) {
  k ^= k >> 30; k *= 0xbf58476d1ce4e5b9ULL;
  k ^= k >> 27; k *= 0x94d049bb133111ebULL;
  k ^= k >> 31;
  return k;
}

#define EP_HASH_EMPTY 0xFFFFFFFFFFFFFFFFULL

// 3-D Morton (Z-order) key of three 16-bit coordinates (offset by 32768 to make them unsigned), batch in the top 16 bits.
// Sorting voxels by this key makes every run of 128 consecutive rows a compact blob, so the 27-neighbour gathers of a
// conv tile stay inside a small, L1-resident window of the feature matrix.
__host__ __device__ __forceinline__ uint64_t ep_spread3(uint32_t v) {  // 16 bits -> every third bit
  uint64_t x = v & 0xFFFFu;
  x = (x | (x << 32)) & 0x00FF00000000FFFFULL;
  x = (x | (x << 16)) & 0x00FF0000FF0000FFULL;
  x = (x | (x << 8)) & 0xF00F00F00F00F00FULL;
  x = (x | (x << 4)) & 0x30C30C30C30C30C3ULL;
  x = (x | (x << 2)) & 0x9249249249249249ULL;
  return x;
}
__host__ __device__ __forceinline__ uint32_t ep_compact3(uint64_t x) {
  x &= 0x9249249249249249ULL;
  x = (x | (x >> 2)) & 0x30C30C30C30C30C3ULL;
  x = (x | (x >> 4)) & 0xF00F00F00F00F00FULL;
  x = (x | (x >> 8)) & 0x00FF0000FF0000FFULL;
  x = (x | (x >> 16)) & 0x00FF00000000FFFFULL;
  x = (x | (x >> 32)) & 0xFFFFULL;
  return (uint32_t)x;
}
__host__ __device__ __forceinline__ uint64_t ep_morton_key(int x, int y, int z, int b) {
  return ((uint64_t)(uint16_t)b << 48) | (ep_spread3((uint32_t)(x + 32768)) << 2) | (ep_spread3((uint32_t)(y + 32768)) << 1) |
         ep_spread3((uint32_t)(z + 32768));
}
// Compact form for coordinates known to lie in [-2^(cb-1), 2^(cb-1)) and 0 <= b < 2^bb: cb bits per axis (offset 2^(cb-1)),
// batch above -> 3 * cb + bb key bits, so the radix sort needs ceil((3 cb + bb) / 8) passes instead of 8.  The ORDER equals
// ep_morton_key's on such coordinates: there the per-axis bits are [sign, 15 - cb copies of !sign, low cb - 1 bits], and the
// repeated levels never decide a comparison that the sign level left open.
__host__ __device__ __forceinline__ bool ep_morton_compact_ok(int x, int y, int z, int b, int cb, int bb) {
  const int h = 1 << (cb - 1);
  return x >= -h && x < h && y >= -h && y < h && z >= -h && z < h && b >= 0 && b < (1 << bb);
}
__host__ __device__ __forceinline__ uint64_t ep_morton_key_compact(int x, int y, int z, int b, int cb) {
  const int h = 1 << (cb - 1);
  return ((uint64_t)(uint32_t)b << (3 * cb)) | (ep_spread3((uint32_t)(x + h)) << 2) | (ep_spread3((uint32_t)(y + h)) << 1) |
         ep_spread3((uint32_t)(z + h));
}
__host__ __device__ __forceinline__ void ep_morton_unkey_compact(uint64_t k, int cb, int& x, int& y, int& z, int& b) {
  const int h = 1 << (cb - 1);
  b = (int)(k >> (3 * cb));
  const uint64_t m = k & ((1ULL << (3 * cb)) - 1);
  x = (int)ep_compact3(m >> 2) - h;
  y = (int)ep_compact3(m >> 1) - h;
  z = (int)ep_compact3(m) - h;
}
__host__ __device__ __forceinline__ void ep_morton_unkey(uint64_t k, int& x, int& y, int& z, int& b) {
  b = (int)(k >> 48);
  const uint64_t m = k & 0x0000FFFFFFFFFFFFULL;
  x = (int)ep_compact3(m >> 2) - 32768;
  y = (int)ep_compact3(m >> 1) - 32768;
  z = (int)ep_compact3(m) - 32768;
}

// Stream-compaction / scan / sort plumbing shared by the sparse kernels.
//
// Replaces the reference's torch.nonzero / boolean indexing / torch.unique(sort) call sites
// (models/neucon_network.py:304,312,492-501; ops/torchsparse_utils.py:20; utils.py:172,179).
// The 64-bit key sort is cub::DeviceRadixSort (CUDA-toolkit header library, like cuBLAS for GEMM);
// the scans and compactions are hand-written.
#include "common.cuh"

#include <cub/device/device_radix_sort.cuh>

namespace {

constexpr int kScanThreads = 256;
constexpr int kItems = 4;                       // items per thread
constexpr int kTile = kScanThreads * kItems;    // 1024 items per CTA

// flags -> per-CTA counts
__global__ void __launch_bounds__(kScanThreads)
flag_count_kernel(const uint8_t* __restrict__ flags, int n, int* __restrict__ block_count) {
  __shared__ int s_scan[33];
  const int base = blockIdx.x * kTile + threadIdx.x * kItems;
  int c = 0;
#pragma unroll
  for (int k = 0; k < kItems; ++k) c += (base + k < n) ? (flags[base + k] != 0) : 0;
  int total;
  ep_block_excl_scan(c, s_scan, &total);
  if (threadIdx.x == 0) block_count[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) scan_blocks_kernel(const int* __restrict__ in, int* __restrict__ out, int nblk,
                                                           int* __restrict__ total_out) {
  __shared__ int s_scan[33];
  int carry = 0;
  for (int base = 0; base < nblk; base += blockDim.x) {
    int i = base + threadIdx.x;
    int v = i < nblk ? in[i] : 0;
    int tot;
    int ex = ep_block_excl_scan(v, s_scan, &tot);
    if (i < nblk) out[i] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0) *total_out = carry;
}

// stable compaction: index list of the set flags (ascending), optional inverse map pos[i] (or -1)
__global__ void __launch_bounds__(kScanThreads)
flag_scatter_kernel(const uint8_t* __restrict__ flags, int n, const int* __restrict__ block_offset,
                    int* __restrict__ out_index, int* __restrict__ out_pos) {
  __shared__ int s_scan[33];
  const int base = blockIdx.x * kTile + threadIdx.x * kItems;
  int f[kItems], c = 0;
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    f[k] = (base + k < n) ? (flags[base + k] != 0) : 0;
    c += f[k];
  }
  int total;
  int ex = ep_block_excl_scan(c, s_scan, &total) + block_offset[blockIdx.x];
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    if (base + k < n) {
      if (f[k]) {
        if (out_index) out_index[ex] = base + k;
        if (out_pos) out_pos[base + k] = ex;
        ++ex;
      } else if (out_pos) {
        out_pos[base + k] = -1;
      }
    }
  }
}

__global__ void iota_kernel(int* __restrict__ p, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = i;
}

// head flags + per-CTA head counts in one pass over the sorted keys (head_flags_kernel + flag_count_kernel)
__global__ void __launch_bounds__(kScanThreads)
head_count_kernel(const uint64_t* __restrict__ keys, int n, uint64_t sentinel, uint8_t* __restrict__ head,
                  int* __restrict__ block_count) {
  __shared__ int s_scan[33];
  const int base = blockIdx.x * kTile + threadIdx.x * kItems;
  int c = 0;
  uint64_t prev = (base > 0 && base - 1 < n) ? keys[base - 1] : 0ull;
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    const int i = base + k;
    if (i < n) {
      const uint64_t key = keys[i];
      const bool h = (key != sentinel) && (i == 0 || prev != key);
      head[i] = h;
      c += h;
      prev = key;
    }
  }
  int total;
  ep_block_excl_scan(c, s_scan, &total);
  if (threadIdx.x == 0) block_count[blockIdx.x] = total;
}

// After compaction of head flags: seg_start[s] = sorted position of segment s's head (= out_index),
// here we additionally emit seg id per original item: seg_of[perm[j]] = (#heads at or before j) - 1.
__global__ void __launch_bounds__(kScanThreads)
segment_ids_kernel(const uint8_t* __restrict__ head, const uint64_t* __restrict__ keys, uint64_t sentinel,
                   const int* __restrict__ perm, int n, const int* __restrict__ block_offset,
                   int* __restrict__ seg_start, int* __restrict__ seg_of_item, int* __restrict__ seg_end) {
  __shared__ int s_scan[33];
  const int base = blockIdx.x * kTile + threadIdx.x * kItems;
  int f[kItems], c = 0;
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    f[k] = (base + k < n) ? (head[base + k] != 0) : 0;
    c += f[k];
  }
  int total;
  int ex = ep_block_excl_scan(c, s_scan, &total) + block_offset[blockIdx.x];
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    if (base + k < n) {
      if (f[k] && seg_start) seg_start[ex] = base + k;           // stable compaction of the heads (flag_scatter_kernel's job)
      ex += f[k];
      const uint64_t key = keys[base + k];
      const bool dropped = key == sentinel;
      if (seg_of_item) seg_of_item[perm[base + k]] = dropped ? -1 : ex - 1;
      if (!dropped && (base + k == n - 1 || keys[base + k + 1] != key)) seg_end[ex - 1] = base + k + 1;
    }
  }
}

}  // namespace

extern "C" {

size_t ep_compact_workspace_bytes(int64_t n) {
  size_t nblk = (size_t)ep_div_up(n > 0 ? n : 1, kTile);
  return 2 * nblk * sizeof(int) + 256;
}

// flags uint8[n] -> out_index int32[total] (ascending positions of set flags), out_pos int32[n] (rank or -1),
// *total_dev = number of set flags.  Either output may be NULL.
int ep_compact_flags(const uint8_t* flags, int64_t n, int32_t* out_index, int32_t* out_pos, int32_t* total_dev,
                     void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (n <= 0 || n > 0x7fffffffLL) return EP_ERR_ARG;
  if (workspace_bytes < ep_compact_workspace_bytes(n)) return EP_ERR_WORKSPACE;
  const int nblk = ep_div_up(n, kTile);
  int* bc = (int*)workspace;
  int* bo = bc + nblk;
  flag_count_kernel<<<nblk, kScanThreads, 0, stream>>>(flags, (int)n, bc);
  scan_blocks_kernel<<<1, 1024, 0, stream>>>(bc, bo, nblk, total_dev);
  flag_scatter_kernel<<<nblk, kScanThreads, 0, stream>>>(flags, (int)n, bo, out_index, out_pos);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

size_t ep_sort_segments_workspace_bytes(int64_t n) {
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                  (const int*)nullptr, (int*)nullptr, (int)(n > 0 ? n : 1), 0, 64, 0);
  size_t nn = (size_t)(n > 0 ? n : 1);
  return cub_bytes + nn * sizeof(int) /*iota*/ + nn /*head*/ + ep_compact_workspace_bytes(n) + 1024;
}

// Group n items by 64-bit key.  Stable radix sort of (key, item id); items whose key == sentinel are dropped.
// Outputs: keys_sorted uint64[n], perm int32[n] (item ids in sorted order), seg_start / seg_end int32[<=n]
// (segment s covers sorted positions [seg_start[s], seg_end[s])), seg_of_item int32[n] (segment id per
// ORIGINAL item, -1 if dropped; may be NULL), *n_segments_dev.
// Segments come out in ascending key order (== torch.unique ordering of the reference).
int ep_sort_segments(const uint64_t* keys, int64_t n, int key_bits, uint64_t sentinel, uint64_t* keys_sorted,
                     int32_t* perm, int32_t* seg_start, int32_t* seg_end, int32_t* seg_of_item,
                     int32_t* n_segments_dev,
                     void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (n <= 0 || n > 0x7fffffffLL || key_bits < 1 || key_bits > 64) return EP_ERR_ARG;
  if (workspace_bytes < ep_sort_segments_workspace_bytes(n)) return EP_ERR_WORKSPACE;
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, keys, keys_sorted, (const int*)nullptr, perm, (int)n, 0,
                                  key_bits, stream);
  char* w = (char*)workspace;
  void* cub_ws = w;
  w += (cub_bytes + 255) / 256 * 256;
  int* iota = (int*)w;
  w += ((size_t)n * sizeof(int) + 255) / 256 * 256;
  uint8_t* head = (uint8_t*)w;
  w += ((size_t)n + 255) / 256 * 256;
  int* bc = (int*)w;
  const int nblk = ep_div_up(n, kTile);
  int* bo = bc + nblk;
  iota_kernel<<<ep_div_up(n, 256), 256, 0, stream>>>(iota, (int)n);
  if (cub::DeviceRadixSort::SortPairs(cub_ws, cub_bytes, keys, keys_sorted, iota, perm, (int)n, 0, key_bits,
                                      stream) != cudaSuccess)
    return EP_ERR_CUDA;
  head_count_kernel<<<nblk, kScanThreads, 0, stream>>>(keys_sorted, (int)n, sentinel, head, bc);
  scan_blocks_kernel<<<1, 1024, 0, stream>>>(bc, bo, nblk, n_segments_dev);
  segment_ids_kernel<<<nblk, kScanThreads, 0, stream>>>(head, keys_sorted, sentinel, perm, (int)n, bo, seg_start, seg_of_item,
                                                        seg_end);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

}  // extern "C"

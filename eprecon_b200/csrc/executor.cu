// Native executor: whole-module launch programs behind ONE C-ABI call each.
//
// The per-kernel entry points of this library are thin (one launch each), which leaves the ORCHESTRATION of a
// module -- ~160 launches for one SPVCNN.forward (models/modules.py:75-175), ~60 for the two ConvGRUs of a level
// (models/gru_fusion.py:339-349, models/modules.py:178-222), 6 per Linear4xTrans head (modules.py:273-311), ~45 for
// the occupancy-initialisation head (models/occupancy_initialization.py:131-176) -- to the caller.  Driven from
// Python that orchestration costs more host time than the kernels take on a B200 and, being serialised by the GIL,
// cannot be overlapped across streams.  The ep_exec_* functions below run those programs natively: they call the
// very same launchers in the very same order with the same arguments as eprecon_b200/modules.py / sparse.py
// (results are bit-identical; tests/test_executor_gpu.py), carve every temporary out of a caller-provided scratch
// arena (stream-ordered bump allocation: no cudaMalloc, no frees) and read the data-dependent sizes (voxel counts)
// back through a pinned 4-byte copy + cudaStreamSynchronize -- the only places where they block.
//
// Unlike the per-kernel entry points these functions DO synchronise the stream (at those read-backs).  Parameters
// arrive as flat int64 descriptors built once per module by eprecon_b200/executor.py (pointers + dimensions).
#include "common.cuh"

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "eprecon_b200.h"

namespace {

std::atomic<size_t> g_launches{0};

// ---- per-launch profiling of the sparse-conv family (bench.py's roofline): when enabled on the calling host thread,
// every spconv() brackets its launches with a CUDA-event pair recorded on the executor's stream (no host gap between
// the record and the launch it brackets: both are issued from this C++ code) and counts the valid neighbour pairs.
struct ProfRec { cudaEvent_t e0, e1; int K, cin, cout, m_in, m_out, impl; int* pairs_dev; };
struct Prof { std::vector<ProfRec> recs; int* pairs_pool = nullptr; int pool_used = 0; };
constexpr int kProfPool = 4096;
thread_local Prof* tl_prof = nullptr;

__global__ void count_valid_kernel(const int* __restrict__ nbr, long long n, int* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int ok = (i < n && nbr[i] >= 0) ? 1 : 0;
  const unsigned b = __ballot_sync(0xffffffffu, ok);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(out, __popc(b));
}

__global__ void csr_expand_kernel(const uint64_t* __restrict__ keys_sorted, const int* __restrict__ seg_start,
                                  const int* __restrict__ seg_end, int s, int* __restrict__ s0, int* __restrict__ s1) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= s) return;
  const int a = seg_start[i];
  const int id = (int)keys_sorted[a];
  s0[id] = a;
  s1[id] = seg_end[i];
}

__global__ void translate_index_kernel(const int* __restrict__ idx, long long n, const int* __restrict__ rank,
                                       int* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int v = idx[i];
  out[i] = v >= 0 ? rank[v] : -1;
}

constexpr uint64_t kU64Max = 0xFFFFFFFFFFFFFFFFULL;

inline int ceil4(int c) { return (c + 3) / 4 * 4; }
inline int bit_length(long long v) {
  int b = 0;
  while (v > 0) { ++b; v >>= 1; }
  return b;
}

struct Conv { const float* w; const float* w_hi; const float* w_lo; const uint16_t* w_hl; const float* bias; int K, cin, cout, npad; };
struct Norm { const float* gamma; const float* beta; float eps; int c; };

struct Reader {
  const int64_t* p;
  int64_t i64() { return *p++; }
  int i32() { return (int)*p++; }
  Conv conv() {
    Conv c;
    c.w = (const float*)(uintptr_t)p[0]; c.w_hi = (const float*)(uintptr_t)p[1]; c.w_lo = (const float*)(uintptr_t)p[2];
    c.w_hl = (const uint16_t*)(uintptr_t)p[3];
    c.bias = (const float*)(uintptr_t)p[4]; c.K = (int)p[5]; c.cin = (int)p[6]; c.cout = (int)p[7]; c.npad = (int)p[8];
    p += 9;
    return c;
  }
  Norm norm() {
    Norm n;
    n.gamma = (const float*)(uintptr_t)p[0]; n.beta = (const float*)(uintptr_t)p[1];
    const int32_t bits = (int32_t)p[2];
    memcpy(&n.eps, &bits, 4);
    n.c = (int)p[3];
    p += 4;
    return n;
  }
};

struct Globals { const int32_t* k3[3]; const int32_t* k2[2]; const int32_t* subm3; };

inline Globals read_globals(const int64_t* g) {
  Globals G;
  for (int i = 0; i < 3; ++i) G.k3[i] = (const int32_t*)(uintptr_t)g[i];
  for (int i = 0; i < 2; ++i) G.k2[i] = (const int32_t*)(uintptr_t)g[3 + i];
  G.subm3 = (const int32_t*)(uintptr_t)g[5];
  return G;
}

inline int stride_slot(int stride) { return stride == 1 ? 0 : (stride == 2 ? 1 : 2); }

struct Exec {
  cudaStream_t st;
  char* base;
  size_t cap, off, peak;
  int32_t* pinned;
  int32_t* counters;
  int err;
  size_t launches;
};

thread_local int32_t* tl_pinned = nullptr;

// zeroed ticket counters for the fused split-K reduction / BatchNorm finalisation of csrc/spconv_hl.cu: one buffer per
// (host thread, stream); the kernels leave it zeroed, so it is cleared exactly once
constexpr int kCounters = 1024;
struct CounterSlot { cudaStream_t st; int32_t* p; };
thread_local std::vector<CounterSlot>* tl_counters = nullptr;

int32_t* counters_for(cudaStream_t st) {
  if (!tl_counters) tl_counters = new std::vector<CounterSlot>();
  for (auto& c : *tl_counters)
    if (c.st == st) return c.p;
  int32_t* p = nullptr;
  if (cudaMalloc((void**)&p, kCounters * sizeof(int32_t)) != cudaSuccess) return nullptr;
  if (cudaMemsetAsync(p, 0, kCounters * sizeof(int32_t), st) != cudaSuccess) { cudaFree(p); return nullptr; }
  tl_counters->push_back({st, p});
  return p;
}

inline bool exec_init(Exec& e, void* arena, size_t arena_bytes, cudaStream_t st) {
  e.st = st; e.base = (char*)arena; e.cap = arena_bytes; e.off = 0; e.peak = 0; e.err = EP_OK; e.launches = 0;
  if (!tl_pinned && cudaHostAlloc((void**)&tl_pinned, 256, cudaHostAllocDefault) != cudaSuccess) { tl_pinned = nullptr; return false; }
  e.pinned = tl_pinned;
  e.counters = counters_for(st);
  return arena != nullptr && ((uintptr_t)arena & 255) == 0;
}

template <class T>
inline T* alloc(Exec& e, size_t n) {
  const size_t bytes = (n * sizeof(T) + 255) & ~(size_t)255;
  if (e.err) return (T*)e.base;
  if (e.off + bytes > e.cap) { e.err = EP_ERR_WORKSPACE; return (T*)e.base; }
  T* p = (T*)(e.base + e.off);
  e.off += bytes;
  if (e.off > e.peak) e.peak = e.off;
  return p;
}

#define RUN(e, k, call)                                \
  do {                                                 \
    if (!(e).err) {                                    \
      const int s__ = (call);                          \
      if (s__ != EP_OK) (e).err = s__;                 \
      else (e).launches += (k);                        \
    }                                                  \
  } while (0)

// `flag` (optional): a second device word fetched in the same stream drain; *flag_out = its value
inline int read_i32(Exec& e, const int32_t* dev, const int32_t* flag = nullptr, int* flag_out = nullptr) {
  if (e.err) return 1;
  if (cudaMemcpyAsync(e.pinned, dev, sizeof(int32_t), cudaMemcpyDeviceToHost, e.st) != cudaSuccess ||
      (flag && cudaMemcpyAsync(e.pinned + 1, flag, sizeof(int32_t), cudaMemcpyDeviceToHost, e.st) != cudaSuccess) ||
      cudaStreamSynchronize(e.st) != cudaSuccess) {
    e.err = EP_ERR_CUDA;
    return 1;
  }
  const int v = e.pinned[0];
  if (flag_out) *flag_out = flag ? e.pinned[1] : 0;
  if (v < 0) { e.err = EP_ERR_CUDA; return 1; }
  return v;
}

// Z-order sort keys come in three widths (csrc/common.cuh): tier 0 = 8 coordinate bits per axis, batch 0 (24 bits, 3 radix
// passes), tier 1 = 9 + 5 batch bits (32 bits, 4 passes), kWideKeys = the 64-bit form (8 passes).  Same row order in all of
// them.  A point cloud that does not fit its tier raises a device flag, is redone one tier up, and the process starts there
// from then on (sticky: fragments of one deployment share their extent).
constexpr int kWideKeys = 2;
constexpr int kTierCoordBits[2] = {8, 9};
constexpr int kTierBatchBits[2] = {0, 5};
std::atomic<int> g_key_tier{[] { const char* v = getenv("EPRECON_KEY_TIER"); return v && v[0] >= '0' && v[0] <= '2' ? v[0] - '0' : 0; }()};

inline void zero_bytes(Exec& e, void* p, size_t bytes) {
  if (e.err || bytes == 0) return;
  if (cudaMemsetAsync(p, 0, bytes, e.st) != cudaSuccess) e.err = EP_ERR_CUDA;
  else e.launches += 1;
}

// ------------------------------------------------------------------------------------------------ sparse structures
struct Table { uint64_t* keys; int32_t* vals; int64_t cap; };
struct VoxelSet { int32_t* coords; int m; int stride; Table t; int32_t* k3; int key_tier; };
struct Csr { const int32_t* perm; const int32_t* s0; const int32_t* s1; int m; };
struct Seg { uint64_t* ks; int32_t* perm; int32_t* s0; int32_t* s1; int32_t* soi; int S; };
struct PointCloud { int n; float* scaled; Csr csr1; VoxelSet vox; int32_t* tap_idx[3]; float* tap_w[3]; };
struct Mat { float* p; int ld; uint16_t* hl; };   // hl: half-pair copy of the rows (csrc/spconv_hl.cu) or nullptr

Table make_table(Exec& e, const int32_t* coords, int m, int batch_first) {
  uint64_t* keys = alloc<uint64_t>(e, m);
  RUN(e, 1, ep_coord_keys(coords, m, batch_first, keys, e.st));
  int sh = bit_length(2LL * m - 1);
  if (sh < 4) sh = 4;
  Table t;
  t.cap = 1LL << sh;
  t.keys = alloc<uint64_t>(e, (size_t)t.cap);
  t.vals = alloc<int32_t>(e, (size_t)t.cap);
  RUN(e, 2, ep_hash_build(keys, m, t.keys, t.vals, t.cap, e.st));
  return t;
}

VoxelSet make_voxelset(Exec& e, int32_t* coords, int m, int stride) {
  VoxelSet v;
  v.coords = coords; v.m = m; v.stride = stride; v.k3 = nullptr; v.key_tier = kWideKeys;
  v.t = make_table(e, coords, m, 0);
  return v;
}

Seg sort_segments(Exec& e, const uint64_t* keys, int n, int key_bits, uint64_t sentinel, bool want_soi, bool want_count,
                  const int32_t* flag = nullptr, int* flag_out = nullptr) {
  Seg s;
  s.ks = alloc<uint64_t>(e, n);
  s.perm = alloc<int32_t>(e, n);
  s.s0 = alloc<int32_t>(e, n);
  s.s1 = alloc<int32_t>(e, n);
  s.soi = want_soi ? alloc<int32_t>(e, n) : nullptr;
  int32_t* nseg = alloc<int32_t>(e, 1);
  const size_t mark = e.off;
  const size_t wsb = ep_sort_segments_workspace_bytes(n);
  void* ws = alloc<char>(e, wsb);
  // iota + (histogram, scan, one onesweep pass per 8 key bits) + head flags / counts + scan + segment ids
  RUN(e, 6 + (key_bits + 7) / 8, ep_sort_segments(keys, n, key_bits, sentinel, s.ks, s.perm, s.s0, s.s1, s.soi, nseg, ws, wsb, e.st));
  if (!e.err) e.off = mark;  // the sort's workspace is dead once it is enqueued (later users are stream-ordered after it)
  s.S = want_count ? read_i32(e, nseg, flag, flag_out) : -1;
  return s;
}

void build_pc(Exec& e, PointCloud& pc, const float* pts, int n, float vres, bool spatial) {
  pc.n = n;
  for (int i = 0; i < 3; ++i) { pc.tap_idx[i] = nullptr; pc.tap_w[i] = nullptr; }
  pc.scaled = alloc<float>(e, 4 * (size_t)n);
  uint64_t* keys = alloc<uint64_t>(e, n);
  int tier = spatial ? g_key_tier.load(std::memory_order_relaxed) : kWideKeys;
  Seg s;
  for (;;) {
    const size_t mark = e.off;
    if (tier == kWideKeys) {
      RUN(e, 1, ep_point_keys(pts, n, vres, spatial ? 1 : 0, pc.scaled, keys, e.st));
      s = sort_segments(e, keys, n, spatial ? 64 : 60, kU64Max, true, true);
      break;
    }
    int32_t* viol = alloc<int32_t>(e, 1);
    zero_bytes(e, viol, sizeof(int32_t));
    const int cb = kTierCoordBits[tier], bb = kTierBatchBits[tier];
    RUN(e, 1, ep_point_keys_compact(pts, n, vres, cb, bb, pc.scaled, keys, viol, e.st));
    int violated = 0;
    s = sort_segments(e, keys, n, 3 * cb + bb, kU64Max, true, true, viol, &violated);
    if (e.err || !violated) break;
    e.off = mark;                                  // redo one tier up; remember it for the calls to come
    ++tier;
    int seen = g_key_tier.load(std::memory_order_relaxed);
    while (seen < tier && !g_key_tier.compare_exchange_weak(seen, tier, std::memory_order_relaxed)) {}
  }
  const int m = s.S;
  pc.csr1 = Csr{s.perm, s.s0, s.s1, m};
  int32_t* vox = alloc<int32_t>(e, 4 * (size_t)m);
  RUN(e, 1, ep_segment_coords(pc.scaled, s.perm, s.s0, m, vox, e.st));
  pc.vox = make_voxelset(e, vox, m, 1);  // keys = sphash(voxel coords): equal to the sorted point keys in hash order
  pc.vox.key_tier = tier;
}

int32_t* kmap_k3(Exec& e, VoxelSet& v, const Globals& G) {
  if (!v.k3) {
    v.k3 = alloc<int32_t>(e, 27 * (size_t)v.m);
    RUN(e, 1, ep_kmap_build(v.coords, v.m, 0, G.k3[stride_slot(v.stride)], 27, v.t.keys, v.t.vals, v.t.cap, 0, 0, 0, v.k3, e.st));
  }
  return v.k3;
}

void downsample(Exec& e, VoxelSet& v, const Globals& G, VoxelSet& coarse, int32_t*& nbr_down, int32_t*& nbr_up) {
  const int step = 2 * v.stride;
  uint64_t* keys = alloc<uint64_t>(e, v.m);
  const int tier = v.key_tier;                     // the fine set fits its tier, so do the coarse sites
  const int cb = tier == kWideKeys ? 0 : kTierCoordBits[tier];
  if (cb) RUN(e, 1, ep_down_keys_compact(v.coords, v.m, step, cb, keys, e.st));
  else RUN(e, 1, ep_down_keys(v.coords, v.m, step, keys, e.st));
  Seg s = sort_segments(e, keys, v.m, cb ? 3 * cb + kTierBatchBits[tier] : 64, kU64Max, false, true);
  const int mc = s.S;
  int32_t* cc = alloc<int32_t>(e, 4 * (size_t)mc);
  if (cb) RUN(e, 1, ep_down_unpack_compact(s.ks, s.s0, mc, cb, cc, e.st));
  else RUN(e, 1, ep_down_unpack(s.ks, s.s0, mc, cc, e.st));
  coarse = make_voxelset(e, cc, mc, step);
  coarse.key_tier = tier;
  nbr_down = alloc<int32_t>(e, 8 * (size_t)mc);
  RUN(e, 1, ep_kmap_build(cc, mc, 0, G.k2[stride_slot(v.stride)], 8, v.t.keys, v.t.vals, v.t.cap, 0, 0, 0, nbr_down, e.st));
  nbr_up = alloc<int32_t>(e, 8 * (size_t)v.m);
  RUN(e, 1, ep_fill_i32(nbr_up, 8LL * v.m, -1, e.st));
  RUN(e, 1, ep_kmap_inverse(nbr_down, mc, 8, nbr_up, e.st));
}

void taps(Exec& e, PointCloud& pc, const VoxelSet& v, const int32_t*& idx, const float*& w) {
  const int slot = stride_slot(v.stride);
  if (!pc.tap_idx[slot]) {
    pc.tap_idx[slot] = alloc<int32_t>(e, 8 * (size_t)pc.n);
    pc.tap_w[slot] = alloc<float>(e, 8 * (size_t)pc.n);
    RUN(e, 1, ep_devox_prepare(pc.scaled, pc.n, v.stride, v.t.keys, v.t.vals, v.t.cap, pc.tap_idx[slot], pc.tap_w[slot], e.st));
  }
  idx = pc.tap_idx[slot];
  w = pc.tap_w[slot];
}

Csr csr_for(Exec& e, PointCloud& pc, const VoxelSet& v) {
  int32_t* idx = alloc<int32_t>(e, pc.n);
  uint64_t* keys = alloc<uint64_t>(e, pc.n);
  RUN(e, 1, ep_point_query(pc.scaled, pc.n, v.stride, v.t.keys, v.t.vals, v.t.cap, idx, keys, v.m, e.st));
  int bits = bit_length(v.m);
  if (bits < 1) bits = 1;
  Seg s = sort_segments(e, keys, pc.n, bits, (uint64_t)v.m, false, true);
  int32_t* s0 = alloc<int32_t>(e, v.m);
  int32_t* s1 = alloc<int32_t>(e, v.m);
  zero_bytes(e, s0, sizeof(int32_t) * (size_t)v.m);
  zero_bytes(e, s1, sizeof(int32_t) * (size_t)v.m);
  if (s.S > 0) RUN(e, 1, ep_csr_expand(s.ks, s.s0, s.s1, s.S, s0, s1, e.st));
  return Csr{s.perm, s0, s1, v.m};
}

// ------------------------------------------------------------------------------------------------------ dense math
Mat voxelize(Exec& e, const Csr& csr, const float* feat, int ld, int c, bool want_hl = false) {
  Mat o{alloc<float>(e, (size_t)csr.m * ceil4(c)), ceil4(c), nullptr};
  if (want_hl) {   // the mean-pooled rows feed a tensor-core conv: write their half-pair copy in the same pass
    o.hl = alloc<uint16_t>(e, (size_t)csr.m * ep_hl_slabs(c) * 64);
    RUN(e, 1, ep_hl_segment_mean(feat, ld, c, csr.perm, csr.s0, csr.s1, csr.m, o.p, o.ld, o.hl, nullptr, e.st));
  } else {
    RUN(e, 1, ep_segment_mean(feat, ld, c, csr.perm, csr.s0, csr.s1, csr.m, o.p, o.ld, e.st));
  }
  return o;
}

Mat devoxelize(Exec& e, const Mat& v, int c, const int32_t* idx, const float* w, int n, const float* add, int ld_add,
               float* out = nullptr, int ld_out = 0) {
  Mat o{out, ld_out, nullptr};
  if (!out) { o.p = alloc<float>(e, (size_t)n * ceil4(c)); o.ld = ceil4(c); }
  RUN(e, 1, ep_devoxelize(v.p, v.ld, c, idx, w, n, add, ld_add, o.p, o.ld, e.st));
  return o;
}

float* bn_ss(Exec& e, const float* part, int m, const Norm& bn);

// out[j,:cout] = bias + sum_k W[k]^T x[nbr[j,k]]; K == 1 runs on the fp32 FFMA kernel, K > 1 on the tensor cores: the TMA
// gather kernel on half-pair operands (csrc/spconv_hl.cu; the input's m_in rows are split right here) when the descriptor
// carries w_hl, else the round-1 3xTF32 kernel
// bn + ss: also produce the train-mode BatchNorm scale / shift of the output (fused into the conv kernel where it pays)
Mat spconv(Exec& e, const float* x, int ldx, int m_in, const Conv& cv, const int32_t* nbr, int m_out, float** part,
           float* out = nullptr, int ld_out = 0, const uint16_t* x_hl = nullptr, const Norm* bn = nullptr,
           const float** ss = nullptr) {
  const int c4 = ceil4(cv.cout);
  Mat o{out, ld_out, nullptr};
  if (!out) {
    o.p = alloc<float>(e, (size_t)m_out * c4);
    o.ld = c4;
    if (c4 != cv.cout) zero_bytes(e, o.p, sizeof(float) * (size_t)m_out * c4);
  }
  float* pp = nullptr;
  if (part) { pp = alloc<float>(e, (size_t)ep_spconv_num_row_tiles(m_out) * 2 * cv.cout); *part = pp; }
  ProfRec* rec = nullptr;
  if (tl_prof && !e.err && tl_prof->pool_used < kProfPool) {
    ProfRec r{};
    r.K = cv.K; r.cin = cv.cin; r.cout = cv.cout; r.m_in = m_in; r.m_out = m_out;
    r.impl = cv.K == 1 ? 0 : (cv.w_hl ? 2 : 1);
    r.pairs_dev = tl_prof->pairs_pool + tl_prof->pool_used++;
    if (nbr) count_valid_kernel<<<ep_div_up((long long)m_out * cv.K, 256), 256, 0, e.st>>>(nbr, (long long)m_out * cv.K, r.pairs_dev);
    if (cudaEventCreate(&r.e0) == cudaSuccess && cudaEventCreate(&r.e1) == cudaSuccess) {
      cudaEventRecord(r.e0, e.st);
      tl_prof->recs.push_back(r);
      rec = &tl_prof->recs.back();
    }
  }
  struct ProfEnd { ProfRec* r; cudaStream_t st; ~ProfEnd() { if (r) cudaEventRecord(r->e1, st); } } prof_end{rec, e.st};
  if (cv.K == 1) {
    RUN(e, 1, ep_spconv_fwd(x, ldx, cv.cin, nbr, 1, cv.w, c4, cv.cout, cv.bias, o.p, o.ld, m_out, pp, e.st));
  } else if (cv.w_hl) {
    float* ssp = bn ? alloc<float>(e, 2 * (size_t)bn->c) : nullptr;
    const size_t mark = e.off;
    const uint16_t* xs = x_hl;
    if (!xs) {      // no producer wrote the half-pair copy (concat buffers, column slices): split here
      uint16_t* t = alloc<uint16_t>(e, (size_t)m_in * ep_hl_slabs(cv.cin) * 64);
      RUN(e, 1, ep_hl_split_rows(x, ldx, cv.cin, m_in, t, nullptr, e.st));
      xs = t;
    }
    const size_t wsb = ep_spconv_hl_workspace_bytes(m_out, cv.npad, cv.K);
    void* ws = wsb ? (void*)alloc<char>(e, wsb) : nullptr;
    RUN(e, ep_spconv_hl_launches(m_out, cv.cin, cv.npad, cv.K, e.counters != nullptr, bn != nullptr),
        ep_spconv_hl_fused_fwd(xs, m_in, cv.cin, nbr, cv.K, cv.w_hl, cv.npad, cv.cout, cv.bias, o.p, o.ld, m_out, pp, ws, wsb, 0,
                               e.counters, kCounters, bn ? bn->gamma : nullptr, bn ? bn->beta : nullptr, bn ? bn->eps : 0.f, ssp,
                               e.st));
    if (!e.err) e.off = mark;
    if (bn) *ss = ssp;
    return o;
  } else {
    const size_t mark = e.off;
    const size_t wsb = ep_spconv_tc_workspace_bytes(m_out, cv.npad, cv.K);
    void* ws = wsb ? (void*)alloc<char>(e, wsb) : nullptr;
    RUN(e, 1, ep_spconv_tc_fwd(x, ldx, cv.cin, nbr, cv.K, cv.w_hi, cv.w_lo, cv.npad, cv.cout, cv.bias, o.p, o.ld, m_out, pp, 3,
                               ws, wsb, e.st));
    if (!e.err) e.off = mark;
  }
  if (bn) *ss = bn_ss(e, pp, m_out, *bn);
  return o;
}

float* bn_ss(Exec& e, const float* part, int m, const Norm& bn) {
  float* ss = alloc<float>(e, 2 * (size_t)bn.c);
  RUN(e, 1, ep_bn_finalize(part, ep_spconv_num_row_tiles(m), bn.c, m, bn.eps, bn.gamma, bn.beta, ss, nullptr, e.st));
  return ss;
}

// BatchNorm apply (+ residual b, + ReLU) in place; with want_hl the same pass writes the rows' half-pair copy for the next conv
void bn_apply(Exec& e, Mat& y, const float* ss, const float* b, int ld_b, const float* ss_b, int m, int c, bool want_hl) {
  if (want_hl) {
    y.hl = alloc<uint16_t>(e, (size_t)m * ep_hl_slabs(c) * 64);
    RUN(e, 1, ep_hl_affine_act(y.p, y.ld, ss, b, ld_b, ss_b, 1, m, c, y.p, y.ld, y.hl, nullptr, e.st));
  } else {
    RUN(e, 1, ep_affine_act(y.p, y.ld, ss, b, ld_b, ss_b, 1, m, c, y.p, y.ld, e.st));
  }
}

// want_hl: the result feeds a K > 1 conv of a descriptor in half-pair mode (the caller checks the consumer's w_hl)
Mat conv_bn_relu(Exec& e, const Mat& x, int m_in, const int32_t* nbr, const Conv& cv, const Norm& bn, int m_out,
                 float* out = nullptr, int ld_out = 0, bool want_hl = false) {
  float* part = nullptr;
  const float* ss = nullptr;
  Mat y = spconv(e, x.p, x.ld, m_in, cv, nbr, m_out, &part, out, ld_out, x.hl, &bn, &ss);
  bn_apply(e, y, ss, nullptr, 0, nullptr, m_out, cv.cout, want_hl);
  return y;
}

struct ResBlock { Conv c1; Norm b1; Conv c2; Norm b2; int has_down; Conv cd; Norm bd; };

ResBlock read_res(Reader& r) {
  ResBlock b;
  b.c1 = r.conv(); b.b1 = r.norm(); b.c2 = r.conv(); b.b2 = r.norm();
  b.has_down = r.i32();
  if (b.has_down) { b.cd = r.conv(); b.bd = r.norm(); }
  return b;
}

Mat residual_block(Exec& e, const Mat& x, const int32_t* nbr, const ResBlock& b, int m, bool want_hl) {
  const bool hl = b.c2.w_hl != nullptr;
  Mat t = conv_bn_relu(e, x, m, nbr, b.c1, b.b1, m, nullptr, 0, hl);
  float* part_u = nullptr;
  const float* ss_u = nullptr;
  Mat u = spconv(e, t.p, t.ld, m, b.c2, nbr, m, &part_u, nullptr, 0, t.hl, &b.b2, &ss_u);
  if (!b.has_down) {
    bn_apply(e, u, ss_u, x.p, x.ld, nullptr, m, b.c2.cout, want_hl && hl);
  } else {
    float* part_d = nullptr;
    const float* ss_d = nullptr;
    Mat d = spconv(e, x.p, x.ld, m, b.cd, nullptr, m, &part_d, nullptr, 0, nullptr, &b.bd, &ss_d);
    bn_apply(e, u, ss_u, d.p, d.ld, ss_d, m, b.c2.cout, want_hl && hl);
  }
  return u;
}

inline void copy_cols(Exec& e, const float* src, int ld_src, int m, int c, float* dst, int ld_dst) {
  RUN(e, 1, ep_gather_rows(src, ld_src, nullptr, 0, 0.f, m, c, dst, ld_dst, e.st));
}

inline void layernorm(Exec& e, const float* x, int ldx, const float* res, int ld_res, int relu_before, const Norm& ln,
                      int relu_after, int m, float* out, int ld_out) {
  RUN(e, 1, ep_layernorm(x, ldx, res, ld_res, relu_before, ln.gamma, ln.beta, ln.eps, relu_after, m, ln.c, out, ld_out, e.st));
}

// SConv3d.run (models/modules.py:178-197): voxelize -> k3 conv -> devoxelize (given taps) + Linear
struct SConv { Conv conv; Conv lin; };

Mat sconv3d(Exec& e, const SConv& s, const float* feat, int ldf, int n, PointCloud& pc, const Globals& G,
            const int32_t* tap_idx, const float* tap_w) {
  Mat x = voxelize(e, pc.csr1, feat, ldf, s.conv.cin, s.conv.w_hl != nullptr);
  Mat y = spconv(e, x.p, x.ld, pc.vox.m, s.conv, kmap_k3(e, pc.vox, G), pc.vox.m, nullptr, nullptr, 0, x.hl);
  Mat lin = spconv(e, feat, ldf, n, s.lin, nullptr, n, nullptr);
  return devoxelize(e, y, s.conv.cout, tap_idx, tap_w, n, lin.p, lin.ld);
}

struct Gru { SConv z, r, q; };

void conv_gru(Exec& e, const Gru& g, const float* h, int ld_h, const float* x, int ld_x, int c, int u, PointCloud& pc1,
              PointCloud& pc2, const Globals& G, const int32_t* idx1, const float* w1, const int32_t* idx1_hash, float* out,
              int ld_out) {
  Mat hx{alloc<float>(e, (size_t)u * 2 * c), 2 * c, nullptr};
  copy_cols(e, h, ld_h, u, c, hx.p, hx.ld);
  copy_cols(e, x, ld_x, u, c, hx.p + c, hx.ld);
  Mat z_pre = sconv3d(e, g.z, hx.p, hx.ld, u, pc1, G, idx1, w1);
  Mat r_pre = sconv3d(e, g.r, hx.p, hx.ld, u, pc2, G, idx1_hash, w1);   // the reference's stale-cache quirk
  Mat rhx{alloc<float>(e, (size_t)u * 2 * c), 2 * c, nullptr};
  RUN(e, 1, ep_gru_rh(r_pre.p, r_pre.ld, h, ld_h, x, ld_x, u, c, rhx.p, rhx.ld, e.st));
  Mat q_pre = sconv3d(e, g.q, rhx.p, rhx.ld, u, pc1, G, idx1, w1);
  RUN(e, 1, ep_gru_out(z_pre.p, z_pre.ld, q_pre.p, q_pre.ld, h, ld_h, u, c, out, ld_out, e.st));
}

// ---------------------------------------------------------------------------------------- descriptor parsing
struct SpvcnnDesc {
  int cs[5], cin0;
  Conv stem, down1, down2, pt0, dec1, dec2, pt1;
  Norm stem_bn, down1_bn, down2_bn, pt0_bn, dec1_bn, dec2_bn, pt1_bn;
  ResBlock s1a, s1b, s2a, s2b, u1a, u1b, u2a, u2b;
};
struct GruDesc { int cv, ci; Gru gv, gi; };
struct Lin4xHead { int c_in, c_out, use_res; Conv l1, l2, l3; Norm n1, n2; };
struct InitDesc { int d; Norm norm0, norm4; Conv ec[7]; Norm en[7]; Conv sc[3]; Norm sn[3]; Conv subm4; };

bool parse_spvcnn(Reader& r, SpvcnnDesc& d) {
  for (int i = 0; i < 5; ++i) d.cs[i] = r.i32();
  d.cin0 = r.i32();
  d.stem = r.conv(); d.stem_bn = r.norm();
  d.down1 = r.conv(); d.down1_bn = r.norm();
  d.s1a = read_res(r); d.s1b = read_res(r);
  d.down2 = r.conv(); d.down2_bn = r.norm();
  d.s2a = read_res(r); d.s2b = read_res(r);
  d.pt0 = r.conv(); d.pt0_bn = r.norm();
  d.dec1 = r.conv(); d.dec1_bn = r.norm();
  d.u1a = read_res(r); d.u1b = read_res(r);
  d.dec2 = r.conv(); d.dec2_bn = r.norm();
  d.u2a = read_res(r); d.u2b = read_res(r);
  d.pt1 = r.conv(); d.pt1_bn = r.norm();
  if (r.i64() != 0x5350564Ell) return false;   // end marker: layout mismatch guard
  // shape consistency of the program (channel bookkeeping of models/modules.py:82-136)
  const int* cs = d.cs;
  return d.stem.cin == d.cin0 && d.stem.cout == cs[0] && d.stem.K == 27 && d.down1.K == 8 && d.down1.cin == cs[0] &&
         d.down1.cout == cs[0] && d.s1a.c1.cin == cs[0] && d.s1b.c2.cout == cs[1] && d.down2.cin == cs[1] &&
         d.s2b.c2.cout == cs[2] && d.pt0.K == 1 && d.pt0.cin == cs[0] && d.pt0.cout == cs[2] && d.dec1.K == 8 &&
         d.dec1.cin == cs[2] && d.dec1.cout == cs[3] && d.u1a.c1.cin == cs[3] + cs[1] && d.u1b.c2.cout == cs[3] &&
         d.dec2.cin == cs[3] && d.dec2.cout == cs[4] && d.u2a.c1.cin == cs[4] + cs[0] && d.u2b.c2.cout == cs[4] &&
         d.pt1.cin == cs[2] && d.pt1.cout == cs[4] && d.stem_bn.c == cs[0] && d.pt1_bn.c == cs[4] &&
         d.u1a.has_down == 1 && d.s1b.has_down == 0;
}

bool parse_gru(Reader& r, GruDesc& d) {
  d.cv = r.i32(); d.ci = r.i32();
  for (Gru* g : {&d.gv, &d.gi})
    for (SConv* s : {&g->z, &g->r, &g->q}) { s->conv = r.conv(); s->lin = r.conv(); }
  if (r.i64() != 0x47525546ll) return false;
  for (int k = 0; k < 2; ++k) {
    const Gru& g = k ? d.gi : d.gv;
    const int c = k ? d.ci : d.cv;
    for (const SConv* s : {&g.z, &g.r, &g.q})
      if (s->conv.K != 27 || s->conv.cin != 2 * c || s->conv.cout != c || s->lin.K != 1 || s->lin.cin != 2 * c ||
          s->lin.cout != c || s->conv.bias || !s->lin.bias)
        return false;
  }
  return true;
}

bool parse_lin4x_head(Reader& r, Lin4xHead& h) {
  h.c_in = r.i32(); h.c_out = r.i32(); h.use_res = r.i32();
  h.l1 = r.conv(); h.n1 = r.norm();
  h.l2 = r.conv(); h.n2 = r.norm();
  h.l3 = r.conv();
  return h.l1.K == 1 && h.l1.cin == h.c_in && h.l1.cout == 4 * h.c_in && h.n1.c == 4 * h.c_in && h.l2.cin == 4 * h.c_in &&
         h.l2.cout == h.c_in && h.n2.c == h.c_in && h.l3.cin == h.c_in && h.l3.cout == h.c_out &&
         (h.use_res == 0 || h.c_in == h.c_out);
}

bool parse_init(Reader& r, InitDesc& d) {
  d.d = r.i32();
  d.norm0 = r.norm();
  for (int i = 0; i < 7; ++i) { d.ec[i] = r.conv(); d.en[i] = r.norm(); }
  for (int i = 0; i < 3; ++i) { d.sc[i] = r.conv(); d.sn[i] = r.norm(); }
  d.subm4 = r.conv(); d.norm4 = r.norm();
  if (r.i64() != 0x494E4954ll) return false;
  const int dd = d.d, hc = dd / 2;
  return d.norm0.c == dd && d.ec[0].K == 1 && d.ec[0].cin == dd && d.ec[2].K == 27 && d.ec[2].cin == dd && d.ec[2].cout == hc &&
         d.ec[5].cout == hc && d.ec[6].K == 1 && d.ec[6].cin == 4 * dd && d.ec[6].cout == dd && d.sc[0].K == 27 &&
         d.sc[2].cout == dd && d.subm4.cout == 1 && d.subm4.K == 27 && d.norm4.c == 1 && d.ec[0].bias && d.subm4.bias;
}

inline int finish(Exec& e, int64_t* stats) {
  g_launches.fetch_add(e.launches, std::memory_order_relaxed);
  if (stats) { stats[0] = (int64_t)e.peak; stats[1] = (int64_t)e.launches; }
  return e.err;
}

}  // namespace

extern "C" {

size_t ep_exec_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

// Per-launch profiling of the sparse-conv family for the CALLING host thread (see ProfRec).  enable(1) starts recording,
// collect() synchronises the recorded events and returns up to `cap` records: meta[i] = {K, cin, cout, m_in, m_out, pairs,
// impl (0 FFMA linear, 1 3xTF32, 2 half-pair TMA)}, ms[i] = device time between the bracketing events; recording stops.
// Z-order sort-key tier new point clouds start in (0 = 24-bit keys, 1 = 32-bit, 2 = 64-bit; raised automatically when a cloud
// does not fit).  set < 0 only queries.  Returns the tier in force before the call.
int ep_exec_key_tier(int set) {
  const int old = g_key_tier.load(std::memory_order_relaxed);
  if (set >= 0 && set <= kWideKeys) g_key_tier.store(set, std::memory_order_relaxed);
  return old;
}

int ep_exec_profile_enable(int on) {
  if (on) {
    if (!tl_prof) tl_prof = new Prof();
    if (!tl_prof->pairs_pool && cudaMalloc((void**)&tl_prof->pairs_pool, kProfPool * sizeof(int)) != cudaSuccess) return EP_ERR_CUDA;
    if (cudaMemset(tl_prof->pairs_pool, 0, kProfPool * sizeof(int)) != cudaSuccess) return EP_ERR_CUDA;
    tl_prof->pool_used = 0;
    for (auto& r : tl_prof->recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    tl_prof->recs.clear();
    tl_prof->recs.reserve(kProfPool);
  } else if (tl_prof) {
    for (auto& r : tl_prof->recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    if (tl_prof->pairs_pool) cudaFree(tl_prof->pairs_pool);
    delete tl_prof;
    tl_prof = nullptr;
  }
  return EP_OK;
}

int ep_exec_profile_collect(int64_t* meta, float* ms, int cap) {
  if (!tl_prof || !meta || !ms) return EP_ERR_ARG;
  if (cudaDeviceSynchronize() != cudaSuccess) return EP_ERR_CUDA;
  std::vector<int> pairs(kProfPool, 0);
  if (cudaMemcpy(pairs.data(), tl_prof->pairs_pool, kProfPool * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return EP_ERR_CUDA;
  int n = 0;
  for (auto& r : tl_prof->recs) {
    if (n < cap) {
      float t = 0.f;
      cudaEventElapsedTime(&t, r.e0, r.e1);
      const int idx = (int)(r.pairs_dev - tl_prof->pairs_pool);
      int64_t* m = meta + 7 * (size_t)n;
      m[0] = r.K; m[1] = r.cin; m[2] = r.cout; m[3] = r.m_in; m[4] = r.m_out;
      m[5] = r.K == 1 ? r.m_out : pairs[idx];
      m[6] = r.impl;
      ms[n] = t;
      ++n;
    }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  tl_prof->recs.clear();
  const int ret = n;
  ep_exec_profile_enable(0);
  return ret;
}

// Host-only: parse a descriptor exactly as the executor would (kind 0 SPVCNN, 1 GRU level, 2 Linear4x, 3 init head) and
// return the number of int64 words consumed, or EP_ERR_ARG on a layout / shape mismatch.  Needs no GPU.
int ep_exec_desc_check(int kind, const int64_t* desc) {
  if (!desc) return EP_ERR_ARG;
  Reader r{desc};
  bool ok = false;
  if (kind == 0) { SpvcnnDesc d; ok = parse_spvcnn(r, d); }
  else if (kind == 1) { GruDesc d; ok = parse_gru(r, d); }
  else if (kind == 2) {
    const int n = r.i32();
    ok = n >= 1 && n <= 16;
    for (int i = 0; ok && i < n; ++i) { Lin4xHead h; ok = parse_lin4x_head(r, h); }
    ok = ok && r.i64() == 0x4C345854ll;
  } else if (kind == 3) { InitDesc d; ok = parse_init(r, d); }
  return ok ? (int)(r.p - desc) : EP_ERR_ARG;
}

int ep_csr_expand(const uint64_t* keys_sorted, const int32_t* seg_start, const int32_t* seg_end, int64_t n_segments,
                  int32_t* s0, int32_t* s1, cudaStream_t stream) {
  if (n_segments <= 0) return EP_ERR_ARG;
  csr_expand_kernel<<<ep_div_up(n_segments, 256), 256, 0, stream>>>(keys_sorted, seg_start, seg_end, (int)n_segments, s0, s1);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

int ep_translate_index(const int32_t* idx, int64_t n, const int32_t* rank, int32_t* out, cudaStream_t stream) {
  if (n <= 0) return EP_ERR_ARG;
  translate_index_kernel<<<ep_div_up(n, 256), 256, 0, stream>>>(idx, n, rank, out);
  EP_CHECK_LAUNCH();
  return EP_OK;
}

// SPVCNN.forward (models/modules.py:138-175).  pts f32 [n,4]=(x,y,z,b) metres; feat f32 [n, ld_feat] (cin columns);
// out f32 [n, ld_out] receives cs[4] columns.  desc: see executor.py::spvcnn_desc.
int ep_exec_spvcnn(const int64_t* desc, const int64_t* globals, const float* pts, const float* feat, int ld_feat,
                   int64_t n64, float vres, float* out, int ld_out, void* arena, size_t arena_bytes, int64_t* stats,
                   cudaStream_t stream) {
  if (!desc || !globals || n64 <= 0 || n64 > 0x7fffffffLL || ld_feat % 4 || ld_out % 4) return EP_ERR_ARG;
  Exec e;
  if (!exec_init(e, arena, arena_bytes, stream)) return EP_ERR_ARG;
  const Globals G = read_globals(globals);
  const int n = (int)n64;
  Reader r{desc};
  SpvcnnDesc D;
  if (!parse_spvcnn(r, D)) return EP_ERR_ARG;
  const int* cs = D.cs;
  const int cin0 = D.cin0;
  if (ld_feat < ceil4(cin0) || ld_out < cs[4]) return EP_ERR_ARG;
  const Conv &stem = D.stem, &down1 = D.down1, &down2 = D.down2, &pt0 = D.pt0, &dec1 = D.dec1, &dec2 = D.dec2, &pt1 = D.pt1;
  const Norm &stem_bn = D.stem_bn, &down1_bn = D.down1_bn, &down2_bn = D.down2_bn, &pt0_bn = D.pt0_bn, &dec1_bn = D.dec1_bn,
             &dec2_bn = D.dec2_bn, &pt1_bn = D.pt1_bn;
  const ResBlock &s1a = D.s1a, &s1b = D.s1b, &s2a = D.s2a, &s2b = D.s2b, &u1a = D.u1a, &u1b = D.u1b, &u2a = D.u2a, &u2b = D.u2b;

  PointCloud pc;
  build_pc(e, pc, pts, n, vres, true);                                             // initial_voxelize
  VoxelSet& v0 = pc.vox;
  const bool hl = stem.w_hl != nullptr;    // half-pair mode: producers write the operand copy their tensor-core consumer gathers
  Mat x0 = voxelize(e, pc.csr1, feat, ld_feat, cin0, hl);
  x0 = conv_bn_relu(e, x0, v0.m, kmap_k3(e, v0, G), stem, stem_bn, v0.m);                  // stem
  const int32_t* idx1; const float* w1;
  taps(e, pc, v0, idx1, w1);
  Mat z0 = devoxelize(e, x0, cs[0], idx1, w1, n, nullptr, 0);                      // voxel_to_point(x0, z)
  Mat x1 = voxelize(e, pc.csr1, z0.p, z0.ld, cs[0], hl);                           // point_to_voxel(x0, z0)
  VoxelSet v1, v2;
  int32_t *down01, *up10, *down12, *up21;
  downsample(e, v0, G, v1, down01, up10);
  x1 = conv_bn_relu(e, x1, v0.m, down01, down1, down1_bn, v1.m, nullptr, 0, hl);
  x1 = residual_block(e, x1, kmap_k3(e, v1, G), s1a, v1.m, true);
  x1 = residual_block(e, x1, kmap_k3(e, v1, G), s1b, v1.m, true);
  downsample(e, v1, G, v2, down12, up21);
  Mat x2 = conv_bn_relu(e, x1, v1.m, down12, down2, down2_bn, v2.m, nullptr, 0, hl);
  x2 = residual_block(e, x2, kmap_k3(e, v2, G), s2a, v2.m, true);
  x2 = residual_block(e, x2, kmap_k3(e, v2, G), s2b, v2.m, false);
  const int32_t* idx4; const float* w4;
  taps(e, pc, v2, idx4, w4);
  Mat z0m{z0.p, z0.ld, nullptr};
  Mat p0 = conv_bn_relu(e, z0m, n, nullptr, pt0, pt0_bn, n);                       // point_transforms[0](z0.F)
  Mat z1 = devoxelize(e, x2, cs[2], idx4, w4, n, p0.p, p0.ld);
  const Csr csr2 = csr_for(e, pc, v2);
  Mat y3 = voxelize(e, csr2, z1.p, z1.ld, cs[2], hl);                              // point_to_voxel(x2, z1)
  // up1: transposed conv to stride 2, concat skip x1, two residual blocks
  Mat cat1{alloc<float>(e, (size_t)v1.m * (cs[3] + cs[1])), cs[3] + cs[1], nullptr};
  conv_bn_relu(e, y3, v2.m, up21, dec1, dec1_bn, v1.m, cat1.p, cat1.ld);
  copy_cols(e, x1.p, x1.ld, v1.m, cs[1], cat1.p + cs[3], cat1.ld);
  y3 = residual_block(e, cat1, kmap_k3(e, v1, G), u1a, v1.m, true);
  y3 = residual_block(e, y3, kmap_k3(e, v1, G), u1b, v1.m, true);
  Mat cat0{alloc<float>(e, (size_t)v0.m * (cs[4] + cs[0])), cs[4] + cs[0], nullptr};
  conv_bn_relu(e, y3, v1.m, up10, dec2, dec2_bn, v0.m, cat0.p, cat0.ld);
  copy_cols(e, x0.p, x0.ld, v0.m, cs[0], cat0.p + cs[4], cat0.ld);
  Mat y4 = residual_block(e, cat0, kmap_k3(e, v0, G), u2a, v0.m, true);
  y4 = residual_block(e, y4, kmap_k3(e, v0, G), u2b, v0.m, false);
  Mat z1m{z1.p, z1.ld, nullptr};
  Mat p1 = conv_bn_relu(e, z1m, n, nullptr, pt1, pt1_bn, n);                       // point_transforms[1](z1.F)
  devoxelize(e, y4, cs[4], idx1, w1, n, p1.p, p1.ld, out, ld_out);
  return finish(e, stats);
}

// The two ConvGRUs of one GRUFusion level (models/gru_fusion.py:339-349; models/modules.py:200-222) on shared
// voxelisations.  pts f32 [u,4] aligned-camera points; h = global (hidden) rows, x = current rows, both [u, ld] with
// the voxel-feature columns [0,cv) followed by the image-feature columns [cv, cv+ci); out [u, ld_out] same layout.
int ep_exec_gru_level(const int64_t* desc, const int64_t* globals, const float* pts, int64_t u64, float vres,
                      const float* h, int ld_h, const float* x, int ld_x, float* out, int ld_out, void* arena,
                      size_t arena_bytes, int64_t* stats, cudaStream_t stream) {
  if (!desc || !globals || u64 <= 0 || u64 > 0x7fffffffLL) return EP_ERR_ARG;
  Exec e;
  if (!exec_init(e, arena, arena_bytes, stream)) return EP_ERR_ARG;
  const Globals G = read_globals(globals);
  const int u = (int)u64;
  Reader r{desc};
  GruDesc D;
  if (!parse_gru(r, D)) return EP_ERR_ARG;
  const int cv = D.cv, ci = D.ci;
  const Gru &gv = D.gv, &gi = D.gi;
  if (ld_h < cv + ci || ld_x < cv + ci || ld_out < cv + ci) return EP_ERR_ARG;

  PointCloud pc1, pc2;
  build_pc(e, pc1, pts, u, vres, true);
  build_pc(e, pc2, pc1.scaled, u, vres, false);   // convr: coordinates divided by vres AGAIN, rows in ascending-hash order
  const int32_t* idx1; const float* w1;
  taps(e, pc1, pc1.vox, idx1, w1);
  // convz's cached taps as convr sees them: voxel ids translated to the reference's ascending-hash row order
  uint64_t* hk = alloc<uint64_t>(e, pc1.vox.m);
  RUN(e, 1, ep_coord_keys(pc1.vox.coords, pc1.vox.m, 0, hk, e.st));
  Seg hs = sort_segments(e, hk, pc1.vox.m, 60, kU64Max, true, false);
  int32_t* idx1_hash = alloc<int32_t>(e, 8 * (size_t)u);
  RUN(e, 1, ep_translate_index(idx1, 8LL * u, hs.soi, idx1_hash, e.st));
  conv_gru(e, gv, h, ld_h, x, ld_x, cv, u, pc1, pc2, G, idx1, w1, idx1_hash, out, ld_out);
  conv_gru(e, gi, h + cv, ld_h, x + cv, ld_x, ci, u, pc1, pc2, G, idx1, w1, idx1_hash, out + cv, ld_out);
  return finish(e, stats);
}

// n_heads Linear4xTrans heads (models/modules.py:273-311) on the same rows.  outs: n_heads device pointers (as int64),
// each [m, ceil4(C_out)] contiguous.
int ep_exec_linear4x(const int64_t* desc, const float* x, int ld_x, int64_t m64, const int64_t* outs, void* arena,
                     size_t arena_bytes, int64_t* stats, cudaStream_t stream) {
  if (!desc || !outs || m64 <= 0 || m64 > 0x7fffffffLL || ld_x % 4) return EP_ERR_ARG;
  Exec e;
  if (!exec_init(e, arena, arena_bytes, stream)) return EP_ERR_ARG;
  const int m = (int)m64;
  Reader r{desc};
  const int n_heads = r.i32();
  for (int hd = 0; hd < n_heads; ++hd) {
    Lin4xHead H;
    if (!parse_lin4x_head(r, H)) return EP_ERR_ARG;
    const int c_out = H.c_out, use_res = H.use_res;
    const Conv &l1 = H.l1, &l2 = H.l2, &l3 = H.l3;
    const Norm &n1 = H.n1, &n2 = H.n2;
    if (ld_x < ceil4(H.c_in)) return EP_ERR_ARG;
    float* out = (float*)(uintptr_t)outs[hd];
    const size_t mark = e.off;
    Mat y1 = spconv(e, x, ld_x, m, l1, nullptr, m, nullptr);
    layernorm(e, y1.p, y1.ld, nullptr, 0, 0, n1, 1, m, y1.p, y1.ld);
    Mat y2 = spconv(e, y1.p, y1.ld, m, l2, nullptr, m, nullptr);
    layernorm(e, y2.p, y2.ld, nullptr, 0, 0, n2, 1, m, y2.p, y2.ld);
    const int c4 = ceil4(c_out);
    if (c4 != c_out) zero_bytes(e, out, sizeof(float) * (size_t)m * c4);
    spconv(e, y2.p, y2.ld, m, l3, nullptr, m, nullptr, out, c4);
    if (use_res) RUN(e, 1, ep_affine_act(out, c4, nullptr, y2.p, y2.ld, nullptr, 0, m, c_out, out, c4, e.st));
    if (!e.err) e.off = mark;
  }
  if (r.i64() != 0x4C345854ll) return EP_ERR_ARG;
  return finish(e, stats);
}

// Sparse head of the occupancy initialisation for one fragment (models/occupancy_initialization.py:131-176):
// BN1d -> Spares3dELAN -> 3 x [SubM k3, ReLU, +res, LayerNorm] -> SubM k3 32->1 -> BN1d.  var f32 [m, ld_var];
// coords int32 [m,4]=(0,x,y,z) in the sx x sy x sz site grid; occ f32 [m,4] (column 0 = logit, the rest zero).
int ep_exec_init_head(const int64_t* desc, const int64_t* globals, const float* var, int ld_var, const int32_t* coords,
                      int64_t m64, int sx, int sy, int sz, float* occ, void* arena, size_t arena_bytes, int64_t* stats,
                      cudaStream_t stream) {
  if (!desc || !globals || m64 <= 0 || m64 > 0x7fffffffLL || ld_var % 4) return EP_ERR_ARG;
  Exec e;
  if (!exec_init(e, arena, arena_bytes, stream)) return EP_ERR_ARG;
  const Globals G = read_globals(globals);
  const int m = (int)m64;
  Reader r{desc};
  InitDesc D;
  if (!parse_init(r, D)) return EP_ERR_ARG;
  const int d = D.d, hc = d / 2;
  const Norm &norm0 = D.norm0, &norm4 = D.norm4;
  const Conv* ec = D.ec; const Norm* en = D.en;
  const Conv* sc = D.sc; const Norm* sn = D.sn;
  const Conv& subm4 = D.subm4;
  if (ld_var < d) return EP_ERR_ARG;

  Table t = make_table(e, coords, m, 1);
  int32_t* k3 = alloc<int32_t>(e, 27 * (size_t)m);
  RUN(e, 1, ep_kmap_build(coords, m, 1, G.subm3, 27, t.keys, t.vals, t.cap, sx, sy, sz, k3, e.st));
  auto nbr_of = [&](const Conv& c) -> const int32_t* { return c.K == 1 ? nullptr : k3; };

  Mat x{alloc<float>(e, (size_t)m * d), d, nullptr};
  float* part0 = alloc<float>(e, (size_t)ep_spconv_num_row_tiles(m) * 2 * d);
  RUN(e, 1, ep_colstats(var, ld_var, m, d, part0, e.st));
  const float* ss0 = bn_ss(e, part0, m, norm0);
  RUN(e, 1, ep_affine_act(var, ld_var, ss0, nullptr, 0, nullptr, 0, m, d, x.p, x.ld, e.st));

  // ELAN: cat = [f1 | f2 | c3 | c4 | c5 | c6]
  Mat cat{alloc<float>(e, (size_t)m * 4 * d), 4 * d, nullptr};
  auto block = [&](int i, const float* in, int ld_in, float* out, int ld_o) {
    Mat y = spconv(e, in, ld_in, m, ec[i], nbr_of(ec[i]), m, nullptr);
    if (!out) { out = y.p; ld_o = y.ld; }
    layernorm(e, y.p, y.ld, nullptr, 0, 0, en[i], 1, m, out, ld_o);
    return Mat{out, ld_o, nullptr};
  };
  block(0, x.p, x.ld, cat.p, cat.ld);
  block(1, x.p, x.ld, cat.p + d, cat.ld);
  block(2, cat.p + d, cat.ld, cat.p + 2 * d, cat.ld);
  block(3, cat.p + 2 * d, cat.ld, cat.p + 2 * d + hc, cat.ld);
  block(4, cat.p + 2 * d + hc, cat.ld, cat.p + 2 * d + 2 * hc, cat.ld);
  block(5, cat.p + 2 * d + 2 * hc, cat.ld, cat.p + 2 * d + 3 * hc, cat.ld);
  x = block(6, cat.p, cat.ld, nullptr, 0);
  for (int i = 0; i < 3; ++i) {
    Mat y = spconv(e, x.p, x.ld, m, sc[i], nbr_of(sc[i]), m, nullptr);
    layernorm(e, y.p, y.ld, x.p, x.ld, 1, sn[i], 0, m, y.p, y.ld);
    x = y;
  }
  zero_bytes(e, occ, sizeof(float) * (size_t)m * 4);
  float* part4 = nullptr;
  spconv(e, x.p, x.ld, m, subm4, nbr_of(subm4), m, &part4, occ, 4);
  const float* ss4 = bn_ss(e, part4, m, norm4);
  RUN(e, 1, ep_affine_act(occ, 4, ss4, nullptr, 0, nullptr, 0, m, 1, occ, 4, e.st));
  return finish(e, stats);
}

}  // extern "C"

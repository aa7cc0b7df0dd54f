// Probe for the next conv-kernel generation (limiter 1): tcgen05.mma with the A operand in TENSOR
// MEMORY, written there with tcgen05.st straight from registers, so that the gathered + split operand never passes
// through shared memory.  One CTA, M = 128, N = 16, K = 8 (kind::tf32).  The probe writes A[m][k] = (m % 7) - 3 + k to
// TMEM lane m, column (A_COL + k) -- the layout ASSUMPTION to be confirmed: one 32-bit element per (lane = row,
// column = k) -- B[k][n] = (k + 1) * ((n % 5) - 2) to shared memory in the canonical K-major / no-swizzle layout the
// production kernel uses, issues one MMA and checks D = A.B exactly (small integers are exact in tf32).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I eprecon_b200/csrc -I include -o /tmp/probe_tmem_a \
//        tools/probes/probe_tmem_a_operand.cu && /tmp/probe_tmem_a
//
// Compiles here (ptxas accepts both instruction forms for sm_100a); NOT yet run on a GPU (round 1 ran out of GPU budget).
#include <cstdio>
#include <cuda_runtime.h>

#include "tc_common.cuh"

using namespace eptc;

constexpr int M = 128, N = 16, K = 8;
constexpr uint32_t D_COL = 0, A_COL = 32, TMEM_COLS = 64;

__global__ void __launch_bounds__(128) probe_kernel(float* __restrict__ d_out, int* __restrict__ n_bad) {
  __shared__ __align__(1024) float s_b[2 * N * 4];   // two 16-byte K chunks x N rows (core matrices of 8 rows x 16 B)
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // B in the production layout: chunk j (k = 4j..4j+3) plane of N rows, 16 bytes per row
  for (int e = tid; e < 2 * N; e += 128) {
    const int j = e / N, n = e - j * N;
    for (int i = 0; i < 4; ++i) s_b[(j * N + n) * 4 + i] = (float)((4 * j + i + 1) * ((n % 5) - 2));
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  // A: thread = row m (TMEM lane), 8 consecutive columns
  {
    uint32_t a[K];
    for (int k = 0; k < K; ++k) a[k] = __float_as_uint((float)((tid % 7) - 3 + k));
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + A_COL;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(a[0]), "r"(a[1]),
                 "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7])
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint64_t db = umma_desc(smem_u32(s_b), N * 16, 128);
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem + D_COL),
        "r"(tmem + A_COL), "l"(db), "r"(idesc), "r"(0u)
        : "memory");
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  float v[16];
  tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + D_COL, v);
  int bad = 0;
  for (int n = 0; n < N; ++n) {
    float want = 0.f;
    for (int k = 0; k < K; ++k) want += (float)((tid % 7) - 3 + k) * (float)((k + 1) * ((n % 5) - 2));
    d_out[tid * N + n] = v[n];
    bad += (v[n] != want);
  }
  if (bad) atomicAdd(n_bad, bad);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
}

int main() {
  float* d;
  int* nb;
  cudaMalloc(&d, M * N * sizeof(float));
  cudaMalloc(&nb, sizeof(int));
  cudaMemset(nb, 0, sizeof(int));
  probe_kernel<<<1, 128>>>(d, nb);
  cudaError_t e = cudaDeviceSynchronize();
  int bad = -1;
  cudaMemcpy(&bad, nb, sizeof(int), cudaMemcpyDeviceToHost);
  float h[2 * N];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("status %s, mismatching elements %d of %d; row 0:", cudaGetErrorString(e), bad, M * N);
  for (int n = 0; n < N; ++n) printf(" %g", h[n]);
  printf("\n");
  return (e == cudaSuccess && bad == 0) ? 0 : 1;
}

// GPU probe (developer tool): tcgen05.mma.cta_group::2 (CTA pair, M = 256 over two SMs, the N = 96 weight operand
// split 48 / 48 between the two CTAs' shared memories) -- correctness of the operand / accumulator placement the
// paired column-scatter conv relies on, and the instruction rate next to the single-CTA M = 128 form.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/probe_pair.cu -o build/probe_pair
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../xmm_superres_denoise_b200/csrc/ptx_sm100.cuh"

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

using namespace xmm;

constexpr int N = 96;

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

// [rows][32 ch] bf16 K-major SWIZZLE_64B: byte offset of element (row, k)
__device__ __forceinline__ uint32_t sw64(int row, int k) {
  const uint32_t chunk = uint32_t(k >> 3) ^ (uint32_t(row >> 1) & 3u);
  return uint32_t(row) * 64u + chunk * 16u + uint32_t(k & 7) * 2u;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
pair_kernel(float* out, long long* cycles, int iters, int check) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const uint32_t rank = cluster_rank();
  uint8_t* a_s = smem;           // [128][32] bf16 = 8 KB
  uint8_t* b_s = smem + 8192;    // [48][32] bf16 = 3 KB (this CTA's half of N)
  for (int i = threadIdx.x; i < 128 * 32; i += blockDim.x) {
    const int m = i / 32, k = i % 32;
    *reinterpret_cast<__nv_bfloat16*>(a_s + sw64(m, k)) = __float2bfloat16(float((m + 3 * k) % 7 - 3 + int(rank)));
  }
  for (int i = threadIdx.x; i < 48 * 32; i += blockDim.x) {
    const int nl = i / 32, k = i % 32, n = nl + 48 * int(rank);
    *reinterpret_cast<__nv_bfloat16*>(b_s + sw64(nl, k)) = __float2bfloat16(float((n + k) % 5 - 2));
  }
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_mbar_init();
  }
  ptx::fence_proxy_async();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(&tmem_ptr)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  ptx::tc_fence_before();
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  ptx::tc_fence_after();
  const uint32_t tb = tmem_ptr;
  if (rank == 0 && threadIdx.x < 32 && ptx::elect_one()) {
    const uint64_t adesc = ptx::umma_smem_desc(ptx::smem_u32(a_s), 16, 512, ptx::UMMA_SW64);
    const uint64_t bdesc = ptx::umma_smem_desc(ptx::smem_u32(b_s), 16, 512, ptx::UMMA_SW64);
    constexpr uint32_t idesc = ptx::umma_idesc_bf16_f32(256, N, 0, 0);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const uint32_t acc = (it == 0 && ks == 0) ? 0u : 1u;
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tb + uint32_t((it & 3) * N)),
            "l"(adesc + uint64_t((ks * 32) >> 4)), "l"(bdesc + uint64_t((ks * 32) >> 4)), "r"(idesc), "r"(check ? acc : 1u)
            : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     ptx::smem_u32(&bar)),
                 "h"(uint16_t(3))
                 : "memory");
    ptx::mbar_wait(&bar, 0);
    cycles[blockIdx.x / 2] = clock64() - t0;
  }
  ptx::mbar_wait(&bar, 0);  // both CTAs: the multicast commit arrives on each CTA's own barrier
  ptx::tc_fence_after();
  if (check) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t r[32];
      ptx::tmem_ld_32x32(tb + (uint32_t(warp * 32) << 16) + uint32_t(c0), r);
      ptx::tmem_ld_wait();
      for (int i = 0; i < 32; ++i)
        out[(size_t(blockIdx.x) * 128 + warp * 32 + lane) * N + c0 + i] = __uint_as_float(r[i]);
    }
  }
  ptx::tc_fence_before();
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (threadIdx.x < 32) {
    ptx::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tb), "n"(512) : "memory");
  }
}

int main() {
  const size_t smem = 1024 + 8192 + 4096;
  CK(cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  float* out;
  long long* cyc;
  CK(cudaMalloc(&out, size_t(148) * 128 * N * sizeof(float)));
  CK(cudaMalloc(&cyc, 148 * sizeof(long long)));
  // correctness: one pair, one iteration (K = 32)
  pair_kernel<<<2, 128, smem>>>(out, cyc, 1, 1);
  CK(cudaDeviceSynchronize());
  std::vector<float> h(size_t(2) * 128 * N);
  CK(cudaMemcpy(h.data(), out, h.size() * sizeof(float), cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int c = 0; c < 2; ++c)
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < N; ++n) {
        float want = 0;
        for (int k = 0; k < 32; ++k) want += float((m + 3 * k) % 7 - 3 + c) * float((n + k) % 5 - 2);
        const float got = h[(size_t(c) * 128 + m) * N + n];
        if (got != want && bad++ < 8) printf("mismatch cta %d m %d n %d: got %g want %g\n", c, m, n, got, want);
      }
  printf("pair MMA M=256 N=%d: %s (%d mismatches)\n", N, bad ? "WRONG" : "correct", bad);
  for (int grid : {2, 148}) {
    const int iters = 20000;
    for (int rep = 0; rep < 2; ++rep) {
      pair_kernel<<<grid, 128, smem>>>(out, cyc, iters, 0);
      CK(cudaDeviceSynchronize());
    }
    std::vector<long long> hc(grid / 2);
    CK(cudaMemcpy(hc.data(), cyc, hc.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (auto v : hc) mx = std::max(mx, v);
    printf("pair rate grid=%3d: %.2f cycles per M=256,N=%d,K=16 MMA (single-CTA M=128: 56.1; tensor ideal 48)\n", grid,
           double(mx) / (2.0 * iters), N);
  }
  return 0;
}

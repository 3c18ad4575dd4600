// GPU probe (developer tool): issue rate of the MN-major tcgen05.mma streams the weight-gradient kernel uses.
//   mode 0: main role as shipped    A = 2 x 64-channel SW128 boxes of a [10][10]-pixel patch (tap views start at
//                                   arbitrary 128-byte rows), B = N-channel dY tile (SW128, 64-channel boxes)
//   mode 1: same, A always at an aligned start with an 8-pixel pitch (is the arbitrary row start costly?)
//   mode 2: tail role as shipped    A = one box aliased twice (LBO = 0), B = N = 32
//   mode 3: stacked tail            A = dY box aliased twice (SW128), B = 32-channel SW64 patch, N = 96 =
//                                   three pixel-shifted views (LBO = 64 B: overlapping swizzle atoms)
//   mode 4: K-major SW128 A and B (what the forward kernels use) for comparison
//   mode 5: operands swapped -- A = dY tile (fixed over the 3 taps of a K-step), B = X tap views (N channels);
//           the A collector buffer is filled by the first tap and reused by the next two (.collector::a::use)
//   mode 6: as 5 without collector reuse
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/probe_mn.cu -o build/probe_mn
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

constexpr int kXBox = 13312, kYBox = 8192;

// collector usage of the A operand: 0 fill, 1 use, 2 lastuse
template <int C>
__device__ __forceinline__ void umma_ss_c(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
  if (C == 0)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(1) : "memory");
  else if (C == 1)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16.collector::a::use [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(1) : "memory");
  else
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(1) : "memory");
}

template <int MODE, int N>
__global__ void __launch_bounds__(128, 1) mn_rate_kernel(long long* cycles, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  for (int i = threadIdx.x; i < (3 * kXBox + 4 * kYBox) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_mbar_init();
  }
  ptx::fence_proxy_async();
  if (threadIdx.x < 32) ptx::tmem_alloc<512>(&tmem_ptr);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tb = tmem_ptr;
  if (threadIdx.x < 32 && ptx::elect_one()) {
    const uint32_t x_addr = ptx::smem_u32(smem);
    const uint32_t y_addr = x_addr + 3 * kXBox;
    constexpr int TAPS = (MODE == 2) ? 9 : 3;
    constexpr uint32_t idesc = ptx::umma_idesc_bf16_f32(128, N, MODE == 4 ? 0 : 1, MODE == 4 ? 0 : 1);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll 1
      for (int s = 0; s < 4; ++s) {
        uint64_t adesc, bdesc;
        if (MODE == 0 || MODE == 1 || MODE == 2)
          bdesc = ptx::umma_smem_desc(y_addr + uint32_t(s * 2 * 8 * 128), kYBox, 8 * 128, ptx::UMMA_SW128);
        if (MODE == 3) adesc = ptx::umma_smem_desc(y_addr + uint32_t(s * 2 * 8 * 128), 0, 8 * 128, ptx::UMMA_SW128);
        if (MODE == 4) {
          adesc = ptx::umma_smem_desc(x_addr + uint32_t(s * 32), 16, 1024, ptx::UMMA_SW128);
          bdesc = ptx::umma_smem_desc(y_addr + uint32_t(s * 32), 16, 1024, ptx::UMMA_SW128);
        }
#pragma unroll 1
        for (int t = 0; t < TAPS; ++t) {
          const int dy = (MODE == 2) ? t / 3 : 1, dx = t % 3;
          if (MODE == 0)
            adesc = ptx::umma_smem_desc(x_addr + uint32_t(((2 * s + dy) * 10 + dx) * 128), kXBox, 10 * 128, ptx::UMMA_SW128);
          if (MODE == 1) adesc = ptx::umma_smem_desc(x_addr + uint32_t(s * 2048), kXBox, 8 * 128, ptx::UMMA_SW128);
          if (MODE == 2)
            adesc = ptx::umma_smem_desc(x_addr + uint32_t(((2 * s + dy) * 10 + dx) * 128), 0, 10 * 128, ptx::UMMA_SW128);
          if (MODE == 3) bdesc = ptx::umma_smem_desc(x_addr + uint32_t(((2 * s + t) * 10) * 64), 64, 10 * 64, ptx::UMMA_SW64);
          if (MODE == 5 || MODE == 6) {
            adesc = ptx::umma_smem_desc(y_addr + uint32_t(s * 2 * 8 * 128), kYBox, 8 * 128, ptx::UMMA_SW128);
            bdesc = ptx::umma_smem_desc(x_addr + uint32_t(((2 * s + dy) * 10 + dx) * 128), kXBox, 10 * 128, ptx::UMMA_SW128);
          }
          const uint32_t dcol = tb + uint32_t(N * 3 <= 512 ? t * N : (t & 1) * N);
          if (MODE == 5) {
            if (t == 0) umma_ss_c<0>(dcol, adesc, bdesc, idesc);
            else if (t == 1) umma_ss_c<1>(dcol, adesc, bdesc, idesc);
            else umma_ss_c<2>(dcol, adesc, bdesc, idesc);
          } else {
            ptx::umma_ss(dcol, adesc, bdesc, idesc, 1u);
          }
        }
      }
    }
    ptx::umma_commit(&bar);
    ptx::mbar_wait(&bar, 0);
    long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tb);
  }
}

template <int MODE, int N>
void run(int grid) {
  const int iters = 300;
  constexpr int TAPS = (MODE == 2) ? 9 : 3;
  long long* d;
  CK(cudaMalloc(&d, 148 * sizeof(long long)));
  size_t smem = 1024 + 3 * kXBox + 4 * kYBox;
  CK(cudaFuncSetAttribute(mn_rate_kernel<MODE, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  for (int rep = 0; rep < 2; ++rep) {
    mn_rate_kernel<MODE, N><<<grid, 128, smem>>>(d, iters);
    CK(cudaDeviceSynchronize());
  }
  std::vector<long long> h(grid);
  CK(cudaMemcpy(h.data(), d, grid * sizeof(long long), cudaMemcpyDeviceToHost));
  double mx = 0;
  for (auto v : h) mx = std::max(mx, double(v));
  const double per = mx / (double(iters) * 4 * TAPS);
  printf("mn_rate mode=%d M=128 N=%3d grid=%3d : %.2f cyc/MMA (tensor ideal %.1f)\n", MODE, N, grid, per, N / 2.0);
  CK(cudaFree(d));
}

int main() {
  for (int grid : {1, 148}) {
    run<0, 160>(grid);
    run<1, 160>(grid);
    run<0, 128>(grid);
    run<0, 64>(grid);
    run<5, 128>(grid);
    run<6, 128>(grid);
    run<5, 160>(grid);
    run<6, 160>(grid);
    run<5, 64>(grid);
    run<6, 64>(grid);
    run<2, 32>(grid);
    run<3, 96>(grid);
    run<4, 160>(grid);
    run<4, 96>(grid);
    run<4, 32>(grid);
  }
  return 0;
}

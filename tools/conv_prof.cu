// Developer tool: where does the conv3x3 kernel wait?  Builds the kernel with XMM_CONV_PROFILE (cycle counters
// around every mbarrier wait) and prints per-tile averages for the dense-block shapes.  Stand-alone: does not link
// libxmm_b200 (same template names, different ConvArgs).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo tools/conv_prof.cu -o build/conv_prof
#define XMM_CONV_PROFILE 1
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <type_traits>
#include <vector>

#include "../include/xmm_b200.h"
#include "../xmm_superres_denoise_b200/csrc/conv3x3_dx.cuh"
#include "../xmm_superres_denoise_b200/csrc/host_common.cuh"

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

using namespace xmm;

template <int KC, int NT, bool DX, bool PAIR = false>
static void run(int B, int H, int W, int k, int stages_cap, int rr = 0) {
  using Cfg = typename std::conditional<DX, DxCfg<KC, NT>, ConvCfg<KC, NT, 0>>::type;
  const int tw = DX ? kDxTileW : kTileW, th = DX ? kDxTileH : kTileH;
  const int cin = k * 32, ctot = 160;
  const size_t npix = size_t(B) * H * W;
  __nv_bfloat16 *in_d, *out_d;
  CK(cudaMalloc(&in_d, npix * ctot * 2));
  CK(cudaMemset(in_d, 0, npix * ctot * 2));
  CK(cudaMalloc(&out_d, npix * ctot * 2));
  ConvArgs a{};
  a.nchunks = cin / KC;
  a.w_bytes = uint32_t(a.nchunks) * 9u * Cfg::kTapBytes;
  void* blob;
  CK(cudaMalloc(&blob, a.w_bytes + Cfg::kBiasBytes));
  CK(cudaMemset(blob, 0, a.w_bytes + Cfg::kBiasBytes));
  a.wblob = blob;
  a.cin_off = 0; a.batch = B; a.height = H; a.width = W;
  a.tiles_x = (W + tw - 1) / tw; a.tiles_y = (H + th - 1) / th;
  a.num_tiles = a.tiles_x * a.tiles_y * B;
  const int max_smem = 232448;
  int stages = int((size_t(max_smem) - Cfg::smem_bytes(a.w_bytes, 0)) / Cfg::kStageBytes);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages > stages_cap) stages = stages_cap;
  a.stages = stages;
  a.strip_rr = rr;
  a.epi.lrelu_slope = 0.2f; a.epi.s0 = 1.f; a.epi.out = out_d; a.epi.out_ctot = ctot; a.epi.out_coff = (k % 5) * 32;
  long long* prof;
  CK(cudaMalloc(&prof, 148 * 16 * sizeof(long long)));
  CUtensorMap tmap;
  if (make_nhwc_tmap(&tmap, in_d, B, H, W, ctot, KC, DX ? kDxTileW : kTileW + 2, DX ? kDxPatchH : kHaloH, false) != 0) { printf("tmap failed\n"); exit(2); }
  a.prof = prof;
  CUtensorMap tmap_o = tmap;
  if (DX && make_nhwc_tmap(&tmap_o, out_d, B, H, W, ctot, NT / 2, kDxTileW, 2, false) != 0) { printf("tmap failed\n"); exit(2); }
  if constexpr (DX) CK(cudaFuncSetAttribute(conv3x3_dx_kernel<KC, NT, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
  else CK(cudaFuncSetAttribute(conv3x3_tc_kernel<KC, NT, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaMemset(prof, 0, 148 * 16 * sizeof(long long)));
    cudaEventRecord(e0);
    if constexpr (DX && !PAIR) { DxSideMaps sm_; sm_.m[0] = sm_.m[1] = sm_.m[2] = tmap; conv3x3_dx_kernel<KC, NT><<<148, kDxThreads, Cfg::smem_bytes(a.w_bytes, stages)>>>(tmap, tmap_o, sm_, a); }
    if constexpr (DX && PAIR) {
      DxSideMaps sm_; sm_.m[0] = sm_.m[1] = sm_.m[2] = tmap;
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(148); cfg.blockDim = dim3(kDxThreads); cfg.dynamicSmemBytes = Cfg::smem_bytes(a.w_bytes, stages);
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      if (rep == 0) {
        int ncl = 0;
        CK(cudaOccupancyMaxActiveClusters(&ncl, conv3x3_dx_kernel<KC, NT, true>, &cfg));
        printf("   max active clusters of 2: %d\n", ncl);
      }
      CK(cudaLaunchKernelEx(&cfg, conv3x3_dx_kernel<KC, NT, true>, tmap, tmap_o, sm_, a));
    }
    if constexpr (!DX) conv3x3_tc_kernel<KC, NT, 0><<<148, kConvThreads, Cfg::smem_bytes(a.w_bytes, stages)>>>(tmap, tmap, a);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    cudaEventElapsedTime(&ms, e0, e1);
  }
  std::vector<long long> h(148 * 16);
  CK(cudaMemcpy(h.data(), prof, h.size() * 8, cudaMemcpyDeviceToHost));
  double s[8] = {0};
  for (int c = 0; c < 148; ++c) for (int i = 0; i < 8; ++i) s[i] += double(h[c * 8 + i]) / 148;
  const double tiles = double(a.num_tiles) / 148;
  long long mn = 1ll << 62, mx = 0;
  for (int c = 0; c < 148; ++c) { mn = std::min(mn, h[c * 8 + 3]); mx = std::max(mx, h[c * 8 + 3]); }
  printf("   whole CTA: %.0f cycles in %.0f ns -> SM clock %.3f GHz\n", s[6], s[7], s[6] / s[7]);
  printf("   mma-loop total cycles per CTA: min %lld  avg %.0f  max %lld  (kernel %.0f)\n", mn, s[3], mx, ms * 1e-3 * 1e9 * s[6] / s[7]);
  printf("%s%s k=%d cin=%d stages=%d: %.3f ms (%.0f cyc/tile at the measured clock) | per tile: mma-loop %.0f  wait_full %.0f  wait_tempty %.0f | producer wait_empty %.0f | epi wait_tfull %.0f  epi busy %.0f\n",
         DX ? "dx" : "tc", PAIR ? "-pair" : (rr ? "-rr" : ""), k, cin, stages, ms, ms * 1e-3 * 1e9 * (s[6] / s[7]) / tiles, s[3] / tiles, s[2] / tiles, s[1] / tiles, s[0] / tiles, s[4] / tiles, s[5] / tiles);
  if (DX) {
    double e[8] = {0};
    for (int c = 0; c < 148; ++c) for (int i = 0; i < 8; ++i) e[i] += double(h[(148 + c) * 8 + i]) / 148;
    printf("   epilogue phases per tile (warp 2): wait staging %.0f | tmem ld + release %.0f | mailbox/shuffle/sum %.0f | math + st.shared %.0f | fence + TMA store %.0f\n",
           e[0] / tiles, e[1] / tiles, e[2] / tiles, e[3] / tiles, e[4] / tiles);
  }
  cudaFree(in_d); cudaFree(out_d); cudaFree(blob); cudaFree(prof);
}

int main(int argc, char** argv) {
  const int cap = argc > 1 ? atoi(argv[1]) : 8;
  if (argc > 2) {  // CTA pairs against single CTAs, both with round-robin strips
    for (int k = 2; k <= 5; ++k) {
      run<32, 32, true>(16, 416, 416, k, cap, 1);
      run<32, 32, true, true>(16, 416, 416, k, cap, 1);
    }
    return 0;
  }
  for (int k = 1; k <= 5; ++k) run<32, 32, true>(16, 416, 416, k, cap);
  for (int k = 1; k <= 5; ++k) run<32, 32, false>(16, 416, 416, k, cap);
  return 0;
}

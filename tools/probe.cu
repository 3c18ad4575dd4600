// GPU probe (developer tool, not part of the product path):
//   1. tcgen05.mma issue rate vs N for SS (A in shared memory) and TS (A in TMEM) operands;
//   2. correctness of xmm_conv3x3_bf16 for every shared-memory tap layout against a CPU loop;
//   3. timing of the dense-block convolution shapes.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/probe.cu \
//         -Lxmm_superres_denoise_b200 -lxmm_b200 -Xlinker -rpath,'$ORIGIN/../xmm_superres_denoise_b200' -o build/probe
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../include/xmm_b200.h"
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

// ------------------------------------------------------------------ 1. UMMA rate
// LAYOUT 0: SW128 K-major canonical tile.  1: SW64 canonical (8-row groups 512 B apart).  2: SW64 haloed
// [18][10]-pixel patch, 9 tap views (SBO = 10 rows) x 2 K-steps -- exactly the conv kernel's A operand stream.
template <int N, bool TS, int KSTEPS, int LAYOUT = 0>
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(long long* cycles, int iters, int nacc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  // zero operands (values do not matter for the rate)
  for (int i = threadIdx.x; i < (16384 + N * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
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
    const uint32_t a_addr = ptx::smem_u32(smem);
    const uint32_t b_addr = a_addr + 16384;
    const uint64_t adesc = LAYOUT == 0   ? ptx::umma_smem_desc(a_addr, 16, 1024, ptx::UMMA_SW128)
                           : LAYOUT == 1 ? ptx::umma_smem_desc(a_addr, 16, 512, ptx::UMMA_SW64)
                                         : ptx::umma_smem_desc(0, 16, 640, ptx::UMMA_SW64);
    const uint64_t bdesc = LAYOUT == 0 ? ptx::umma_smem_desc(b_addr, 16, 1024, ptx::UMMA_SW128)
                                       : ptx::umma_smem_desc(b_addr, 16, 512, ptx::UMMA_SW64);
    constexpr uint32_t idesc = ptx::umma_idesc_bf16_f32(128, N, 0, 0);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = tb + uint32_t((it % nacc) * N);
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ++ks) {
        if (TS) {
          ptx::umma_ts(d, tb + 448u + uint32_t((ks & 3) * 8), bdesc + uint64_t(((ks & 3) * 32) >> 4), idesc, 1u);
        } else if (LAYOUT == 0) {
          ptx::umma_ss(d, adesc + uint64_t(((ks & 3) * 32) >> 4), bdesc + uint64_t(((ks & 3) * 32) >> 4), idesc, 1u);
        } else if (LAYOUT == 1) {
          ptx::umma_ss(d, adesc + uint64_t(((ks & 1) * 32) >> 4), bdesc + uint64_t(((ks & 1) * 32) >> 4), idesc, 1u);
        } else {
          const int tap = (ks >> 1) % 9, dy = tap / 3, dx = tap % 3;
          const uint32_t a = a_addr + uint32_t((dy * 10 + dx) * 64 + (ks & 1) * 32);
          ptx::umma_ss(d, adesc | uint64_t((a & 0x3FFFFu) >> 4), bdesc + uint64_t(((ks & 1) * 32) >> 4), idesc, 1u);
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

template <int N, bool TS, int LAYOUT = 0>
void run_rate(int nacc) {
  constexpr int KSTEPS = 36;  // one conv K-chunk worth of MMAs between loop overheads
  const int iters = 400;
  long long* d;
  CK(cudaMalloc(&d, 148 * sizeof(long long)));
  size_t smem = 1024 + 16384 + 256 * 128;
  CK(cudaFuncSetAttribute(umma_rate_kernel<N, TS, KSTEPS, LAYOUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  for (int rep = 0; rep < 2; ++rep) {
    umma_rate_kernel<N, TS, KSTEPS, LAYOUT><<<148, 128, smem>>>(d, iters, nacc);
    CK(cudaDeviceSynchronize());
  }
  std::vector<long long> h(148);
  CK(cudaMemcpy(h.data(), d, 148 * sizeof(long long), cudaMemcpyDeviceToHost));
  double mx = 0, mn = 1e30;
  for (auto v : h) {
    mx = std::max(mx, double(v));
    mn = std::min(mn, double(v));
  }
  const double per = mx / (double(iters) * KSTEPS);
  printf("umma_rate layout=%d M=128 N=%3d %s nacc=%d : %.2f cyc/MMA (ideal %.1f) -> %.1f%% of tensor peak  [min-cta %.2f]\n", LAYOUT, N,
         TS ? "TS" : "SS", nacc, per, N / 2.0, 100.0 * (N / 2.0) / per, mn / (double(iters) * KSTEPS));
  CK(cudaFree(d));
}

// ------------------------------------------------------------------ 2/3. conv
static float bf(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

struct ConvCase {
  int B, H, W, in_ctot, in_coff, cin, kc, cout, out_ctot, out_coff;
  bool bias, mask, r1, r2, shuffle;
  float slope;
};

static std::vector<__nv_bfloat16> rand_bf16(size_t n, std::mt19937& g, float lo, float hi) {
  std::uniform_real_distribution<float> d(lo, hi);
  std::vector<__nv_bfloat16> v(n);
  for (auto& x : v) x = __float2bfloat16_rn(d(g));
  return v;
}

template <class T>
T* to_dev(const std::vector<T>& h) {
  T* d;
  CK(cudaMalloc(&d, h.size() * sizeof(T)));
  CK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}

// returns max abs error / max abs reference
static double run_conv_case(const ConvCase& c, int tap_mode, bool verbose) {
  std::mt19937 g(1234);
  const size_t npix = size_t(c.B) * c.H * c.W;
  auto in_h = rand_bf16(npix * c.in_ctot, g, -1.f, 1.f);
  std::vector<float> w_h(size_t(c.cout) * c.cin * 9), b_h(c.cout);
  {
    std::uniform_real_distribution<float> d(-0.1f, 0.1f);
    for (auto& x : w_h) x = d(g);
    for (auto& x : b_h) x = c.bias ? d(g) * 5.f : 0.f;
  }
  const int out_c = c.shuffle ? c.cout / 4 : c.cout;
  const size_t opix = c.shuffle ? npix * 4 : npix;
  auto mask_h = rand_bf16(npix * c.cout, g, -1.f, 1.f);
  auto r1_h = rand_bf16(npix * c.cout, g, -1.f, 1.f);
  auto r2_h = rand_bf16(npix * c.cout, g, -1.f, 1.f);
  std::vector<__nv_bfloat16> out_h(opix * c.out_ctot, __float2bfloat16_rn(-7.f));

  __nv_bfloat16* in_d = to_dev(in_h);
  float* w_d = to_dev(w_h);
  float* b_d = to_dev(b_h);
  __nv_bfloat16* mask_d = to_dev(mask_h);
  __nv_bfloat16* r1_d = to_dev(r1_h);
  __nv_bfloat16* r2_d = to_dev(r2_h);
  __nv_bfloat16* out_d = to_dev(out_h);

  const int nchunks = c.cin / c.kc;
  void* blob;
  CK(cudaMalloc(&blob, xmm_pack_blob_bytes(c.cout, c.kc, nchunks)));
  xmm_pack_job job{};
  job.dst = blob;
  job.bias = c.bias ? b_d : nullptr;
  job.nt = c.cout;
  job.kc = c.kc;
  job.nchunks = nchunks;
  job.nseg = 1;
  job.perm = c.shuffle ? 1 : 0;
  job.n_valid = c.cout;
  job.bias_n = c.cout;
  job.seg[0] = xmm_pack_segment{w_d, c.cin, 0, 0, 0, 0, c.cin, 1.0f, 0, c.cout};
  xmm_pack_job* job_d;
  CK(cudaMalloc(&job_d, sizeof(job)));
  CK(cudaMemcpy(job_d, &job, sizeof(job), cudaMemcpyHostToDevice));
  if (xmm_pack_weights(job_d, 1, nullptr) != 0) {
    printf("pack failed: %s\n", xmm_last_error());
    exit(2);
  }
  void* blob_row = nullptr;  // the same layer in the row-hop kernel's block order
  if (!c.shuffle && c.cout == c.kc) {
    CK(cudaMalloc(&blob_row, xmm_pack_blob_bytes(c.cout, c.kc, nchunks)));
    xmm_pack_job job2 = job;
    job2.dst = blob_row;
    job2.tap_order = 1;
    CK(cudaMemcpy(job_d, &job2, sizeof(job2), cudaMemcpyHostToDevice));
    if (xmm_pack_weights(job_d, 1, nullptr) != 0) { printf("pack (row) failed: %s\n", xmm_last_error()); exit(2); }
  }

  xmm_conv3x3_params p{};
  p.in = in_d; p.in_ctot = c.in_ctot; p.in_coff = c.in_coff; p.cin = c.cin;
  p.wblob = blob; p.kc = c.kc; p.cout = c.cout;
  p.batch = c.B; p.height = c.H; p.width = c.W;
  p.lrelu_slope = c.slope;
  p.mask = c.mask ? mask_d : nullptr; p.mask_ctot = c.cout; p.mask_coff = 0; p.mask_slope = 0.2f;
  p.s0 = c.r1 ? 0.2f : 1.0f;
  p.r1 = c.r1 ? r1_d : nullptr; p.r1_ctot = c.cout; p.r1_coff = 0; p.s1 = 1.0f;
  p.r2 = c.r2 ? r2_d : nullptr; p.r2_ctot = c.cout; p.r2_coff = 0; p.s2 = 0.5f;
  p.out = out_d; p.out_ctot = c.out_ctot; p.out_coff = c.out_coff;
  p.pixel_shuffle = c.shuffle ? 1 : 0;
  p.tap_mode = tap_mode;
  p.wblob_row = blob_row;
  int rc = xmm_conv3x3_bf16(&p, nullptr);
  if (rc != 0) {
    printf("conv launch failed rc=%d: %s\n", rc, xmm_last_error());
    return 1e9;
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("conv kernel failed: %s\n", cudaGetErrorString(e));
    exit(3);
  }
  CK(cudaMemcpy(out_h.data(), out_d, out_h.size() * 2, cudaMemcpyDeviceToHost));

  // CPU reference
  double max_err = 0, max_ref = 0;
  size_t bad_untouched = 0;
  std::vector<float> wq(w_h.size());
  for (size_t i = 0; i < w_h.size(); ++i) wq[i] = bf(w_h[i]);
  for (int b = 0; b < c.B; ++b)
    for (int y = 0; y < c.H; ++y)
      for (int x = 0; x < c.W; ++x) {
        const size_t pix = (size_t(b) * c.H + y) * c.W + x;
        for (int n = 0; n < c.cout; ++n) {
          double acc = 0;
          for (int dy = 0; dy < 3; ++dy) {
            const int yy = y + dy - 1;
            if (yy < 0 || yy >= c.H) continue;
            for (int dx = 0; dx < 3; ++dx) {
              const int xx = x + dx - 1;
              if (xx < 0 || xx >= c.W) continue;
              const __nv_bfloat16* ip = &in_h[((size_t(b) * c.H + yy) * c.W + xx) * c.in_ctot + c.in_coff];
              const float* wp = &wq[size_t(n) * c.cin * 9 + dy * 3 + dx];
              for (int k = 0; k < c.cin; ++k) acc += double(__bfloat162float(ip[k])) * wp[size_t(k) * 9];
            }
          }
          float v = float(acc) + b_h[n];
          v = v > 0 ? v : v * c.slope;
          if (c.mask) v *= (__bfloat162float(mask_h[pix * c.cout + n]) > 0 ? 1.f : 0.2f);
          v *= (c.r1 ? 0.2f : 1.0f);
          if (c.r1) v += __bfloat162float(r1_h[pix * c.cout + n]);
          if (c.r2) v += 0.5f * __bfloat162float(r2_h[pix * c.cout + n]);
          float got;
          if (c.shuffle) {
            const int cc = n / 4, gq = n % 4;
            const size_t hp = (size_t(b) * 2 * c.H + 2 * y + (gq >> 1)) * (2 * c.W) + 2 * x + (gq & 1);
            got = __bfloat162float(out_h[hp * c.out_ctot + c.out_coff + cc]);
          } else {
            got = __bfloat162float(out_h[pix * c.out_ctot + c.out_coff + n]);
          }
          max_err = std::max(max_err, double(std::fabs(got - v)));
          max_ref = std::max(max_ref, double(std::fabs(v)));
        }
        // channels outside the output window must be untouched
        if (!c.shuffle)
          for (int ch = 0; ch < c.out_ctot; ++ch)
            if ((ch < c.out_coff || ch >= c.out_coff + out_c) && __bfloat162float(out_h[pix * c.out_ctot + ch]) != -7.f)
              ++bad_untouched;
      }
  if (verbose)
    printf("  conv B%d %dx%d cin=%d(@%d/%d) kc=%d cout=%d mode=%d bias%d mask%d r1%d r2%d shuf%d : max_err=%.4g max_ref=%.4g rel=%.3g untouched_bad=%zu %s\n",
           c.B, c.H, c.W, c.cin, c.in_coff, c.in_ctot, c.kc, c.cout, tap_mode, c.bias, c.mask, c.r1, c.r2, c.shuffle,
           max_err, max_ref, max_err / max_ref, bad_untouched, (max_err / max_ref < 2e-2 && bad_untouched == 0) ? "OK" : "FAIL");
  cudaFree(in_d); cudaFree(w_d); cudaFree(b_d); cudaFree(mask_d); cudaFree(r1_d); cudaFree(r2_d); cudaFree(out_d);
  cudaFree(blob); cudaFree(job_d); if (blob_row) cudaFree(blob_row);
  return bad_untouched ? 1e9 : max_err / max_ref;
}

static void time_conv(int B, int H, int W, int F, int k, int kc, int tap_mode, int cout_override = 0, bool shuffle = false) {
  const int cin = k * F, ctot = 5 * F;
  const int cout = cout_override ? cout_override : F;
  const size_t npix = size_t(B) * H * W;
  __nv_bfloat16 *in_d, *out_d;
  CK(cudaMalloc(&in_d, npix * ctot * 2));
  CK(cudaMemset(in_d, 0, npix * ctot * 2));
  const size_t out_elems = shuffle ? npix * 4 * F : npix * ctot;
  if (shuffle || k == 5) {
    CK(cudaMalloc(&out_d, out_elems * 2));
  } else {
    out_d = in_d;
  }
  std::vector<float> w_h(size_t(cout) * cin * 9, 0.01f);
  float* w_d = to_dev(w_h);
  const int nchunks = cin / kc;
  void* blob;
  CK(cudaMalloc(&blob, xmm_pack_blob_bytes(cout, kc, nchunks)));
  xmm_pack_job job{};
  job.dst = blob; job.nt = cout; job.kc = kc; job.nchunks = nchunks; job.nseg = 1; job.n_valid = cout;
  job.perm = shuffle;
  job.bias_n = cout;
  job.seg[0] = xmm_pack_segment{w_d, cin, 0, 0, 0, 0, cin, 1.0f, 0, cout};
  xmm_pack_job* job_d;
  CK(cudaMalloc(&job_d, sizeof(job)));
  CK(cudaMemcpy(job_d, &job, sizeof(job), cudaMemcpyHostToDevice));
  xmm_pack_weights(job_d, 1, nullptr);
  void* blob_row = nullptr;
  if (!shuffle && cout == kc) {
    CK(cudaMalloc(&blob_row, xmm_pack_blob_bytes(cout, kc, nchunks)));
    job.dst = blob_row; job.tap_order = 1;
    CK(cudaMemcpy(job_d, &job, sizeof(job), cudaMemcpyHostToDevice));
    xmm_pack_weights(job_d, 1, nullptr);
  }
  xmm_conv3x3_params p{};
  p.wblob_row = blob_row;
  p.in = in_d; p.in_ctot = ctot; p.in_coff = 0; p.cin = cin; p.wblob = blob; p.kc = kc; p.cout = cout;
  p.batch = B; p.height = H; p.width = W; p.lrelu_slope = 0.2f; p.s0 = 1.f;
  p.out = out_d; p.out_ctot = shuffle ? F : ctot; p.out_coff = (shuffle || k == 5) ? 0 : k * F;
  p.pixel_shuffle = shuffle; p.tap_mode = tap_mode;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    int rc = xmm_conv3x3_bf16(&p, nullptr);
    cudaEventRecord(e1);
    if (rc) { printf("time_conv launch failed: %s\n", xmm_last_error()); break; }
    CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0) best = std::min(best, ms);
  }
  const double flop = 2.0 * 9 * cin * cout * double(npix);
  printf("  time conv B%d %dx%d F=%d k=%d cin=%d cout=%d kc=%d mode=%d : %.3f ms  %.1f TFLOP/s (%.1f%% of 1667.8)\n", B, H,
         W, F, k, cin, cout, kc, tap_mode, best, flop / best * 1e-9, 100.0 * flop / best * 1e-9 / 1667.8);
  cudaFree(in_d); if (out_d != in_d) cudaFree(out_d); cudaFree(w_d); cudaFree(blob); cudaFree(job_d);
  if (blob_row) cudaFree(blob_row);
}

// ------------------------------------------------------------------ dense-block chain: pipelined vs layer by layer
// Forward dense block (rrdb_blocks.py:37-54): conv_k reads slots [0,k) of buf and writes slot k; conv5 writes slot 0
// of `nxt` with the block residual.  Both runs use the column-scatter kernel (tap_mode 4), so the pipelined launch
// must reproduce the layer-by-layer result BIT FOR BIT.
static void run_chain_case(int B, int H, int W, bool time_it) {
  const int F = 32, ctot = 5 * F;
  std::mt19937 g(99);
  const size_t npix = size_t(B) * H * W;
  std::vector<__nv_bfloat16> buf_h(npix * ctot, __float2bfloat16_rn(-7.f));
  {
    std::uniform_real_distribution<float> d(-1.f, 1.f);
    for (size_t p = 0; p < npix; ++p)
      for (int c = 0; c < F; ++c) buf_h[p * ctot + c] = __float2bfloat16_rn(d(g));
  }
  __nv_bfloat16* buf_d[2] = {to_dev(buf_h), to_dev(buf_h)};
  __nv_bfloat16* nxt_d[2];
  for (int i = 0; i < 2; ++i) { CK(cudaMalloc(&nxt_d[i], npix * ctot * 2)); CK(cudaMemset(nxt_d[i], 0, npix * ctot * 2)); }
  void* blobs[5];
  std::vector<void*> frees;
  for (int k = 1; k <= 5; ++k) {
    const int cin = k * F;
    std::vector<float> w_h(size_t(F) * cin * 9), b_h(F);
    std::uniform_real_distribution<float> d(-0.06f, 0.06f);
    for (auto& x : w_h) x = d(g);
    for (auto& x : b_h) x = d(g);
    float* w_d = to_dev(w_h);
    float* b_d = to_dev(b_h);
    CK(cudaMalloc(&blobs[k - 1], xmm_pack_blob_bytes(F, 32, cin / 32)));
    xmm_pack_job job{};
    job.dst = blobs[k - 1]; job.bias = b_d; job.nt = F; job.kc = 32; job.nchunks = cin / 32; job.nseg = 1;
    job.n_valid = F; job.bias_n = F;
    job.seg[0] = xmm_pack_segment{w_d, cin, 0, 0, 0, 0, cin, 1.0f, 0, F};
    xmm_pack_job* job_d;
    CK(cudaMalloc(&job_d, sizeof(job)));
    CK(cudaMemcpy(job_d, &job, sizeof(job), cudaMemcpyHostToDevice));
    if (xmm_pack_weights(job_d, 1, nullptr) != 0) { printf("pack failed: %s\n", xmm_last_error()); exit(2); }
    frees.push_back(w_d); frees.push_back(b_d); frees.push_back(job_d);
  }
  auto make = [&](int which, xmm_conv3x3_params* L) {
    for (int k = 1; k <= 5; ++k) {
      xmm_conv3x3_params p{};
      p.in = buf_d[which]; p.in_ctot = ctot; p.in_coff = 0; p.cin = k * F; p.wblob = blobs[k - 1]; p.kc = 32; p.cout = F;
      p.batch = B; p.height = H; p.width = W; p.tap_mode = 4;
      if (k < 5) {
        p.lrelu_slope = 0.2f; p.s0 = 1.f; p.out = buf_d[which]; p.out_ctot = ctot; p.out_coff = k * F;
      } else {
        p.lrelu_slope = 1.f; p.s0 = 0.2f; p.r1 = buf_d[which]; p.r1_ctot = ctot; p.r1_coff = 0; p.s1 = 1.f;
        p.out = nxt_d[which]; p.out_ctot = ctot; p.out_coff = 0;
      }
      L[k - 1] = p;
    }
  };
  xmm_conv3x3_params La[5], Lb[5];
  make(0, La);
  make(1, Lb);
  const size_t ws_bytes = xmm_conv3x3_chain_workspace_bytes(5, B, H);
  void* ws;
  CK(cudaMalloc(&ws, ws_bytes));
  int rc = xmm_conv3x3_chain_bf16(La, 5, 2, ws, ws_bytes, nullptr);
  if (rc) { printf("chain (layer by layer) failed: %s\n", xmm_last_error()); exit(2); }
  rc = xmm_conv3x3_chain_bf16(Lb, 5, 1, ws, ws_bytes, nullptr);
  if (rc) { printf("chain (pipelined) failed: %s\n", xmm_last_error()); exit(2); }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("chain kernel failed: %s\n", cudaGetErrorString(e)); exit(3); }
  std::vector<__nv_bfloat16> a(npix * ctot), b(npix * ctot), na(npix * ctot), nb(npix * ctot);
  CK(cudaMemcpy(a.data(), buf_d[0], a.size() * 2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(b.data(), buf_d[1], b.size() * 2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(na.data(), nxt_d[0], a.size() * 2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(nb.data(), nxt_d[1], b.size() * 2, cudaMemcpyDeviceToHost));
  size_t diff = 0, diff_n = 0, untouched = 0;
  double amax = 0;
  for (size_t i = 0; i < a.size(); ++i) {
    if (memcmp(&a[i], &b[i], 2) != 0) ++diff;
    if (memcmp(&na[i], &nb[i], 2) != 0) ++diff_n;
    if (__bfloat162float(b[i]) == -7.f) ++untouched;
    amax = std::max(amax, double(std::fabs(__bfloat162float(nb[i]))));
  }
  printf("  chain B%d %dx%d: pipelined vs layer-by-layer: %zu / %zu differing values in the block buffer, %zu in the output (max |out| %.3g, unwritten %zu) %s\n",
         B, H, W, diff, a.size(), diff_n, amax, untouched, (diff == 0 && diff_n == 0 && untouched == 0 && amax > 0) ? "OK" : "FAIL");
  if (time_it) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 2; mode >= 1; --mode) {
      float best = 1e30f;
      for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        rc = xmm_conv3x3_chain_bf16(Lb, 5, mode, ws, ws_bytes, nullptr);
        cudaEventRecord(e1);
        if (rc) { printf("chain launch failed: %s\n", xmm_last_error()); break; }
        CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0) best = std::min(best, ms);
      }
      const double flop = 2.0 * 9 * F * F * 15 * double(npix);
      printf("  time chain B%d %dx%d mode=%d (%s): %.3f ms  %.1f TFLOP/s (%.1f%% of 1667.8)\n", B, H, W, mode,
             mode == 1 ? "pipelined" : "layer by layer", best, flop / best * 1e-9, 100.0 * flop / best * 1e-9 / 1667.8);
    }
  }
  for (int i = 0; i < 2; ++i) { cudaFree(buf_d[i]); cudaFree(nxt_d[i]); }
  for (void* q : frees) cudaFree(q);
  for (void* q : blobs) cudaFree(q);
  cudaFree(ws);
}

int main(int argc, char** argv) {
  const bool do_rate = argc < 2 || strstr(argv[1], "rate");
  const bool do_conv = argc < 2 || strstr(argv[1], "conv");
  const bool do_time = argc < 2 || strstr(argv[1], "time");
  const bool do_chain = argc < 2 || strstr(argv[1], "chain");
  if (xmm_check_device() != 0) { printf("device check failed: %s\n", xmm_last_error()); return 1; }
  if (do_rate) {
    run_rate<32, false>(1); run_rate<32, false>(2); run_rate<64, false>(1); run_rate<64, false>(2);
    run_rate<96, false>(2); run_rate<128, false>(2); run_rate<256, false>(2);
    run_rate<32, true>(2); run_rate<64, true>(2); run_rate<128, true>(2);
    run_rate<32, false, 1>(2); run_rate<32, false, 2>(2); run_rate<64, false, 1>(2); run_rate<64, false, 2>(2);
    run_rate<96, false, 1>(2); run_rate<96, false, 2>(2); run_rate<96, false, 2>(1);
  }
  int good_mode[2] = {0, 0};  // per kc
  if (do_conv) {
    for (int mode : {4, 1, 0}) {
      printf("tap_mode %d (4 = column-scatter, 1 = haloed tap views, 0 = auto)\n", mode);
      // plain conv, interior + ragged borders, channel windows
      ConvCase a{2, 40, 24, 160, 32, 96, 32, 32, 160, 128, false, false, false, false, false, 1.0f};
      double ea = run_conv_case(a, mode, true);
      ConvCase b{1, 32, 16, 320, 64, 128, 64, 64, 320, 256, true, false, false, false, false, 0.2f};
      double eb = run_conv_case(b, mode, true);
      (void)ea; (void)eb;
      ConvCase c{3, 416, 416, 160, 0, 160, 32, 32, 160, 0, true, false, false, false, false, 0.2f};
      if (mode == 4) run_conv_case(c, mode, true);  // full-size image: every CTA range starts/ends mid-strip
      ConvCase d{2, 37, 50, 160, 32, 64, 32, 32, 160, 96, true, false, false, false, false, 0.2f};
      run_conv_case(d, mode, true);  // ragged strip (37 = 4*8 + 5) and ragged last tile (50 = 3*16 + 2)
      ConvCase s1{1, 8, 16, 160, 0, 32, 32, 32, 160, 32, true, false, false, false, false, 0.2f};
      run_conv_case(s1, mode, true);  // exactly one tile: strip-end flush only
      ConvCase s2{1, 5, 64, 160, 0, 96, 32, 32, 160, 96, true, false, false, false, false, 0.2f};
      run_conv_case(s2, mode, true);  // one strip of 4 tiles on 4 CTAs: every range but the first has a pre-tile
    }
    const int m32 = 0, m64 = 0;
    // epilogue features
    ConvCase e1{2, 33, 19, 160, 0, 160, 32, 32, 32, 0, true, false, true, true, false, 1.0f};
    run_conv_case(e1, m32, true);
    ConvCase e2{2, 33, 19, 160, 32, 128, 32, 32, 160, 0, false, true, true, false, false, 1.0f};
    run_conv_case(e2, m32, true);
    ConvCase e3{1, 24, 24, 32, 0, 32, 32, 128, 32, 0, true, false, false, false, true, 0.01f};
    run_conv_case(e3, m32, true);
    ConvCase e4{1, 16, 16, 64, 0, 64, 64, 256, 64, 0, true, false, false, false, true, 0.01f};
    run_conv_case(e4, m64, true);
    ConvCase e5{3, 48, 40, 64, 0, 64, 64, 64, 64, 0, true, false, true, false, false, 0.2f};
    run_conv_case(e5, m64, true);
  } else {
    good_mode[0] = good_mode[1] = 1;
  }
  if (argc >= 2 && strstr(argv[1], "row")) {  // row-hop form (tap_mode 9): correctness, then timing against 4 / 1
    const int m = 9;
    ConvCase r0{1, 16, 8, 32, 0, 32, 32, 32, 32, 0, true, false, false, false, false, 0.2f};
    run_conv_case(r0, m, true);   // one column, one row per band
    ConvCase r1{2, 48, 40, 160, 32, 96, 32, 32, 160, 128, true, false, false, false, false, 0.2f};
    run_conv_case(r1, m, true);   // channel windows, 3 chunks
    ConvCase r2{3, 96, 21, 160, 0, 160, 32, 32, 160, 0, true, false, false, false, false, 1.0f};
    run_conv_case(r2, m, true);   // ragged width, 5 chunks
    ConvCase r3{1, 40, 24, 64, 0, 64, 32, 32, 32, 0, true, false, true, true, false, 1.0f};
    run_conv_case(r3, m, true);   // 10 bands; two residuals
    ConvCase r4{2, 80, 19, 160, 32, 128, 32, 32, 160, 0, false, true, false, false, false, 1.0f};
    run_conv_case(r4, m, true);   // LeakyReLU' mask
    ConvCase r5{2, 64, 32, 64, 0, 64, 64, 64, 64, 0, true, false, true, false, false, 1.0f};
    run_conv_case(r5, m, true);   // 64 filters (SWIZZLE_128B, 8 slots)
    ConvCase r6{3, 416, 832, 32, 0, 32, 32, 32, 160, 0, true, false, false, false, false, 0.2f};
    run_conv_case(r6, m, true);   // 312 columns: round-robin rounds + tail ranges
    ConvCase r7{1, 416, 416, 64, 0, 64, 32, 32, 64, 32, true, false, false, false, false, 0.2f};
    run_conv_case(r7, m, true);   // batch 1: every CTA a partial column
    for (int B : {16, 64}) {
      for (int mode : {9, 4, 1}) {
        for (int k = 1; k <= 5; ++k) time_conv(B, 416, 416, 32, k, 32, mode);
      }
    }
    for (int mode : {9, 1})
      for (int k = 1; k <= 2; ++k) time_conv(8, 416, 416, 64, k, 64, mode);
    time_conv(1, 416, 416, 32, 2, 32, 9); time_conv(1, 416, 416, 32, 2, 32, 4);
    time_conv(16, 832, 832, 32, 1, 32, 9); time_conv(16, 832, 832, 32, 1, 32, 1);
    return 0;
  }
  if (argc >= 2 && strstr(argv[1], "sweep")) {  // timing only (env sweeps of XMM_CHAIN_SPLIT / XMM_CHAIN_SEGS)
    run_chain_case(argc > 2 ? atoi(argv[2]) : 16, 416, 416, true);
    return 0;
  }
  if (do_chain) {
    run_chain_case(2, 40, 50, false);
    run_chain_case(1, 8, 16, false);
    run_chain_case(3, 100, 416, false);
    run_chain_case(16, 416, 416, true);
  }
  if (do_time) {
    for (int mode : {4, 1}) {
      for (int k = 1; k <= 5; ++k) time_conv(16, 416, 416, 32, k, 32, mode);
      if (mode != 4) time_conv(16, 416, 416, 32, 1, 32, mode, 128, true);
      for (int k = 1; k <= 2; ++k) time_conv(8, 416, 416, 64, k, 64, mode);
    }
  }
  return 0;
}

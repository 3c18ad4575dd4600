// HBM-bound loss kernels: L1 / Poisson-NLL / MSE reductions with their gradients, and the
// (MS-)SSIM forward + backward (utils/loss_functions.py:11-47, metrics/metrics.py:30-39 and the
// torchmetrics terms restated in oracle/rrdb_oracle.py).
//
// SSIM facts the kernels rely on (torchmetrics functional/image/ssim.py, sigma = 2.5):
//  * the Gaussian window has 19 taps (2*int(3.5*sigma+0.5)+1), not kernel_size=13;
//  * the image is reflect-padded by 9 and the SSIM map cropped by 9 again, so every value that
//    survives the crop is a VALID 19x19 correlation of the un-padded image -- no padding code;
//  * data_range = max(p.max-p.min, t.max-t.min) is recomputed per scale from the live tensors and
//    carries gradient into the arg-max / arg-min pixels of `preds` when preds' range is the larger;
//  * variances are clamped at 0, the per-image means pass through relu, scales are 2x2 avg-pools.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace xmm {

constexpr int kSsimTaps = 19;
constexpr int kSsimPad = 9;
constexpr int kSsimTileW = 32, kSsimTileH = 16;      // output tile (static shared memory stays < 48 KB)
constexpr int kSsimInW = kSsimTileW + 2 * kSsimPad;   // 50
constexpr int kSsimInH = kSsimTileH + 2 * kSsimPad;   // 34
constexpr int kMaxScales = 5;

// ---- per-scale device record (floats), written by range_finalize_kernel / ssim kernels ----
struct ScaleStats {
  float minp, maxp, mint, maxt;  // ranges of preds / target at this scale
  float nmaxp, nminp;            // number of preds pixels equal to max / min (ties share the gradient)
  float dr, c1, c2;              // data range and the two SSIM constants
  float use_p;                   // 1: data range came from preds (gradient flows), 0: from target
  float d_dr;                    // dL/d(data_range) at this scale (filled by msssim_finalize_kernel)
  float pad;
};

struct Window {
  float w[kSsimTaps];
};

__device__ __forceinline__ float warp_sum(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ------------------------------------------------------------------ elementwise reductions
// partial[block][8] = { sum|p-t|, sum(p - t*log(p+1e-8)), sum (p-t)^2, min_p, max_p, min_t, max_t, - }
__global__ void __launch_bounds__(256) loss_reduce_kernel(const float* __restrict__ p, const float* __restrict__ t,
                                                          size_t n, float* __restrict__ partial) {
  float s_abs = 0.f, s_poi = 0.f, s_sq = 0.f, mnp = INFINITY, mxp = -INFINITY, mnt = INFINITY, mxt = -INFINITY;
  const size_t stride = size_t(gridDim.x) * blockDim.x * 4;
  for (size_t i = (size_t(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    float pv[4], tv[4];
    int cnt = 4;
    if (i + 4 <= n) {
      const float4 a = *reinterpret_cast<const float4*>(p + i);
      const float4 b = *reinterpret_cast<const float4*>(t + i);
      pv[0] = a.x; pv[1] = a.y; pv[2] = a.z; pv[3] = a.w;
      tv[0] = b.x; tv[1] = b.y; tv[2] = b.z; tv[3] = b.w;
    } else {
      cnt = int(n - i);
      for (int j = 0; j < cnt; ++j) { pv[j] = p[i + j]; tv[j] = t[i + j]; }
    }
    for (int j = 0; j < cnt; ++j) {
      const float d = pv[j] - tv[j];
      s_abs += fabsf(d);
      s_poi += pv[j] - tv[j] * logf(pv[j] + 1e-8f);  // F.poisson_nll_loss(log_input=False, eps=1e-8)
      s_sq = fmaf(d, d, s_sq);
      mnp = fminf(mnp, pv[j]); mxp = fmaxf(mxp, pv[j]);
      mnt = fminf(mnt, tv[j]); mxt = fmaxf(mxt, tv[j]);
    }
  }
  __shared__ float red[8][7];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  s_abs = warp_sum(s_abs); s_poi = warp_sum(s_poi); s_sq = warp_sum(s_sq);
  mnp = warp_min(mnp); mxp = warp_max(mxp); mnt = warp_min(mnt); mxt = warp_max(mxt);
  if (lane == 0) {
    red[warp][0] = s_abs; red[warp][1] = s_poi; red[warp][2] = s_sq;
    red[warp][3] = mnp; red[warp][4] = mxp; red[warp][5] = mnt; red[warp][6] = mxt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float o[7] = {0.f, 0.f, 0.f, INFINITY, -INFINITY, INFINITY, -INFINITY};
    for (int w = 0; w < 8; ++w) {
      o[0] += red[w][0]; o[1] += red[w][1]; o[2] += red[w][2];
      o[3] = fminf(o[3], red[w][3]); o[4] = fmaxf(o[4], red[w][4]);
      o[5] = fminf(o[5], red[w][5]); o[6] = fmaxf(o[6], red[w][6]);
    }
    for (int k = 0; k < 7; ++k) partial[size_t(blockIdx.x) * 8 + k] = o[k];
  }
}

// sums[0..6] = fixed-order reduction of the block partials (deterministic); also fills the range
// part of a ScaleStats record when `st` != nullptr.
__global__ void loss_reduce_finalize_kernel(const float* __restrict__ partial, int nblocks, float* __restrict__ sums,
                                            ScaleStats* __restrict__ st) {
  if (blockIdx.x != 0 || threadIdx.x >= 32) return;
  // one warp: lane l takes partials l, l + 32, ... in order, then a fixed butterfly -- deterministic, and ~30x
  // shorter than one thread walking all partials (65..90 us per call, five calls per MS-SSIM step)
  const int lane = threadIdx.x;
  double a = 0, b = 0, c = 0;
  float mnp = INFINITY, mxp = -INFINITY, mnt = INFINITY, mxt = -INFINITY;
  for (int i = lane; i < nblocks; i += 32) {
    const float* q = partial + size_t(i) * 8;
    a += q[0]; b += q[1]; c += q[2];
    mnp = fminf(mnp, q[3]); mxp = fmaxf(mxp, q[4]); mnt = fminf(mnt, q[5]); mxt = fmaxf(mxt, q[6]);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, d);
    b += __shfl_xor_sync(0xffffffffu, b, d);
    c += __shfl_xor_sync(0xffffffffu, c, d);
    mnp = fminf(mnp, __shfl_xor_sync(0xffffffffu, mnp, d));
    mxp = fmaxf(mxp, __shfl_xor_sync(0xffffffffu, mxp, d));
    mnt = fminf(mnt, __shfl_xor_sync(0xffffffffu, mnt, d));
    mxt = fmaxf(mxt, __shfl_xor_sync(0xffffffffu, mxt, d));
  }
  if (lane != 0) return;
  if (sums != nullptr) {
    sums[0] = float(a); sums[1] = float(b); sums[2] = float(c);
    sums[3] = mnp; sums[4] = mxp; sums[5] = mnt; sums[6] = mxt;
  }
  if (st != nullptr) {
    st->minp = mnp; st->maxp = mxp; st->mint = mnt; st->maxt = mxt;
    const float rp = mxp - mnp, rt = mxt - mnt;
    st->use_p = (rt > rp) ? 0.f : 1.f;  // python max(a, b) keeps a unless b > a
    st->dr = st->use_p != 0.f ? rp : rt;
    st->nmaxp = 0.f; st->nminp = 0.f; st->d_dr = 0.f;
  }
}

// SSIM constants depend on k1, k2 (set once the range is known) and the tie counts need the range.
__global__ void ssim_constants_kernel(ScaleStats* __restrict__ st, float k1, float k2) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const float a = k1 * st->dr, b = k2 * st->dr;
    st->c1 = a * a;
    st->c2 = b * b;
  }
}

__global__ void count_ties_kernel(const float* __restrict__ p, size_t n, ScaleStats* __restrict__ st) {
  const float mx = st->maxp, mn = st->minp;
  float cmax = 0.f, cmin = 0.f;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    const float v = p[i];
    cmax += (v == mx) ? 1.f : 0.f;
    cmin += (v == mn) ? 1.f : 0.f;
  }
  cmax = warp_sum(cmax);
  cmin = warp_sum(cmin);
  if ((threadIdx.x & 31) == 0) {
    if (cmax != 0.f) atomicAdd(&st->nmaxp, cmax);  // integer-valued floats: exact, order independent
    if (cmin != 0.f) atomicAdd(&st->nminp, cmin);
  }
}

// grad[i] (=|+=) gl * ( a*sign(p-t) + b*(1 - t/(p+1e-8)) + c*(p-t) ),  coef = {a, b, c} on the device
__global__ void loss_grad_kernel(const float* __restrict__ p, const float* __restrict__ t, size_t n,
                                 const float* __restrict__ coef, const float* __restrict__ gl,
                                 float* __restrict__ grad, int accumulate) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float a = coef[0], b = coef[1], c = coef[2], g = gl ? *gl : 1.f;
  const float pv = p[i], tv = t[i];
  const float d = pv - tv;
  const float sgn = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
  const float v = g * (a * sgn + b * (1.f - tv / (pv + 1e-8f)) + c * d);
  grad[i] = accumulate ? grad[i] + v : v;
}

// ------------------------------------------------------------------ avg-pool (next scale)
__global__ void avgpool2_pair_kernel(const float* __restrict__ p, const float* __restrict__ t, float* __restrict__ po,
                                     float* __restrict__ to, int nimg, int h, int w) {
  const int oh = h / 2, ow = w / 2;
  const size_t n = size_t(nimg) * oh * ow;
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int x = int(i % ow);
  const size_t r = i / ow;
  const int y = int(r % oh);
  const size_t img = r / oh;
  const float* a = p + (img * h + 2 * y) * w + 2 * x;
  const float* b = t + (img * h + 2 * y) * w + 2 * x;
  po[i] = (a[0] + a[1] + a[w] + a[w + 1]) * 0.25f;
  to[i] = (b[0] + b[1] + b[w] + b[w + 1]) * 0.25f;
}

// ------------------------------------------------------------------ SSIM statistics
struct SsimArgs {
  const float* p;
  const float* t;
  int nimg, h, w;            // images of h x w; valid outputs (h-18) x (w-18)
  const ScaleStats* st;
  Window win;
  // forward: per-image sums  acc[img][4] = { sum sim, sum cs, sum d(sel)/dc1, sum d(sel)/dc2 }
  float* acc;
  int use_sim;               // which map feeds this scale's value: 1 = ssim (last scale / plain SSIM), 0 = cs
  // backward: maps of  k_img * d(sel)/d{mu_p total, E[pp], E[pt]}  over the valid region
  const float* kimg;         // per-image coefficient (nullptr in the forward pass)
  float* ga; float* gb; float* gc;
};

__global__ void __launch_bounds__(256) ssim_stats_kernel(const SsimArgs a) {
  __shared__ float sp[kSsimInH][kSsimInW + 1];
  __shared__ float stt[kSsimInH][kSsimInW + 1];
  __shared__ float hz[5][kSsimInH][kSsimTileW + 1];
  __shared__ float red[8][4];
  const int vh = a.h - 2 * kSsimPad, vw = a.w - 2 * kSsimPad;
  const int img = blockIdx.z;
  const int oy0 = blockIdx.y * kSsimTileH, ox0 = blockIdx.x * kSsimTileW;  // valid-region coordinates
  const float* pi = a.p + size_t(img) * a.h * a.w;
  const float* ti = a.t + size_t(img) * a.h * a.w;
  for (int i = threadIdx.x; i < kSsimInH * kSsimInW; i += blockDim.x) {
    const int r = i / kSsimInW, c = i - r * kSsimInW;
    const int y = oy0 + r, x = ox0 + c;  // input coordinates (valid coord + [0, 18])
    float pv = 0.f, tv = 0.f;
    if (y < a.h && x < a.w) { pv = pi[size_t(y) * a.w + x]; tv = ti[size_t(y) * a.w + x]; }
    sp[r][c] = pv;
    stt[r][c] = tv;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kSsimInH * kSsimTileW; i += blockDim.x) {
    const int r = i / kSsimTileW, c = i - r * kSsimTileW;
    float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f, m4 = 0.f;
#pragma unroll
    for (int k = 0; k < kSsimTaps; ++k) {
      const float wv = a.win.w[k], pv = sp[r][c + k], tv = stt[r][c + k];
      m0 = fmaf(wv, pv, m0); m1 = fmaf(wv, tv, m1);
      m2 = fmaf(wv, pv * pv, m2); m3 = fmaf(wv, tv * tv, m3); m4 = fmaf(wv, pv * tv, m4);
    }
    hz[0][r][c] = m0; hz[1][r][c] = m1; hz[2][r][c] = m2; hz[3][r][c] = m3; hz[4][r][c] = m4;
  }
  __syncthreads();
  const float c1 = a.st->c1, c2 = a.st->c2;
  const float kc = a.kimg ? a.kimg[img] : 0.f;
  float s_sim = 0.f, s_cs = 0.f, s_d1 = 0.f, s_d2 = 0.f;
  for (int i = threadIdx.x; i < kSsimTileH * kSsimTileW; i += blockDim.x) {
    const int r = i / kSsimTileW, c = i - r * kSsimTileW;
    const int oy = oy0 + r, ox = ox0 + c;
    if (oy >= vh || ox >= vw) continue;
    float mp = 0.f, mt = 0.f, epp = 0.f, ett = 0.f, ept = 0.f;
#pragma unroll
    for (int k = 0; k < kSsimTaps; ++k) {
      const float wv = a.win.w[k];
      mp = fmaf(wv, hz[0][r + k][c], mp); mt = fmaf(wv, hz[1][r + k][c], mt);
      epp = fmaf(wv, hz[2][r + k][c], epp); ett = fmaf(wv, hz[3][r + k][c], ett);
      ept = fmaf(wv, hz[4][r + k][c], ept);
    }
    const float vpp_raw = epp - mp * mp, vtt_raw = ett - mt * mt;
    const float vpp = fmaxf(vpp_raw, 0.f), vtt = fmaxf(vtt_raw, 0.f);
    const float vpt = ept - mp * mt;
    const float up = 2.f * vpt + c2, lo = vpp + vtt + c2;
    const float lum_n = 2.f * mp * mt + c1, lum_d = mp * mp + mt * mt + c1;
    const float cs = up / lo, lum = lum_n / lum_d;
    const float sim = lum * cs;
    s_sim += sim;
    s_cs += cs;
    // derivatives of the selected map
    const float dcs_dvpt = 2.f / lo;
    const float dcs_dvpp = (vpp_raw >= 0.f) ? -up / (lo * lo) : 0.f;  // torch.clamp(min=0) passes grad at == 0
    const float dcs_dc2 = (lo - up) / (lo * lo);
    const float dlum_dmp = (2.f * mt * lum_d - lum_n * 2.f * mp) / (lum_d * lum_d);
    const float dlum_dc1 = (lum_d - lum_n) / (lum_d * lum_d);
    float d_mp, d_vpp, d_vpt;
    if (a.use_sim) {
      d_mp = dlum_dmp * cs; d_vpp = lum * dcs_dvpp; d_vpt = lum * dcs_dvpt;
      s_d1 += dlum_dc1 * cs; s_d2 += lum * dcs_dc2;
    } else {
      d_mp = 0.f; d_vpp = dcs_dvpp; d_vpt = dcs_dvpt;
      s_d2 += dcs_dc2;
    }
    if (a.kimg != nullptr) {
      // vpp = E[pp] - mp^2, vpt = E[pt] - mp*mt  ->  total derivative w.r.t. mu_p
      const size_t o = (size_t(img) * vh + oy) * vw + ox;
      a.ga[o] = kc * (d_mp - 2.f * mp * d_vpp - mt * d_vpt);
      a.gb[o] = kc * d_vpp;
      a.gc[o] = kc * d_vpt;
    }
  }
  if (a.acc != nullptr) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    s_sim = warp_sum(s_sim); s_cs = warp_sum(s_cs); s_d1 = warp_sum(s_d1); s_d2 = warp_sum(s_d2);
    if (lane == 0) { red[warp][0] = s_sim; red[warp][1] = s_cs; red[warp][2] = s_d1; red[warp][3] = s_d2; }
    __syncthreads();
    if (threadIdx.x < 4) {
      float v = 0.f;
      for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
      // per-(tile) partial in a fixed slot -> deterministic second-stage sum in the finalize kernel
      const size_t tiles = size_t(gridDim.x) * gridDim.y;
      a.acc[((size_t(img) * tiles) + size_t(blockIdx.y) * gridDim.x + blockIdx.x) * 4 + threadIdx.x] = v;
    }
  }
}

// ------------------------------------------------------------------ SSIM gradient w.r.t. preds
struct SsimGradArgs {
  const float* p;
  const float* t;
  int nimg, h, w;
  const ScaleStats* st;
  Window win;
  const float* ga; const float* gb; const float* gc;  // valid-region maps from ssim_stats_kernel
  const float* coarse;  // gradient w.r.t. the next (coarser) scale's preds [nimg][h/2][w/2], or nullptr
  float* grad;          // [nimg][h][w]
  int accumulate;       // add to grad instead of overwriting (scale 0 on top of the L1/Poisson gradient)
};

__global__ void __launch_bounds__(256) ssim_grad_kernel(const SsimGradArgs a) {
  __shared__ float m[3][kSsimInH][kSsimInW + 1];
  __shared__ float hz[3][kSsimInH][kSsimTileW + 1];
  const int vh = a.h - 2 * kSsimPad, vw = a.w - 2 * kSsimPad;
  const int img = blockIdx.z;
  const int y0 = blockIdx.y * kSsimTileH, x0 = blockIdx.x * kSsimTileW;  // input-pixel tile
  // grad(x) = sum_q w(q - x) M(q), q valid: q in [x-9, x+9]; valid index = q - 9 in [x-18, x]
  for (int i = threadIdx.x; i < kSsimInH * kSsimInW; i += blockDim.x) {
    const int r = i / kSsimInW, c = i - r * kSsimInW;
    const int vy = y0 + r - 2 * kSsimPad, vx = x0 + c - 2 * kSsimPad;
    float v0 = 0.f, v1 = 0.f, v2 = 0.f;
    if (vy >= 0 && vy < vh && vx >= 0 && vx < vw) {
      const size_t o = (size_t(img) * vh + vy) * vw + vx;
      v0 = a.ga[o]; v1 = a.gb[o]; v2 = a.gc[o];
    }
    m[0][r][c] = v0; m[1][r][c] = v1; m[2][r][c] = v2;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kSsimInH * kSsimTileW; i += blockDim.x) {
    const int r = i / kSsimTileW, c = i - r * kSsimTileW;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < kSsimTaps; ++k) {
      const float wv = a.win.w[kSsimTaps - 1 - k];  // symmetric window; written as a true correlation transpose
      s0 = fmaf(wv, m[0][r][c + k], s0); s1 = fmaf(wv, m[1][r][c + k], s1); s2 = fmaf(wv, m[2][r][c + k], s2);
    }
    hz[0][r][c] = s0; hz[1][r][c] = s1; hz[2][r][c] = s2;
  }
  __syncthreads();
  const ScaleStats st = *a.st;
  for (int i = threadIdx.x; i < kSsimTileH * kSsimTileW; i += blockDim.x) {
    const int r = i / kSsimTileW, c = i - r * kSsimTileW;
    const int y = y0 + r, x = x0 + c;
    if (y >= a.h || x >= a.w) continue;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < kSsimTaps; ++k) {
      const float wv = a.win.w[kSsimTaps - 1 - k];
      s0 = fmaf(wv, hz[0][r + k][c], s0); s1 = fmaf(wv, hz[1][r + k][c], s1); s2 = fmaf(wv, hz[2][r + k][c], s2);
    }
    const size_t o = (size_t(img) * a.h + y) * a.w + x;
    const float pv = a.p[o], tv = a.t[o];
    float g = s0 + 2.f * pv * s1 + tv * s2;
    if (st.use_p != 0.f && st.d_dr != 0.f) {  // data-range path: torch.max / torch.min spread over ties
      if (pv == st.maxp) g += st.d_dr / st.nmaxp;
      if (pv == st.minp) g -= st.d_dr / st.nminp;
    }
    if (a.coarse != nullptr) {
      const int ch = a.h / 2, cw = a.w / 2;
      if ((y >> 1) < ch && (x >> 1) < cw) g += 0.25f * a.coarse[(size_t(img) * ch + (y >> 1)) * cw + (x >> 1)];
    }
    a.grad[o] = a.accumulate ? a.grad[o] + g : g;
  }
}

// ------------------------------------------------------------------ (MS-)SSIM finalize
// acc layout: [scale][img][tile][4] partial sums; tiles[s] tiles per image at scale s.
// Outputs: value[0] = mean over batch of prod_s v_s^beta_s  (nscales == 1, beta = 1: plain SSIM mean);
//          per-image values (for torchmetrics-style state) in img_val[img];
//          kimg[s][img] = gl * d value / d v_{s,img} / Nvalid_s  (backward coefficient of the selected map);
//          st[s].d_dr   = dL/d(data_range_s).
struct MsFinalizeArgs {
  const float* acc[kMaxScales];
  int tiles[kMaxScales];
  int nvalid[kMaxScales];
  int nscales;
  int batch, channels;        // images = batch * channels; per-batch value = mean over its channels
  float betas[kMaxScales];
  float k1, k2;
  ScaleStats* st;             // [nscales]
  float* value;               // scalar
  float* img_val;             // [batch]
  float* kimg;                // [nscales][batch*channels]
  const float* gl;            // upstream gradient scalar (device) times the term's weight, or nullptr (=1)
  float weight;               // multiplies gl in the backward coefficients
};

__device__ __forceinline__ double warp_sum_d(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One block of 256 threads.  Every (scale, image) pair is summed over its tiles by ONE warp (lanes stride the tiles,
// double accumulation, butterfly reduction: a fixed order, so the result is deterministic); the per-pair results
// live in dynamic shared memory: vsum / p1 / p2 [nscales][nimg] floats.  (The first version walked all tiles with
// one thread per batch element and then once more with a single thread: 1.6 ms per call at 16 x 832x832.)
__global__ void __launch_bounds__(256) msssim_finalize_kernel(const MsFinalizeArgs a) {
  extern __shared__ float fin_sm[];
  const int nimg = a.batch * a.channels;
  float* vsum = fin_sm;                       // [nscales][nimg]  sum of the selected map (sim on the last scale, cs before)
  float* p1 = vsum + a.nscales * nimg;        // [nscales][nimg]  kimg * sum of d/dc1 terms
  float* p2 = p1 + a.nscales * nimg;          // [nscales][nimg]  kimg * sum of d/dc2 terms
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int npairs = a.nscales * nimg;
  for (int pr = warp; pr < npairs; pr += nwarps) {
    const int s = pr / nimg, i = pr - s * nimg;
    const bool last = (s == a.nscales - 1);
    const float* q = a.acc[s] + size_t(i) * a.tiles[s] * 4;
    double sum = 0.0;
    for (int tI = lane; tI < a.tiles[s]; tI += 32) sum += q[tI * 4 + (last ? 0 : 1)];
    sum = warp_sum_d(sum);
    if (lane == 0) vsum[pr] = float(sum);
  }
  __syncthreads();
  // per batch element: value and the backward coefficients of its maps
  for (int b = threadIdx.x; b < a.batch; b += blockDim.x) {
    float v[kMaxScales];
    float prod = 1.f;
    for (int s = 0; s < a.nscales; ++s) {
      double sum = 0.0;
      for (int c = 0; c < a.channels; ++c) sum += vsum[s * nimg + b * a.channels + c];
      float mean = float(sum / (double(a.nvalid[s]) * a.channels));
      if (a.nscales > 1) mean = fmaxf(mean, 0.f);  // normalize="relu" (MS-SSIM only)
      v[s] = mean;
      prod *= (a.nscales > 1) ? powf(mean, a.betas[s]) : mean;
    }
    a.img_val[b] = prod;
    const float g = (a.gl ? *a.gl : 1.f) * a.weight / float(a.batch);
    for (int s = 0; s < a.nscales; ++s) {
      float dv;  // d prod / d v_s
      if (a.nscales > 1) dv = (v[s] > 0.f) ? a.betas[s] * prod / v[s] : 0.f;  // 0^beta: emit 0, not inf*0
      else dv = 1.f;
      const float k = g * dv / (float(a.nvalid[s]) * a.channels);
      for (int c = 0; c < a.channels; ++c) a.kimg[size_t(s) * nimg + b * a.channels + c] = k;
    }
  }
  __syncthreads();  // kimg / img_val of this block's threads are visible to the block (single-block grid)
  // data-range derivatives: per pair, kimg * (sum of the c1 / c2 sensitivity partials)
  for (int pr = warp; pr < npairs; pr += nwarps) {
    const int s = pr / nimg, i = pr - s * nimg;
    const float* q = a.acc[s] + size_t(i) * a.tiles[s] * 4;
    double s1 = 0.0, s2 = 0.0;
    for (int tI = lane; tI < a.tiles[s]; tI += 32) { s1 += q[tI * 4 + 2]; s2 += q[tI * 4 + 3]; }
    s1 = warp_sum_d(s1);
    s2 = warp_sum_d(s2);
    if (lane == 0) {
      const float k = a.kimg[size_t(s) * nimg + i];
      p1[pr] = float(double(k) * s1);
      p2[pr] = float(double(k) * s2);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int i = 0; i < a.batch; ++i) tot += a.img_val[i];
    a.value[0] = float(tot / a.batch);
  }
  if (threadIdx.x < a.nscales) {
    const int s = threadIdx.x;
    double d1 = 0.0, d2 = 0.0;
    for (int i = 0; i < nimg; ++i) { d1 += p1[s * nimg + i]; d2 += p2[s * nimg + i]; }
    const float dr = a.st[s].dr;
    a.st[s].d_dr = float(d1 * 2.0 * a.k1 * a.k1 * dr + d2 * 2.0 * a.k2 * a.k2 * dr);
  }
}

}  // namespace xmm

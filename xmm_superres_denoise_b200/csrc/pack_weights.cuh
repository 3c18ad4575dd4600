// Repack fp32 OIHW convolution weights (the nn.Conv2d parameters the reference keeps,
// generator_rrdb.py:31-54, rrdb_blocks.py:27-31) into the bf16 shared-memory image the
// tensor-core kernel bulk-copies: [chunk][tap][n][k] with the UMMA K-major 64B/128B
// swizzle already applied, followed by NT fp32 biases.  The tap blocks come in one of two orders
// (xmm_pack_job::tap_order): dy*3+dx, or dx*3+(2-dy) for the row-hop kernel (conv3x3_row.cuh).
//
// One launch repacks every layer of the model: jobs live in a device-side table that is
// built once (parameter storage is stable across optimizer steps).
#pragma once
#include <cuda_bf16.h>
#include <cstdint>

#include "../../include/xmm_b200.h"

namespace xmm {

__global__ void pack_jobs_kernel(const xmm_pack_job* __restrict__ jobs) {
  const xmm_pack_job& job = jobs[blockIdx.y];
  const int nblocks = job.nchunks * 9;
  const int rowb = job.kc * 2;
  const int swz_mask = (job.kc == 64) ? 7 : 3;
  uint8_t* dst = static_cast<uint8_t*>(job.dst);
  for (int blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
    const int ch = blk / 9, tap = blk - ch * 9;
    int dy = tap / 3, dx = tap - dy * 3;
    if (job.tap_order == 1) {  // row-hop order: block = dx * 3 + (2 - dy)
      dx = tap / 3;
      dy = 2 - (tap - dx * 3);
    }
    uint8_t* bdst = dst + size_t(blk) * job.nt * rowb;
    for (int e = threadIdx.x; e < job.nt * job.kc; e += blockDim.x) {
      const int n = e / job.kc, k = e - n * job.kc;
      const int kg = ch * job.kc + k;  // position in the packed K dimension
      float val = 0.f;
      if (n < job.n_valid) {
        for (int s = 0; s < job.nseg; ++s) {
          const xmm_pack_segment& sg = job.seg[s];
          if (kg >= sg.k_off && kg < sg.k_off + sg.k_count && n >= sg.n_off && n < sg.n_off + sg.n_count) {
            const int c = kg - sg.k_off;
            const int nn = n - sg.n_off;
            int o, i, t;
            if (!sg.transpose) {
              o = sg.o_off + (job.perm ? shuffle_perm(nn, sg.n_count / 4) : nn);
              i = sg.i_off + c;
              t = dy * 3 + dx;
            } else {
              o = sg.o_off + (job.perm ? shuffle_perm(c, sg.k_count / 4) : c);
              i = sg.i_off + nn;
              t = (2 - dy) * 3 + (2 - dx);
            }
            val = sg.scale * sg.src[(size_t(o) * sg.src_cin + i) * 9 + t];
            if (sg.part == 1) val -= __bfloat162float(__float2bfloat16_rn(val));  // low-order half
          }
        }
      }
      uint32_t byte = uint32_t(n * rowb + k * 2);
      byte ^= ((byte >> 7) & swz_mask) << 4;
      *reinterpret_cast<__nv_bfloat16*>(bdst + byte) = __float2bfloat16_rn(val);
    }
  }
  if (blockIdx.x == 0) {
    float* bdst = reinterpret_cast<float*>(dst + size_t(nblocks) * job.nt * rowb);
    for (int n = threadIdx.x; n < job.nt; n += blockDim.x) {
      float b = 0.f;
      if (job.bias != nullptr && n < job.n_valid && n < job.bias_n && !job.seg[0].transpose) {
        const int o = job.seg[0].o_off + (job.perm ? shuffle_perm(n, job.nt / 4) : n);
        b = job.bias[o];
      }
      bdst[n] = b;
    }
  }
}

}  // namespace xmm

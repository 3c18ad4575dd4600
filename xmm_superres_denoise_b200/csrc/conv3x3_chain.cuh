// A chain of dependent 3x3 convolutions -- one dense block of the reference (the five convs of
// ResidualDenseBlock_5C.forward, rrdb_blocks.py:37-54, or the five data gradients of its backward) -- as
// ONE persistent kernel in which the layers run as a pipeline over groups of SMs.
//
// Why.  Layer by layer the dense block is HBM-bound on B200: conv_k re-reads x_0..x_{k-1}, 1280 B per pixel and
// block against 0.28 MFLOP, i.e. 0.55 ms of traffic at batch 16 against 0.46 ms of tensor time, and 3.5 GB per
// block never fits the 126 MB L2.  Here layer k trails layer k-1 by two 8-row strips, so every x_j is consumed
// from L2 a few microseconds after it was written: HBM sees x_0 read once and x_1..x_5 written once (384 B per
// pixel).
//
// How.  The grid is one CTA per SM, split into one group per layer in proportion to the layer's cost
// (K = 9 * cin).  Every CTA keeps ITS layer's packed weights resident in shared memory and is otherwise the
// column-scatter conv of conv3x3_dx.cuh (TMA producer warp, tcgen05 issuer warp, 8 epilogue warps, carried
// column sums).  Work items are strip segments (8 rows x tiles_x/segs tiles), dealt round-robin inside a group so
// that a group always works on a compact band of strips.  A layer's epilogue warps publish "segment stored" with
// a release-add (after their TMA stores have been performed) on a per-(layer, strip) counter; the next layer's TMA producer acquires strips s-1, s, s+1 of
// the previous layer before it requests a tile of strip s (generic-proxy writes -> async-proxy reads: the
// acquire is followed by fence.proxy.async).  Layer 0 depends on nothing and no layer waits on a later one, so
// with all CTAs co-resident (cooperative launch) the pipeline cannot deadlock.
#pragma once
#include "conv3x3_dx.cuh"

namespace xmm {

constexpr int kChainMaxLayers = 5;

struct ChainLayer {
  const void* wblob;
  uint32_t w_bytes;
  int nchunks;
  int cin_off;
  int cta_begin, cta_count;  // this layer's group of CTAs
  int stages;                // activation pipeline depth that fits next to this layer's weights
  ConvEpilogue epi;
};

struct ChainArgs {
  ChainLayer layer[kChainMaxLayers];
  int nlayers;
  int batch, height, width;
  int tiles_x, tiles_y;  // tiles per strip, strips per image
  int segs;              // segments per strip (work item = one segment)
  int* done;             // [nlayers][batch * tiles_y] segment-completion counters, zero on entry
  long long* prof;       // optional [gridDim.x][4] cycle counters: CTA total, dependency wait, epilogue idle, tiles
};

struct ChainTmaps {
  CUtensorMap m[kChainMaxLayers];    // layer inputs
  CUtensorMap out[kChainMaxLayers];  // layer outputs (TMA store)
};

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

template <int KC, int NT>
__global__ void __launch_bounds__(kDxThreads, 1)
conv3x3_chain_kernel(const __grid_constant__ ChainTmaps tmaps, const __grid_constant__ ChainArgs args) {
  using Cfg = DxCfg<KC, NT>;
  // which layer does this CTA serve?
  int li = 0;
#pragma unroll
  for (int l = 1; l < kChainMaxLayers; ++l)
    if (l < args.nlayers && int(blockIdx.x) >= args.layer[l].cta_begin) li = l;
  const ChainLayer& L = args.layer[li];
  const CUtensorMap* tmap = &tmaps.m[li];
  const CUtensorMap* tmap_out = &tmaps.out[li];

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_s = smem;
  float* bias_s = reinterpret_cast<float*>(smem + L.w_bytes);
  uint8_t* stage_s = smem + ((L.w_bytes + Cfg::kBiasBytes + 1023) & ~1023u);
  uint8_t* out_s = stage_s + size_t(L.stages) * Cfg::kStageBytes;  // per-warp staging tiles for the TMA stores
  uint8_t* mail_s = out_s + Cfg::kOutBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(mail_s + Cfg::kMailBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kMaxStages;
  uint64_t* tfull_bar = bars + 2 * kMaxStages;
  uint64_t* tempty_bar = tfull_bar + 4;
  uint64_t* w_bar = tempty_bar + 4;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const long long prof_k0 = clock64();

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(tmap);
    ptx::prefetch_tmap(tmap_out);
    for (int s = 0; s < L.stages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < Cfg::kAccStages; ++a) {
      ptx::mbar_init(&tfull_bar[a], 1);
      ptx::mbar_init(&tempty_bar[a], kDxEpiWarps);
    }
    ptx::mbar_init(w_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc<Cfg::kTmemCols>(tmem_ptr_s);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  // Work items of this CTA: item = strip * segs + seg, dealt round-robin inside the layer's group.
  const int nstrips = args.batch * args.tiles_y;
  const int nitems = nstrips * args.segs;
  const int item0 = int(blockIdx.x) - L.cta_begin;
  const int item_step = L.cta_count;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (+ dependency acquire)
    if (ptx::elect_one()) {
      const uint32_t wtot = L.w_bytes + Cfg::kBiasBytes;
      ptx::mbar_expect_tx(w_bar, wtot);
      const uint8_t* gsrc = static_cast<const uint8_t*>(L.wblob);
      for (uint32_t off = 0; off < wtot; off += 32768u) {
        const uint32_t n = (wtot - off < 32768u) ? (wtot - off) : 32768u;
        ptx::bulk_load(w_s + off, gsrc + off, n, w_bar);
      }
      const int* dep = li > 0 ? args.done + size_t(li - 1) * nstrips : nullptr;
      const int dep_target = args.segs * kDxEpiWarps;  // one arrival per epilogue warp and stored segment
      int stage = 0;
      uint32_t phase = 0;
      long long dep_wait = 0;
      for (int item = item0; item < nitems; item += item_step) {
        const int strip = item / args.segs, seg = item - strip * args.segs;
        const int b = strip / args.tiles_y, ty = strip - b * args.tiles_y;
        const int ta = seg * args.tiles_x / args.segs, tb = (seg + 1) * args.tiles_x / args.segs;
        if (dep != nullptr) {
          const long long w0 = clock64();
          const int lo = ty > 0 ? strip - 1 : strip, hi = ty + 1 < args.tiles_y ? strip + 1 : strip;
          for (int s = lo; s <= hi; ++s)
            while (ld_acquire_gpu(dep + s) < dep_target) __nanosleep(64);
          fence_proxy_async_all();
          dep_wait += clock64() - w0;
        }
        const int y0 = ty * kDxTileH - 1;
        for (int tx = (ta > 0 ? ta - 1 : ta); tx < tb; ++tx) {
          for (int ch = 0; ch < L.nchunks; ++ch) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
            ptx::mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
            ptx::tma_load_4d(stage_s + size_t(stage) * Cfg::kStageBytes, tmap, &full_bar[stage], L.cin_off + ch * KC,
                             tx * kDxTileW, y0, b);
            if (++stage == L.stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
      if (args.prof != nullptr) args.prof[size_t(blockIdx.x) * 4 + 1] = dep_wait;
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (ptx::elect_one()) {
      ptx::mbar_wait(w_bar, 0);
      ptx::tc_fence_after();
      const uint64_t bdesc0 = ptx::umma_smem_desc(ptx::smem_u32(w_s), 16, 8 * Cfg::kRowB, Cfg::kLayout);
      const uint64_t adesc0 = ptx::umma_smem_desc(ptx::smem_u32(stage_s), 16, 8 * Cfg::kRowB, Cfg::kLayout);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int item = item0; item < nitems; item += item_step) {
        const int seg = item % args.segs;
        const int ta = seg * args.tiles_x / args.segs, tb = (seg + 1) * args.tiles_x / args.segs;
        for (int tx = (ta > 0 ? ta - 1 : ta); tx < tb; ++tx) {
          ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
          ptx::tc_fence_after();
          const uint32_t d_addr = tmem_base + uint32_t(acc * Cfg::kAccCols);
          for (int ch = 0; ch < L.nchunks; ++ch) {
            ptx::mbar_wait(&full_bar[stage], phase);
            ptx::tc_fence_after();
            const uint64_t adesc_st = adesc0 + uint64_t((uint32_t(stage) * Cfg::kStageBytes) >> 4);
            const uint64_t bdesc_ch = bdesc0 + uint64_t((uint32_t(ch) * 9u * Cfg::kTapBytes) >> 4);
            dx_issue_chunk<KC, NT>(d_addr, adesc_st, bdesc_ch, ch == 0);
            ptx::umma_commit(&empty_bar[stage]);
            if (++stage == L.stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
          ptx::umma_commit(&tfull_bar[acc]);
          if (++acc == Cfg::kAccStages) {
            acc = 0;
            acc_phase ^= 1u;
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue
    ptx::mbar_wait(w_bar, 0);
    DxEpiWarp<KC, NT> w;
    w.init(warp, lane, bias_s, out_s, mail_s);
    int* done = args.done + size_t(li) * nstrips;
    int acc = 0;
    uint32_t acc_phase = 0;
    long long epi_idle = 0, ntiles = 0;
    for (int item = item0; item < nitems; item += item_step) {
      const int strip = item / args.segs, seg = item - strip * args.segs;
      const int b = strip / args.tiles_y, ty = strip - b * args.tiles_y;
      const int ta = seg * args.tiles_x / args.segs, tb = (seg + 1) * args.tiles_x / args.segs;
      const int t_first = ta > 0 ? ta - 1 : ta;
      ntiles += tb - t_first;
      for (int tx = t_first; tx < tb; ++tx) {
        const long long w0 = clock64();
        ptx::mbar_wait(&tfull_bar[acc], acc_phase);
        epi_idle += clock64() - w0;
        ptx::tc_fence_after();
        const uint32_t t_addr = tmem_base + (uint32_t(w.q * 32) << 16) + uint32_t(acc * Cfg::kAccCols);
        dx_epilogue_tile<KC, NT>(w, L.epi, tmap_out, t_addr, &tempty_bar[acc], b, ty, tx, args.tiles_x,
                                 /*pre=*/tx < ta, /*has_pend=*/tx > t_first, /*direct=*/false, args.height, args.width);
        if (++acc == Cfg::kAccStages) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
      // Publish this warp's share of the segment: wait until its TMA stores have been performed, then release.
      // (Direct stores of tile 0 / the strip-end column by the other lanes are ordered before the release by the
      // warp barrier and the cumulativity of the fence.)
      __syncwarp();
      if (ptx::elect_one()) {
        ptx::bulk_wait<0>();
        fence_proxy_async_all();
        __threadfence();
        atomicAdd(done + strip, 1);
      }
    }
    if (L.epi.colsum != nullptr) colsum_flush<Cfg::kWarpCols>(L.epi, w.csum, w.col_w, lane);
    if (args.prof != nullptr && warp == 2 && lane == 0) {
      args.prof[size_t(blockIdx.x) * 4 + 2] = epi_idle;
      args.prof[size_t(blockIdx.x) * 4 + 3] = ntiles;
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
  if (args.prof != nullptr && threadIdx.x == 0) args.prof[size_t(blockIdx.x) * 4] = clock64() - prof_k0;
}

}  // namespace xmm

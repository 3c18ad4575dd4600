// 3x3 / stride 1 / pad 1 convolution, "row-gather / column-scatter" form, for narrow layers
// (Cout = 32 or 64: every conv of the reference's dense blocks, rrdb_blocks.py:27-31, and their
// data gradients).
//
// Why a second form.  With pixels on M and Cout on N (conv3x3_tc.cuh) one tcgen05.mma reads a
// 128-row A slice (4 KB) from shared memory to feed only N = 32 columns; the instruction is
// shared-memory bound at 32 + N/4 = 40 cycles for 16 cycles of math (measured, tools/probe.cu).
// Here the three taps of one filter ROW share one A read:
//
//     D_dx[y][x'] = sum_dy sum_ci X[y + dy - 1][x'][ci] * W[dy][dx][co][ci]        (N = 3 * Cout)
//     out[y][x]   = D_0[y][x - 1] + D_1[y][x] + D_2[y][x + 1]
//
// so a (dy, k-step) pair is ONE MMA with N = 96 (56 cycles for 48 of math) instead of three.
// dy stays a gather: the M = 128 rows are an [8 rows][16 columns] pixel block of the [10][16]
// row-haloed patch, and a row shift is a 1 KB-aligned start address (canonical K-major layout).
//
// The +-1 column shift is applied in the epilogue.  Neighbouring columns are neighbouring TMEM lanes =
// neighbouring threads of an epilogue warp (two image rows of 16 columns per warp): two warp shuffles per
// output value.  Across the 16-column tile boundary the sums are CARRIED: a CTA walks a strip of 8 image
// rows left to right, lane 15 of each row keeps D_0[15] (the left term of the next tile's column 0) and the
// unfinished sum D_0[14] + D_1[15] of its own column, which it completes -- and stores -- one tile later when
// lane 0 holds D_2[16].  No column halo is ever loaded or multiplied: every MMA row is a useful pixel
// (D_0[-1] and D_2[W] are products with the zero padding).
//
// Work split: the B * ceil(H/8) * ceil(W/16) tiles, strip-major, are cut into gridDim.x equal contiguous
// ranges.  A range that starts mid-strip first runs the tile before it as a "pre-tile" (outputs suppressed)
// to establish the carry; a range that ends mid-strip leaves its last column to the next CTA's pre-tile.
//
// Warp roles (576 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM owner), warps 2..17 = epilogue:
// TMEM lane quarter = warp_id % 4, and the four warps of a quarter take a quarter of the Cout columns each.  The
// epilogue is ~20 dependent instructions per output value (measured: latency-bound with 2 warps per scheduler),
// so it is spread over 4 warps per scheduler to keep up with the MMAs of the narrow layers.
//
// Everything else -- resident weights, TMA pipeline, double-buffered TMEM accumulators, fused
// bias / LeakyReLU / mask / residual / inverse-pixel-shuffle epilogue -- is conv3x3_tc.cuh's.
// The packed weight image is the same one ([chunk][tap][Cout][KC]: the three dx taps of a row
// are adjacent, so they are one N = 3*Cout operand).
#pragma once
#include "conv3x3_tc.cuh"

namespace xmm {

constexpr int kDxTileH = 8;
constexpr int kDxTileW = 16;
constexpr int kDxPatchH = kDxTileH + 2;
constexpr int kDxEpiWarps = 16;
constexpr int kDxThreads = 64 + 32 * kDxEpiWarps;

template <int KC, int NT>
struct DxCfg {
  static_assert(KC == 32 || KC == 64, "K chunk is 32 (SWIZZLE_64B) or 64 (SWIZZLE_128B) channels");
  static_assert(NT == 32 || NT == 64, "column-scatter form is for Cout 32 / 64");
  static constexpr int kRowB = KC * 2;
  static constexpr uint32_t kLayout = (KC == 64) ? ptx::UMMA_SW128 : ptx::UMMA_SW64;
  static constexpr int kStageBytes = kDxPatchH * kDxTileW * kRowB;  // 10 KB / 20 KB: a multiple of 1024
  static constexpr int kKSteps = KC / 16;
  static constexpr int kTapBytes = NT * kRowB;
  static constexpr int kAccCols = 3 * NT;
  static constexpr int kAccStages = 512 / kAccCols >= 4 ? 4 : 2;  // accumulator ring in TMEM (4 x 96 or 2 x 192 columns)
  static constexpr int kTmemCols = 512;
  static constexpr uint32_t kIdesc = ptx::umma_idesc_bf16_f32(128, 3 * NT, 0, 0);
  static constexpr int kBiasBytes = NT * 4;
  static constexpr int kBarBytes = (2 * kMaxStages + 2 * 4 + 1) * 8 + 16;
  static constexpr int kWarpCols = NT / 4;          // Cout columns owned by one epilogue warp
  static constexpr int kChunk = kWarpCols < 16 ? kWarpCols : 16;  // columns per tcgen05.ld / store group
  static constexpr int kWarpChunks = kWarpCols / kChunk;
  static constexpr int kOutRowB = NT * 2;                          // one output pixel in the store staging tile
  static constexpr int kOutTileBytes = kDxTileH * kDxTileW * kOutRowB;  // 8 KB / 16 KB, [8][16] pixels, TMA swizzled
  static size_t smem_bytes(uint32_t w_bytes, int stages) {
    return 1024 + w_bytes + kBiasBytes + 1024 + size_t(stages) * kStageBytes + 2 * kOutTileBytes + kBarBytes;
  }
};

// Byte offset of 16-byte chunk k16 of pixel p inside a staging tile whose rows are 64 B (SWIZZLE_64B: chunk bits
// [4,6) ^= address bits [7,9)) or 128 B (SWIZZLE_128B: chunk bits [4,7) ^= address bits [7,10)) -- the layout the
// output tensor map expects, and conflict-free for a warp's 16-byte stores.
template <int NT>
__device__ __forceinline__ uint32_t dx_out_offset(int p, int k16) {
  if (NT == 32) return uint32_t(p * 64 + ((k16 ^ ((p >> 1) & 3)) << 4));
  return uint32_t(p * 128 + ((k16 ^ (p & 7)) << 4));
}

// Tile g of the strip-major order -> image, strip row, tile column.
struct DxTile {
  int b, ty, tx;
  __device__ __forceinline__ DxTile(int g, int tiles_x, int tiles_y) {
    const int strip = g / tiles_x;
    tx = g - strip * tiles_x;
    b = strip / tiles_y;
    ty = strip - b * tiles_y;
  }
  __device__ __forceinline__ void next(int tiles_x, int tiles_y) {
    if (++tx == tiles_x) {
      tx = 0;
      if (++ty == tiles_y) {
        ty = 0;
        ++b;
      }
    }
  }
};

// The MMAs of one K-chunk (KC channels) of one tile: per filter row dy, KC/16 instructions of N = 3*NT.
template <int KC, int NT>
__device__ __forceinline__ void dx_issue_chunk(uint32_t d_addr, uint64_t adesc_st, uint64_t bdesc_ch, bool first_chunk) {
  using Cfg = DxCfg<KC, NT>;
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
    for (int ks = 0; ks < Cfg::kKSteps; ++ks) {
      const uint64_t adesc = adesc_st + uint64_t((uint32_t(dy * kDxTileW * Cfg::kRowB) + uint32_t(ks * 32)) >> 4);
      const uint64_t bdesc = bdesc_ch + uint64_t((uint32_t(dy * 3 * Cfg::kTapBytes) + uint32_t(ks * 32)) >> 4);
      ptx::umma_ss(d_addr, adesc, bdesc, Cfg::kIdesc, (first_chunk && dy == 0 && ks == 0) ? 0u : 1u);
    }
  }
}

// Epilogue of one tile for one warp: read this warp's columns of the three partial sums, form
// out[c] = D_0[c-1] + D_1[c] + D_2[c+1] with two shuffles (the tile-boundary terms come from / go to the
// carry registers of lane 15 of each row), release the accumulator, apply the fused epilogue and store.
// `x` is this lane's output column: the current tile's for lanes 0..14, the previous tile's column 15 for lane 15.
template <int KC, int NT>
__device__ __forceinline__ void dx_epilogue_tile(const ConvEpilogue& epi,
                                                 const float (&bias_r)[NT / 4], uint32_t t_addr, uint64_t* tempty,
                                                 float (&carry)[NT / 4], float (&pend)[NT / 4], int lane, int col_w, int b, int y, int x,
                                                 bool valid, bool strip_end, int flush_x, int H, int W,
                                                 uint8_t* out_tile,  // != nullptr: stage for a TMA store
                                                 int prow            // this lane's row inside the tile
#ifdef XMM_CONV_PROFILE
                                                 , long long* pa
#endif
                                                 ) {
  using Cfg = DxCfg<KC, NT>;
#ifdef XMM_CONV_PROFILE
  long long pt = clock64(), pn;
#define XMM_EPI_MARK(i) pn = clock64(); pa[i] += pn - pt; pt = pn
#else
#define XMM_EPI_MARK(i)
#endif
  const bool last_col = (lane & 15) == 15;
  const int src_l = (lane & 16) | ((lane + 15) & 15);
  const int src_r = (lane & 16) | ((lane + 1) & 15);
  constexpr int CH = Cfg::kChunk;
#pragma unroll
  for (int cc = 0; cc < Cfg::kWarpChunks; ++cc) {
    uint32_t d0[CH], d1[CH], d2[CH];
    ptx::tmem_ld_cols<CH>(t_addr + uint32_t(cc * CH), d0);
    ptx::tmem_ld_cols<CH>(t_addr + uint32_t(NT + cc * CH), d1);
    ptx::tmem_ld_cols<CH>(t_addr + uint32_t(2 * NT + cc * CH), d2);
    ptx::tmem_ld_wait();
    XMM_EPI_MARK(0);
    if (cc == Cfg::kWarpChunks - 1) {
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(tempty);
    }
    float v[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const float c0 = __uint_as_float(d0[i]);
      const float send = last_col ? carry[cc * CH + i] : c0;
#ifdef XMM_EXP_NOSHFL
      const float left = send, right = __uint_as_float(d2[i]);
#else
      const float left = __shfl_sync(0xffffffffu, send, src_l);
      const float right = __shfl_sync(0xffffffffu, __uint_as_float(d2[i]), src_r);
#endif
      const float lsum = left + (__uint_as_float(d1[i]) + bias_r[cc * CH + i]);
      v[i] = (last_col ? pend[cc * CH + i] : lsum) + right;
      pend[cc * CH + i] = lsum;   // meaningful on lane 15 only: D_0[14] + D_1[15] (+ bias)
      carry[cc * CH + i] = c0;    // meaningful on lane 15 only: D_0[15]
    }
    XMM_EPI_MARK(1);
    if (out_tile != nullptr) {
      if (valid) {
        conv_epilogue_math<NT, CH>(epi, nullptr, v, col_w + cc * CH, b, y, x, H, W);
        // box pixel: image row of this lane, column = lane column + 1 (lane 15 is the box's column 0)
        const int p = prow * kDxTileW + (((lane & 15) + 1) & 15);
#pragma unroll
        for (int j = 0; j < CH / 8; ++j)
          *reinterpret_cast<uint4*>(out_tile + dx_out_offset<NT>(p, (col_w + cc * CH) / 8 + j)) = pack8(v + j * 8);
      }
    } else if (valid) {
      conv_epilogue_cols<NT, CH>(epi, nullptr, v, col_w + cc * CH, b, y, x, H, W);
    }
    XMM_EPI_MARK(2);
  }
  // End of the strip: column 15 of the last tile has no right neighbour (D_2[W] = 0).
  if (strip_end && last_col && y < H && flush_x < W) {
#pragma unroll
    for (int cc = 0; cc < Cfg::kWarpChunks; ++cc) {
      float v[CH];
#pragma unroll
      for (int i = 0; i < CH; ++i) v[i] = pend[cc * CH + i];
      conv_epilogue_cols<NT, CH>(epi, nullptr, v, col_w + cc * CH, b, y, flush_x, H, W);
    }
  }
}

// After every epilogue warp has staged its part of a tile: one TMA store of the [8][16]-pixel x NT-channel box
// whose column 0 is the previous tile's column 15 (box x = 16*tx - 1; out-of-image rows / columns are clipped by
// the tensor map on the high side, which is what makes the carried column and the ragged edges free; the box may
// not start below 0, so tile 0 of a strip uses direct stores instead).  Called by all epilogue
// warps.  The staging tile is double buffered: the elected thread first makes sure the store issued two tiles ago
// has finished reading this buffer (it waits for the reads of all its earlier stores; they are a tile old).
__device__ __forceinline__ void dx_store_tile(const CUtensorMap* tmap_out, const uint8_t* out_tile, int warp, int c0,
                                              int x0, int y0, int b) {
  ptx::fence_proxy_async();  // this thread's st.shared -> visible to the async proxy
  if (warp == 2 && ptx::elect_one()) ptx::bulk_wait_read<0>();
  ptx::named_bar_sync(1, 32 * kDxEpiWarps);
  if (warp == 2 && ptx::elect_one()) {
    ptx::tma_store_4d(tmap_out, out_tile, c0, x0, y0, b);
    ptx::bulk_commit();
  }
}

template <int KC, int NT>
__global__ void __launch_bounds__(kDxThreads, 1)
conv3x3_dx_kernel(const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ CUtensorMap tmap_out,
                  const ConvArgs args) {
  using Cfg = DxCfg<KC, NT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_s = smem;
  float* bias_s = reinterpret_cast<float*>(smem + args.w_bytes);
  uint8_t* stage_s = smem + ((args.w_bytes + Cfg::kBiasBytes + 1023) & ~1023u);
  uint8_t* out_s = stage_s + size_t(args.stages) * Cfg::kStageBytes;  // 2 staging tiles for the TMA store
  uint64_t* bars = reinterpret_cast<uint64_t*>(out_s + 2 * Cfg::kOutTileBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kMaxStages;
  uint64_t* tfull_bar = bars + 2 * kMaxStages;
  uint64_t* tempty_bar = tfull_bar + 4;
  uint64_t* w_bar = tempty_bar + 4;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
#ifdef XMM_CONV_PROFILE
  long long prof_t0_ = 0;
  long long prof_acc_[6] = {0, 0, 0, 0, 0, 0};
  long long prof_epi_[3] = {0, 0, 0};
  const long long prof_k0_ = clock64();
  unsigned long long prof_g0_;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(prof_g0_));
#endif

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_in);
    ptx::prefetch_tmap(&tmap_out);
    for (int s = 0; s < args.stages; ++s) {
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

  // This CTA's contiguous tile range [t0, t1) of the strip-major order; g0 < t0 adds the pre-tile.
  const long long total = args.num_tiles;
  const int t0 = int(total * blockIdx.x / gridDim.x);
  const int t1 = int(total * (blockIdx.x + 1) / gridDim.x);
  const int g0 = (t0 % args.tiles_x != 0) ? t0 - 1 : t0;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (ptx::elect_one()) {
      const uint32_t wtot = args.w_bytes + Cfg::kBiasBytes;
      ptx::mbar_expect_tx(w_bar, wtot);
      const uint8_t* gsrc = static_cast<const uint8_t*>(args.wblob);
      for (uint32_t off = 0; off < wtot; off += 32768u) {
        const uint32_t n = (wtot - off < 32768u) ? (wtot - off) : 32768u;
        ptx::bulk_load(w_s + off, gsrc + off, n, w_bar);
      }
      int stage = 0;
      uint32_t phase = 0;
      DxTile t(g0, args.tiles_x, args.tiles_y);
      for (int g = g0; g < t1; ++g, t.next(args.tiles_x, args.tiles_y)) {
        const int y0 = t.ty * kDxTileH - 1, x0 = t.tx * kDxTileW;
        for (int ch = 0; ch < args.nchunks; ++ch) {
          XMM_PROF_T0();
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
          XMM_PROF_ADD(0);
          ptx::mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          ptx::tma_load_4d(stage_s + size_t(stage) * Cfg::kStageBytes, &tmap_in, &full_bar[stage],
                           args.cin_off + ch * KC, x0, y0, t.b);
          if (++stage == args.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
      XMM_PROF_FLUSH(0);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (ptx::elect_one()) {
      ptx::mbar_wait(w_bar, 0);
      ptx::tc_fence_after();
      const uint32_t w_addr = ptx::smem_u32(w_s);
      const uint32_t st_addr = ptx::smem_u32(stage_s);
      const uint64_t bdesc0 = ptx::umma_smem_desc(w_addr, 16, 8 * Cfg::kRowB, Cfg::kLayout);
      const uint64_t adesc0 = ptx::umma_smem_desc(st_addr, 16, 8 * Cfg::kRowB, Cfg::kLayout);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      XMM_PROF_START(3);
      for (int g = g0; g < t1; ++g) {
        XMM_PROF_T0();
        ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        XMM_PROF_ADD(1);
        ptx::tc_fence_after();
        const uint32_t d_addr = tmem_base + uint32_t(acc * Cfg::kAccCols);
        for (int ch = 0; ch < args.nchunks; ++ch) {
          XMM_PROF_T0();
          ptx::mbar_wait(&full_bar[stage], phase);
          XMM_PROF_ADD(2);
          ptx::tc_fence_after();
          const uint64_t adesc_st = adesc0 + uint64_t((uint32_t(stage) * Cfg::kStageBytes) >> 4);
          const uint64_t bdesc_ch = bdesc0 + uint64_t((uint32_t(ch) * 9u * Cfg::kTapBytes) >> 4);
          dx_issue_chunk<KC, NT>(d_addr, adesc_st, bdesc_ch, ch == 0);
          ptx::umma_commit(&empty_bar[stage]);
          if (++stage == args.stages) {
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
      XMM_PROF_STOP(3);
      XMM_PROF_FLUSH(1); XMM_PROF_FLUSH(2); XMM_PROF_FLUSH(3);
    }
  } else {
    // ------------------------------------------------------------ epilogue
    const int q = warp & 3;                   // TMEM lane quarter = image rows 2q, 2q+1 of the strip
    const int part = (warp - 2) >> 2;         // which quarter of the Cout columns
    const int prow = 2 * q + (lane >> 4);
    const int pcol = lane & 15;
    const bool last_col = pcol == 15;         // finishes the PREVIOUS tile's column 15
    const int col_w = part * Cfg::kWarpCols;
    float carry[Cfg::kWarpCols], pend[Cfg::kWarpCols], bias_r[Cfg::kWarpCols];
#pragma unroll
    for (int i = 0; i < Cfg::kWarpCols; ++i) carry[i] = pend[i] = 0.f;
    ptx::mbar_wait(w_bar, 0);
#pragma unroll
    for (int i = 0; i < Cfg::kWarpCols; ++i) bias_r[i] = bias_s[col_w + i];
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool use_tma = args.epi.pixel_shuffle == 0;  // (inverse) pixel shuffle scatters: direct stores
    int obuf = 0;
    DxTile t(g0, args.tiles_x, args.tiles_y);  // advanced incrementally: no divisions in the tile loop
    for (int g = g0; g < t1; ++g) {
      const int y = t.ty * kDxTileH + prow;
      const bool pre = g < t0;
      const bool has_pend = (t.tx > 0) && (g != g0);
      if (t.tx == 0) {
#pragma unroll
        for (int i = 0; i < Cfg::kWarpCols; ++i) carry[i] = 0.f;  // D_0[-1]: zero padding
      }
      const int x = last_col ? t.tx * kDxTileW - 1 : t.tx * kDxTileW + pcol;
      const bool valid = (y < args.height) && (last_col ? has_pend : (!pre && x < args.width));
      XMM_PROF_T0();
      ptx::mbar_wait(&tfull_bar[acc], acc_phase);
      XMM_PROF_ADD(4);
      XMM_PROF_T0();
      ptx::tc_fence_after();
      const uint32_t t_addr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * Cfg::kAccCols + col_w);
      // a TMA store may not start at a negative coordinate (x = -1 is an illegal instruction on B200): the first
      // tile of a strip, whose box column 0 is outside the image anyway, stores directly
      uint8_t* out_tile = (use_tma && !pre && t.tx > 0) ? out_s + obuf * Cfg::kOutTileBytes : nullptr;
      dx_epilogue_tile<KC, NT>(args.epi, bias_r, t_addr, &tempty_bar[acc], carry, pend, lane, col_w, t.b, y, x, valid,
                               t.tx == args.tiles_x - 1, t.tx * kDxTileW + 15, args.height, args.width, out_tile, prow
#ifdef XMM_CONV_PROFILE
                               , prof_epi_
#endif
                               );
      if (out_tile != nullptr) {
        dx_store_tile(&tmap_out, out_tile, warp, args.epi.out_coff, t.tx * kDxTileW - 1, t.ty * kDxTileH, t.b);
        obuf ^= 1;
      }
      XMM_PROF_ADD(5);
      if (++acc == Cfg::kAccStages) {
        acc = 0;
        acc_phase ^= 1u;
      }
      t.next(args.tiles_x, args.tiles_y);
    }
    if (warp == 2 && ptx::elect_one()) ptx::bulk_wait<0>();  // stores complete before the CTA (and its smem) goes away
    if (warp == 2 && lane == 0) { XMM_PROF_FLUSH(4); XMM_PROF_FLUSH(5); }
#ifdef XMM_CONV_PROFILE
    if (warp == 2 && lane == 0)
      for (int i = 0; i < 3; ++i) args.prof[size_t(gridDim.x) * 8 + size_t(blockIdx.x) * 4 + i] = prof_epi_[i];
#endif
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
#ifdef XMM_CONV_PROFILE
  if (threadIdx.x == 0) {
    unsigned long long g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    args.prof[size_t(blockIdx.x) * 8 + 6] = clock64() - prof_k0_;
    args.prof[size_t(blockIdx.x) * 8 + 7] = (long long)(g1 - prof_g0_);
  }
#endif
}

}  // namespace xmm

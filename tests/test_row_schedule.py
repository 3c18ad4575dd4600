"""CPU: the row-hop convolution (csrc/conv3x3_row.cuh) restated in Python and checked against ``F.conv2d``.

What is emulated, following the C++ line by line (same integer arithmetic):

* ``RowSched``: the pieces (column, first row, end row) of every CTA -- every output row of every column is produced
  by exactly one piece;
* the TMA producer's banded loads (band coordinate -1 / +1 for the rows above / below a band, zero fill outside the
  tensor);
* the MMA issuer's slot ring: which accumulator slots an input row's N = 3*Cout instruction targets, where it is
  split (ring end, overwriting first touch), which weight block goes where (pack order dx*3 + (2 - dy));
* the epilogue's slot / parity sequence: a slot is never overwritten before its previous occupant was drained.

The GPU tests check the kernel itself (tests/test_gpu_forward.py::test_conv3x3_row_hop_*)."""

import numpy as np
import pytest
import torch
import torch.nn.functional as F

K_SLOTS = {32: 16, 64: 8}


def row_bands(height):
    for n in range(16, 7, -1):
        if height % n == 0:
            return n, height // n
    return None


def sched_pieces(ncols, band_h, tiles_x, grid, cta, rr_rounds):
    """RowSched::init + get for one CTA -> [(b, xt, ra, rb)]."""
    tail_col0 = rr_rounds * grid
    total = (ncols - tail_col0) * band_h
    t0, t1 = total * cta // grid, total * (cta + 1) // grid
    tail_c0 = t0 // band_h
    ntail = (t1 - 1) // band_h - tail_c0 + 1 if t1 > t0 else 0
    out = []
    for i in range(rr_rounds + ntail):
        if i < rr_rounds:
            col, ra, rb = cta + i * grid, 0, band_h
        else:
            c = tail_c0 + (i - rr_rounds)
            col = tail_col0 + c
            base = c * band_h
            ra = t0 - base if t0 > base else 0
            rb = t1 - base if t1 < base + band_h else band_h
        b = col // tiles_x
        out.append((b, col - b * tiles_x, ra, rb))
    return out


def launch_split(batch, height, width, sm_count, rr=True):
    nbands, band_h = row_bands(height)
    tiles_x = (width + 7) // 8
    ncols = tiles_x * batch
    total_rows = ncols * band_h
    grid = min(total_rows, sm_count)
    rr_rounds = ncols // grid if (rr and ncols >= 2 * grid) else 0
    return nbands, band_h, tiles_x, ncols, grid, rr_rounds


@pytest.mark.parametrize("batch,height,width,sms", [(1, 416, 416, 148), (16, 416, 416, 148), (64, 416, 416, 148),
                                                     (2, 48, 40, 148), (3, 96, 80, 148), (1, 832, 832, 148),
                                                     (5, 33 * 3, 50, 7), (2, 40, 19, 148), (1, 16, 8, 148)])
def test_every_output_row_belongs_to_exactly_one_piece(batch, height, width, sms):
    nbands, band_h, tiles_x, ncols, grid, rr_rounds = launch_split(batch, height, width, sms)
    seen = np.zeros((ncols, band_h), dtype=np.int32)
    steps = []
    for cta in range(grid):
        n = 0
        for b, xt, ra, rb in sched_pieces(ncols, band_h, tiles_x, grid, cta, rr_rounds):
            assert 0 <= ra < rb <= band_h and 0 <= b < batch and 0 <= xt < tiles_x
            seen[b * tiles_x + xt, ra:rb] += 1
            n += rb - ra + 2
        steps.append(n)
    assert (seen == 1).all()
    # balance: no CTA walks more than the mean + one column's worth of halo rows (+ one row of rounding)
    assert max(steps) <= (ncols * band_h) / grid + 2 * (rr_rounds + 2) + 1


def emulate_cta(x, w_packed, cout, pieces, nbands, band_h, width, kslots, out, drained_log):
    """One CTA of conv3x3_row_kernel on float64 numpy data.

    x: [B, H, W, Cin]; w_packed[dx][blk] = [Cout, Cin] with blk <-> dy = 2 - blk (tap_order 1);
    out: [B, H, W, Cout] (written through the banded view)."""
    cin = x.shape[3]
    tmem = np.full((kslots, 128, cout), np.nan)   # garbage until overwritten
    occupant = [None] * kslots                    # (v) of the output row a slot holds, None = drained / never used
    v = 0

    def load_patch(b, xt, r):
        row, band0 = r, 0
        if r < 0:
            row, band0 = band_h - 1, -1
        elif r >= band_h:
            row, band0 = 0, 1
        patch = np.zeros((16, 10, cin))
        for k in range(16):
            band = band0 + k
            if not 0 <= band < nbands:
                continue                              # TMA zero fill (band outside the tensor)
            y = band * band_h + row
            for p in range(10):
                xx = xt * 8 - 1 + p
                if 0 <= xx < width:
                    patch[k, p] = x[b, y, xx]
        return patch

    def drain(b, xt, ra, j, vv):
        slot = vv & (kslots - 1)
        assert occupant[slot] == vv, "epilogue reads a slot that does not hold its row"
        acc = tmem[slot]
        for m in range(128):
            band, px = m >> 3, m & 7
            xx = xt * 8 + px
            if band < nbands and xx < width:          # TMA store clips the rest
                out[b, band * band_h + j, xx] = acc[m]
        occupant[slot] = None
        drained_log.append(vv)

    for b, xt, ra, rb in pieces:
        for r in range(ra - 1, rb + 1):
            j0, j1 = max(r - 1, ra), min(r + 1, rb - 1)
            nblk = j1 - j0 + 1
            jb = j0 - (r - 1)
            vj0 = v + (j0 - ra)
            has_new = r + 1 <= rb - 1
            wrap = kslots - (vj0 & (kslots - 1))
            nold = nblk - 1 if has_new else nblk
            a0 = min(nold, wrap)
            a1 = nold - a0
            b0 = min(nblk, wrap)
            b1 = nblk - b0
            if has_new:
                vnew = v + (r + 1 - ra)
                assert occupant[vnew & (kslots - 1)] is None, "overwriting a slot that was not drained"
                occupant[vnew & (kslots - 1)] = vnew
            patch = load_patch(b, xt, r)

            def mma(dx, k, length, accumulate):
                if length <= 0:
                    return
                a = patch[:, dx:dx + 8].reshape(128, cin)          # descriptor view: start + dx pixels
                slot0 = (vj0 + k) & (kslots - 1)
                assert slot0 + length <= kslots, "instruction runs past the end of TMEM"
                for t in range(length):
                    d = a @ w_packed[dx][jb + k + t].T
                    if accumulate:
                        tmem[slot0 + t] += d
                    else:
                        tmem[slot0 + t] = d

            first = True
            for dx in range(3):
                if first:
                    mma(dx, 0, a0, True)
                    mma(dx, a0, a1, True)
                    if has_new:
                        mma(dx, nblk - 1, 1, False)
                    first = False
                else:
                    mma(dx, 0, b0, True)
                    mma(dx, b0, b1, True)
            if r - 1 >= ra:
                drain(b, xt, ra, r - 1, v + (r - 1 - ra))          # row r-1 is complete: the epilogue may take it
        v += rb - ra


@pytest.mark.parametrize("batch,height,width,cin,cout,sms", [
    (1, 32, 16, 8, 32, 5),      # two bands rows each, ring wraps many times (16 slots, 2 rows per column... several columns)
    (2, 48, 21, 4, 32, 7),      # ragged width (21 = 2*8 + 5), tail pieces that start / end mid-column
    (1, 40, 8, 4, 32, 3),       # 10 bands only (40 = 10 * 4): lanes of missing bands idle
    (1, 144, 8, 4, 64, 2),      # 8 slots (Cout = 64), 9-row columns: every wrap position
    (3, 160, 24, 4, 32, 148),   # more CTAs than rows per CTA = 1: single-row pieces
    (9, 64, 40, 4, 32, 4),      # round-robin rounds + tail
])
def test_emulated_kernel_equals_conv2d(batch, height, width, cin, cout, sms):
    g = torch.Generator().manual_seed(height * 131 + width)
    x = torch.randn(batch, cin, height, width, generator=g, dtype=torch.float64)
    w = torch.randn(cout, cin, 3, 3, generator=g, dtype=torch.float64)
    want = F.conv2d(x, w, padding=1).permute(0, 2, 3, 1).numpy()
    nbands, band_h, tiles_x, ncols, grid, rr_rounds = launch_split(batch, height, width, sms)
    # tap_order 1: block (dx, blk) holds W[:, :, dy = 2 - blk, dx]
    w_packed = [[w[:, :, 2 - blk, dx].numpy() for blk in range(3)] for dx in range(3)]
    xn = x.permute(0, 2, 3, 1).contiguous().numpy()
    out = np.full((batch, height, width, cout), np.nan)
    for cta in range(grid):
        log = []
        pieces = sched_pieces(ncols, band_h, tiles_x, grid, cta, rr_rounds)
        emulate_cta(xn, w_packed, cout, pieces, nbands, band_h, width, K_SLOTS[cout], out, log)
        assert log == list(range(len(log)))   # rows are completed (and drained) in slot-ring order
    assert not np.isnan(out).any()
    np.testing.assert_allclose(out, want, rtol=1e-10, atol=1e-10)


def test_pack_order_of_the_row_hop_weight_image():
    """pack_jobs_kernel with tap_order = 1: block index tap -> (dx, dy) = (tap // 3, 2 - tap % 3)."""
    seen = set()
    for tap in range(9):
        dx = tap // 3
        dy = 2 - (tap - dx * 3)
        seen.add((dy, dx))
        assert tap == dx * 3 + (2 - dy)
    assert len(seen) == 9

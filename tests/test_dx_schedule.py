"""CPU: the work split of the column-scatter convolution (csrc/conv3x3_dx.cuh, struct DxRuns and the issue order of
conv3x3_dx_kernel) restated in Python and checked for its invariants over many problem shapes:

* every tile's outputs are stored by exactly one CTA (pre-tiles and dummy strips store nothing);
* a range that starts mid-strip begins with exactly one pre-tile, the tile before it in the same strip;
* CTA pairs (cta_group::2) issue the same number of tiles in both CTAs of a pair;
* with two epilogue groups the two streams of a unit are two whole strips, the issue order alternates between them and
  every accumulator / side-tile ring position is used by exactly one tile.

The restatement follows the C++ line by line (same integer arithmetic); the GPU tests check the kernel itself
(tests/test_gpu_forward.py::test_conv3x3_cta_pairs_bit_identical_to_single_ctas)."""

import pytest


def dx_runs(num_tiles, tiles_x, grid, b, strip_rr, pair):
    """DxRuns::init for CTA b: list of runs (r0, len, tvalid, dummy)."""
    nstrips = num_tiles // tiles_x
    rr_step, rr_len, rr_r0 = grid * tiles_x, tiles_x, b * tiles_x
    runs = []
    if pair:
        lead = b & ~1
        rr_runs = (nstrips - lead + grid - 1) // grid if nstrips > lead else 0
        dummy_from = (nstrips - b + grid - 1) // grid if nstrips > b else 0
        for run in range(rr_runs):
            dummy = run >= dummy_from
            r0 = rr_r0 if dummy else rr_r0 + run * rr_step
            runs.append((r0, rr_len, r0, dummy))
        return runs
    rr_runs = nstrips // grid if strip_rr else 0
    done = rr_runs * rr_step
    rem = num_tiles - done
    t0, t1 = done + rem * b // grid, done + rem * (b + 1) // grid
    for run in range(rr_runs):
        r0 = rr_r0 + run * rr_step
        runs.append((r0, rr_len, r0, False))
    if t1 > t0:
        g0 = t0 - 1 if t0 % tiles_x != 0 else t0
        runs.append((g0, t1 - g0, t0, False))
    return runs


def issue_order(runs, rr_runs, g2):
    """The kernel's issue order of CTA tiles: [(tile, stored?, group)], units of one run or (G2) two rr runs."""
    order = []
    npairs = rr_runs // 2 if g2 else 0
    nunits = len(runs) - npairs
    for u in range(nunits):
        if u < npairs:
            (r0a, la, ta, da), (r0b, lb, tb, db) = runs[2 * u], runs[2 * u + 1]
            assert la == lb
            for i in range(la):
                order.append((r0a + i, r0a + i >= ta and not da, 0))
                order.append((r0b + i, r0b + i >= tb and not db, 1))
        else:
            r0, ln, tv, dm = runs[u + npairs]
            for i in range(ln):
                order.append((r0 + i, r0 + i >= tv and not dm, 0))
    return order


SHAPES = [(26, 52, 16), (26, 52, 64), (3, 601, 1), (3, 592, 1), (4, 2400, 1), (2, 5, 3), (1, 1, 1), (26, 52, 1), (7, 13, 5)]


@pytest.mark.parametrize("tiles_x,tiles_y,batch", SHAPES)
@pytest.mark.parametrize("grid", [148, 37, 2])
@pytest.mark.parametrize("strip_rr,pair,g2", [(0, 0, 0), (1, 0, 0), (1, 1, 0), (1, 0, 1)])
def test_every_tile_is_stored_exactly_once(tiles_x, tiles_y, batch, grid, strip_rr, pair, g2):
    num_tiles = tiles_x * tiles_y * batch
    nstrips = tiles_y * batch
    grid = min(grid, num_tiles)
    if pair and (grid % 2 or nstrips < grid):
        pytest.skip("pairs need an even grid and at least one strip per CTA (launch_conv_dx)")
    stored = [0] * num_tiles
    per_cta_tiles = []
    for b in range(grid):
        runs = dx_runs(num_tiles, tiles_x, grid, b, strip_rr, pair)
        rr_runs = len(runs) if pair else (nstrips // grid if strip_rr else 0)
        order = issue_order(runs, rr_runs, bool(g2))
        per_cta_tiles.append(len(order))
        for r0, ln, tv, dm in runs:
            assert ln >= 1 and 0 <= r0 and r0 + ln <= num_tiles
            assert tv in (r0, r0 + 1)  # at most one pre-tile ...
            if tv == r0 + 1:           # ... the tile before the range start, in the same strip
                assert tv % tiles_x != 0 and not dm
        for tile, is_stored, group in order:
            if is_stored:
                stored[tile] += 1
        if g2:  # ring positions: tile k of the issue order uses accumulator k % 4; both groups see disjoint positions
            pos = {0: [], 1: []}
            for k, (_, _, group) in enumerate(order):
                pos[group].append(k)
            assert not set(pos[0]) & set(pos[1])
            npairs = (nstrips // grid) // 2
            assert len(pos[1]) == npairs * tiles_x
    assert all(c == 1 for c in stored), (stored.count(0), max(stored))
    if pair:  # lock step: both CTAs of a pair issue the same number of tiles
        for p in range(0, grid, 2):
            assert per_cta_tiles[p] == per_cta_tiles[p + 1]
    else:     # balance: no CTA has more than one strip's worth of tiles above the mean (+ its pre-tile)
        mean = num_tiles / grid
        assert max(per_cta_tiles) <= mean + 2


def test_round_robin_rounds_are_full():
    """While rounds remain every CTA takes exactly one strip per round; 832 strips on 148 CTAs leave 92 strips that are
    cut into 148 contiguous ranges of ~16 tiles."""
    tiles_x, nstrips, grid = 26, 832, 148
    num_tiles = tiles_x * nstrips
    lens = []
    for b in range(grid):
        runs = dx_runs(num_tiles, tiles_x, grid, b, 1, 0)
        assert [r[1] for r in runs[:5]] == [26] * 5 and len(runs) == 6
        lens.append(runs[5][1])
    assert min(lens) >= 16 and max(lens) <= 18

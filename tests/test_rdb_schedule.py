"""CPU: the work split and step schedule of the fused dense-block kernels (csrc/conv3x3_rdb.cuh: rdb_tile_origin,
RdbSched, rdb_walk), restated in Python and checked for the properties the kernel's barriers rely on -- no GPU.

* tiles: the pixels each tile stores partition the image row, and lie inside the lanes that are still exact after
  NL fused layers (halo recompute instead of halo exchange);
* pieces: every (column, row) is stored by exactly one CTA;
* walk: layer l takes its step on row r only after layer l-1 took its step on row r+1 in an EARLIER round; every row a
  layer reads from an in-CTA map exists; a ring slot is given back (by the row's last reader) before the issuer can
  block on the row that re-uses it; every accumulator row that is touched is completed (flush steps);
* accumulator slots: the rows alive in a step's window occupy three distinct, consecutive slots and a slot is only
  re-used after its row was drained (the one-row-in-flight rule of the issuer).
"""
import itertools

import pytest

P = 63  # kRdbP


def tile_origin(t, nl):
    return 0 if t == 0 else (P - (nl - 1)) + (t - 1) * (P - 2 * (nl - 1))


def tiles_x(width, nl):
    n = 1
    while (0 if n == 1 else tile_origin(n - 1, nl) - (nl - 1)) + P < width:
        n += 1
    return n


def piece(col, t0, t1, band_h, ntx, width, nl):
    base = col * band_h
    ra = max(t0 - base, 0)
    rb = min(t1 - base, band_h)
    b, t = divmod(col, ntx)
    own_lo = tile_origin(t, nl)
    x0 = 0 if t == 0 else own_lo - (nl - 1)
    own_hi = width if t == ntx - 1 else tile_origin(t + 1, nl)
    return dict(b=b, t=t, x0=x0, own_lo=own_lo, own_hi=own_hi, ra=ra, rb=rb)


def cta_pieces(cta, grid, batch, band_h, width, nl):
    ntx = tiles_x(width, nl)
    total = batch * ntx * band_h
    t0, t1 = total * cta // grid, total * (cta + 1) // grid
    if t1 <= t0:
        return []
    c0 = t0 // band_h
    n = (t1 - 1) // band_h - c0 + 1
    return [piece(c0 + i, t0, t1, band_h, ntx, width, nl) for i in range(n)]


def walk_reference(pieces, nl):
    """The schedule as a rule: in every round, on the state at its start, layer l steps when layer l-1 took its
    step on row r+1 in an earlier round.  Yields (round, layer, piece index, r, flush, n, seq[list])."""
    if not pieces:
        return
    pi, r, hi, n, seq, done = [0] * nl, [0] * nl, [0] * nl, [0] * nl, [[0] * nl for _ in range(nl)], [False] * nl
    for l in range(nl):
        e = nl - 1 - l
        r[l] = pieces[0]["ra"] - e - 1
        hi[l] = pieces[0]["rb"] + e + 2
    rnd = 0
    while not done[nl - 1]:
        sp, sr, sd = list(pi), list(r), list(done)
        for l in range(nl):
            go = not done[l]
            if l > 0 and go:
                go = sd[l - 1] or sp[l - 1] > pi[l] or (sp[l - 1] == pi[l] and sr[l - 1] > r[l] + 1)
            if not go:
                continue
            e = nl - 1 - l
            pc = pieces[pi[l]]
            yield rnd, l, pi[l], r[l], r[l] > pc["rb"] + e, n[l], list(seq[l])
            n[l] += 1
            r[l] += 1
            if r[l] > hi[l]:
                for m in range(nl):
                    seq[l][m] += (pc["rb"] - pc["ra"]) + 2 * (nl - 1 - m)
                pi[l] += 1
                if pi[l] >= len(pieces):
                    done[l] = True
                else:
                    r[l] = pieces[pi[l]]["ra"] - e - 1
                    hi[l] = pieces[pi[l]]["rb"] + e + 2
        rnd += 1
        assert rnd < 10 ** 6


def walk(pieces, nl):
    """rdb_walk as the kernel runs it: the same rule, with the steady state (all layers inside one piece, two rows
    apart) taken `reps` rounds at a time without re-deciding."""
    if not pieces:
        return
    pi, r, hi, n, seq, done = [0] * nl, [0] * nl, [0] * nl, [0] * nl, [[0] * nl for _ in range(nl)], [False] * nl
    for l in range(nl):
        e = nl - 1 - l
        r[l] = pieces[0]["ra"] - e - 1
        hi[l] = pieces[0]["rb"] + e + 2
    rnd = 0
    while not done[nl - 1]:
        go = [not done[0]] + [False] * (nl - 1)
        steady = not done[0]
        reps = pieces[pi[0]]["rb"] + (nl - 1) - r[0] + 1 if not done[0] else 0
        for l in range(1, nl):
            ready = (not done[l]) and (done[l - 1] or pi[l - 1] > pi[l] or (pi[l - 1] == pi[l] and r[l - 1] > r[l] + 1))
            go[l] = ready
            steady = steady and ready and pi[l] == pi[0]
            if not done[l]:
                reps = min(reps, pieces[pi[l]]["rb"] + (nl - 1 - l) - r[l] + 1)
        if not steady or reps < 1:
            reps = 1
        for _ in range(reps):
            for l in range(nl):
                if not go[l]:
                    continue
                e = nl - 1 - l
                pc = pieces[pi[l]]
                yield rnd, l, pi[l], r[l], r[l] > pc["rb"] + e, n[l], list(seq[l])
                n[l] += 1
                r[l] += 1
                if r[l] > hi[l]:
                    for m in range(nl):
                        seq[l][m] += (pc["rb"] - pc["ra"]) + 2 * (nl - 1 - m)
                    pi[l] += 1
                    if pi[l] >= len(pieces):
                        done[l] = True
                    else:
                        r[l] = pieces[pi[l]]["ra"] - e - 1
                        hi[l] = pieces[pi[l]]["rb"] + e + 2
            rnd += 1
        assert rnd < 10 ** 6


@pytest.mark.parametrize("nl,batch,band_h,width,grid", [(3, 64, 208, 416, 148), (2, 64, 208, 416, 148), (3, 2, 24, 40, 48),
                                                        (2, 3, 4, 130, 7), (3, 1, 4, 8, 4), (3, 3, 5, 70, 5), (2, 2, 9, 64, 13)])
def test_steady_state_shortcut_is_the_same_schedule(nl, batch, band_h, width, grid):
    ntx = tiles_x(width, nl)
    grid = min(grid, batch * ntx * band_h)
    for cta in sorted(set([0, 1, grid // 3, grid - 1])):
        pieces = cta_pieces(cta, grid, batch, band_h, width, nl)
        assert list(walk(pieces, nl)) == list(walk_reference(pieces, nl))


@pytest.mark.parametrize("nl", [2, 3])
@pytest.mark.parametrize("width", [1, 8, 40, 61, 62, 63, 64, 120, 122, 123, 416, 417, 832, 1000])
def test_tiles_partition_the_row_inside_the_exact_lanes(nl, width):
    ntx = tiles_x(width, nl)
    sh = nl - 1
    covered = []
    for t in range(ntx):
        pc = piece(t, 0, 1, 1, ntx, width, nl)
        assert 0 <= pc["x0"] and pc["own_lo"] < pc["own_hi"] <= width
        covered += list(range(pc["own_lo"], pc["own_hi"]))
        # lanes exact after nl layers: sh pixels are lost on every side that is not an image border
        lo = pc["x0"] + (0 if pc["x0"] == 0 else sh)
        hi = pc["x0"] + P - (0 if pc["x0"] + P >= width else sh)
        assert lo <= pc["own_lo"] and min(pc["own_hi"], width) <= hi, (t, pc, lo, hi)
    assert covered == list(range(width))
    if width == 416:
        assert ntx == 7  # 93 % of the lanes useful at the reference's image width


@pytest.mark.parametrize("nl,batch,band_h,width,grid", [(3, 64, 208, 416, 148), (2, 64, 208, 416, 148), (3, 1, 208, 416, 148),
                                                        (3, 2, 24, 40, 48), (2, 3, 4, 130, 7), (3, 1, 4, 8, 4),
                                                        (3, 16, 208, 416, 148), (2, 5, 33, 200, 148)])
def test_every_row_of_every_column_is_stored_once(nl, batch, band_h, width, grid):
    ntx = tiles_x(width, nl)
    grid = min(grid, batch * ntx * band_h)
    seen = {}
    for cta in range(grid):
        for pc in cta_pieces(cta, grid, batch, band_h, width, nl):
            for row in range(pc["ra"], pc["rb"]):
                key = (pc["b"], pc["t"], row)
                assert key not in seen
                seen[key] = cta
    assert len(seen) == batch * ntx * band_h


RINGS = {3: (5, 3), 2: (2, 0)}


@pytest.mark.parametrize("nl,batch,band_h,width,grid", [(3, 64, 208, 416, 148), (2, 64, 208, 416, 148), (3, 2, 24, 40, 48),
                                                        (2, 3, 4, 130, 7), (3, 1, 4, 8, 4), (3, 3, 5, 70, 5),
                                                        (3, 1, 208, 416, 148), (2, 2, 9, 64, 13)])
def test_walk_dependencies_rings_and_accumulator_slots(nl, batch, band_h, width, grid):
    ntx = tiles_x(width, nl)
    grid = min(grid, batch * ntx * band_h)
    ctas = range(grid) if grid <= 16 else [0, 1, grid // 2, grid - 2, grid - 1]
    for cta in ctas:
        pieces = cta_pieces(cta, grid, batch, band_h, width, nl)
        steps = list(walk(pieces, nl))
        issued = {}          # (layer, piece, r) -> (index in issue order, round)
        for idx, (rnd, l, p, r, flush, n, seq) in enumerate(steps):
            assert (l, p, r) not in issued
            issued[(l, p, r)] = (idx, rnd)
        per_layer = {l: [s for s in steps if s[1] == l] for l in range(nl)}
        for l in range(nl):
            e = nl - 1 - l
            assert [s[5] for s in per_layer[l]] == list(range(len(per_layer[l])))  # n counts the layer's steps
            # every piece: rows ra-e-1 .. rb+e real, then two flush steps
            for p, pc in enumerate(pieces):
                rows = [(s[3], s[4]) for s in per_layer[l] if s[2] == p]
                want = [(r, False) for r in range(pc["ra"] - e - 1, pc["rb"] + e + 1)] + \
                       [(pc["rb"] + e + 1, True), (pc["rb"] + e + 2, True)]
                assert rows == want
        # dependencies and map rows
        produced = {}        # (map, seq) -> issue index of the step that completed it
        for idx, (rnd, l, p, r, flush, n, seq) in enumerate(steps):
            pc = pieces[p]
            e = nl - 1 - l
            j = r - 1
            if l < nl - 1 and pc["ra"] - e <= j < pc["rb"] + e:
                sq = seq[l] + (j - (pc["ra"] - e))
                assert (l, sq) not in produced
                produced[(l, sq)] = idx
            if flush:
                continue
            if l > 0:
                dep = issued[(l - 1, p, r + 1)]
                assert dep[1] < rnd, "the producing step must be a full round earlier"
        # sequence numbers are dense per map
        for m in range(nl - 1):
            sqs = sorted(s for (mm, s) in produced if mm == m)
            assert sqs == list(range(len(sqs)))
        # reads: the row exists; ring slot re-use is safe
        readers = {}         # (map, seq) -> list of issue indices
        release = {}
        for idx, (rnd, l, p, r, flush, n, seq) in enumerate(steps):
            if flush:
                continue
            pc = pieces[p]
            e = nl - 1 - l
            for m in range(l):
                sq = seq[m] + (r - (pc["ra"] - (nl - 1 - m)))
                assert (m, sq) in produced and produced[(m, sq)] < idx
                readers.setdefault((m, sq), []).append(idx)
                if l == nl - 1 or r < pc["ra"] - e or r > pc["rb"] + e - 1:
                    assert (m, sq) not in release
                    release[(m, sq)] = idx
        for key, rd in readers.items():
            assert key in release and release[key] == max(rd), "exactly the last reader gives the tile back"
        assert set(readers) == set(produced)
        for (m, sq), first_read in ((k, min(v)) for k, v in readers.items()):
            ring = RINGS[nl][m]
            if sq >= ring:
                # the issuer blocks on row sq (first reader) until the epilogue wrote it, which needs the release of
                # row sq - ring: that release must come earlier in the issuer's program order
                assert release[(m, sq - ring)] < first_read
        # accumulator slots of each layer (5 per layer: window positions 0..2, carry slots 3, 4)
        for l in range(nl):
            live = {}        # slot -> row
            prev_completed = None
            for (rnd, ll, p, r, flush, n, seq) in per_layer[l]:
                c = (r - 1) % 3
                if not flush:
                    for k, row in enumerate((r - 1, r, r + 1)):
                        slot = c + k
                        assert live.get(slot, (p, row)) == (p, row), "slot still holds another row"
                        live[slot] = (p, row)
                # the step completes row r-1: drained (main slot + carry slot) before the next step of this layer
                done_row = (p, r - 1)
                phi = (r - 1) % 3
                slots = [phi] + ([3 + phi] if phi < 2 else [])
                for s in list(live):
                    if live[s] == done_row:
                        assert s in slots
                        del live[s]
            assert not live, "every touched accumulator row is completed by the flush steps"

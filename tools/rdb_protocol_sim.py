"""Developer tool (CPU): discrete-event model of the barrier protocol of csrc/conv3x3_rdb.cuh -- TMA producer, MMA
issuer, tensor pipe, two epilogue groups -- over the walk restated in tests/test_rdb_schedule.py.  Detects deadlocks
and out-of-protocol barrier use before any GPU time is spent.  python tools/rdb_protocol_sim.py"""
import os
import sys
from collections import deque

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_rdb_schedule import RINGS, cta_pieces, tiles_x, walk  # noqa: E402


class Bar:
    def __init__(self, count, name):
        self.count, self.pending, self.phase, self.name = count, count, 0, name

    def arrive(self):
        self.pending -= 1
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def ready(self, parity):
        return (self.phase & 1) != parity


def simulate(nl, g, pieces, stages, verbose=False):
    ring = RINGS[nl]
    full = [[Bar(1, f"full{l}.{s}") for s in range(8)] for l in range(nl)]
    empty = [Bar(1, f"empty{s}") for s in range(stages)]
    tfull = [[Bar(1, f"tfull{l}.{g}") for g in range(2)] for l in range(nl)]
    tdrain = [[Bar(4, f"tdrain{l}.{g}") for g in range(2)] for l in range(nl)]
    mfull = [[Bar(4, f"mfull{m}.{s}") for s in range(max(ring[m], 1))] for m in range(nl - 1)]
    mempty = [[Bar(1, f"mempty{m}.{s}") for s in range(max(ring[m], 1))] for m in range(nl - 1)]
    pipe = deque()       # tensor pipe FIFO: ("mma",) or ("commit", bar)
    tma = deque()        # in-flight loads: (ticks left, bar)
    steps = list(walk(pieces, nl))

    def producer():
        stage, phase = 0, 0
        fills = [0] * nl
        for (rnd, l, p, r, flush, n, seq) in steps:
            if flush:
                continue
            for _ in range(g):
                while not empty[stage].ready(phase ^ 1):
                    yield ("empty", stage)
                tma.append([3, full[l][fills[l] % 8]])
                fills[l] += 1
                stage += 1
                if stage == stages:
                    stage, phase = 0, phase ^ 1
        return

    def mma(my_layer):
        stage, fills = 0, 0
        cnt = [[0, 0] for _ in range(nl)]
        prev = [(0, 0)] * nl
        for task, (rnd, l, p, r, flush, n, seq) in enumerate(steps):
            pc = pieces[p]
            grp = task & 1
            if l != my_layer:
                if not flush:
                    stage += g
                    if stage >= stages:
                        stage -= stages
                continue
            if n > 0:
                while not tdrain[l][prev[l][0]].ready(prev[l][1] & 1):
                    yield ("tdrain", l, n)
            if not flush:
                for _ in range(g):
                    while not full[l][fills % 8].ready((fills // 8) & 1):
                        yield ("full", l, fills)
                    fills += 1
                    pipe.append(("mma",))
                    pipe.append(("commit", empty[stage]))
                    stage += 1
                    if stage == stages:
                        stage = 0
                for m in range(nl - 1):
                    if m < l:
                        sq = seq[m] + (r - (pc["ra"] - (nl - 1 - m)))
                        slot = sq % ring[m]
                        while not mfull[m][slot].ready((sq // ring[m]) & 1):
                            yield ("mfull", m, sq)
                        pipe.append(("mma",))
                        e = nl - 1 - l
                        if l == nl - 1 or r < pc["ra"] - e or r > pc["rb"] + e - 1:
                            pipe.append(("commit", mempty[m][slot]))
            pipe.append(("commit", tfull[l][grp]))
            prev[l] = (grp, cnt[l][grp])
            cnt[l][grp] += 1
        return

    def epilogue(group):
        cnt = [0] * nl
        for task, (rnd, l, p, r, flush, n, seq) in enumerate(steps):
            if (task & 1) != group:
                continue
            kth = cnt[l]
            cnt[l] += 1
            pc = pieces[p]
            j = r - 1
            e = nl - 1 - l
            real = pc["ra"] - e <= j < pc["rb"] + e
            while not tfull[l][group].ready(kth & 1):
                yield ("tfull", l, n)
            for _ in range(4):
                tdrain[l][group].arrive()
            if not real:
                continue
            if l < nl - 1:
                sq = seq[l] + (j - (pc["ra"] - e))
                slot = sq % ring[l]
                while not mempty[l][slot].ready(((sq // ring[l]) & 1) ^ 1):
                    yield ("mempty", l, sq)
                for _ in range(4):
                    mfull[l][slot].arrive()
        return

    agents = {"producer": producer(), "epi0": epilogue(0), "epi1": epilogue(1)}
    for l in range(nl):
        agents[f"mma{l}"] = mma(l)
    blocked = {}
    ticks = 0
    while agents:
        progressed = False
        for name in list(agents):
            try:
                blocked[name] = next(agents[name])
            except StopIteration:
                del agents[name]
                blocked.pop(name, None)
                progressed = True
        # tensor pipe: one op per tick; TMA: count down
        if pipe:
            op = pipe.popleft()
            if op[0] == "commit":
                op[1].arrive()
            progressed = True
        for t in list(tma):
            t[0] -= 1
            if t[0] <= 0:
                t[1].arrive()
                tma.remove(t)
            progressed = True
        ticks += 1
        if not progressed and not pipe and not tma:
            # agents that yield are blocked; if all are and nothing is in flight: deadlock
            return False, dict(blocked)
        if ticks > 5_000_000:
            return False, {"timeout": dict(blocked)}
    return True, ticks


def main():
    cases = [(3, 1, 1, 4, 16, 4), (2, 4, 1, 4, 16, 4), (3, 1, 2, 24, 40, 48), (2, 4, 2, 24, 40, 48), (3, 1, 64, 208, 416, 148),
             (2, 4, 64, 208, 416, 148), (3, 1, 3, 5, 70, 5), (2, 4, 3, 5, 70, 5)]
    ok_all = True
    for nl, g, batch, band_h, width, grid in cases:
        ntx = tiles_x(width, nl)
        grid = min(grid, batch * ntx * band_h)
        stages = 5 if nl == 3 else 4
        for cta in sorted(set([0, 1, grid // 2, grid - 1])):
            pieces = cta_pieces(cta, grid, batch, band_h, width, nl)
            ok, info = simulate(nl, g, pieces, stages)
            print(f"NL={nl} G={g} batch={batch} band_h={band_h} W={width} grid={grid} cta={cta}: "
                  f"{'ok, ticks ' + str(info) if ok else 'DEADLOCK ' + str(info)}")
            ok_all &= ok
    sys.exit(0 if ok_all else 1)


if __name__ == "__main__":
    main()

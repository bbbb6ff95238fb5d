"""Discrete-event model of dw_tc_wgrad3.cu's warp-role pipeline: which hand-off chain bounds a phase?

Hop latencies are the measured ones (tools/hop_probe.cu): mbarrier.arrive -> waiter runs again 178 clk, tcgen05.commit ->
waiter 236 clk, tcgen05.ld + wait 160 clk.  The tensor pipe executes MMAs in issue order (16 clk per N = 32 product MMA,
32 per N = 64 transposing MMA).  Everything else (TMA, regrouping) is modelled as never late unless --regroup is given.

    python tools/w3_pipeline_sim.py [--dt 2] [--aslots 2] [--work 1] ...
"""
import argparse
import heapq

ARRIVE, COMMIT, LD, ST, TRY = 178, 236, 160, 100, 25


class Bar:
    def __init__(self, count):
        self.count, self.arr, self.done, self.waiters = count, {}, {}, {}


class Sim:
    def __init__(self):
        self.now, self.q, self.seq = 0, [], 0
        self.pipe_free, self.last_mma = 0, {}

    def spawn(self, gen, name):
        self._push(0, gen, name)

    def _push(self, t, gen, name):
        self.seq += 1
        heapq.heappush(self.q, (t, self.seq, gen, name))

    def arrive(self, bar, use, t):
        a = bar.arr.setdefault(use, [])
        a.append(t)
        if len(a) == bar.count:
            bar.done[use] = max(a)
            for (gen, name) in bar.waiters.pop(use, []):
                self._push(bar.done[use] + TRY, gen, name)

    def run(self):
        while self.q:
            t, _, gen, name = heapq.heappop(self.q)
            self.now = t
            try:
                op = next(gen)
            except StopIteration:
                continue
            kind = op[0]
            if kind == "wait":
                _, bar, use = op
                if use < 0 or (use in bar.done and bar.done[use] <= t):
                    self._push(t + TRY, gen, name)
                elif use in bar.done:
                    self._push(bar.done[use] + TRY, gen, name)
                else:
                    bar.waiters.setdefault(use, []).append((gen, name))
            elif kind == "arrive":
                self.arrive(op[1], op[2], t + ARRIVE)
                self._push(t + 10, gen, name)
            elif kind == "mma":
                start = max(self.pipe_free, t)
                self.pipe_free = start + op[1]
                self.last_mma[name] = self.pipe_free
                # issue blocks while more than `depth` clocks of work are queued
                self._push(max(t + 4, self.pipe_free - op[2]), gen, name)
            elif kind == "commit":
                self.arrive(op[1], op[2], max(t, self.last_mma.get(name, 0)) + COMMIT)
                self._push(t + 8, gen, name)
            elif kind == "work":
                self._push(t + op[1], gen, name)


def build(args):
    P = args.planes * 5
    sim = Sim()
    DT, AS = args.dt, args.aslots
    dt_full = [Bar(1) for _ in range(DT)]
    dt_empty = [Bar(args.conv) for _ in range(DT)]
    a_full = [Bar(args.conv) for _ in range(AS)]
    a_empty = [Bar(1) for _ in range(AS)]
    qd = args.qdepth * 16
    w = args.work

    def sel():
        for n in range(P):
            s, u = n % DT, n // DT
            yield ("wait", dt_empty[s], u - 1)
            if w:
                for _ in range(4):
                    yield ("mma", 32, qd)
            yield ("commit", dt_full[s], u)

    def main():
        for n in range(P):
            s, u = n % AS, n // AS
            yield ("wait", a_full[s], u)
            if w:
                for _ in range(26):
                    yield ("mma", 16, qd)
            yield ("commit", a_empty[s], u)
        main.end = sim.now

    def conv():
        for n in range(P):
            s, u = n % DT, n // DT
            yield ("wait", dt_full[s], u)
            if w:
                yield ("work", LD)
            yield ("arrive", dt_empty[s], u)
            s2, u2 = n % AS, n // AS
            yield ("wait", a_empty[s2], u2 - 1)
            if w:
                yield ("work", ST + 40)
            yield ("arrive", a_full[s2], u2)

    def one():   # single issuer: transposition `ahead` phases before the product
        for n in range(min(args.ahead, P)):
            yield from sel_phase(n)
        for n in range(P):
            if n + args.ahead < P:
                yield from sel_phase(n + args.ahead)
            s, u = n % AS, n // AS
            yield ("wait", a_full[s], u)
            if w:
                for _ in range(26):
                    yield ("mma", 16, qd)
            yield ("commit", a_empty[s], u)
        main.end = sim.now

    def sel_phase(n):
        s, u = n % DT, n // DT
        yield ("wait", dt_empty[s], u - 1)
        if w:
            for _ in range(4):
                yield ("mma", 32, qd)
        yield ("commit", dt_full[s], u)

    if args.issuers == 2:
        sim.spawn(sel(), "sel")
        sim.spawn(main(), "main")
    else:
        sim.spawn(one(), "main")
    for c in range(args.conv):
        sim.spawn(conv(), "conv%d" % c)
    sim.run()
    return sim.now / P


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--planes", type=int, default=40)
    ap.add_argument("--dt", type=int, default=1, help="transposed tiles in flight (phases)")
    ap.add_argument("--aslots", type=int, default=2)
    ap.add_argument("--conv", type=int, default=8)
    ap.add_argument("--issuers", type=int, default=2)
    ap.add_argument("--ahead", type=int, default=1)
    ap.add_argument("--work", type=int, default=1)
    ap.add_argument("--qdepth", type=int, default=8, help="MMAs the issuer may run ahead of the tensor pipe")
    a = ap.parse_args()
    print("clk per phase: %.0f  (tensor work alone: %d)" % (build(a), 4 * 32 + 26 * 16 if a.work else 0))

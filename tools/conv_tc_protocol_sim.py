"""Protocol simulator for csrc/conv_tc.cu (the warp-specialised fused conv forward, not yet run on a GPU): the mbarrier
waits / arrivals / tcgen05.commit completions / TMA completions of one CTA are replayed with randomised timing, with
every shared-memory tile and TMEM accumulator tracked as a resource, so that

  * a parity or ordering mistake shows up as a deadlock (no agent can make progress), and
  * a missing dependency shows up as a hazard (a buffer overwritten while the async proxy may still read it, or read
    before it was produced).

The agent programs below restate the kernel's control flow line by line (same barrier names, same parity expressions);
keep them in sync with the kernel.   python tools/conv_tc_protocol_sim.py [--mode stats|apply] [--trials N]
"""
import argparse
import random

N_CH = 63
WS_RING = 4


class MBar:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.phase = name, count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, f"{self.name}: too many arrivals"
        if self.pending == 0:
            self.pending = self.count
            self.phase ^= 1

    def passed(self, parity):
        # mbarrier.try_wait.parity succeeds when the phase with that parity has completed
        return self.phase != parity


class Sim:
    def __init__(self, mode, rng):
        self.mode, self.rng = mode, rng
        B = {}
        for nm, cnt in (("tile_full", 1), ("tile_empty", 1), ("c1_full", 1), ("c1_empty", 4), ("a1_full", 4), ("a1_empty", 1)):
            B[nm] = [MBar(f"{nm}[{i}]", cnt) for i in range(2)]
        B["ws_full"] = [MBar(f"ws_full[{i}]", 1) for i in range(WS_RING)]
        B["ws_empty"] = [MBar(f"ws_empty[{i}]", 1) for i in range(WS_RING)]
        B["y2_full"] = [MBar("y2_full", 1)]
        self.B = B
        self.async_ops = []            # (remaining delay, callback)
        # resource states: what each buffer currently holds / who may still be reading it
        self.im = [None, None]         # channel whose im2col tile is in the buffer
        self.im_reading = [0, 0]       # outstanding conv UMMAs reading it
        self.c1 = [None, None]         # channel whose conv result is in the TMEM buffer
        self.c1_writing = [0, 0]
        self.a1 = [None, None]
        self.a1_written = [0, 0]       # epilogue warps that have written their rows
        self.a1_reading = [0, 0]
        self.ws = [None] * WS_RING
        self.ws_reading = [0] * WS_RING
        self.y2_count = 0              # channels accumulated into Y2
        self.done_epi = 0

    # ---- async engines ----
    def commit(self, bars, on_done=None):
        """tcgen05.commit: arrives on `bars` once every UMMA issued so far has completed"""
        self.async_ops.append([self.rng.randint(1, 6), bars, on_done])

    def tma(self, bar, on_done):
        self.async_ops.append([self.rng.randint(1, 8), [bar], on_done])

    def tick_async(self):
        # UMMA / commits complete in issue order (the tensor pipe is in-order); TMA completions may be reordered
        if not self.async_ops:
            return False
        self.async_ops[0][0] -= 1
        progressed = False
        while self.async_ops and self.async_ops[0][0] <= 0:
            _, bars, on_done = self.async_ops.pop(0)
            if on_done:
                on_done()
            for b in bars:
                b.arrive()
            progressed = True
        return progressed or bool(self.async_ops)

    # ---- agents (generators yield a predicate to wait for) ----
    def builders(self):
        B = self.B
        for c in range(N_CH):
            bi, n = c & 1, c >> 1
            yield lambda: True                                   # pooled sums (smem scratch private to the builders)
            yield (lambda bi=bi, n=n: B["tile_empty"][bi].passed((n & 1) ^ 1))
            assert self.im_reading[bi] == 0, f"builders overwrite im2col buffer {bi} (channel {c}) while UMMAs read channel {self.im[bi]}"
            self.im[bi] = c
            B["tile_full"][bi].arrive()

    def epilogue(self, q):
        B = self.B
        for c in range(N_CH):
            bi, n = c & 1, c >> 1
            yield (lambda bi=bi, n=n: B["c1_full"][bi].passed(n & 1))
            assert self.c1[bi] == c and self.c1_writing[bi] == 0, f"epilogue {q} reads C1[{bi}] for channel {c}, holds {self.c1[bi]}"
            B["c1_empty"][bi].arrive()
            if self.mode == "apply":
                yield (lambda bi=bi, n=n: B["a1_empty"][bi].passed((n & 1) ^ 1))
                assert self.a1_reading[bi] == 0, f"epilogue {q} overwrites A1[{bi}] (channel {c}) while spatial UMMAs read channel {self.a1[bi]}"
                if self.a1_written[bi] == 0:
                    self.a1[bi] = c
                assert self.a1[bi] == c, f"A1[{bi}] mixes channels {self.a1[bi]} and {c}"
                self.a1_written[bi] += 1
                B["a1_full"][bi].arrive()
        if self.mode == "apply":
            yield lambda: B["y2_full"][0].passed(0)
            assert self.y2_count == N_CH, f"Y2 read after {self.y2_count} channels"
        self.done_epi += 1

    def control(self):
        B = self.B

        def load_ws(c):
            wi, n = c % WS_RING, c // WS_RING
            yield (lambda: B["ws_empty"][wi].passed((n & 1) ^ 1))
            assert self.ws_reading[wi] == 0, f"TMA overwrites Ws ring slot {wi} (channel {c}) while UMMAs read channel {self.ws[wi]}"

            def landed(wi=wi, c=c):
                self.ws[wi] = c
            self.tma(B["ws_full"][wi], landed)

        def spatial(c):
            bi, wi = c & 1, c % WS_RING
            yield (lambda: B["a1_full"][bi].passed((c >> 1) & 1))
            yield (lambda: B["ws_full"][wi].passed((c // WS_RING) & 1))
            assert self.a1[bi] == c and self.a1_written[bi] == 4, f"spatial({c}) reads A1[{bi}] = channel {self.a1[bi]}, {self.a1_written[bi]} warps"
            assert self.ws[wi] == c, f"spatial({c}) reads Ws slot {wi} holding channel {self.ws[wi]}"
            self.a1_reading[bi] += 1
            self.ws_reading[wi] += 1

            def done(bi=bi, wi=wi):
                self.a1_reading[bi] -= 1
                self.a1_written[bi] = 0
                self.ws_reading[wi] -= 1
                self.y2_count += 1
            self.commit([B["a1_empty"][bi], B["ws_empty"][wi]], done)

        if self.mode == "apply":
            yield from load_ws(0)
            yield from load_ws(1)
        for c in range(N_CH):
            bi, n = c & 1, c >> 1
            yield (lambda bi=bi, n=n: B["tile_full"][bi].passed(n & 1))
            yield (lambda bi=bi, n=n: B["c1_empty"][bi].passed((n & 1) ^ 1))
            assert self.im[bi] == c, f"conv({c}) reads im2col buffer {bi} holding channel {self.im[bi]}"
            self.im_reading[bi] += 1
            self.c1_writing[bi] += 1

            def done(bi=bi, c=c):
                self.im_reading[bi] -= 1
                self.c1_writing[bi] -= 1
                self.c1[bi] = c
            self.commit([B["tile_empty"][bi], B["c1_full"][bi]], done)
            if self.mode == "apply":
                if c + 2 < N_CH:
                    yield from load_ws(c + 2)
                if c > 0:
                    yield from spatial(c - 1)
        if self.mode == "apply":
            yield from spatial(N_CH - 1)
            self.commit([B["y2_full"][0]])

    def run(self):
        agents = [("builders", self.builders())] + [(f"epilogue{q}", self.epilogue(q)) for q in range(4)] + \
                 [("control", self.control())]
        waiting = {}
        for name, g in agents:
            waiting[name] = (g, next(g))
        steps = 0
        while waiting:
            steps += 1
            assert steps < 2_000_000, "livelock"
            ready = [n for n, (g, pred) in waiting.items() if pred()]
            if not ready:
                if not self.tick_async():
                    states = {n: "blocked" for n in waiting}
                    raise RuntimeError(f"DEADLOCK with agents {states}")
                continue
            if self.rng.random() < 0.3:
                self.tick_async()
            name = self.rng.choice(ready)
            g, _ = waiting[name]
            try:
                waiting[name] = (g, next(g))
            except StopIteration:
                del waiting[name]
        while self.tick_async():
            pass
        assert self.done_epi == 4
        if self.mode == "apply":
            assert self.y2_count == N_CH


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="both")
    ap.add_argument("--trials", type=int, default=200)
    args = ap.parse_args()
    modes = ["stats", "apply"] if args.mode == "both" else [args.mode]
    for mode in modes:
        for t in range(args.trials):
            Sim(mode, random.Random(t)).run()
        print(f"{mode}: {args.trials} randomised schedules, no deadlock, no hazard")


if __name__ == "__main__":
    main()

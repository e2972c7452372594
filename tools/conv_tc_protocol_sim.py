"""Protocol replay of the mbarrier handshakes of csrc/conv_tc.cu (the backward kernels B1 / B2) on the CPU: every role is
a coroutine issuing the same wait / arrive / commit sequence as the CUDA code, barriers implement the mbarrier phase /
parity / arrival-count rules (a wait on parity P succeeds iff the phase of parity P has completed and the waiter is at
most one phase behind), tcgen05.commit arrives after a random delay.  Finds deadlocks and parity mistakes before a GPU
run.     python tools/conv_tc_protocol_sim.py"""
import random
import sys


class Bar:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.phase = name, count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, f"{self.name}: too many arrivals in phase {self.phase}"
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def test(self, parity):            # mbarrier.try_wait.parity
        return (self.phase & 1) != parity


N_B1_ISSUERS = 2        # EEGB200_B1_ISSUERS (csrc/conv_tc.cu): 2 = dWs UMMAs of even / odd iterations on two threads


def run(mode, gc, n_tiles, seed, verbose=False):
    rnd = random.Random(seed)
    BS = mode == "B1"
    total = n_tiles * gc
    B = {}
    for nm, cnt in (("im_full", 1), ("im_empty", 1), ("dyk_full", 1), ("dyk_empty", 3 if BS else 1), ("c_full", 1),
                    ("c_empty", 16), ("op_full", 16), ("op_empty", 1 if BS else 3), ("gg_full", 1), ("gg_empty", 4),
                    ("dym_full", 1), ("dym_empty", 1)):
        B[nm] = [Bar(f"{nm}[{i}]", cnt) for i in range(2)]
    B["im4_full"] = [Bar(f"im4_full[{i}]", 1) for i in range(4)]
    B["im4_empty"] = [Bar(f"im4_empty[{i}]", 3) for i in range(4)]
    for nm in ("final_a", "final_b", "final_c0", "final_c1"):
        B[nm] = Bar(nm, 1)
    delayed = []            # (time, barrier): tcgen05.commit arrivals
    now = [0]

    def commit(bar):
        delayed.append((now[0] + rnd.randint(1, 40), bar))

    def dec(it):
        tl = it // gc
        return tl, it - tl * gc

    def builder(gb, lanes=1):
        # one coroutine stands for the 128 threads of the group (they move in lock step through the named barriers)
        pending = -1
        if gb == 0 and total > 0:
            yield ("wait", B["dyk_empty"][0], 1)
            pending = 0
        for it in range(gb, total, 2):
            tl, ci = dec(it)
            bi, n = (gb, it >> 1) if BS else (it & 3, it >> 2)
            imf, ime = ("im_full", "im_empty") if BS else ("im4_full", "im4_empty")
            if ci == gc - 1 and tl + 1 < n_tiles:
                tb = (tl + 1) & 1
                yield ("wait", B["dyk_empty"][tb], (((tl + 1) >> 1) & 1) ^ 1)
                pending = tb
            yield ("wait", B[ime][bi], (n & 1) ^ 1)
            if pending >= 0:
                yield ("arrive", B["dyk_full"][pending])
            yield ("arrive", B[imf][bi])
            pending = -1

    def epilogue(ge, w):
        for it in range(total):
            bi, n = it & 1, it >> 1
            yield ("wait", B["c_full"][bi], n & 1)
            yield ("arrive", B["c_empty"][bi])
            if BS:
                yield ("wait", B["op_empty"][bi], (n & 1) ^ 1)
                yield ("arrive", B["op_full"][bi])
            else:
                yield ("wait", B["op_empty"][0], (it & 1) ^ 1)
                yield ("arrive", B["op_full"][0])
        yield ("wait", B["final_a"], 0)
        yield ("wait", B["final_b"], 0)
        yield ("wait", B["final_c0"], 0)
        yield ("wait", B["final_c1"], 0)

    def scatter(w):
        for it in range(total):
            bi, n = it & 1, it >> 1
            yield ("wait", B["gg_full"][bi], n & 1)
            yield ("arrive", B["gg_empty"][bi])

    def ctrl_a():
        for it in range(total):
            tl, ci = dec(it)
            bi, tb, n = it & 1, tl & 1, it >> 1
            if BS:
                yield ("wait", B["im_full"][bi], n & 1)
            else:
                yield ("wait", B["im4_full"][it & 3], (it >> 2) & 1)
            yield ("wait", B["c_empty"][bi], (n & 1) ^ 1)
            yield ("commit", B["im_empty"][bi] if BS else B["im4_empty"][it & 3])
            if ci == 0:
                yield ("wait", B["dyk_full"][tb], (tl >> 1) & 1)
            yield ("commit", B["c_full"][bi])
            if ci == gc - 1:
                yield ("commit", B["dyk_empty"][tb])
        yield ("commit", B["final_a"])

    def ctrl_b(cb):
        # second-stage issuers.  B1: cb 0 / 1 = dWs of the even / odd iterations (the 3-channel group: cb 0 alone);
        # B2: cb 0 = G, cb 1 / 2 = the two halves of the dwt chain
        solo = BS and (gc != 4 or N_B1_ISSUERS == 1)
        for it in range(total):
            tl, ci = dec(it)
            bj = it & 1
            if BS:
                if (cb != 0) if solo else (bj != cb):
                    continue
                if ci == (0 if solo else cb):
                    yield ("wait", B["dyk_full"][tl & 1], (tl >> 1) & 1)
                yield ("wait", B["op_full"][bj], (it >> 1) & 1)
                yield ("commit", B["op_empty"][bj])
                if ci == (gc - 1 if solo else 2 + cb):
                    yield ("commit", B["dyk_empty"][tl & 1])
                    if solo:
                        yield ("commit", B["dyk_empty"][tl & 1])
            else:
                yield ("wait", B["op_full"][0], it & 1)
                if cb == 0:
                    yield ("wait", B["gg_empty"][bj], ((it >> 1) & 1) ^ 1)
                    yield ("commit", B["gg_full"][bj])
                    yield ("commit", B["op_empty"][0])
                else:
                    yield ("commit", B["op_empty"][0])
                    yield ("commit", B["im4_empty"][it & 3])
        yield ("commit", B["final_b"] if cb == 0 else B["final_c0" if cb == 1 else "final_c1"])
        if cb == 0:
            for i in range(2 if BS else 3, 3):
                yield ("commit", B["final_c0" if i == 1 else "final_c1"])

    roles = {"builder0": builder(0), "builder1": builder(1), "ctrlA": ctrl_a()}
    for cb in range(2 if BS else 3):
        roles[f"ctrlB{cb}"] = ctrl_b(cb)
    for ge in range(2):
        for w in range(8):
            roles[f"epi{ge}.{w}"] = epilogue(ge, w)
    if not BS:
        for w in range(4):
            roles[f"scat{w}"] = scatter(w)
    cur = {k: None for k in roles}
    done = set()
    while len(done) < len(roles):
        progressed = False
        now[0] += 1
        for t, bar in [d for d in delayed if d[0] <= now[0]]:
            bar.arrive()
            delayed.remove((t, bar))
            progressed = True
        names = [k for k in roles if k not in done]
        rnd.shuffle(names)
        for k in names:
            for _ in range(rnd.randint(1, 3)):
                if cur[k] is None:
                    try:
                        cur[k] = next(roles[k])
                    except StopIteration:
                        done.add(k)
                        progressed = True
                        break
                op = cur[k]
                if op[0] == "wait":
                    if not op[1].test(op[2]):
                        break
                elif op[0] == "arrive":
                    op[1].arrive()
                else:
                    commit(op[1])
                cur[k] = None
                progressed = True
        if not progressed and not delayed:
            stuck = {k: (cur[k][1].name, cur[k][2], "phase", cur[k][1].phase) for k in roles if k not in done and cur[k]}
            return False, stuck
    return True, None


def run_fwd(n_items, seed, F_LEN=21, RING=4):
    """forward kernel F2 (conv_tc_fwd_kernel<MODE_APPLY>): same engine, roles of the forward protocol"""
    rnd = random.Random(seed)
    total = n_items * F_LEN
    B = {}
    for nm, cnt in (("tile_full", 1), ("tile_empty", 1), ("c1_full", 1), ("c1_empty", 16), ("a1_full", 16), ("a1_empty", 1),
                    ("y2_full", 1), ("y2_empty", 16)):
        B[nm] = [Bar(f"{nm}[{i}]", cnt) for i in range(2)]
    B["ws_full"] = [Bar(f"ws_full[{i}]", 1) for i in range(RING)]
    B["ws_empty"] = [Bar(f"ws_empty[{i}]", 1) for i in range(RING)]
    delayed, now = [], [0]

    def commit(bar):
        delayed.append((now[0] + rnd.randint(1, 40), bar))

    def builder(gb):
        for it in range(gb, total, 2):
            yield ("wait", B["tile_empty"][gb], ((it >> 1) & 1) ^ 1)
            yield ("arrive", B["tile_full"][gb])

    def epilogue(ge, w):
        for it in range(total):
            bi, n, item, l = it & 1, it >> 1, it // F_LEN, it % F_LEN
            yield ("wait", B["c1_full"][bi], n & 1)
            yield ("arrive", B["c1_empty"][bi])
            yield ("wait", B["a1_empty"][bi], (n & 1) ^ 1)
            yield ("arrive", B["a1_full"][bi])
            if l == F_LEN - 1:
                ib = item & 1
                yield ("wait", B["y2_full"][ib], (item >> 1) & 1)
                yield ("arrive", B["y2_empty"][ib])

    def ctrl_a():
        for it in range(total):
            bi, n = it & 1, it >> 1
            yield ("wait", B["tile_full"][bi], n & 1)
            yield ("wait", B["c1_empty"][bi], (n & 1) ^ 1)
            yield ("commit", B["tile_empty"][bi])
            yield ("commit", B["c1_full"][bi])

    def ctrl_b():
        def load(it):
            wi, n = it % RING, it // RING
            yield ("wait", B["ws_empty"][wi], (n & 1) ^ 1)
            yield ("commit", B["ws_full"][wi])          # TMA completion = delayed arrival
        for it in range(min(3, total)):
            yield from load(it)
        for it in range(total):
            bi, wi, item, l = it & 1, it % RING, it // F_LEN, it % F_LEN
            ib = item & 1
            if l == 0:
                yield ("wait", B["y2_empty"][ib], ((item >> 1) & 1) ^ 1)
            yield ("wait", B["a1_full"][bi], (it >> 1) & 1)
            yield ("wait", B["ws_full"][wi], (it // RING) & 1)
            yield ("commit", B["a1_empty"][bi])
            yield ("commit", B["ws_empty"][wi])
            if l == F_LEN - 1:
                yield ("commit", B["y2_full"][ib])
            if it + 3 < total:
                yield from load(it + 3)

    roles = {"builder0": builder(0), "builder1": builder(1), "ctrlA": ctrl_a(), "ctrlB": ctrl_b()}
    for ge in range(2):
        for w in range(8):
            roles[f"epi{ge}.{w}"] = epilogue(ge, w)
    return _drive(roles, delayed, now, commit, rnd)


def _drive(roles, delayed, now, commit, rnd):
    cur = {k: None for k in roles}
    done = set()
    while len(done) < len(roles):
        progressed = False
        now[0] += 1
        for t, bar in [d for d in delayed if d[0] <= now[0]]:
            bar.arrive()
            delayed.remove((t, bar))
            progressed = True
        names = [k for k in roles if k not in done]
        rnd.shuffle(names)
        for k in names:
            for _ in range(rnd.randint(1, 3)):
                if cur[k] is None:
                    try:
                        cur[k] = next(roles[k])
                    except StopIteration:
                        done.add(k)
                        progressed = True
                        break
                op = cur[k]
                if op[0] == "wait":
                    if not op[1].test(op[2]):
                        break
                elif op[0] == "arrive":
                    op[1].arrive()
                else:
                    commit(op[1])
                cur[k] = None
                progressed = True
        if not progressed and not delayed:
            return False, {k: (cur[k][1].name, cur[k][2], "phase", cur[k][1].phase) for k in roles if k not in done and cur[k]}
    return True, None


if __name__ == "__main__":
    bad = 0
    for mode in ("B1", "B2"):
        for gc in (4, 3):
            for n_tiles in (1, 2, 3, 4, 7, 12, 38):
                for seed in range(60):
                    ok, stuck = run(mode, gc, n_tiles, seed)
                    if not ok:
                        bad += 1
                        print(f"DEADLOCK mode={mode} gc={gc} tiles={n_tiles} seed={seed}")
                        for k, v in sorted(stuck.items()):
                            print("   ", k, v)
                        break
    for n_items in (1, 2, 3, 7):
        for seed in range(40):
            ok, stuck = run_fwd(n_items, seed)
            if not ok:
                bad += 1
                print(f"DEADLOCK forward items={n_items} seed={seed}")
                for k, v in sorted(stuck.items()):
                    print("   ", k, v)
                break
    print("protocol sim:", "FAIL" if bad else "PASS")
    sys.exit(1 if bad else 0)

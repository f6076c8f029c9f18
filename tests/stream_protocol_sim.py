"""CPU model + adversarial simulation of the streaming kernel's scheduling (tfoptflow_b200/csrc/fwd_stream.cuh).

`segments` / `lane_tasks` restate the kernel's segment construction and lane -> work mapping.  `simulate` runs the
ring protocol with exact mbarrier parity semantics - try_wait.parity(P) succeeds iff P is NOT the parity of the
barrier's current (incomplete) phase, so a waiter two phases ahead passes prematurely - under random schedules with
long stalls and checks that nothing deadlocks, no consumer reads a slot that does not hold its slab / pair, and no
producer overwrites data that still has readers.  Barriers are indexed j % (BARMUL * N)."""
import collections
import random

GS = 4


def segments(B, H, W, TW, NXG, grid, cta):
    NH = 16 // NXG
    TPG = 8 * GS // NH + 1
    NPAIR, nsx = (H + 1) // 2, (W + TW - 1) // TW
    Q = B * nsx * NPAIR
    qlo, qhi = Q * cta // grid, Q * (cta + 1) // grid
    segs, q, task, j, a = [], qlo, 0, 0, 0
    while q < qhi:
        strip = q // NPAIR
        p0 = q - strip * NPAIR
        rem = qhi - q
        p1 = NPAIR if NPAIR - p0 < rem else p0 + rem
        s = dict(b=strip // nsx, x0=(strip % nsx) * TW, p0=p0, p1=p1, nslab=p1 - p0 + 5, task0=task, j0=j, a0=a, H=H)
        task += TPG * ((s["nslab"] + GS - 1) // GS)
        j += s["nslab"]
        a += p1 - p0
        segs.append(s)
        q += p1 - p0
    return segs, task


def lane_tasks(sg, tl, NXG):
    """The 32 lane-tasks of segment-local task `tl` exactly as the kernel decodes them."""
    NH = 16 // NXG
    NF = 8 * GS // NH
    TPG, KP = NF + 1, 4 // NXG
    g, rr = divmod(tl, TPG)
    is_edge = rr == NF
    out = []
    if not is_edge and GS * g + ((rr * NH) >> 3) >= sg["nslab"]:
        return out
    for lane in range(32):
        lo = lane & 1
        if not is_edge:
            # full tasks: (warped row, c1 row, x-group) so that every aligned group of 4 lanes shares operands
            hi, xg = (lane >> 1) % NH, lane // (2 * NH)
            c = rr * NH + hi
            l = GS * g + (c >> 3)
            r8 = c & 7
            k = sg["p0"] - 2 + l - 2 + (r8 >> 1)
            sub, bsel, part = r8 & 1, lo, 0
            arrive_b = lo == 0 and xg == 0 and (hi == 0 or r8 == 0)
            arrive_a = lo == 0 and xg == 0 and sub == 0
            kfrac = 1
        else:
            xg, hi = (lane >> 1) % NXG, lane // (2 * NXG)
            ms, part = divmod(hi, KP)
            l = GS * g + ms
            k = sg["p0"] - 2 + l + (2 if lo == 0 else -3)
            sub, bsel = lo, 1 - lo
            arrive_b = lo == 0 and xg == 0 and part == 0
            arrive_a = xg == 0 and part == 0
            kfrac = 1
        slab_ok = l < sg["nslab"]
        m = sg["p0"] - 2 + l
        pair_ok = slab_ok and sg["p0"] <= k < sg["p1"]
        ya, yb = 2 * k + sub, 2 * m - 1 + bsel
        out.append(dict(l=l, k=k, xg=xg, ya=ya, yb=yb, slab_ok=slab_ok, pair_ok=pair_ok,
                        arrive_b=slab_ok and arrive_b, arrive_a=pair_ok and arrive_a,
                        store=pair_ok and ya < sg["H"], writer=part == 0, kfrac=kfrac))
    return out


class MBar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def try_wait(self, parity):
        return (self.phase & 1) != parity


def segment_tasks(npairs, NXG):
    """[(per-slab arrivals, per-pair arrivals, slabs waited on, pairs waited on)] of one segment in grab order."""
    sg = dict(b=0, x0=0, p0=0, p1=npairs, nslab=npairs + 5, task0=0, j0=0, a0=0, H=2 * npairs)
    NH = 16 // NXG
    TPG = 8 * GS // NH + 1
    tasks = []
    for tl in range(TPG * ((sg["nslab"] + GS - 1) // GS)):
        lts = lane_tasks(sg, tl, NXG)
        if not lts:
            continue
        sl, pr = collections.Counter(), collections.Counter()
        for lt in lts:
            if lt["slab_ok"]:
                sl[lt["l"]] += 1 if lt["arrive_b"] else 0
            if lt["pair_ok"]:
                pr[lt["k"]] += 1 if lt["arrive_a"] else 0
        tasks.append((dict(sl), dict(pr)))
    return sg["nslab"], tasks


def simulate(npairs, NSLAB, NAP, NPW, NCW, NXG=4, BARMUL=2, seed=0, stall_p=0.3, max_steps=3_000_000):
    rnd = random.Random(seed)
    nslab, tasks = segment_tasks(npairs, NXG)
    nb, na = BARMUL * NSLAB, BARMUL * NAP
    eb_count = 3 if NXG == 4 else 2
    full_b = [MBar(1) for _ in range(nb)]
    empty_b = [MBar(eb_count) for _ in range(nb)]
    full_a = [MBar(1) for _ in range(na)]
    empty_a = [MBar(6) for _ in range(na)]
    slot_b, slot_a = [None] * NSLAB, [None] * NAP
    readers = collections.Counter()
    tot_b, tot_a = collections.Counter(), collections.Counter()
    for sl, pr in tasks:
        for j, c in sl.items():
            tot_b[j] += c
        for a, c in pr.items():
            tot_a[a] += c
    assert set(tot_b.values()) == {eb_count} and set(tot_a.values()) == {6}, (set(tot_b.values()), set(tot_a.values()))
    prod = []
    for w in range(NPW):
        items = []
        for l in range(nslab):
            if l % NPW == w:
                if l < npairs:
                    items.append(("A", l))
                items.append(("B", l))
        prod.append(items)
    ppos, stall = [0] * NPW, [0] * (NPW + NCW)
    cons, next_task, done = [None] * NCW, 0, 0
    for _ in range(max_steps):
        if done == len(tasks) and all(ppos[w] == len(prod[w]) for w in range(NPW)):
            return "ok"
        ag = rnd.randrange(NPW + NCW)
        if stall[ag] > 0:
            stall[ag] -= 1
            continue
        if rnd.random() < stall_p:
            stall[ag] = rnd.randrange(1, 400)
            continue
        if ag < NPW:
            w = ag
            if ppos[w] >= len(prod[w]):
                continue
            kind, idx = prod[w][ppos[w]]
            N, full, empty, slots, tot = (NAP, full_a, empty_a, slot_a, tot_a) if kind == "A" else \
                                         (NSLAB, full_b, empty_b, slot_b, tot_b)
            prev = idx - N
            if prev >= 0 and not empty[prev % (BARMUL * N)].try_wait((prev // (BARMUL * N)) & 1):
                continue
            s = idx % N
            if slots[s] is not None and readers[(kind, slots[s])] > 0:
                return f"hazard: producer overwrites {kind}{slots[s]} with {kind}{idx} while it has readers"
            slots[s] = idx
            readers[(kind, idx)] = tot[idx]
            full[idx % (BARMUL * N)].arrive()
            ppos[w] += 1
        else:
            c = ag - NPW
            if cons[c] is None:
                if next_task >= len(tasks):
                    continue
                cons[c] = next_task
                next_task += 1
            sl, pr = tasks[cons[c]]
            if not (all(full_b[j % nb].try_wait((j // nb) & 1) for j in sl) and
                    all(full_a[a % na].try_wait((a // na) & 1) for a in pr)):
                continue
            for j in sl:
                if slot_b[j % NSLAB] != j:
                    return f"hazard: task {cons[c]} reads slab slot {j % NSLAB} expecting {j}, holds {slot_b[j % NSLAB]}"
            for a in pr:
                if slot_a[a % NAP] != a:
                    return f"hazard: task {cons[c]} reads pair slot {a % NAP} expecting {a}, holds {slot_a[a % NAP]}"
            for j, cnt in sl.items():
                for _ in range(cnt):
                    empty_b[j % nb].arrive()
                readers[("B", j)] -= cnt
            for a, cnt in pr.items():
                for _ in range(cnt):
                    empty_a[a % na].arrive()
                readers[("A", a)] -= cnt
            cons[c] = None
            done += 1
    return "deadlock"

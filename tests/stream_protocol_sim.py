"""Adversarial simulation of the streaming kernel's ring protocol (tfoptflow_b200/csrc/fwd_stream.cuh) with exact
mbarrier parity semantics: try_wait.parity(P) succeeds iff P is NOT the parity of the barrier's current (incomplete)
phase - so a waiter that is two phases ahead passes prematurely.  The simulation checks, under random schedules
with long random stalls, that (a) nothing deadlocks, (b) no consumer reads a slot that does not hold its slab /
pair, (c) no producer overwrites data that still has readers.

Barriers are indexed j % (BARMUL * N) (BARMUL = 1: one barrier per data slot; 2: two alternating barrier sets)."""
import collections
import random


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


def segment_tasks(npairs):
    """(slab set, pair set, per-slab arrivals, per-pair arrivals) of one segment in the kernel's grab order."""
    nslab = npairs + 5
    tasks = []
    for g in range((nslab + 3) // 4):
        for rr in range(9):
            sl, pr = collections.Counter(), collections.Counter()
            for lane in range(32):
                lo, hi = lane & 1, lane >> 3
                if rr < 8:
                    l = 4 * g + (rr >> 1)
                    if l >= nslab:
                        break
                    ki = l - 4 + 2 * (rr & 1) + (hi >> 1)
                    arr_b, arr_a = lane == 0, lane & 15 == 0
                else:
                    l = 4 * g + hi
                    ki = l - 2 + (2 if lo == 0 else -3)
                    arr_b, arr_a = lane & 7 == 0, lane & 6 == 0
                if l < nslab:
                    sl[l] += 1 if arr_b else 0
                    if 0 <= ki < npairs:
                        pr[ki] += 1 if arr_a else 0
            if sl:
                tasks.append((dict(sl), dict(pr)))
    return nslab, tasks


def simulate(npairs, NSLAB, NAP, NPW, NCW, BARMUL=2, seed=0, stall_p=0.3, max_steps=2_000_000):
    rnd = random.Random(seed)
    nslab, tasks = segment_tasks(npairs)
    nb, na = BARMUL * NSLAB, BARMUL * NAP
    full_b = [MBar(1) for _ in range(nb)]
    empty_b = [MBar(3) for _ in range(nb)]
    full_a = [MBar(1) for _ in range(na)]
    empty_a = [MBar(6) for _ in range(na)]
    slot_b, slot_a = [None] * NSLAB, [None] * NAP      # which slab / pair a data slot holds
    readers_b, readers_a = collections.Counter(), collections.Counter()  # outstanding readers of the current content
    tot_b, tot_a = collections.Counter(), collections.Counter()
    for sl, pr in tasks:
        for j, c in sl.items():
            tot_b[j] += c
        for a, c in pr.items():
            tot_a[a] += c
    assert set(tot_b.values()) == {3} and set(tot_a.values()) == {6}
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
    for step in range(max_steps):
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
            N, full, empty, slots, readers = (NAP, full_a, empty_a, slot_a, readers_a) if kind == "A" else \
                                             (NSLAB, full_b, empty_b, slot_b, readers_b)
            prev = idx - N
            if prev >= 0:
                nbar = BARMUL * N
                if not empty[prev % nbar].try_wait((prev // nbar) & 1):
                    continue
            s = idx % N
            if slots[s] is not None and readers[(kind, slots[s])] > 0:
                return f"hazard: producer overwrites {kind}{slots[s]} with {kind}{idx} while it has readers"
            slots[s] = idx
            readers[(kind, idx)] = (tot_a if kind == "A" else tot_b)[idx]
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
            ok = all(full_b[j % nb].try_wait((j // nb) & 1) for j in sl) and \
                all(full_a[a % na].try_wait((a // na) & 1) for a in pr)
            if not ok:
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
                readers_b[("B", j)] -= cnt
            for a, cnt in pr.items():
                for _ in range(cnt):
                    empty_a[a % na].arrive()
                readers_a[("A", a)] -= cnt
            cons[c] = None
            done += 1
    return "deadlock"

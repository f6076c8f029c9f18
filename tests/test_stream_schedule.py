"""CPU model of the streaming kernel's work decomposition (tfoptflow_b200/csrc/fwd_stream.cuh): segments,
slab / pair rings, full tasks F and edge tasks E.  Checks that every (pixel, dy) of the cost volume is produced
exactly once from the right (c1 row, warped row) combination and that every ring slot gets exactly the number of
consumer arrivals its mbarrier was initialised with (3 per slab, 6 per pair)."""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import pytest

TW, PX = 32, 8


TPG = 9  # tasks per group of 4 slabs: Fa, Fb per slab + one edge task


def segments(B, H, W, grid, cta):
    NPAIR, nsx = (H + 1) // 2, (W + TW - 1) // TW
    Q = B * nsx * NPAIR
    qlo, qhi = Q * cta // grid, Q * (cta + 1) // grid
    segs, q, task, j, a = [], qlo, 0, 0, 0
    while q < qhi:
        strip = q // NPAIR
        p0 = q - strip * NPAIR
        rem = qhi - q
        p1 = NPAIR if NPAIR - p0 < rem else p0 + rem
        s = dict(b=strip // nsx, x0=(strip % nsx) * TW, p0=p0, p1=p1, nslab=p1 - p0 + 5, task0=task, j0=j, a0=a)
        task += TPG * ((s["nslab"] + 3) // 4)
        j += s["nslab"]
        a += p1 - p0
        segs.append(s)
        q += p1 - p0
    return segs, task


def run_cta(B, H, W, grid, cta, written, slab_arr, pair_arr):
    segs, total = segments(B, H, W, grid, cta)
    for t in range(total):
        sg = [s for s in segs if s["task0"] <= t][-1]
        tl = t - sg["task0"]
        g, rr = divmod(tl, TPG)
        for lane in range(32):
            xg, lo, hi = (lane >> 1) & 3, lane & 1, lane >> 3
            if rr < 8:
                l = 4 * g + (rr >> 1)
                k = sg["p0"] - 2 + l - 2 + 2 * (rr & 1) + (hi >> 1)
                sub, bsel = hi & 1, lo
                if l >= sg["nslab"]:
                    break
            else:
                l = 4 * g + hi
                k = sg["p0"] - 2 + l + (2 if lo == 0 else -3)
                sub, bsel = lo, 1 - lo
            slab_ok = l < sg["nslab"]
            m = sg["p0"] - 2 + l
            pair_ok = slab_ok and sg["p0"] <= k < sg["p1"]
            if rr < 8:
                if lane == 0:
                    slab_arr[(cta, sg["j0"] + l)] += 1
                if pair_ok and lane & 15 == 0:
                    pair_arr[(cta, sg["a0"] + k - sg["p0"])] += 1
            else:
                if slab_ok and lane & 7 == 0:
                    slab_arr[(cta, sg["j0"] + l)] += 1
                if pair_ok and lane & 6 == 0:
                    pair_arr[(cta, sg["a0"] + k - sg["p0"])] += 1
            yb, ya = 2 * m - 1 + bsel, 2 * k + sub
            if not (pair_ok and ya < H):
                continue
            dyi = yb - ya + 4
            assert 0 <= dyi <= 8
            for p in range(PX):
                x = sg["x0"] + xg * PX + p
                if x < W:
                    written[(sg["b"], ya, x, dyi)] += 1
    return segs


@pytest.mark.parametrize("B,H,W,grid", [(2, 12, 40, 3), (1, 7, 33, 5), (3, 10, 64, 7), (1, 112, 32, 4), (2, 5, 8, 1),
                                        (1, 2, 2, 1), (4, 6, 70, 11)])
def test_stream_schedule_covers_every_output_once(B, H, W, grid):
    written, slab_arr, pair_arr = collections.Counter(), collections.Counter(), collections.Counter()
    nslab = npair = 0
    for cta in range(grid):
        segs = run_cta(B, H, W, grid, cta, written, slab_arr, pair_arr)
        nslab += sum(s["nslab"] for s in segs)
        npair += sum(s["p1"] - s["p0"] for s in segs)
    assert len(written) == B * H * W * 9 and set(written.values()) == {1}
    assert len(slab_arr) == nslab and set(slab_arr.values()) == {3}
    assert len(pair_arr) == npair and set(pair_arr.values()) == {6}
    assert npair == B * ((W + TW - 1) // TW) * ((H + 1) // 2)


@pytest.mark.parametrize("cfg", [(7, 12, 8, 8), (7, 12, 4, 8), (12, 18, 8, 8)])
def test_ring_protocol_has_no_deadlock_or_hazard(cfg):
    """Random adversarial schedules with long stalls against exact mbarrier parity semantics (two barrier sets)."""
    from stream_protocol_sim import simulate
    for seed in range(4):
        assert simulate(40, *cfg, BARMUL=2, seed=seed) == "ok"


def test_single_barrier_set_is_unsafe():
    """Documents why the kernel uses two alternating barrier sets: one set lets a consumer pass a parity wait early."""
    from stream_protocol_sim import simulate
    res = [simulate(40, 7, 12, 8, 8, BARMUL=1, seed=s) for s in range(6)]
    assert any(r.startswith("hazard") for r in res)

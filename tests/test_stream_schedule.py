"""CPU model of the streaming kernel's work decomposition (tfoptflow_b200/csrc/fwd_stream.cuh): segments,
slab / pair rings, full tasks F and edge tasks E.  Checks that every (pixel, dy) of the cost volume is produced
exactly once from the right (c1 row, warped row) combination and that every ring slot gets exactly the number of
consumer arrivals its mbarrier was initialised with (2 per slab, 6 per pair)."""
import collections

import pytest

TW, PX = 32, 8


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
        task += 5 * ((s["nslab"] + 3) // 4)
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
        g, rr = divmod(tl, 5)
        if rr < 4:
            l = 4 * g + rr
            if l >= sg["nslab"]:
                continue
            m = sg["p0"] - 2 + l
            slab_arr[(cta, sg["j0"] + l)] += 1
            for lane in range(32):
                xg, bsel, pk = lane & 3, (lane >> 2) & 1, lane >> 3
                k = m - 2 + pk
                if not (sg["p0"] <= k < sg["p1"]):
                    continue
                if lane & 7 == 0:
                    pair_arr[(cta, sg["a0"] + k - sg["p0"])] += 1
                yb = 2 * m - 1 + bsel
                for s in range(2):
                    ya = 2 * k + s
                    dyi = 7 - 2 * pk - s + bsel
                    assert yb - ya == dyi - 4
                    for p in range(PX):
                        x = sg["x0"] + xg * PX + p
                        if ya < H and x < W:
                            written[(sg["b"], ya, x, dyi)] += 1
        else:
            for lane in range(32):
                xg, half, ms = lane & 3, (lane >> 2) & 1, lane >> 3
                l = 4 * g + ms
                if l >= sg["nslab"]:
                    continue
                m = sg["p0"] - 2 + l
                if lane & 7 == 0:
                    slab_arr[(cta, sg["j0"] + l)] += 1
                k = m + 2 if half == 0 else m - 3
                if not (sg["p0"] <= k < sg["p1"]):
                    continue
                if lane & 3 == 0:
                    pair_arr[(cta, sg["a0"] + k - sg["p0"])] += 1
                ya, yb, dyi = 2 * k + half, 2 * m - 1 + (1 - half), (0 if half == 0 else 8)
                assert yb - ya == dyi - 4
                for p in range(PX):
                    x = sg["x0"] + xg * PX + p
                    if ya < H and x < W:
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
    assert len(slab_arr) == nslab and set(slab_arr.values()) == {2}
    assert len(pair_arr) == npair and set(pair_arr.values()) == {6}
    assert npair == B * ((W + TW - 1) // TW) * ((H + 1) // 2)

"""CPU model of the streaming kernel's work decomposition (tfoptflow_b200/csrc/fwd_stream.cuh): segments,
slab / pair rings, full tasks and edge tasks for the three strip widths (NXG = 4, 2, 1).  Checks that every
(pixel, dy) of the cost volume is produced exactly once from the right (c1 row, warped row) combination and that
every ring slot gets exactly the number of consumer arrivals its mbarrier was initialised with.  The ring
protocol itself is exercised by an adversarial simulation with exact mbarrier parity semantics."""
import collections
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from stream_protocol_sim import lane_tasks, segments, simulate  # noqa: E402

def run_cta(B, H, W, NXG, grid, cta, written, slab_arr, pair_arr, PX=8):
    TW = PX * NXG
    segs, total = segments(B, H, W, TW, NXG, grid, cta)
    for t in range(total):
        sg = [s for s in segs if s["task0"] <= t][-1]
        for lt in lane_tasks(sg, t - sg["task0"], NXG):
            if lt["arrive_b"]:
                slab_arr[(cta, sg["j0"] + lt["l"])] += 1
            if lt["arrive_a"]:
                pair_arr[(cta, sg["a0"] + lt["k"] - sg["p0"])] += 1
            if not lt["store"] or not lt["writer"]:
                continue
            ya, yb = lt["ya"], lt["yb"]
            dyi = yb - ya + 4
            assert 0 <= dyi <= 8 and ya < H
            for p in range(PX):
                x = sg["x0"] + lt["xg"] * PX + p
                if x < W:
                    written[(sg["b"], ya, x, dyi)] += lt["kfrac"]
    return segs


@pytest.mark.parametrize("NXG,PX", [(4, 8), (4, 6), (4, 4), (2, 6), (2, 8), (1, 8), (1, 6)])
@pytest.mark.parametrize("B,H,W,grid", [(2, 12, 40, 3), (1, 7, 33, 5), (3, 10, 64, 7), (1, 112, 32, 4), (2, 5, 8, 1),
                                        (1, 2, 2, 1), (4, 6, 70, 11)])
def test_stream_schedule_covers_every_output_once(B, H, W, grid, NXG, PX):
    written, slab_arr, pair_arr = collections.Counter(), collections.Counter(), collections.Counter()
    nslab = npair = 0
    for cta in range(grid):
        segs = run_cta(B, H, W, NXG, grid, cta, written, slab_arr, pair_arr, PX)
        nslab += sum(s["nslab"] for s in segs)
        npair += sum(s["p1"] - s["p0"] for s in segs)
    assert len(written) == B * H * W * 9 and set(written.values()) == {1}
    assert len(slab_arr) == nslab and set(slab_arr.values()) == {3 if NXG == 4 else 2}
    assert len(pair_arr) == npair and set(pair_arr.values()) == {6}
    assert npair == B * ((W + PX * NXG - 1) // (PX * NXG)) * ((H + 1) // 2)


@pytest.mark.parametrize("cfg", [(7, 12, 4, 8, 4), (6, 12, 4, 8, 2), (7, 12, 4, 8, 1), (6, 10, 4, 8, 1), (7, 12, 8, 8, 4)])
def test_ring_protocol_has_no_deadlock_or_hazard(cfg):
    """Random adversarial schedules with long stalls against exact mbarrier parity semantics (two barrier sets)."""
    NSLAB, NAP, NPW, NCW, NXG = cfg
    for seed in range(3):
        assert simulate(40, NSLAB, NAP, NPW, NCW, NXG=NXG, BARMUL=2, seed=seed) == "ok"


def test_single_barrier_set_is_unsafe():
    """Documents why the kernel uses two alternating barrier sets: one set lets a consumer pass a parity wait early."""
    res = [simulate(40, 7, 12, 8, 8, NXG=4, BARMUL=1, seed=s) for s in range(6)]
    assert any(r.startswith("hazard") for r in res)

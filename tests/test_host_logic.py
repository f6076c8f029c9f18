"""CPU tests of the host-side logic: synthetic inputs, batch sharding and the gradient all-reduce that
replaces tfoptflow/multi_gpus.py:average_gradients (world_size 2 over gloo)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tfoptflow_b200 import multi_gpus, pyramid, synth


def test_flow_scalers_match_reference():
    # tfoptflow/model_pwcnet.py:1560 "scaler values are 0.625, 1.25, 2.5, 5.0"
    assert [pyramid.flow_scaler(l) for l in (5, 4, 3, 2)] == [0.625, 1.25, 2.5, 5.0]


def test_level_shapes_cfg():
    assert synth.level_shapes(448, 1024) == [(6, 7, 16, 196), (5, 14, 32, 128), (4, 28, 64, 96), (3, 56, 128, 64),
                                             (2, 112, 256, 32)]
    assert synth.level_shapes(384, 512)[0] == (6, 6, 8, 196)


def test_synth_is_deterministic_and_top_level_has_no_flow():
    a = synth.make_pyramid(2, 128, 192, np.float32)
    b = synth.make_pyramid(2, 128, 192, np.float32)
    assert a[0]["flow"] is None and a[0]["lvl"] == 6 and a[-1]["lvl"] == 2
    for x, y in zip(a, b):
        assert np.array_equal(x["c1"], y["c1"]) and np.array_equal(x["c2"], y["c2"])
        if x["flow"] is not None:
            assert np.array_equal(x["flow"], y["flow"])
            assert x["scaler"] == pyramid.flow_scaler(x["lvl"])


def test_shard_batch_is_a_partition():
    for bs in (32, 8, 5, 1):
        for world in (1, 2, 3, 4, 8):
            cuts = [multi_gpus.shard_batch(bs, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == bs
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    r, w = multi_gpus.init_from_env("gloo")
    assert (r, w) == (rank, world)
    g = torch.Generator().manual_seed(rank)
    grads = [torch.randn(7, 3, generator=g), None, torch.randn(1000, generator=g), torch.randn(2, 2, 2, generator=g).half()]
    ref = []
    for i in range(len(grads)):
        if grads[i] is None:
            ref.append(None)
            continue
        acc = 0
        for rr in range(world):
            gg = torch.Generator().manual_seed(rr)
            tower = [torch.randn(7, 3, generator=gg), None, torch.randn(1000, generator=gg),
                     torch.randn(2, 2, 2, generator=gg).half()]
            acc = acc + tower[i].float()
        ref.append(acc / world)
    out = multi_gpus.average_gradients(grads, bucket_bytes=2048)  # forces several buckets
    ok = all((o is None and e is None) or torch.allclose(o.float(), e, atol=2e-3) for o, e in zip(out, ref))
    # persistent bucket: per-rank unscale (1/128) + clip to global norm 5.0 BEFORE the mean (model_pwcnet.py:311-314)
    def tower(rr):
        gg = torch.Generator().manual_seed(100 + rr)
        return [torch.randn(50, 4, generator=gg) * 128 * 3, None, torch.randn(333, generator=gg) * 128 * 3]
    mine = tower(rank)
    bucket = multi_gpus.GradientBucket(mine)
    bucket.average(mine, loss_scale=128.0, clip_norm=5.0)
    exp = None
    for rr in range(world):
        t = [x / 128.0 for x in tower(rr) if x is not None]
        gn = torch.sqrt(sum((x.double() ** 2).sum() for x in t)).float()
        t = [x * (5.0 / torch.clamp(gn, min=5.0)) for x in t]
        exp = t if exp is None else [e + x for e, x in zip(exp, t)]
    exp = [e / world for e in exp]
    ok = ok and mine[1] is None and all(torch.allclose(o, e, atol=1e-5) for o, e in zip([mine[0], mine[2]], exp))
    # batch sharding covers the batch exactly once across ranks
    lo, hi = multi_gpus.shard_batch(9, rank, world)
    cnt = torch.zeros(9)
    cnt[lo:hi] += 1
    dist.all_reduce(cnt)
    ok = ok and bool((cnt == 1).all())
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_average_gradients_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=100) for _ in ps]
    for p in ps:
        p.join(timeout=30)
    assert sorted(res) == [(0, True), (1, True)]


def test_average_gradients_single_process_is_identity():
    g = [torch.ones(3), None]
    assert multi_gpus.average_gradients(g) is g

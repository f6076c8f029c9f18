"""Known-answer vectors of SURVEY.md section 8c for oracle/pwc_oracle.py, and the committed golden fixtures
(tests/golden/*.npz, generated from the reference's own source by tests/golden/make_golden_from_reference.py).
CPU only.  The bit-for-bit pin against the imported reference is tests/test_oracle_vs_reference.py."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import pwc_oracle as O
from oracle import ref_graph_torch as T
from tfoptflow_b200 import synth


def _rand(shape, seed, dtype=np.float32):
    return np.random.default_rng(seed).standard_normal(shape).astype(dtype)


# (1) zero flow => identity, bit exact away from the last row/col, <= 1ulp there
def test_zero_flow_identity():
    img = _rand((2, 7, 9, 5), 0)
    out = O.dense_image_warp(img, np.zeros((2, 7, 9, 2), np.float32))
    assert np.array_equal(out[:, :-1, :-1], img[:, :-1, :-1])
    np.testing.assert_allclose(out, img, rtol=2e-7, atol=1e-7)


# (2) channel 0 displaces rows, sign is minus
def test_unit_flow_shifts_rows_then_cols():
    img = _rand((1, 6, 8, 3), 1)
    f = np.zeros((1, 6, 8, 2), np.float32)
    f[..., 0] = 1.0
    out = O.dense_image_warp(img, f)
    np.testing.assert_allclose(out[:, 1:], img[:, :-1], rtol=2e-7, atol=1e-7)
    np.testing.assert_allclose(out[:, 0], img[:, 0], rtol=2e-7, atol=1e-7)  # row 0 replicated
    f = np.zeros((1, 6, 8, 2), np.float32)
    f[..., 1] = 1.0
    out = O.dense_image_warp(img, f)
    np.testing.assert_allclose(out[:, :, 1:], img[:, :, :-1], rtol=2e-7, atol=1e-7)
    np.testing.assert_allclose(out[:, :, 0], img[:, :, 0], rtol=2e-7, atol=1e-7)


# (3) huge flow => clamped border pixel
def test_huge_flow_clamps_to_border():
    img = _rand((1, 5, 6, 2), 2)
    out = O.dense_image_warp(img, np.full((1, 5, 6, 2), 1e4, np.float32))
    assert np.allclose(out, img[:, :1, :1])
    out = O.dense_image_warp(img, np.full((1, 5, 6, 2), -1e4, np.float32))
    assert np.allclose(out, img[:, -1:, -1:])


# (4) independent second oracle: grid_sample(border, align_corners=True)
def test_warp_matches_grid_sample():
    d = synth.make_level(2, 13, 17, 6, 3, np.float32, "smooth", sigma=3.0)
    img, flow = d["c2"], d["flow"]
    out = O.dense_image_warp(img, flow)
    b, h, w, c = img.shape
    gy, gx = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    qy, qx = gy[None] - flow[..., 0], gx[None] - flow[..., 1]
    grid = torch.from_numpy(np.stack([2 * qx / (w - 1) - 1, 2 * qy / (h - 1) - 1], -1).astype(np.float32))
    ref = torch.nn.functional.grid_sample(torch.from_numpy(img).permute(0, 3, 1, 2), grid, mode="bilinear",
                                          padding_mode="border", align_corners=True).permute(0, 2, 3, 1).numpy()
    np.testing.assert_allclose(out, ref, atol=2e-5)


# (5) centre channel of cost_volume(c1, c1) is mean_c(c1^2)
def test_centre_channel_is_mean_square():
    c1 = _rand((1, 6, 8, 16), 4)
    cv = O.cost_volume(c1, c1, 4)
    np.testing.assert_allclose(cv[..., 40], np.mean(c1 * c1, -1), rtol=1e-6)


# (6) all-ones: pins zero padding AFTER warping and the channel order (dy outer)
def test_all_ones_padding_and_channel_order():
    one = np.ones((1, 6, 8, 3), np.float32)
    cv = O.cost_volume(one, one, 4)
    assert np.all(cv[..., 40] == 1)
    nz = np.argwhere(cv[0, :, :, 0] != 0)  # dy = dx = -4
    assert len(nz) == 8 and nz[:, 0].min() == 4 and nz[:, 1].min() == 4
    # channel 8 is dy=-4, dx=+4: needs i >= 4 and j <= 3
    nz = np.argwhere(cv[0, :, :, 8] != 0)
    assert nz[:, 0].min() == 4 and nz[:, 1].max() == 3
    # channel 72 is dy=+4, dx=-4: needs i <= 1 and j >= 4
    nz = np.argwhere(cv[0, :, :, 72] != 0)
    assert nz[:, 0].max() == 1 and nz[:, 1].min() == 4


# (7) unfold-based correlation
def test_cost_volume_matches_unfold():
    c1, w = _rand((2, 7, 9, 10), 5), _rand((2, 7, 9, 10), 6)
    r, n = 3, 7
    cv = O.cost_volume(c1, w, r)
    cols = torch.nn.functional.unfold(torch.from_numpy(w).permute(0, 3, 1, 2), n, padding=r).reshape(2, 10, n * n, 7, 9)
    ref = (cols * torch.from_numpy(c1).permute(0, 3, 1, 2)[:, :, None]).mean(1).permute(0, 2, 3, 1).numpy()
    ref = np.maximum(0.1 * ref, ref)
    np.testing.assert_allclose(cv, ref, atol=3e-7, rtol=1e-5)


# (8) leaky-relu slope, fp32 and fp16
def test_negative_preactivation_scaled():
    c1 = np.ones((1, 4, 4, 2), np.float32)
    cv = O.cost_volume(c1, -c1, 1)
    assert np.allclose(cv[0, 1, 1], -0.1)
    cvh = O.cost_volume(c1.astype(np.float16), -c1.astype(np.float16), 1)
    assert cvh.dtype == np.float16
    assert cvh[0, 1, 1, 4] == np.float16(-1) * np.float16(0.1)


# (9) backward: fp64 finite differences at non-integer flows; zero-flow => dflow == 0
def test_backward_finite_differences():
    rng = np.random.default_rng(7)
    b, h, w, c, r = 1, 5, 6, 3, 2
    c1 = rng.standard_normal((b, h, w, c))
    c2 = rng.standard_normal((b, h, w, c))
    flow = rng.uniform(-1.5, 1.5, (b, h, w, 2))
    flow += 0.05 * np.sign(flow - np.round(flow)) * (np.abs(flow - np.round(flow)) < 0.05)  # stay off the ties
    g = rng.standard_normal((b, h, w, (2 * r + 1) ** 2))
    dc1, dc2, dflow = O.warp_cost_volume_bwd(g, c1, c2, flow, r, 1.25)

    def loss(c1_, c2_, f_):
        return float(np.sum(g * O.warp_cost_volume(c1_, c2_, f_, r, 1.25)))

    eps = 1e-6
    for arr, grad, which in ((c1, dc1, 0), (c2, dc2, 1), (flow, dflow, 2)):
        idxs = [tuple(rng.integers(0, s) for s in arr.shape) for _ in range(25)]
        for idx in idxs:
            p, m = arr.copy(), arr.copy()
            p[idx] += eps
            m[idx] -= eps
            args_p = [c1, c2, flow]
            args_m = [c1, c2, flow]
            args_p[which], args_m[which] = p, m
            fd = (loss(*args_p) - loss(*args_m)) / (2 * eps)
            assert abs(fd - grad[idx]) < 1e-5 * max(1.0, abs(fd)), (which, idx, fd, grad[idx])


def test_zero_flow_has_zero_flow_gradient():
    d = synth.make_level(1, 6, 8, 4, 8, np.float32, "zero", with_grad=True)
    _, _, dflow = O.warp_cost_volume_bwd(d["g"], d["c1"], d["c2"], d["flow"], 4)
    # alpha == 0 exactly => no gradient through the clamp (TF tie rule) ...
    assert np.all(dflow[:, :-1, :, 0] == 0) and np.all(dflow[:, :, :-1, 1] == 0)
    # ... except on the last row / column where floor clamps to size-2 and alpha == 1 (0 < 1 <= 1)
    assert np.any(dflow[:, -1, :, 0] != 0) and np.any(dflow[:, :, -1, 1] != 0)


def test_backward_matches_torch_autograd():
    d = synth.make_level(2, 8, 10, 6, 9, np.float32, "smooth", with_grad=True)
    dc1, dc2, dflow = O.warp_cost_volume_bwd(d["g"], d["c1"], d["c2"], d["flow"], 4, 2.5)
    t = [torch.from_numpy(d[k]).double().requires_grad_() for k in ("c1", "c2", "flow")]
    T.cost_volume(t[0], T.dense_image_warp(t[1], t[2] * 2.5), 4).backward(torch.from_numpy(d["g"]).double())
    np.testing.assert_allclose(dc1, t[0].grad.numpy(), atol=2e-5)
    np.testing.assert_allclose(dc2, t[1].grad.numpy(), atol=2e-5)
    np.testing.assert_allclose(dflow, t[2].grad.numpy(), atol=2e-3)


# (10) fp16 runs op by op in fp16 and stays close to fp32
def test_fp16_regime():
    d = synth.make_level(1, 12, 16, 32, 10, np.float16, "smooth")
    o16 = O.warp_cost_volume(d["c1"], d["c2"], d["flow"], 4, 1.25)
    assert o16.dtype == np.float16
    o32 = O.warp_cost_volume(d["c1"].astype(np.float32), d["c2"].astype(np.float32), d["flow"].astype(np.float32), 4, 1.25)
    assert np.abs(o16.astype(np.float32) - o32).max() < 2e-2


def test_port_matches_oracle():
    d = synth.make_level(2, 9, 12, 8, 11, np.float32, "stress")
    o = O.warp_cost_volume(d["c1"], d["c2"], d["flow"], 4, 0.625)
    t = T.cost_volume(torch.from_numpy(d["c1"]), T.dense_image_warp(torch.from_numpy(d["c2"]), torch.from_numpy(d["flow"]) * 0.625), 4)
    np.testing.assert_allclose(o, t.numpy(), atol=1e-6, rtol=1e-5)


def test_error_behaviour():
    with pytest.raises(ValueError):
        O.interpolate_bilinear(np.zeros((2, 3, 4)), np.zeros((1, 2, 2)))
    with pytest.raises(ValueError):
        O.interpolate_bilinear(np.zeros((1, 2, 3, 4)), np.zeros((1, 2)))
    with pytest.raises(ValueError):
        O.interpolate_bilinear(np.zeros((1, 2, 3, 4)), np.zeros((1, 2, 2)), indexing="zz")


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))),
                         ids=lambda p: os.path.basename(p)[:-4])
def test_golden_fixtures(path):
    z = np.load(path)
    flow = z["flow"] if "flow" in z.files else None
    r, fs = int(z["r"]), float(z["flow_scale"])
    out = O.warp_cost_volume(z["c1"], z["c2"], flow, r, fs)
    assert out.dtype == z["out"].dtype
    assert np.array_equal(out, z["out"])
    if "dc1" in z.files:
        # fixtures hold float64 autograd through the reference source (or the oracle itself for the tie case)
        dc1, dc2, dflow = O.warp_cost_volume_bwd(z["g"], z["c1"], z["c2"], flow, r, fs)
        for got, want in ((dc1, z["dc1"]), (dc2, z["dc2"])) + (((dflow, z["dflow"]),) if dflow is not None else ()):
            assert np.abs(got - want).max() <= 2e-6 * np.abs(want).max()


def test_synth_byte_counts_match_survey():
    # SURVEY.md section 8d per-pair totals
    assert abs(synth.fwd_bytes_per_pair(448, 1024, 4) / 1e6 - 25.700) < 1e-3
    assert abs(synth.bwd_bytes_per_pair(448, 1024, 4) / 1e6 - 51.399) < 1e-3
    assert abs(synth.fwd_bytes_per_pair(384, 512, 4) / 1e6 - 11.014) < 1e-3
    assert abs(synth.fwd_bytes_per_pair(1088, 1920, 2, 8) / 1e6 - 130.859) < 1e-3

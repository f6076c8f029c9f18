"""CPU tests of the widened rows: the PyTorch PWC-Net's structure against the reference's
published parameter counts, its loop against an injected CPU stand-in of the operator, the loss, `.flo` IO and the
`adapt_x` padding rule.  The operator itself has no CPU path; `corr_fn` injects the oracle here."""
import os

import numpy as np
import pytest
import torch

from oracle import pwc_oracle as O
from tfoptflow_b200 import model_pwcnet as M
from tfoptflow_b200 import optflow


def test_parameter_counts_match_the_reference():
    # README.md:55 and the comments at tfoptflow/model_pwcnet.py:100-103
    assert M.ModelPWCNet(M.pwcnet_options("sm")).num_parameters() == 4705064
    assert M.ModelPWCNet(M.pwcnet_options("lg")).num_parameters() == 14079050
    assert M.ModelPWCNet(M.pwcnet_options("sm", use_dense_cx=True)).num_parameters() == 9374274
    assert M.ModelPWCNet(M.pwcnet_options("sm", use_res_cx=True)).num_parameters() == 6774064


def _oracle_corr(calls):
    def fn(c1, c2, up_flow, scaler, lvl):
        calls.append((lvl, scaler, tuple(c1.shape), None if up_flow is None else tuple(up_flow.shape)))
        fl = None if up_flow is None else up_flow.detach().numpy()
        out = O.warp_cost_volume(c1.detach().numpy(), c2.detach().numpy(), fl, 4, scaler)
        return torch.from_numpy(out)
    return fn


@pytest.mark.parametrize("variant", ["sm", "lg"])
def test_forward_loop_calls_the_hot_path_like_the_reference(variant):
    calls = []
    m = M.ModelPWCNet(M.pwcnet_options(variant), corr_fn=_oracle_corr(calls)).eval()
    x = torch.rand(1, 2, 64, 128, 3)
    with torch.no_grad():
        flow_pred, flow_pyr = m(x)
    assert tuple(flow_pred.shape) == (1, 64, 128, 2)
    assert [tuple(f.shape[1:3]) for f in flow_pyr] == [(1, 2), (2, 4), (4, 8), (8, 16), (16, 32)]
    # model_pwcnet.py:1550-1564: no warp at the top level, scalers 0.625, 1.25, 2.5, 5.0 below
    assert [c[0] for c in calls] == [6, 5, 4, 3, 2]
    assert calls[0][3] is None and [c[1] for c in calls[1:]] == [0.625, 1.25, 2.5, 5.0]
    assert [c[2][3] for c in calls] == [196, 128, 96, 64, 32]
    assert torch.isfinite(flow_pred).all()


def test_same_padding_matches_tensorflow_for_stride_2():
    # TF 'same' with stride 2 on an even input pads 0 before / 1 after: output[0,0] sees input rows 0..2, not -1..1
    conv = M._ConvSame(1, 1, stride=2)
    with torch.no_grad():
        conv.conv.weight.zero_()
        conv.conv.weight[0, 0, 0, 0] = 1.0  # picks the top-left tap
        conv.conv.bias.zero_()
    x = torch.arange(36, dtype=torch.float32).reshape(1, 1, 6, 6) + 1
    y = conv(x)
    assert tuple(y.shape) == (1, 1, 3, 3)
    assert y[0, 0, 0, 0] == x[0, 0, 0, 0] and y[0, 0, 1, 1] == x[0, 0, 2, 2]


def test_resize_bilinear_tf1_semantics():
    x = torch.arange(8, dtype=torch.float32).reshape(1, 1, 2, 4)
    up = M.resize_bilinear_tf1(x, (4, 8))
    # src = dst * 0.5 with no half-pixel offset; the last row / column replicate
    assert torch.allclose(up[0, 0, 0], torch.tensor([0, 0.5, 1, 1.5, 2, 2.5, 3, 3.0]))
    assert torch.allclose(up[0, 0, 1], up[0, 0, 0] + 2.0) and torch.allclose(up[0, 0, 3], up[0, 0, 2])
    down = M.resize_bilinear_tf1(torch.arange(16, dtype=torch.float32).reshape(1, 1, 4, 4), (2, 2))
    assert torch.equal(down[0, 0], torch.tensor([[0.0, 2.0], [8.0, 10.0]]))


def test_pwcnet_loss_known_answer():
    opts = M.pwcnet_options("sm")
    y = torch.zeros(2, 64, 64, 2)
    pyr = [torch.zeros(2, 64 >> l, 64 >> l, 2) for l in (6, 5, 4, 3, 2)]
    assert float(M.pwcnet_loss(y, pyr, opts)) == 0.0
    # a constant ground-truth flow of (3, 4) px: level with height h sees (3, 4) / (64 / h), norm 5 * h / 64 per pixel
    y[...] = torch.tensor([3.0, 4.0])
    want = sum(a * (5.0 * h / 64) * h * h for a, h in zip(opts["alphas"], (1, 2, 4, 8, 16)))
    assert abs(float(M.pwcnet_loss(y, pyr, opts)) - want) < 1e-4 * want
    opts_r = M.pwcnet_options("sm", loss_fn="loss_robust", q=0.4, epsilon=0.01)
    want_r = sum(a * (7.0 * h / 64 * h * h + 0.01) ** 0.4 for a, h in zip(opts_r["alphas"], (1, 2, 4, 8, 16)))
    assert abs(float(M.pwcnet_loss(y, pyr, opts_r)) - want_r) < 1e-4 * want_r


def test_training_step_runs_end_to_end_with_injected_operator():
    calls = []
    m = M.ModelPWCNet(M.pwcnet_options("sm"), corr_fn=_oracle_corr(calls))
    opt = torch.optim.Adam(m.parameters(), lr=1e-4, eps=1e-8)  # model_pwcnet.py:266-270
    x, y = torch.rand(1, 2, 64, 64, 3), torch.randn(1, 64, 64, 2)
    flow_pred, flow_pyr = m(x)
    loss = M.pwcnet_loss(y, flow_pyr, m.opts) + M.l2_regulariser(m, m.opts["gamma"])
    loss.backward()
    assert all(p.grad is not None for n, p in m.named_parameters() if n.startswith(("est.2", "head.2", "ctx.2")))
    opt.step()
    assert torch.isfinite(loss)


def test_adapt_x_pads_sintel_to_448(tmp_path):
    x = np.random.default_rng(0).integers(0, 256, (2, 2, 436, 1024, 3), dtype=np.uint8)
    xa, info = M.adapt_x(x)
    assert xa.shape == (2, 2, 448, 1024, 3) and info == (2, 2, 436, 1024, 3) and xa.dtype == np.float32
    assert np.all(xa[:, :, 436:] == 0) and abs(float(xa[0, 0, 0, 0, 0]) - x[0, 0, 0, 0, 0] / 255.0) < 1e-7
    xa2, info2 = M.adapt_x(list(x[:, :, :384, :512]))
    assert xa2.shape == (2, 2, 384, 512, 3) and info2 is None
    assert M.adapt_x(x, M.pwcnet_options("sm", use_mixed_precision=True))[0].dtype == np.float16
    ya, yinfo = M.adapt_y(np.zeros((2, 436, 1024, 2), np.float32))
    assert ya.shape == (2, 448, 1024, 2) and yinfo == (2, 436, 1024, 2)
    flows, pyr = M.postproc_y_hat_test([np.ones((2, 448, 1024, 2)), [np.zeros((2, 7 << k, 16 << k, 2)) for k in range(5)]], yinfo)
    assert flows.shape == (2, 436, 1024, 2) and len(pyr) == 2 and len(pyr[0]) == 5


def test_flo_round_trip_and_layout(tmp_path):
    flow = np.random.default_rng(1).standard_normal((5, 7, 2)).astype(np.float32)
    p = str(tmp_path / "sub" / "a.flo")
    optflow.flow_write(flow, p)
    raw = open(p, "rb").read()
    assert raw[:4] == b"PIEH" and len(raw) == 12 + 5 * 7 * 2 * 4
    assert np.frombuffer(raw[4:12], "<i4").tolist() == [7, 5]  # width first
    assert np.array_equal(optflow.flow_read(p), flow)
    assert optflow.flow_mag_stats(np.zeros((2, 2, 2)))[2] == 0.0
    with open(str(tmp_path / "bad.flo"), "wb") as f:
        f.write(b"\x00" * 16)
    with pytest.raises(IOError):
        optflow.flow_read(str(tmp_path / "bad.flo"))
    # PFM (FlyingThings3D): 3 channels, bottom-up rows
    pf = np.zeros((5, 7, 3), "<f4")
    pf[:, :, :2] = flow
    with open(str(tmp_path / "a.pfm"), "wb") as f:
        f.write(b"PF\n7 5\n-1.0\n")
        np.flipud(pf).tofile(f)
    assert np.array_equal(optflow.flow_read(str(tmp_path / "a.pfm")), flow)

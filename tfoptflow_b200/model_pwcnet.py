"""PyTorch PWC-Net (sm and lg) whose coarse-to-fine loop calls the B200 warp + cost-volume boundary.

Mirrors the reference's graph builder, tfoptflow/model_pwcnet.py:
  `nn`               :1525-1593  the loop: corr at the top level, warp(c2, up_flow * 20/2^lvl) + corr below, deconv x2
  `extract_features` :1012-1101  6-level siamese pyramid, 3 convs per level, channels 16/32/64/96/128/196 (:1084)
  `predict_flow`     :1282-1448  128-128-96-64-32 estimator (+ dense connections in lg), 2-channel flow head
  `refine_flow`      :1453-1520  7-layer dilated context network
  `deconv`           :1180-1221  learned 4x4 stride-2 transposed conv to 2 channels (flow AND features)
  `adapt_x` / `postproc_y_hat_test` :371-409, :449-474  /255, zero-pad H, W to multiples of 64, crop back
The dense convolutions are cuDNN's (torch.nn.Conv2d on channels_last tensors, i.e. NHWC memory); only the
warp + cost-volume step of every level is this repo's own kernels.  A channels_last NCHW tensor permuted to
[B,H,W,C] is contiguous, so features go to the operator and cost volumes come back without a copy.

pwcnet-sm-6-2 = no dense / residual connections (4,705,064 parameters), pwcnet-lg-6-2 = both (14,079,050), README.md:55.
"""
from __future__ import annotations

import copy

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .core_costvol import cost_volume
from .core_warp import dense_image_warp

# tfoptflow/model_pwcnet.py:33-105 (the keys that shape the network and its loss)
_DEFAULT_PWCNET_OPTIONS = {
    "pyr_lvls": 6, "flow_pred_lvl": 2, "search_range": 4,
    "use_dense_cx": False, "use_res_cx": False,
    "use_mixed_precision": False, "loss_scaler": 128.0,
    "loss_fn": "loss_multiscale", "alphas": [0.32, 0.08, 0.02, 0.01, 0.005, 0.0025], "gamma": 0.0004, "q": 1.0,
    "epsilon": 0.0, "batch_size": 8,
}
NUM_CHANN = [None, 16, 32, 64, 96, 128, 196]  # model_pwcnet.py:1084


def pwcnet_options(variant: str = "sm", **overrides):
    """`sm`: plain estimator, `lg`: dense + residual connections (the reference's two published models)."""
    o = copy.deepcopy(_DEFAULT_PWCNET_OPTIONS)
    if variant == "lg":
        o["use_dense_cx"] = o["use_res_cx"] = True
    elif variant != "sm":
        raise ValueError("variant must be 'sm' or 'lg'")
    o.update(overrides)
    return o


class _ConvSame(nn.Module):
    """tf.layers.conv2d(x, f, 3, stride, 'same', dilation_rate=d) (+ leaky-relu 0.1).  TensorFlow's 'same' padding of a
    stride-2 3x3 convolution on an even input is asymmetric (0 before, 1 after), unlike padding=1."""

    def __init__(self, cin, cout, stride=1, dilation=1, act=True):
        super().__init__()
        self.stride, self.act = stride, act
        self.conv = nn.Conv2d(cin, cout, 3, stride, padding=dilation if stride == 1 else 0, dilation=dilation)
        nn.init.kaiming_normal_(self.conv.weight, mode="fan_in", nonlinearity="relu")  # he_normal, model_pwcnet.py:1086
        nn.init.zeros_(self.conv.bias)

    def forward(self, x):
        if self.stride == 2:
            h, w = x.shape[-2:]
            ph = max((-(-h // 2) - 1) * 2 + 3 - h, 0)
            pw = max((-(-w // 2) - 1) * 2 + 3 - w, 0)
            x = F.pad(x, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2))
        x = self.conv(x)
        return F.leaky_relu(x, 0.1) if self.act else x


def resize_bilinear_tf1(x: torch.Tensor, size) -> torch.Tensor:
    """tf.image.resize_bilinear of TF 1.x with its default align_corners=False: src = dst * (in / out), no half-pixel
    offset (model_pwcnet.py:1590, losses.py:56).  x: [B,C,H,W]."""
    b, c, h, w = x.shape
    oh, ow = int(size[0]), int(size[1])
    if (oh, ow) == (h, w):
        return x
    ys = torch.arange(oh, device=x.device, dtype=torch.float32) * (h / oh)
    xs = torch.arange(ow, device=x.device, dtype=torch.float32) * (w / ow)
    y0, x0 = ys.floor().long().clamp_(max=h - 1), xs.floor().long().clamp_(max=w - 1)
    y1, x1 = (y0 + 1).clamp_(max=h - 1), (x0 + 1).clamp_(max=w - 1)
    ay, ax = (ys - y0).to(x.dtype).view(1, 1, oh, 1), (xs - x0).to(x.dtype).view(1, 1, 1, ow)
    top = x[:, :, y0][:, :, :, x0] * (1 - ax) + x[:, :, y0][:, :, :, x1] * ax
    bot = x[:, :, y1][:, :, :, x0] * (1 - ax) + x[:, :, y1][:, :, :, x1] * ax
    return top * (1 - ay) + bot * ay


class ModelPWCNet(nn.Module):
    """`nn(x)` of the reference as a torch module.  forward(x) takes image pairs [B,2,H,W,3] in 0..1 (after `adapt_x`)
    and returns (flow_pred [B,H,W,2], flow_pyr list of [B,h,w,2], coarse to fine) like model_pwcnet.py:1593.

    `corr_fn(c1, c2, up_flow, scaler, lvl)` may replace the operator (tests inject a CPU stand-in); the default is
    `cost_volume(c1, dense_image_warp(c2, up_flow * scaler))` - the reference's two calls, fused by this library."""

    def __init__(self, options=None, corr_fn=None):
        super().__init__()
        self.opts = options or pwcnet_options("sm")
        o = self.opts
        self.corr_fn = corr_fn
        L, l0, r = o["pyr_lvls"], o["flow_pred_lvl"], o["search_range"]
        D = (2 * r + 1) ** 2
        # feature pyramid: conv{lvl}a (stride 2), conv{lvl}aa, conv{lvl}b, shared by both images
        self.featpyr = nn.ModuleList()
        cin = 3
        for lvl in range(1, L + 1):
            f = NUM_CHANN[lvl]
            self.featpyr.append(nn.Sequential(_ConvSame(cin, f, 2), _ConvSame(f, f), _ConvSame(f, f)))
            cin = f
        # per-level estimator, context network, up-samplers
        self.est, self.head, self.ctx = nn.ModuleDict(), nn.ModuleDict(), nn.ModuleDict()
        self.up_flow, self.up_feat = nn.ModuleDict(), nn.ModuleDict()
        for lvl in range(L, l0 - 1, -1):
            cin = D if lvl == L else D + NUM_CHANN[lvl] + 2 + 2
            layers = []
            for f in (128, 128, 96, 64, 32):
                layers.append(_ConvSame(cin, f))
                cin = cin + f if o["use_dense_cx"] else f
            self.est[str(lvl)] = nn.ModuleList(layers)
            self.head[str(lvl)] = _ConvSame(cin, 2, act=False)
            if lvl == l0 or o["use_res_cx"]:
                dil = (1, 2, 4, 8, 16, 1)
                chans = (128, 128, 128, 96, 64, 32)
                ctx, c = [], cin
                for f, d in zip(chans, dil):
                    ctx.append(_ConvSame(c, f, dilation=d))
                    c = f
                ctx.append(_ConvSame(c, 2, act=False))
                self.ctx[str(lvl)] = nn.Sequential(*ctx)
            if lvl != l0:  # conv2d_transpose(x, 2, 4, 2, 'same'), model_pwcnet.py:1221
                self.up_flow[str(lvl)] = nn.ConvTranspose2d(2, 2, 4, 2, 1)
                self.up_feat[str(lvl)] = nn.ConvTranspose2d(cin, 2, 4, 2, 1)

    # -- the two wrappers of the reference (model_pwcnet.py:1106-1178, :1226-1277), NHWC in / out
    def warp(self, c2, sc_up_flow, lvl, name="warp"):
        return dense_image_warp(c2, sc_up_flow, name=f"{name}{lvl}")

    def corr(self, c1, warp, lvl, name="corr"):
        return cost_volume(c1, warp, self.opts["search_range"], f"{name}{lvl}")

    def _hot_path(self, c1, c2, up_flow, lvl):
        """NCHW channels_last in, NCHW channels_last out; the operator sees contiguous NHWC views."""
        a, b = c1.permute(0, 2, 3, 1), c2.permute(0, 2, 3, 1)
        scaler = 20.0 / 2 ** lvl  # model_pwcnet.py:1560
        fl = None if up_flow is None else up_flow.permute(0, 2, 3, 1)
        if self.corr_fn is not None:
            out = self.corr_fn(a, b, fl, scaler, lvl)
        elif fl is None:
            out = self.corr(a.contiguous(), b.contiguous(), lvl)
        else:
            out = self.corr(a.contiguous(), self.warp(b.contiguous(), fl.contiguous() * scaler, lvl), lvl)
        return out.permute(0, 3, 1, 2)

    def extract_features(self, x_tnsr):
        """x_tnsr [B,2,H,W,3] -> 1-based lists c1, c2 of NCHW channels_last features (model_pwcnet.py:1012-1101)."""
        c1, c2 = [None], [None]
        for pyr, idx in ((c1, 0), (c2, 1)):
            x = x_tnsr[:, idx].permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
            for blk in self.featpyr:
                x = blk(x)
                pyr.append(x)
        return c1, c2

    def predict_flow(self, corr, c1, up_flow, up_feat, lvl):
        x = corr if c1 is None else torch.cat([corr, c1, up_flow, up_feat], dim=1)  # model_pwcnet.py:1424
        for conv in self.est[str(lvl)]:
            act = conv(x)
            x = torch.cat([act, x], dim=1) if self.opts["use_dense_cx"] else act
        return x, self.head[str(lvl)](x)

    def refine_flow(self, feat, flow, lvl):
        return flow + self.ctx[str(lvl)](feat)

    def forward(self, x_tnsr):
        o = self.opts
        c1, c2 = self.extract_features(x_tnsr)
        flow_pyr, up_flow, up_feat = [], None, None
        for lvl in range(o["pyr_lvls"], o["flow_pred_lvl"] - 1, -1):
            if lvl == o["pyr_lvls"]:
                corr = self._hot_path(c1[lvl], c2[lvl], None, lvl)
                upfeat, flow = self.predict_flow(corr, None, None, None, lvl)
            else:
                corr = self._hot_path(c1[lvl], c2[lvl], up_flow, lvl)
                upfeat, flow = self.predict_flow(corr, c1[lvl], up_flow, up_feat, lvl)
            if lvl != o["flow_pred_lvl"]:
                if o["use_res_cx"]:
                    flow = self.refine_flow(upfeat, flow, lvl)
                flow_pyr.append(flow.permute(0, 2, 3, 1))
                up_flow = self.up_flow[str(lvl)](flow)
                up_feat = self.up_feat[str(lvl)](upfeat)
            else:
                flow = self.refine_flow(upfeat, flow, lvl)
                flow_pyr.append(flow.permute(0, 2, 3, 1))
                scaler = 2 ** o["flow_pred_lvl"]
                h, w = flow.shape[-2:]
                flow_pred = resize_bilinear_tf1(flow, (h * scaler, w * scaler)) * scaler  # model_pwcnet.py:1590
                return flow_pred.permute(0, 2, 3, 1), flow_pyr

    def num_parameters(self):
        return sum(p.numel() for p in self.parameters())


def pwcnet_loss(y, y_hat_pyr, opts):
    """tfoptflow/losses.py:27-73.  y [B,H,W,2] ground truth, y_hat_pyr list of [B,h,w,2] predictions (coarse to fine,
    like `flow_pyr`).  The reference indexes alphas and the pyramid with the same `lvl` counter; so does this."""
    norm_order = 2 if opts["loss_fn"] == "loss_multiscale" else 1
    total = 0.0
    gt = y.permute(0, 3, 1, 2).float()
    gt_h = gt.shape[2]
    for lvl in range(opts["pyr_lvls"] - opts["flow_pred_lvl"] + 1):
        pred = y_hat_pyr[lvl].float()
        lh, lw = pred.shape[1], pred.shape[2]
        scaled = resize_bilinear_tf1(gt, (lh, lw)).permute(0, 2, 3, 1) / (gt_h / lh)
        diff = scaled - pred
        norm = torch.linalg.vector_norm(diff, ord=norm_order, dim=3)
        level_loss = norm.sum(dim=(1, 2)).mean()
        total = total + opts["alphas"][lvl] * torch.pow(level_loss + opts["epsilon"], opts["q"])
    return total


def l2_regulariser(model: nn.Module, gamma: float):
    """`gamma * sum(tf.nn.l2_loss(var))` over all trainable variables (model_pwcnet.py:299-300); l2_loss = sum(v^2) / 2."""
    return gamma * sum((p.float() ** 2).sum() for p in model.parameters()) * 0.5


# ---- pre / post processing (the driver part of the reference that shapes the operator's inputs) -----------------
def adapt_x(x, opts=None):
    """model_pwcnet.py:371-409: uint8 image pairs [N,2,H,W,3] (array or list) -> float 0..1, zero-padded at the bottom /
    right to multiples of 2**pyr_lvls; returns (x_adapt, original shape or None)."""
    opts = opts or _DEFAULT_PWCNET_OPTIONS
    assert isinstance(x, (np.ndarray, list))
    if isinstance(x, np.ndarray):
        assert x.ndim == 5 and x.shape[1] == 2 and x.shape[4] == 3
    else:
        assert x[0].ndim == 4 and (x[0].shape[0] == 2 or x[0].shape[3] == 3)
    dt = np.float16 if opts.get("use_mixed_precision") else np.float32
    x_adapt = np.array(x, dtype=dt) if isinstance(x, list) else x.astype(dt)
    x_adapt /= 255.0
    m = 2 ** opts["pyr_lvls"]
    pad_h, pad_w = (-x_adapt.shape[2]) % m, (-x_adapt.shape[3]) % m
    info = None
    if pad_h or pad_w:
        info = x_adapt.shape
        x_adapt = np.pad(x_adapt, [(0, 0), (0, 0), (0, pad_h), (0, pad_w), (0, 0)], mode="constant", constant_values=0.0)
    return x_adapt, info


def adapt_y(y, opts=None):
    """model_pwcnet.py:411-447: flows [N,H,W,2] padded the same way."""
    opts = opts or _DEFAULT_PWCNET_OPTIONS
    y_adapt = np.array(y, dtype=np.float32) if isinstance(y, list) else y
    assert y_adapt.ndim == 4 and y_adapt.shape[3] == 2
    m = 2 ** opts["pyr_lvls"]
    pad_h, pad_w = (-y_adapt.shape[1]) % m, (-y_adapt.shape[2]) % m
    info = None
    if pad_h or pad_w:
        info = y_adapt.shape
        y_adapt = np.pad(y_adapt, [(0, 0), (0, pad_h), (0, pad_w), (0, 0)], mode="constant", constant_values=0.0)
    return y_adapt, info


def postproc_y_hat_test(y_hat, adapt_info=None, opts=None):
    """model_pwcnet.py:449-474: crop predicted flows back to (N,H,W,2) `adapt_info`, split the pyramid per sample."""
    opts = opts or _DEFAULT_PWCNET_OPTIONS
    assert isinstance(y_hat, (list, tuple)) and len(y_hat) == 2
    pred_flows = y_hat[0]
    if adapt_info is not None:
        pred_flows = pred_flows[:, 0:adapt_info[1], 0:adapt_info[2], :]
    pyramids = y_hat[1]
    per_sample = []
    for idx in range(len(pred_flows)):
        per_sample.append([pyramids[lvl][idx] for lvl in range(opts["pyr_lvls"] - opts["flow_pred_lvl"] + 1)])
    return pred_flows, per_sample


@torch.no_grad()
def predict_from_img_pairs(model: ModelPWCNet, img_pairs, batch_size=1, device=None):
    """model_pwcnet.py:957-1007: list of (img1, img2) uint8 HxWx3 pairs -> list of [H,W,2] float32 flows."""
    device = device or next(model.parameters()).device
    preds, n = [], len(img_pairs)
    for s in range(0, n, batch_size):
        idx = [i % n for i in range(s, s + batch_size)]  # the reference wraps around to fill the last batch
        x = np.array([img_pairs[i] for i in idx])
        x_adapt, info = adapt_x(x, model.opts)
        y_info = None if info is None else (info[0], info[2], info[3], 2)
        xt = torch.from_numpy(x_adapt).to(device)
        flow_pred, flow_pyr = model(xt)
        y_hats, _ = postproc_y_hat_test([flow_pred.float().cpu().numpy(), [f.float().cpu().numpy() for f in flow_pyr]],
                                        y_info, model.opts)
        preds.extend(list(y_hats))
    return preds[:n]

"""Generates tests/golden/*.npz - small input/output vectors for the warp + cost-volume path.

The reference (TensorFlow 1.10.1) cannot be imported in this image, so the outputs
come from oracle/pwc_oracle.py AFTER each case is cross-checked here against an
independently written formulation:
  * dense_image_warp  vs torch.nn.functional.grid_sample(border, align_corners=True)
  * cost_volume       vs an unfold-based correlation
  * backward          vs torch autograd through oracle/ref_graph_torch.py
Run:  python tests/golden/make_golden.py      (rewrites the .npz files)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pwc_oracle as O  # noqa: E402
from oracle import ref_graph_torch as T  # noqa: E402
from tfoptflow_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def grid_sample_warp(image, flow):
    b, h, w, c = image.shape
    img = torch.from_numpy(image.astype(np.float64)).permute(0, 3, 1, 2)
    gy, gx = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    qy = gy[None] - flow[..., 0].astype(np.float64)
    qx = gx[None] - flow[..., 1].astype(np.float64)
    grid = np.stack([2 * qx / (w - 1) - 1, 2 * qy / (h - 1) - 1], -1)
    out = torch.nn.functional.grid_sample(img, torch.from_numpy(grid), mode="bilinear",
                                          padding_mode="border", align_corners=True)
    return out.permute(0, 2, 3, 1).numpy()


def unfold_corr(c1, warp, r):
    b, h, w, c = c1.shape
    n = 2 * r + 1
    wp = torch.from_numpy(warp.astype(np.float64)).permute(0, 3, 1, 2)
    cols = torch.nn.functional.unfold(wp, n, padding=r).reshape(b, c, n * n, h, w)
    a = torch.from_numpy(c1.astype(np.float64)).permute(0, 3, 1, 2)[:, :, None]
    cv = (cols * a).mean(1).permute(0, 2, 3, 1).numpy()
    return np.maximum(0.1 * cv, cv)


CASES = [
    # name, B, H, W, C, r, dtype, flow kind, flow_scale
    ("f32_r4_smooth", 2, 12, 16, 32, 4, np.float32, "smooth", 1.25),
    ("f32_r4_stress", 1, 9, 11, 20, 4, np.float32, "stress", 1.0),
    ("f32_r4_noflow_c196", 2, 6, 8, 196, 4, np.float32, None, 1.0),
    ("f32_r2_odd", 1, 5, 7, 4, 2, np.float32, "smooth", 0.625),
    ("f32_r8_smooth", 1, 17, 30, 16, 8, np.float32, "smooth", 5.0),
    ("f32_r4_zero", 1, 6, 8, 8, 4, np.float32, "zero", 1.0),
    ("f16_r4_smooth", 2, 12, 16, 32, 4, np.float16, "smooth", 2.5),
    ("f16_r4_noflow", 1, 7, 16, 196, 4, np.float16, None, 1.0),
]


def main():
    for i, (name, b, h, w, c, r, dt, fk, fs) in enumerate(CASES):
        d = synth.make_level(b, h, w, c, 4242 + i, dt, fk, with_grad=True, search_range=r)
        c1, c2, flow, g = d["c1"], d["c2"], d["flow"], d["g"]
        out = O.warp_cost_volume(c1, c2, flow, r, fs)
        save = dict(c1=c1, c2=c2, g=g, out=out, r=r, flow_scale=np.float32(fs))
        if flow is not None:
            save["flow"] = flow
            scaled = flow * np.asarray(fs).astype(flow.dtype)
            warp = O.dense_image_warp(c2, scaled)
            save["warp"] = warp
        else:
            warp = c2
        if dt == np.float32:
            if flow is not None:
                gs = grid_sample_warp(c2, scaled)
                assert np.abs(gs - warp).max() < 2e-5, (name, np.abs(gs - warp).max())
            uc = unfold_corr(c1, warp, r)
            assert np.abs(uc - out).max() < 5e-6 * max(1.0, np.abs(uc).max()), (name, np.abs(uc - out).max())
            dc1, dc2, dflow = O.warp_cost_volume_bwd(g, c1, c2, flow, r, fs)
            # autograd cross-check (non-integer flows only; ties differ by design)
            if fk != "zero":
                tc1 = torch.from_numpy(c1).double().requires_grad_()
                tc2 = torch.from_numpy(c2).double().requires_grad_()
                if flow is not None:
                    tf_ = torch.from_numpy(flow).double().requires_grad_()
                    o = T.cost_volume(tc1, T.dense_image_warp(tc2, tf_ * fs), r)
                else:
                    tf_ = None
                    o = T.cost_volume(tc1, tc2, r)
                o.backward(torch.from_numpy(g).double())
                assert np.abs(tc1.grad.numpy() - dc1).max() < 1e-4, name
                assert np.abs(tc2.grad.numpy() - dc2).max() < 1e-4, name
                if tf_ is not None:
                    assert np.abs(tf_.grad.numpy() - dflow).max() < 2e-3 * max(1.0, np.abs(dflow).max()), \
                        (name, np.abs(tf_.grad.numpy() - dflow).max())
            save.update(dc1=dc1, dc2=dc2)
            if dflow is not None:
                save["dflow"] = dflow
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **save)
        print("wrote", name, out.shape, out.dtype)


if __name__ == "__main__":
    main()

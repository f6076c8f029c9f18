"""Generates tests/golden/*.npz from the REFERENCE'S OWN SOURCE.

`/root/reference/tfoptflow/core_warp.py` and `core_costvol.py` are imported unmodified on a numpy stand-in for
the TensorFlow symbols they use (tests/golden/tf_numpy_shim.py; TF 1.10.1 itself cannot run in this image) and
executed on seeded synthetic inputs.  Fields:

  c1, c2, flow, g, r, flow_scale   inputs (tfoptflow_b200/synth.py, seeds 4242 + case)
  warp, out                        forward results of the imported reference, in the tensor dtype (fp32 / fp16),
                                   through the loop's call pattern corr(c1, warp(c2, up_flow * scaler))
                                   (tfoptflow/model_pwcnet.py:1552-1564)
  dc1, dc2, dflow (fp32 cases)     float64 torch.autograd THROUGH THE REFERENCE'S SOURCE (torch stand-in).  The
                                   reference has no backward code (TF autodiff, model_pwcnet.py:547-550); at exact
                                   ties TF and torch route gradients differently, so the one tie case (`zero` flow:
                                   integer coordinates) takes its gradients from oracle/pwc_oracle.py, which
                                   implements TF 1.x's documented tie rules; `bwd_source` says which.

Before a file is written the oracle restatement (oracle/pwc_oracle.py) must agree with the imported reference:
bit-for-bit forward, 2e-6 (relative to the largest gradient) backward.

Run here (needs /root/reference):  python tests/golden/make_golden_from_reference.py
The .npz files travel to the GPU box; the reference does not."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import tf_numpy_shim as shim  # noqa: E402
from oracle import pwc_oracle as O  # noqa: E402
from tfoptflow_b200 import synth  # noqa: E402

CASES = [
    # name, B, H, W, C, r, dtype, flow kind, flow_scale
    ("f32_r4_smooth", 2, 12, 16, 32, 4, np.float32, "smooth", 1.25),
    ("f32_r4_stress", 1, 9, 11, 20, 4, np.float32, "stress", 1.0),
    ("f32_r4_noflow_c196", 2, 6, 8, 196, 4, np.float32, None, 1.0),
    ("f32_r2_odd", 1, 5, 7, 4, 2, np.float32, "smooth", 0.625),
    ("f32_r8_smooth", 1, 17, 30, 16, 8, np.float32, "smooth", 5.0),
    ("f32_r4_zero", 1, 6, 8, 8, 4, np.float32, "zero", 1.0),
    ("f32_r4_tall_c32", 1, 30, 52, 32, 4, np.float32, "smooth", 5.0),   # streaming-kernel shape (H >= 28)
    ("f32_r4_c64_stress", 1, 16, 24, 64, 4, np.float32, "stress", 2.5),
    ("f16_r4_smooth", 2, 12, 16, 32, 4, np.float16, "smooth", 2.5),
    ("f16_r4_noflow", 1, 7, 16, 196, 4, np.float16, None, 1.0),
    ("f16_r8_stress", 1, 17, 30, 32, 8, np.float16, "stress", 5.0),
]


def main():
    ref = shim.Reference("numpy")
    ref_t = shim.Reference("torch")
    for i, (name, b, h, w, c, r, dt, fk, fs) in enumerate(CASES):
        d = synth.make_level(b, h, w, c, 4242 + i, dt, fk, with_grad=True, search_range=r)
        c1, c2, flow, g = d["c1"], d["c2"], d["flow"], d["g"]
        out = ref.warp_cost_volume(c1, c2, flow, r, fs)
        assert out.dtype == dt
        assert np.array_equal(out, O.warp_cost_volume(c1, c2, flow, r, fs)), f"{name}: oracle != reference (fwd)"
        save = dict(c1=c1, c2=c2, g=g, out=out, r=r, flow_scale=np.float32(fs),
                    fwd_source="reference source on tf_numpy_shim (numpy backend)")
        if flow is not None:
            scaled = flow * np.asarray(fs).astype(flow.dtype)
            save["flow"] = flow
            save["warp"] = ref.dense_image_warp(c2, scaled)
            assert np.array_equal(save["warp"], O.dense_image_warp(c2, scaled)), f"{name}: oracle != reference (warp)"
        if dt == np.float32:
            o1, o2, o3 = O.warp_cost_volume_bwd(g, c1, c2, flow, r, fs)
            if fk == "zero":
                dc1, dc2, dflow = o1, o2, o3
                save["bwd_source"] = "oracle/pwc_oracle.py (TF 1.x tie rules at integer coordinates)"
            else:
                out64, dc1, dc2, dflow = ref_t.backward_fp64(g, c1, c2, flow, r, fs)
                assert np.abs(out64 - out).max() <= 2e-6 * max(1.0, np.abs(out64).max())
                for nm, a, o in (("dc1", dc1, o1), ("dc2", dc2, o2), ("dflow", dflow, o3)):
                    if a is not None:
                        e, s = np.abs(a - o).max(), np.abs(a).max()
                        assert e <= 2e-6 * s, f"{name}: oracle {nm} off by {e:.3e} (scale {s:.3e})"
                save["bwd_source"] = "float64 torch.autograd through the reference source (tf_numpy_shim torch backend)"
            save.update(dc1=dc1, dc2=dc2)
            if dflow is not None:
                save["dflow"] = dflow
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **save)
        print("wrote", name, out.shape, out.dtype, save.get("bwd_source", "")[:30])


if __name__ == "__main__":
    main()

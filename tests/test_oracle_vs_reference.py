"""Pins oracle/pwc_oracle.py to the REFERENCE'S OWN SOURCE: `/root/reference/tfoptflow/core_warp.py` and
`core_costvol.py` are imported unmodified on a numpy stand-in for the TensorFlow symbols they use
(tests/golden/tf_numpy_shim.py) and must agree with the oracle bit for bit on the level shapes of every
BASELINE.json config (fp32 and fp16, search_range 4 and 8, smooth / stress / zero / no flow); the backward is
checked against float64 torch.autograd through the same unmodified source.

CPU only.  Skipped where /root/reference does not exist (the GPU box); the committed tests/golden/*.npz were
generated from the same import (tests/golden/make_golden_from_reference.py) and travel instead."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import tf_numpy_shim as shim  # noqa: E402

from oracle import pwc_oracle as O  # noqa: E402
from tfoptflow_b200 import synth  # noqa: E402

pytestmark = pytest.mark.skipif(not shim.reference_available(), reason="reference checkout not present")


@pytest.fixture(scope="module")
def ref():
    return shim.Reference("numpy")


@pytest.fixture(scope="module")
def ref_torch():
    return shim.Reference("torch")


def test_reference_modules_are_the_unmodified_files(ref):
    for m, f in ((ref.core_warp, "core_warp.py"), (ref.core_costvol, "core_costvol.py")):
        assert os.path.samefile(m.__file__, os.path.join(shim.REFERENCE_DIR, f))
    assert "tensorflow" not in sys.modules or not getattr(sys.modules["tensorflow"], "__pwc_shim__", False)


# BASELINE.json configs: [0]/[1] 384x512 (levels 6..2, sweep adds level 1), [2] 448x1024, [4] 1088x1920 r=8
def _level_cases():
    cases = []
    for (cfg, H, W, r, levels) in (("cfg0", 384, 512, 4, (6, 5, 4, 3, 2, 1)), ("cfg2", 448, 1024, 4, (6, 5, 4, 3, 2)),
                                   ("cfg4", 1088, 1920, 8, (6, 5, 4, 3))):
        for (l, h, w, c) in synth.level_shapes(H, W, levels):
            for dt in (np.float32, np.float16):
                if dt == np.float16 and h * w * c * (2 * r + 1) ** 2 > 3e8:
                    continue  # numpy fp16 arithmetic is ~10x slower; the big fp16 levels are covered in fp32
                cases.append(pytest.param(h, w, c, r, dt, l, id=f"{cfg}-L{l}-{h}x{w}x{c}-r{r}-{np.dtype(dt).name}"))
    return cases


@pytest.mark.parametrize("h,w,c,r,dt,lvl", _level_cases())
def test_oracle_equals_reference_on_config_level_shapes(ref, h, w, c, r, dt, lvl):
    fs = synth.FLOW_SCALERS[lvl]
    for fk in (("smooth", "stress") if h * w <= 64 * 128 else ("smooth",)):
        d = synth.make_level(1, h, w, c, 9000 + lvl, dt, fk)
        # the op sees up_flow (pre-scaler), like synth.make_pyramid
        flow = np.ascontiguousarray((d["flow"].astype(np.float32) / fs).astype(dt))
        got = O.warp_cost_volume(d["c1"], d["c2"], flow, r, fs)
        want = ref.warp_cost_volume(d["c1"], d["c2"], flow, r, fs)
        assert got.dtype == want.dtype == dt
        assert np.array_equal(got, want), f"{fk}: oracle differs from the reference source"
    if lvl == 6:  # the loop's top level correlates un-warped features (model_pwcnet.py:1552-1554)
        assert np.array_equal(O.cost_volume(d["c1"], d["c2"], r), ref.cost_volume(d["c1"], d["c2"], r))


@pytest.mark.parametrize("dt", [np.float32, np.float16])
@pytest.mark.parametrize("fk", ["zero", "stress", "huge", "integer", "halfway"])
def test_oracle_equals_reference_on_edge_flows(ref, fk, dt):
    d = synth.make_level(2, 9, 14, 12, 77, dt, "stress" if fk == "stress" else "zero")
    if fk == "huge":
        d["flow"][0], d["flow"][1] = 1e4, -1e4
    elif fk == "integer":
        d["flow"][..., 0], d["flow"][..., 1] = 3.0, -2.0
    elif fk == "halfway":
        d["flow"][:] = 0.5
    w_o, w_r = O.dense_image_warp(d["c2"], d["flow"]), ref.dense_image_warp(d["c2"], d["flow"])
    assert w_o.dtype == w_r.dtype == dt and np.array_equal(w_o, w_r)
    for r in (0, 1, 4):
        assert np.array_equal(O.cost_volume(d["c1"], w_o, r), ref.cost_volume(d["c1"], w_r, r))


def test_mixed_dtypes_like_the_reference_allows(ref):
    # core_warp.py:174-175: image and flow need not share a dtype (fp16 features, fp32 flow)
    d = synth.make_level(1, 10, 12, 8, 5, np.float16, "smooth")
    f32 = d["flow"].astype(np.float32) * np.float32(1.37)
    a, b = O.dense_image_warp(d["c2"], f32), ref.dense_image_warp(d["c2"], f32)
    assert a.dtype == b.dtype == np.float16 and np.array_equal(a, b)


def test_reference_error_behaviour_matches(ref):
    for bad in ((np.zeros((2, 3, 4), np.float32), np.zeros((1, 2, 2), np.float32)),
                (np.zeros((1, 2, 3, 4), np.float32), np.zeros((1, 2), np.float32))):
        with pytest.raises(ValueError):
            ref.core_warp._interpolate_bilinear(*bad)
        with pytest.raises(ValueError):
            O.interpolate_bilinear(*bad)
    with pytest.raises(ValueError):
        ref.core_warp._interpolate_bilinear(np.zeros((1, 2, 3, 4), np.float32), np.zeros((1, 2, 2), np.float32), indexing="zz")


@pytest.mark.parametrize("case", [(2, 12, 16, 32, 4, "smooth", 1.25), (1, 9, 11, 20, 4, "stress", 1.0),
                                  (1, 10, 14, 16, 8, "smooth", 5.0), (1, 6, 8, 196, 4, None, 1.0)])
def test_oracle_backward_equals_autograd_of_reference_source(ref_torch, case):
    b, h, w, c, r, fk, fs = case
    d = synth.make_level(b, h, w, c, 31, np.float32, fk, with_grad=True, search_range=r)
    _, dc1, dc2, dflow = ref_torch.backward_fp64(d["g"], d["c1"], d["c2"], d["flow"], r, fs)
    o1, o2, o3 = O.warp_cost_volume_bwd(d["g"], d["c1"], d["c2"], d["flow"], r, fs)
    for nm, a, o in (("dc1", dc1, o1), ("dc2", dc2, o2), ("dflow", dflow, o3)):
        if a is None:
            assert o is None
            continue
        err, scale = float(np.abs(a - o).max()), float(np.abs(a).max())
        print(f"{nm}: max err {err:.3e} (scale {scale:.3e})")
        assert err <= 2e-6 * scale, nm

"""Ring-protocol stress of the streaming kernel on the GPU: every geometry (C = 16 .. 128), ragged shapes, worst-case
flows, odd CTA counts (different segment cuts), fp32 and fp16, repeated launches.  The generic kernel is the
comparison.  scripts/stress_stream.py is the long version; it found the 8-pair-slot deadlock of an early C = 128
geometry (an edge task needs 9 c1 pairs resident at once), now excluded by a static_assert."""
import itertools

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

SHAPES = [(3, 60, 50, 32), (2, 57, 129, 64), (5, 30, 24, 96), (2, 41, 38, 128), (3, 113, 25, 16)]


@pytest.mark.timeout(300)
@pytest.mark.parametrize("shape,dt", list(itertools.product(SHAPES, (np.float32, np.float16))))
def test_stream_kernel_under_odd_cta_counts(shape, dt):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from tfoptflow_b200 import _ffi, synth, warp_cost_volume
    dev = torch.device("cuda", 0)
    b, h, w, c = shape
    d = synth.make_level(b, h, w, c, 3 + h, dt, "stress")
    c1, c2, fl = (torch.from_numpy(d[k]).to(dev) for k in ("c1", "c2", "flow"))
    try:
        _ffi.set_option("fwd_impl", 1)
        _ffi.set_option("stream_ctas", 0)
        ref = warp_cost_volume(c1, c2, fl, 4, 1.7).float()
        _ffi.set_option("fwd_impl", 3)
        tol = (2e-6 if dt == np.float32 else 4e-3) * float(ref.abs().max())
        for ctas in (0, 1, 5, 37):
            _ffi.set_option("stream_ctas", ctas)
            for _ in range(2):
                got = warp_cost_volume(c1, c2, fl, 4, 1.7).float()
                torch.cuda.synchronize()
                assert float((got - ref).abs().max()) <= tol, (shape, ctas)
    finally:
        _ffi.set_option("fwd_impl", 0)
        _ffi.set_option("stream_ctas", 0)

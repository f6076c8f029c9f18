"""CPU model of the index arithmetic of the 16-bit band GEMM (tfoptflow_b200/csrc/fwd_mma16.cuh): which accumulator
element of the m16n8k16 fragments is which (pixel, displacement), and that the band GEMM of a 16-pixel block against
its 16 + 2R neighbours reproduces the reference's cost volume row (core_costvol.py:29-36) on random data."""
import numpy as np
import pytest

from oracle import pwc_oracle as O


@pytest.mark.parametrize("R", [4, 8])
def test_every_pixel_displacement_pair_is_extracted_exactly_once(R):
    ND, ntiles = 2 * R + 1, (16 + 2 * R) // 8
    seen = {}
    for lane in range(32):
        g, tq = lane >> 2, lane & 3
        for nt in range(ntiles):
            for e in range(4):          # accumulator register e of mma.m16n8k16: row g + 8 (e >> 1), column 2 tq + (e & 1)
                m, n = g + 8 * (e >> 1), 8 * nt + 2 * tq + (e & 1)
                dxi = n - m
                if 0 <= dxi < ND:
                    assert (m, dxi) not in seen
                    seen[(m, dxi)] = (lane, nt, e)
    assert set(seen) == {(m, d) for m in range(16) for d in range(ND)}


@pytest.mark.parametrize("R", [4, 8])
def test_band_gemm_of_a_block_is_the_cost_volume_row(R):
    rng = np.random.default_rng(R)
    C, W, ND = 32, 48, 2 * R + 1
    c1 = rng.standard_normal((1, 2 * R + 1, W, C)).astype(np.float32)
    wp = rng.standard_normal((1, 2 * R + 1, W, C)).astype(np.float32)
    ref = O.cost_volume(c1, wp, R)                      # [1, H, W, ND * ND]
    y, x0 = R, 16                                       # the middle row has all 2R + 1 displacement rows inside the image
    for dyi in range(ND):
        # S[m, n] = sum_c c1[y, x0 + m, c] * warped[y + dyi - R, x0 - R + n, c], n = 0 .. 16 + 2R - 1 (zero outside the image)
        nb = np.zeros((16 + 2 * R, C), np.float32)
        for n in range(16 + 2 * R):
            x = x0 - R + n
            if 0 <= x < W:
                nb[n] = wp[0, y + dyi - R, x]
        S = c1[0, y, x0:x0 + 16].astype(np.float64) @ nb.astype(np.float64).T
        for m in range(16):
            for dxi in range(ND):
                v = S[m, m + dxi] / C
                v = max(0.1 * v, v)
                assert abs(v - ref[0, y, x0 + m, dyi * ND + dxi]) <= 1e-5 * max(1.0, abs(v))

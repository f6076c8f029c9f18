"""CPU model of the fused backward's data flow (tfoptflow_b200/csrc/bwd_fused.cuh): the A-row layout with its bank
swizzle, the target-row-indexed workspace tensor K1 stores its plane blocks into and K2 loads shifted boxes from, the 9-row window and the skewed row-pair
walk of the consumer.  It is checked against the oracle, so an indexing mistake in the design shows up here, on the
CPU, before any GPU time is spent; the CUDA code itself is checked by tests/test_gpu_parity.py."""
import numpy as np
import pytest

from oracle import pwc_oracle as O
from tfoptflow_b200 import synth

R, ND, D, TW = 4, 9, 81, 32


def a_index(px, d, ex):
    return d * TW + ((((px >> 2) ^ (ex & 7)) << 2) | (px & 3))


def test_a_index_is_a_bijection_and_builder_stores_are_conflict_free():
    seen = {a_index(px, d, d % ND) for px in range(TW) for d in range(D)}
    assert seen == set(range(D * TW))
    # builder: lanes = 4 consecutive pixels x 8 consecutive planes -> 32 distinct banks unless the octet wraps ex 8 -> 0
    for grp in range(8):
        for octet in range(10):
            banks = [a_index(grp * 4 + (l >> 3), octet * 8 + (l & 7), (octet * 8 + (l & 7)) % ND) % 32 for l in range(32)]
            assert len(set(banks)) >= 28


def store_row_blocks(ws, gp, b, x0, y, H, W):
    """bwd_fused.cuh produce_a_k1: plane block dyi of SOURCE row y goes, at the source's own x, to TARGET row y + dyi - 4
    of the workspace tensor [B][H][dyi][dxi][x] (one tensor store per block, clipped at the image edge); the blocks of
    target row y whose source row y + 4 - dyi is outside the image are written as zeros."""
    n = min(TW, W - x0)
    for dyi in range(ND):
        yq = y + dyi - R
        if 0 <= yq < H:
            assert np.isnan(ws[b, yq, dyi, :, x0:x0 + n]).all(), "workspace block written twice"
            ws[b, yq, dyi, :, x0:x0 + n] = gp[b, y, x0:x0 + n, dyi * ND:(dyi + 1) * ND].T
        if not 0 <= y + R - dyi < H:
            assert np.isnan(ws[b, y, dyi, :, x0:x0 + n]).all(), "workspace block written twice"
            ws[b, y, dyi, :, x0:x0 + n] = 0.0


def load_a_row(ws, b, x0, y, W):
    """produce_a_k2: 9 tensor loads, box (x0 + ex - 4 .. + 31, dxi = 8 - ex, dyi = 0 .. 8, row y), zero outside [0, W)"""
    A = np.zeros(D * TW)
    for ex in range(ND):
        for px in range(TW):
            xs = x0 + px + ex - R
            for ey in range(ND):
                v = ws[b, y, 2 * R - ey, 2 * R - ex, xs] if 0 <= xs < W else 0.0
                A[a_index(px, ey * ND + ex, ex)] = v
    return A


def consumer(A_rows, win, P):
    """A_rows[j] : [81*32] swizzled planes of band row j; win[u] : [40, C] window rows -> out[j][px][c].
    Row pair p: half h handles row 2p + h with ey = t - h at step t, window row u = 2p + t (same for both halves)."""
    C = win.shape[-1]
    out = np.zeros((2 * P, TW, C), np.float64)
    for p in range(P):
        for t in range(ND + 1):
            u = 2 * p + t
            for h in range(2):
                ey = t - h
                if not 0 <= ey < ND:
                    continue
                for ex in range(ND):
                    a = np.array([A_rows[2 * p + h][a_index(px, ey * ND + ex, ex)] for px in range(TW)])
                    out[2 * p + h] += a[:, None] * win[u][ex:ex + TW]
    return out


@pytest.mark.parametrize("shape", [(1, 10, 40, 32, 6), (2, 7, 33, 32, 4), (1, 6, 64, 32, 4), (1, 9, 20, 32, 6), (1, 3, 35, 32, 2)])
def test_model_matches_oracle(shape):
    B, H, W, C, RB = shape
    d = synth.make_level(B, H, W, C, 5, np.float32, "smooth", with_grad=True)
    c1, c2, flow, g = (d[k] for k in ("c1", "c2", "flow", "g"))
    warp = O.dense_image_warp(c2, flow)
    outf = O.cost_volume(c1, warp, R)
    dc1_ref, dwarp_ref = O.cost_volume_bwd(g, outf, c1, warp, R)
    gp = (g * np.where(outf > 0, 1.0, 0.1)).astype(np.float64) / C
    nsx = (W + TW - 1) // TW
    nb = (H + RB - 1) // RB
    ws = np.full((B, H, ND, ND, W), np.nan)   # never-written entries must not be used
    dc1 = np.zeros_like(dc1_ref, dtype=np.float64)
    dwarp = np.zeros_like(dc1)

    def units():
        for b in range(B):
            for band in range(nb):
                for sx in range(nsx):
                    y0 = band * RB
                    rows = min(RB, H - y0)
                    yield b, sx, sx * TW, y0, rows, (rows + 1) // 2

    def window(src, b, x0, y0, P):
        win = np.zeros((2 * P + 2 * R, TW + 2 * R, C))
        for u in range(2 * P + 2 * R):
            y = y0 - R + u
            for px in range(TW + 2 * R):
                x = x0 - R + px
                if 0 <= y < H and 0 <= x < W:
                    win[u, px] = src[b, y, x]
        return win

    # ---- K1
    for b, sx, x0, y0, rows, P in units():
        A = np.zeros((2 * P, D * TW))
        for j in range(rows):
            for px in range(TW):
                if x0 + px < W:
                    for dd in range(D):
                        A[j][a_index(px, dd, dd % ND)] = gp[b, y0 + j, x0 + px, dd]
            store_row_blocks(ws, gp, b, x0, y0 + j, H, W)
        o = consumer(A, window(warp, b, x0, y0, P), P)
        for j in range(rows):
            n = min(TW, W - x0)
            dc1[b, y0 + j, x0:x0 + n] = o[j][:n]
    # ---- K2: the A row of dwarp row y, strip sx is the workspace tile as it is (one bulk copy)
    for b, sx, x0, y0, rows, P in units():
        A = np.zeros((2 * P, D * TW))
        for j in range(rows):
            A[j] = load_a_row(ws, b, x0, y0 + j, W)
            assert np.isfinite(A[j]).all(), "K2 read a workspace entry K1 never wrote"
        o = consumer(A, window(c1, b, x0, y0, P), P)
        for j in range(rows):
            n = min(TW, W - x0)
            dwarp[b, y0 + j, x0:x0 + n] = o[j][:n]
    assert np.abs(dc1 - dc1_ref).max() <= 2e-6 * np.abs(dc1_ref).max()
    assert np.abs(dwarp - dwarp_ref).max() <= 2e-6 * np.abs(dwarp_ref).max()


def test_k2_tensor_load_decomposition_matches_the_shifted_row():
    """produce_a_k2 / k2_issue_row / k2_finish_row: per ex one aligned box (32 px x 9 dyi) at x0 + 4 floor((ex - 4) / 4), two
    remainder boxes (4 px x 3 planes x 9 dyi: x0 + 28 for ex = 1..3, x0 + 32 for ex = 5..7) and an in-place shift by
    r = ex & 3 must give A[ex][dyi][px] = ws[y][dyi][8 - ex][x0 + px + ex - 4], zero outside [0, W)."""
    rng = np.random.default_rng(3)
    for W, x0 in ((64, 0), (64, 32), (40, 32), (33, 32), (20, 0)):
        ws_row = rng.standard_normal((ND, ND, W))          # [dyi][dxi][x] of one target row

        def box(xs, nx, dxi0, ndxi):                        # TMA tensor load: zero fill outside the tensor
            out = np.zeros((ND, ndxi, nx))
            for k in range(nx):
                if 0 <= xs + k < W:
                    out[:, :, k] = ws_row[:, dxi0:dxi0 + ndxi, xs + k]
            return out

        A = np.zeros((ND, ND, TW))                          # [ex][dyi][px]
        for ex in range(ND):
            c = 4 * ((ex - R) // 4)
            assert c % 4 == 0 and c in (-4, 0, 4)
            A[ex] = box(x0 + c, TW, 2 * R - ex, 1)[:, 0, :]
        rem = [box(x0 + TW - R, 4, 5, 3), box(x0 + TW, 4, 1, 3)]   # planes dxi 5..7 / 1..3
        for k, ex in enumerate((1, 2, 3, 5, 6, 7)):
            r = ex & 3
            rbox = rem[0 if k < 3 else 1]
            new = np.empty((ND, TW))
            for lane in range(TW):
                for dyi in range(ND):
                    new[dyi, lane] = A[ex][dyi, lane + r] if lane + r < TW else rbox[dyi, 3 - r, lane + r - TW]
            A[ex] = new
        for ex in range(ND):
            for px in range(TW):
                xs = x0 + px + ex - R
                want = ws_row[:, 2 * R - ex, xs] if 0 <= xs < W else np.zeros(ND)
                assert np.array_equal(A[ex][:, px], want), (W, x0, ex, px)

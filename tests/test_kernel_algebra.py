"""fp64 CPU emulations of the index algebra of the two kernels added late in round 1, thread-for-thread as the CUDA code
addresses its operands, against ATen / the oracle:

  * csrc/wgrad_sm100.cu — weight gradient as a GEMM over pixels: plain layers (gz box shifted against an unshifted x tile,
    one tap group per kernel row) and up layers (parity planes of G = FIR^T(gz), tap = (plane, shift) pairs);
  * csrc/fir_producer.cuh — the FIR pass of an up layer evaluated per consumer halo tile from a zero-filled window of the
    parity planes, horizontal pass first (He/Ho row results), then the vertical taps, zero padding outside the image.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import stylegan2_oracle as orc


def _zero_window(t, r0, c0, nr, nc):
    """t[..., r0:r0+nr, c0:c0+nc] with out-of-range rows / columns zero-filled (what a TMA box load returns)."""
    H, W = t.shape[-2:]
    out = t.new_zeros(t.shape[:-2] + (nr, nc))
    rs, re, cs, ce = max(r0, 0), min(r0 + nr, H), max(c0, 0), min(c0 + nc, W)
    if rs < re and cs < ce:
        out[..., rs - r0:re - r0, cs - c0:ce - c0] = t[..., rs:re, cs:ce]
    return out


@pytest.mark.parametrize('h,w,bw,bh', [(8, 8, 8, 8), (20, 20, 16, 4), (5, 7, 8, 8)])
def test_wgrad_plain_tap_groups(h, w, bw, bh):
    """wgrad_launch(up == 0): group ky loads the gz box at origin (y0 + 1 - ky, x0 - 1), size (bh, bw + 2); tap kx starts
    (2 - kx) pixels into it; the x tile is the unshifted (bh, bw) box; every pixel tile adds into the same accumulator."""
    g = torch.Generator().manual_seed(h * 100 + w)
    B, cin, cout = 2, 3, 4
    x = torch.randn(B, cin, h, w, generator=g, dtype=torch.float64)
    gz = torch.randn(B, cout, h, w, generator=g, dtype=torch.float64)
    gw = torch.zeros(cout, cin, 3, 3, dtype=torch.float64)
    for y0 in range(0, h, bh):
        for x0 in range(0, w, bw):
            xt = _zero_window(x, y0, x0, bh, bw)
            for ky in range(3):
                box = _zero_window(gz, y0 + 1 - ky, x0 - 1, bh, bw + 2)
                for kx in range(3):
                    a = box[..., :, 2 - kx:2 - kx + bw]
                    gw[:, :, ky, kx] += torch.einsum('boyx,biyx->oi', a, xt)
    ref = torch.nn.grad.conv2d_weight(x, (cout, cin, 3, 3), gz, padding=1)
    assert float((gw - ref).abs().max()) <= 1e-12


@pytest.mark.parametrize('h,w,bw,bh', [(8, 8, 8, 8), (12, 12, 16, 4)])
def test_wgrad_up_parity_plane_groups(h, w, bw, bh):
    """wgrad_launch(up == 2): operand = the four parity planes of G (gradient of the (2h+1)x(2w+1) conv_transpose2d output),
    plane 2*pu+pv at (I,J) = G[2I+pu, 2J+pv]; tap (ky,kx) reads plane (ky&1, kx&1) shifted by (ky==2, kx==2) inside a
    (bh+1, bw+1) box at the tile origin."""
    g = torch.Generator().manual_seed(h + 7)
    B, cin, cout = 2, 3, 4
    x = torch.randn(B, cin, h, w, generator=g, dtype=torch.float64)
    G = torch.randn(B, cout, 2 * h + 1, 2 * w + 1, generator=g, dtype=torch.float64)
    Gp = F.pad(G, (0, 1, 0, 1))
    planes = [Gp[:, :, pu::2, pv::2] for pu in (0, 1) for pv in (0, 1)]           # each (h+1) x (w+1)
    gw = torch.zeros(cout, cin, 3, 3, dtype=torch.float64)
    for y0 in range(0, h, bh):
        for x0 in range(0, w, bw):
            xt = _zero_window(x, y0, x0, bh, bw)
            for pl in range(4):
                pu, pv = pl >> 1, pl & 1
                box = _zero_window(planes[pl], y0, x0, bh + 1, bw + 1)
                for ky in range(pu, 3, 2):
                    for kx in range(pv, 3, 2):
                        a, b = int(ky == 2), int(kx == 2)
                        gw[:, :, ky, kx] += torch.einsum('boyx,biyx->oi', box[..., a:a + bh, b:b + bw], xt)
    ref = torch.nn.grad.conv2d_weight(G, (cin, cout, 3, 3), x, stride=2).transpose(0, 1)     # model.py:248-254 weight layout
    assert float((gw - ref).abs().max()) <= 1e-12


@pytest.mark.parametrize('hin,tile_w', [(8, 16), (8, 8), (16, 16)])
def test_fused_fir_producer_tile_algebra(hin, tile_w):
    """fir_produce_group_smem: for every consumer halo tile (origin (Y0, X0) odd, 18 x (tile_w + 2) pixels) the window of
    plane rows m_first-1 .. m_first+10 and columns n_first-1 .. is loaded zero-filled; row results
        He_r(px0) = gx3 ee[r][n+1] + gx2 eo[r][n] + gx1 ee[r][n] + gx0 eo[r][n-1],   He_r(px1) = gx3 eo[r][n+1] + gx2 ee[r][n+1] + ...
    (Ho from oe / oo) feed z(2m) = gy3 He[m+1] + gy2 Ho[m] + gy1 He[m] + gy0 Ho[m-1], z(2m+1) = gy3 Ho[m+1] + gy2 He[m+1] + ...;
    pixels outside the image are the convolution's zero padding.  Reference: upfirdn2d(pad=(1,1)) of the interleaved planes
    (= Blur after conv_transpose2d, model.py:72-88,256-257), then zero-padded by one pixel."""
    g = torch.Generator().manual_seed(hin * 10 + tile_w)
    C, win = 3, hin
    T = torch.randn(1, C, 2 * hin + 1, 2 * win + 1, generator=g, dtype=torch.float64)       # conv_transpose2d output
    fir = (orc.make_fir_kernel([1, 3, 3, 1]) * 4).double()
    ref = orc.upfirdn2d(T, fir, 1, 1, (1, 1))                                                # [1, C, 2hin, 2win]
    Ho, Wo = 2 * hin, 2 * win
    refp = F.pad(ref, (1, 1, 1, 1))                                                          # consumer's zero padding
    Tp = F.pad(T, (0, 1, 0, 1))
    plane = {(pu, pv): Tp[0, :, pu::2, pv::2] for pu in (0, 1) for pv in (0, 1)}             # [C, hin+1, win+1]
    # separable taps exactly as the kernel derives them: gy[a] = sum_i fir[3-a][i], gx[a] = sum_i fir[i][3-a] / total
    gy = [float(fir[3 - a].sum()) for a in range(4)]
    gx = [float(fir[:, 3 - a].sum() / fir.sum()) for a in range(4)]
    kHH, kHW = 18, tile_w + 2
    kcols = kHW // 2 + 1
    for ty in range((Ho + 15) // 16):
        for tx in range((Wo + tile_w - 1) // tile_w):
            Y0, X0 = ty * 16 - 1, tx * tile_w - 1
            m_first, n_first = (Y0 - 1) >> 1, (X0 - 1) >> 1
            win_ = {k: _zero_window(v, m_first - 1, n_first - 1, 12, kcols + 2) for k, v in plane.items()}
            tile = torch.zeros(C, kHH, kHW, dtype=torch.float64)
            for c in range(kcols):                       # lane's plane column n = n_first + c, window column c + 1
                def hrow(odd, r):
                    e, o = win_[(odd, 0)], win_[(odd, 1)]                 # even-column / odd-column plane of this row parity
                    p0 = gx[3] * e[:, r, c + 2] + gx[2] * o[:, r, c + 1] + gx[1] * e[:, r, c + 1] + gx[0] * o[:, r, c]
                    p1 = gx[3] * o[:, r, c + 2] + gx[2] * e[:, r, c + 2] + gx[1] * o[:, r, c + 1] + gx[0] * e[:, r, c + 1]
                    return p0, p1
                for mi in range(10):                     # plane row m = m_first + mi, window row mi + 1
                    r, m = mi + 1, m_first + mi
                    ho_prev, he_cur, ho_cur = hrow(1, r - 1), hrow(0, r), hrow(1, r)
                    he_next, ho_next = hrow(0, r + 1), hrow(1, r + 1)
                    for py in range(2):
                        Y = 2 * m + py
                        rr = Y - Y0
                        if rr < 0 or rr >= kHH:
                            continue
                        a3, a2, a1, a0 = (ho_next, he_next, ho_cur, he_cur) if py else (he_next, ho_cur, he_cur, ho_prev)
                        for px in range(2):
                            X = 2 * (n_first + c) + px
                            pc = X - X0
                            if pc < 0 or pc >= kHW:
                                continue
                            z = gy[3] * a3[px] + gy[2] * a2[px] + gy[1] * a1[px] + gy[0] * a0[px]
                            inside = 0 <= Y < Ho and 0 <= X < Wo
                            tile[:, rr, pc] = z if inside else 0.0
            want = _zero_window(refp[0], Y0 + 1, X0 + 1, kHH, kHW)        # refp index = image index + 1
            assert float((tile - want).abs().max()) <= 1e-12, (ty, tx)

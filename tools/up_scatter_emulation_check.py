"""fp64 emulation of the scatter up-conv (csrc/modconv_sm100.cu mode 2 + csrc/up_finish_sm100.cu) against
conv_transpose2d(stride 2) + upfirdn2d(blur, pad=(1,1)) as the reference computes it (model.py:246-257).
Checks the parity-plane / shift bookkeeping and the FIR indexing that the kernels implement.  CPU only."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import stylegan2_oracle as orc  # noqa: E402


def scatter_planes(x, w):
    """x [B,Cin,H,W], w [Cout,Cin,3,3] -> planes [B,4(oe,ee,eo,oo),Cout,H+1,W+1] exactly as the GEMM accumulates them."""
    B, Cin, H, W = x.shape
    Cout = w.shape[0]
    xp = F.pad(x, (1, 1, 1, 1))                               # x[-1..H, -1..W] with zeros (TMA OOB fill)
    t = x.new_zeros(B, 4, Cout, H + 1, W + 1)
    blocks = [(1, 0), (0, 0), (0, 1), (1, 1)]                 # (pu, pv) of oe, ee, eo, oo
    for sft in range(4):
        a, b = sft >> 1, sft & 1
        xs = xp[:, :, 1 - a:1 - a + H + 1, 1 - b:1 - b + W + 1]      # x[I-a, J-b] for I in 0..H, J in 0..W
        first = 1 if sft >= 2 else 0
        nblk = [4, 2, 2, 1][sft]
        for blk in range(first, first + nblk):
            pu, pv = blocks[blk]
            ky, kx = pu + 2 * a, pv + 2 * b
            assert ky <= 2 and kx <= 2
            t[:, blk] += torch.einsum('bchw,oc->bohw', xs, w[:, :, ky, kx])
    return t


def finish(t, fir, H, W):
    """FIR over the parity planes, thread-for-thread like up_finish_kernel."""
    B, _, C, Hp, Wp = t.shape
    out = t.new_zeros(B, C, 2 * H, 2 * W)
    kf = torch.flip(fir, [0, 1])
    for m in range(H):
        for n in range(W):
            for du in range(-1, 4):
                iu = m + (du >> 1)
                if iu < 0:
                    continue
                for dv in range(-1, 4):
                    iv = n + (dv >> 1)
                    if iv < 0:
                        continue
                    pu, pv = du & 1, dv & 1
                    plane = (3 if pv else 0) if pu else (2 if pv else 1)
                    v = t[:, plane, :, iu, iv]
                    for py in range(2):
                        a = du - py + 1
                        if a < 0 or a > 3:
                            continue
                        for px in range(2):
                            bq = dv - px + 1
                            if bq < 0 or bq > 3:
                                continue
                            out[:, :, 2 * m + py, 2 * n + px] += kf[a, bq] * v
    return out


def main():
    torch.manual_seed(0)
    worst = 0.0
    for (B, Cin, Cout, H, W) in [(2, 5, 7, 4, 4), (1, 3, 4, 5, 7), (1, 2, 2, 1, 1)]:
        x = torch.randn(B, Cin, H, W, dtype=torch.float64)
        w = torch.randn(Cout, Cin, 3, 3, dtype=torch.float64)
        fir = (orc.make_fir_kernel([1, 3, 3, 1]) * 4).double()
        fir = fir + 0.01 * torch.randn(4, 4, dtype=torch.float64)        # asymmetric taps: catches a missing flip
        ref = F.conv_transpose2d(x, w.transpose(0, 1), stride=2)
        ref = orc.upfirdn2d(ref, fir, 1, 1, (1, 1))
        got = finish(scatter_planes(x, w), fir, H, W)
        e = (got - ref).abs().max().item()
        worst = max(worst, e)
        print('B=%d Cin=%d Cout=%d %dx%d: max err %.2e' % (B, Cin, Cout, H, W, e))
    assert worst < 1e-12, worst
    print('ok')
    return worst


if __name__ == '__main__':
    main()

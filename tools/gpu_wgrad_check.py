"""GPU check of sgr_modconv_wgrad (csrc/wgrad_sm100.cu) against ATen in fp64, per tap, with timing.
   python tools/gpu_wgrad_check.py [quick | big | ncu]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stylegan_directions_face_reenactment_b200 import _native as N  # noqa: E402


def to_c8(x):
    b, c, h, w = x.shape
    out = torch.empty(2 * b * c * h * w, dtype=torch.bfloat16, device=x.device)
    N.check(N.lib().sgr_nchw_to_c8(N.ptr(x), None, N.ptr(out), b, c, h, w, 0, N.FMT_BF16, N.stream()), 'sgr_nchw_to_c8')
    return out


def planes_of(G, h, w):
    """G [B,C,2h+1,2w+1] -> [B,4C,h+1,w+1], channel (2pu+pv)*C + c at (I,J) = G[c,2I+pu,2J+pv] (0 outside)."""
    b, c = G.shape[:2]
    Gp = torch.zeros(b, c, 2 * h + 2, 2 * w + 2, device=G.device, dtype=G.dtype)
    Gp[:, :, :2 * h + 1, :2 * w + 1] = G
    return torch.cat([Gp[:, :, pu::2, pv::2] for pu in (0, 1) for pv in (0, 1)], 1).contiguous()


def wgrad(x, g, up, cout):
    b, cin, h, w = x.shape
    lib = N.lib()
    xc8 = to_c8(x)
    gc8 = to_c8(planes_of(g, h, w) if up else g)
    gw = torch.full((cout, cin, 3, 3), float('nan'), device=x.device)
    nbytes = lib.sgr_wgrad_scratch_bytes(cout, cin)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    a = N.WgradArgs()
    a.batch, a.cin, a.cout, a.h_in, a.w_in, a.up = b, cin, cout, h, w, 2 if up else 0
    a.x_c8, a.gz_c8, a.gw, a.scratch, a.scratch_bytes = N.ptr(xc8), N.ptr(gc8), N.ptr(gw), N.ptr(scratch), nbytes

    keep = (xc8, gc8, scratch)

    def run(_keep=keep):
        N.check(lib.sgr_modconv_wgrad(C.byref(a), N.stream()), 'sgr_modconv_wgrad')
    run()
    torch.cuda.synchronize()
    return gw, run


def reference(x, g, up, cout):
    cin = x.shape[1]
    xd, gd = x.double(), g.double()
    if up:
        return torch.nn.grad.conv2d_weight(gd, (cin, cout, 3, 3), xd, stride=2).transpose(0, 1).contiguous()
    return torch.nn.grad.conv2d_weight(xd, (cout, cin, 3, 3), gd, padding=1)


def case(b, cin, cout, h, up, timing=True):
    torch.manual_seed(b * 1000 + cin + cout + h + up)
    dev = 'cuda'
    x = torch.randn(b, cin, h, h, device=dev)
    g = torch.randn(b, cout, 2 * h + 1, 2 * h + 1, device=dev) if up else torch.randn(b, cout, h, h, device=dev)
    gw, run = wgrad(x, g, up, cout)
    ref = reference(x, g, up, cout)
    err = (gw.double() - ref).abs()
    scale = ref.abs().max().item()
    per_tap = (err.amax((0, 1)) / scale).flatten().tolist()
    rel = err.max().item() / scale
    ms = ms_lib = float('nan')
    if timing:
        xs, gs = x.contiguous(), g.contiguous()

        def lib_run():
            if up:
                return torch.nn.grad.conv2d_weight(gs, (cin, cout, 3, 3), xs, stride=2)
            return torch.nn.grad.conv2d_weight(xs, (cout, cin, 3, 3), gs, padding=1)
        for _ in range(3):
            lib_run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            lib_run()
        e1.record()
        torch.cuda.synchronize()
        ms_lib = e0.elapsed_time(e1) / 10
        for _ in range(3):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
    flops = 2.0 * 9 * cin * cout * b * h * h
    print('B=%d %d->%d %dx%d up=%d  rel_err=%.2e nan=%d  per-tap %s  %.3f ms  %.1f TFLOP/s  (cuDNN via ATen, TF32 %s: %.3f ms)'
          % (b, cin, cout, h, h, up, rel, int(torch.isnan(gw).sum()), ' '.join('%.0e' % t for t in per_tap), ms,
             flops / ms / 1e9, torch.backends.cudnn.allow_tf32, ms_lib), flush=True)
    return rel


def main():
    quick = len(sys.argv) > 1 and sys.argv[1] == 'quick'
    if len(sys.argv) > 1 and sys.argv[1] in ('big', 'ncu'):     # ncu: one launch per case (profiler capture)
        for c in [(8, 512, 512, 32, 0), (8, 256, 256, 64, 0), (8, 64, 64, 256, 0), (8, 128, 64, 128, 1), (32, 512, 512, 32, 0)]:
            case(*c, timing=sys.argv[1] == 'big')
        return
    cases = [(1, 128, 128, 16, 0), (1, 128, 128, 16, 1), (2, 64, 64, 8, 0), (2, 32, 32, 4, 0), (1, 256, 128, 32, 1),
             (3, 512, 512, 4, 0), (2, 512, 512, 8, 1)]
    if not quick:
        cases += [(1, 512, 512, 32, 0), (1, 512, 256, 32, 1), (1, 256, 256, 64, 0), (1, 128, 128, 128, 0),
                  (1, 128, 64, 128, 1), (1, 64, 64, 256, 0), (8, 512, 512, 32, 0), (8, 64, 64, 256, 0),
                  (1, 64, 64, 20, 0), (1, 64, 32, 12, 1)]
    worst = 0.0
    for c in cases:
        worst = max(worst, case(*c))
    print('worst rel err %.2e' % worst)


if __name__ == '__main__':
    main()

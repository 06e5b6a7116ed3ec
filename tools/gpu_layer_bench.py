"""Times single modulated-conv layers through the C ABI (sgr_modconv_forward) on random operands.
    python tools/gpu_layer_bench.py [B]
Prints per layer: GEMM kernel ms (libsgr's own event pair around the tcgen05 launch) and whole-call ms."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stylegan_directions_face_reenactment_b200 import _native as N  # noqa: E402

LAYERS = [  # (name, cin, cout, h_in, up)
    ('L5 up 16>32 512>512', 512, 512, 16, 2), ('L6 32 512>512', 512, 512, 32, 0), ('L7 up 32>64 512>256', 512, 256, 32, 2),
    ('L8 64 256>256', 256, 256, 64, 0), ('L9 up 64>128 256>128', 256, 128, 64, 2), ('L10 128 128>128', 128, 128, 128, 0),
    ('L11 up 128>256 128>64', 128, 64, 128, 2), ('L12 256 64>64', 64, 64, 256, 0)]


def run(B, only=None, reps=5):
    lib = N.lib()
    dev = torch.device('cuda')
    st = torch.cuda.current_stream().cuda_stream
    fir = (torch.outer(torch.tensor([1., 3., 3., 1.]), torch.tensor([1., 3., 3., 1.])) / 16).to(dev)
    for name, cin, cout, h, up in LAYERS:
        if only and only not in name:
            continue
        w = torch.randn(cout, cin, 3, 3, device=dev)
        nb = lib.sgr_packed_weight_bytes(cout, cin, 3, up, 0)
        packed = torch.empty(nb, dtype=torch.uint8, device=dev)
        wsq = torch.empty(cout, cin, device=dev)
        nt = 0 if up else lib.sgr_choose_column_tile(B, h, h, cout)
        N.check(lib.sgr_pack_modconv_weight(N.ptr(w), N.ptr(fir), cout, cin, 3, up, 0, 0, nt, N.ptr(packed), N.ptr(wsq), st), 'pack')
        x = torch.randn(B, cin, h, h, device=dev)
        xc8 = torch.empty(2 * B * cin * h * h, dtype=torch.bfloat16, device=dev)
        N.check(lib.sgr_nchw_to_c8(N.ptr(x), None, N.ptr(xc8), B, cin, h, h, 0, 0, st), 'c8')
        ho = 2 * h if up else h
        demod = torch.rand(B, cout, device=dev) + 0.5
        s2 = torch.rand(B, cout, device=dev) + 0.5
        bias = torch.randn(cout, device=dev)
        noise = torch.randn(ho, ho, device=dev)
        nw = torch.randn(1, device=dev)
        out = torch.empty(2 * B * cout * ho * ho, dtype=torch.bfloat16, device=dev)
        a = N.ConvArgs()
        a.batch, a.cin, a.cout, a.h_in, a.w_in = B, cin, cout, h, h
        a.ksize, a.up, a.act, a.act_gain = 3, up, 1, 2 ** 0.5
        a.operand_format = a.out_format = 0
        a.column_tile = nt
        a.single_pass = N.single_pass()
        a.x_c8, a.w_packed, a.demod, a.bias = N.ptr(xc8), N.ptr(packed), N.ptr(demod), N.ptr(bias)
        a.noise, a.noise_weight, a.s2, a.out_c8 = N.ptr(noise), N.ptr(nw), N.ptr(s2), N.ptr(out)
        if up == 2:
            scratch = torch.empty(lib.sgr_up_scratch_bytes(B, cout, h, h), dtype=torch.uint8, device=dev)
            a.t_scratch, a.fir = N.ptr(scratch), N.ptr(fir)
        for _ in range(2):
            N.check(lib.sgr_modconv_forward(C.byref(a), st), 'conv')
        torch.cuda.synchronize()
        lib.sgr_profile_enable(1)
        lib.sgr_profile_collect(None, 0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            N.check(lib.sgr_modconv_forward(C.byref(a), st), 'conv')
        e1.record()
        torch.cuda.synchronize()
        buf = (C.c_float * 64)()
        tags = (C.c_int * 64)()
        n = lib.sgr_profile_collect_tagged(buf, tags, 64)
        lib.sgr_profile_enable(0)
        gemm = sum(buf[i] for i in range(n) if tags[i] == 0) / reps
        total = e0.elapsed_time(e1) / reps
        fl = 2 * cin * cout * 9 * h * h * B
        print('%-26s gemm %.4f ms (%.0f algo TF/s)  call %.4f ms  [box %s]' % (name, gemm, fl / gemm / 1e9, total,
                                                                           os.environ.get('SGR_UP_BOX', 'auto')), flush=True)


if __name__ == '__main__':
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 32, sys.argv[2] if len(sys.argv) > 2 else None)

"""Host-side mirrors of the reference's two native operators, backed by libsgr.so.

Same names, argument meaning and error behaviour as libs/gan/StyleGAN2/op/{upfirdn2d.py:149-165, fused_act.py:73-86}
of the reference (inputs must be CUDA tensors; a CPU tensor raises, as the reference's fused op does at
op/fused_act.py:53-55).
"""
import math

import torch
from torch import nn

from . import _native as N


def _as_f32(t):
    if t.dtype != torch.float32:
        raise RuntimeError('libsgr kernels are fp32 (got %s)' % t.dtype)
    return t.contiguous()


def _upfirdn2d_raw(x, kernel, up, down, pad0, pad1):
    x = _as_f32(x)
    kernel = _as_f32(kernel)
    b, c, h, w = x.shape
    kh, kw = kernel.shape
    out_h = (h * up + pad0 + pad1 - kh + down) // down
    out_w = (w * up + pad0 + pad1 - kw + down) // down
    y = torch.empty(b, c, out_h, out_w, device=x.device, dtype=torch.float32)
    if y.numel() == 0 or x.numel() == 0:
        if y.numel():
            y.zero_()
        return y
    with torch.cuda.device(x.device):
        N.check(N.lib().sgr_upfirdn2d(N.ptr(x), N.ptr(y), N.ptr(kernel), b * c, h, w, up, down, pad0, pad1, kh, kw,
                                      N.stream()), 'sgr_upfirdn2d')
    return y


class _UpFirDn2d(torch.autograd.Function):
    """Autograd rule of op/upfirdn2d.py:89-146: the gradient is the same operator with up/down swapped,
    flipped taps and the g_pad of :112-117."""

    @staticmethod
    def forward(ctx, x, kernel, up, down, pad0, pad1):
        y = _upfirdn2d_raw(x, kernel, up, down, pad0, pad1)
        ctx.save_for_backward(kernel)
        ctx.cfg = (up, down, pad0, pad1, x.shape[2], x.shape[3], y.shape[2], y.shape[3])
        return y

    @staticmethod
    def backward(ctx, gy):
        kernel, = ctx.saved_tensors
        up, down, pad0, pad1, in_h, in_w, out_h, out_w = ctx.cfg
        kh, kw = kernel.shape
        # g_pad per axis as op/upfirdn2d.py:112-117: g_pad_x0 = kernel_w - pad_x0 - 1, g_pad_y0 = kernel_h - pad_y0 - 1,
        # g_pad_x1 = in_w * up - out_w * down + pad_x0 - up + 1 (likewise y).  sgr_upfirdn2d takes ONE (pad0, pad1) pair for
        # both axes (as the reference's Python entry point does), so the axes must agree — they do for every square FIR.
        gx0, gy0 = kw - pad0 - 1, kh - pad0 - 1
        gx1 = in_w * up - out_w * down + pad0 - up + 1
        gy1 = in_h * up - out_h * down + pad0 - up + 1
        if gx0 != gy0:
            raise NotImplementedError('upfirdn2d backward: non-square FIR kernels (%dx%d) need per-axis padding' % (kh, kw))
        # the trailing pads can differ by the decimation remainder: the larger one only appends outputs, sliced off below
        gx = _UpFirDn2d.apply(gy, torch.flip(kernel, [0, 1]), down, up, gx0, max(gx1, gy1))
        return gx[:, :, :in_h, :in_w], None, None, None, None, None


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
    if not input.is_cuda:
        raise RuntimeError('upfirdn2d: input must be a CUDA tensor (no CPU fallback)')
    return _UpFirDn2d.apply(input, kernel, up, down, pad[0], pad[1])


def _bias_act_raw(x, bias, ref, grad, slope, scale):
    x = _as_f32(x)
    y = torch.empty_like(x)
    if x.numel() == 0:
        return y
    outer = x.shape[0]
    channels = x.shape[1] if x.ndim > 1 else 1
    inner = 1
    for d in x.shape[2:]:
        inner *= d
    if x.ndim == 1:
        outer, channels = 1, x.shape[0]
    with torch.cuda.device(x.device):
        N.check(N.lib().sgr_fused_bias_act(N.ptr(x), N.ptr(bias), N.ptr(ref), N.ptr(y), outer, channels, inner, grad,
                                           slope, scale, N.stream()), 'sgr_fused_bias_act')
    return y


class _FusedLeakyReLU(torch.autograd.Function):
    """op/fused_act.py:19-70: forward saves the OUTPUT; backward masks on its sign; bias grad = sum over non-channel dims."""

    @staticmethod
    def forward(ctx, x, bias, slope, scale):
        y = _bias_act_raw(x, None if bias is None else _as_f32(bias), None, 0, slope, scale)
        ctx.save_for_backward(y)
        ctx.cfg = (slope, scale, bias is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
        y, = ctx.saved_tensors
        slope, scale, has_bias = ctx.cfg
        gx = _FusedLeakyReLUBackward.apply(gy, y, slope, scale)
        gb = None
        if has_bias:
            dims = [0] + list(range(2, gx.ndim))
            gb = gx.sum(dims)
        return gx, gb, None, None


class _FusedLeakyReLUBackward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gy, y, slope, scale):
        ctx.save_for_backward(y)
        ctx.cfg = (slope, scale)
        return _bias_act_raw(gy, None, y, 1, slope, scale)

    @staticmethod
    def backward(ctx, ggx):
        y, = ctx.saved_tensors
        slope, scale = ctx.cfg
        return _bias_act_raw(ggx, None, y, 1, slope, scale), None, None, None


def fused_leaky_relu(input, bias=None, negative_slope=0.2, scale=2 ** 0.5):
    if not input.is_cuda:
        raise RuntimeError('fused_leaky_relu: input must be a CUDA tensor (no CPU fallback)')
    return _FusedLeakyReLU.apply(input, bias, negative_slope, scale)


class FusedLeakyReLU(nn.Module):
    def __init__(self, channel, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel))
        self.negative_slope = negative_slope
        self.scale = scale

    def forward(self, input):
        return fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)


SQRT2 = math.sqrt(2.0)

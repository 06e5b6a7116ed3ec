"""dL/d(latent) through the synthesis network: one C call (sgr_synthesis_backward, csrc/backward.cu)."""
import ctypes as C

import torch

from . import _native as N
from .synthesis import _descriptor, _workspace


def synthesis_backward(g, lat, feats, noise, grad_image):
    if not grad_image.is_cuda:
        raise RuntimeError('synthesis backward needs CUDA tensors (no CPU fallback)')
    batch = lat.shape[0]
    dev = lat.device
    gimg = grad_image.contiguous().float()
    with torch.cuda.device(dev):
        desc = _descriptor(g, noise, batch, backward=True)
        ws = _workspace(g, desc, batch, dev, backward=True)
        dlat = torch.empty_like(lat)
        arr = (C.c_void_p * len(feats))(*[f.data_ptr() for f in feats])
        N.check(N.lib().sgr_synthesis_backward(C.byref(desc.struct), N.ptr(lat), batch, arr, N.ptr(gimg), N.ptr(dlat),
                                               N.ptr(ws), ws.numel(), N.stream()), 'sgr_synthesis_backward')
    return dlat

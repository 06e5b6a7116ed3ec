"""dL/d(latent) through the synthesis network: one C call (sgr_synthesis_backward, csrc/backward.cu).

Generator-parameter gradients (the `optimize_g` fine-tuning of libs/optimization.py:25-72) are
written by the same C call (sgr_backward_extras.params): the weight gradients by the tcgen05 weight-gradient GEMM
(csrc/wgrad_sm100.cu) on the operands the backward pass already holds, the per-channel ones (modulation linears, noise
weight, biases, ToRGB, constant input) by small reduction kernels (csrc/backward.cu).  Only a network with a layer packed
in polyphase mode (non-separable blur kernel, never the case for the reference's [1,3,3,1]) takes the fallback below,
which assembles the same gradients from dL/d(activation), dL/d(style) and dL/d(modulated constant input) with ATen
operators (conv2d_weight = cuDNN).  Parameter gradients are only formed in train() mode (optimize_g calls
generator.train(), :29); in eval() mode (A-matrix training, libs/trainer.py:111,144) the generator is frozen.
"""
import ctypes as C
import math

import torch
import torch.nn.functional as F

from . import _native as N
from .synthesis import _descriptor, _workspace, synthesis_param_list  # noqa: F401  (re-exported)

SQRT2 = math.sqrt(2.0)
FORCE_ATEN_WGRAD = False      # tools/gpu_optimize_g_bench.py: time the ATen/cuDNN weight gradients against the library's


def synthesis_backward(g, lat, feats, noise, grad_image, want_param_grads=False):
    if not grad_image.is_cuda:
        raise RuntimeError('synthesis backward needs CUDA tensors (no CPU fallback)')
    batch = lat.shape[0]
    dev = lat.device
    gimg = grad_image.contiguous().float()
    styled, rgbs = g.styled_layers(), g.rgb_layers()
    with torch.cuda.device(dev):
        desc = _descriptor(g, noise, batch, backward=True)
        ws = _workspace(g, desc, batch, dev, backward=True)
        dlat = torch.empty_like(lat)
        arr = (C.c_void_p * len(feats))(*[f.data_ptr() for f in feats])
        if not want_param_grads:
            N.check(N.lib().sgr_synthesis_backward_ex(C.byref(desc.struct), N.ptr(lat), batch, arr, N.ptr(gimg), N.ptr(dlat),
                                                      N.ptr(ws), ws.numel(), None, N.stream()), 'sgr_synthesis_backward')
            return dlat, None
        ex = N.BackwardExtras()
        if all(l.conv.up_mode() in (0, 2) for l in styled) and not FORCE_ATEN_WGRAD:
            # native path: every parameter gradient is written by the C call, straight into tensors shaped like the parameters
            params = synthesis_param_list(g)
            grads = [torch.empty_like(p, dtype=torch.float32) for p in params]
            weights = [l.conv.weight.detach().contiguous().float() for l in styled]
            pg = N.ParamGrads()
            pg.g_const_input = grads[0].data_ptr()
            k = 1
            for i, l in enumerate(styled):
                e = pg.styled[i]
                e.weight = weights[i].data_ptr()
                e.g_weight, e.g_mod_weight, e.g_mod_bias, e.g_noise_weight, e.g_act_bias = [t.data_ptr() for t in grads[k:k + 5]]
                k += 5
            for i, l in enumerate(rgbs):
                e = pg.rgb[i]
                e.g_weight, e.g_mod_weight, e.g_mod_bias, e.g_bias = [t.data_ptr() for t in grads[k:k + 4]]
                k += 4
            nscratch = N.lib().sgr_synthesis_wgrad_scratch_bytes(C.byref(desc.struct), batch)
            wscratch = g._workspace.get((batch, dev.index, 'wgrad'))
            if wscratch is None or wscratch.numel() < nscratch:
                for k in [k for k in g._workspace if k[2] == 'wgrad']:      # one live scratch (as for the workspaces)
                    del g._workspace[k]
                wscratch = torch.empty(nscratch, dtype=torch.uint8, device=dev)
                g._workspace[(batch, dev.index, 'wgrad')] = wscratch
            ex.params, ex.wgrad_scratch, ex.wgrad_scratch_bytes = C.pointer(pg), wscratch.data_ptr(), nscratch
            N.check(N.lib().sgr_synthesis_backward_ex(C.byref(desc.struct), N.ptr(lat), batch, arr, N.ptr(gimg), N.ptr(dlat),
                                                      N.ptr(ws), ws.numel(), C.byref(ex), N.stream()), 'sgr_synthesis_backward')
            return dlat, grads
        # fallback (a layer packed in polyphase mode = non-separable blur kernel, or FORCE_ATEN_WGRAD): assemble the
        # gradients from the extra outputs with ATen operators
        gfeats = [torch.empty_like(f) for f in feats]
        ds_styled = [torch.empty(batch, l.conv.in_channel, device=dev) for l in styled]
        ds_rgb = [torch.empty(batch, l.conv.in_channel, device=dev) for l in rgbs]
        g_input = torch.empty(batch, styled[0].conv.in_channel, 4, 4, device=dev)
        a1 = (C.c_void_p * len(gfeats))(*[t.data_ptr() for t in gfeats])
        a2 = (C.c_void_p * len(ds_styled))(*[t.data_ptr() for t in ds_styled])
        a3 = (C.c_void_p * len(ds_rgb))(*[t.data_ptr() for t in ds_rgb])
        ex.gfeats, ex.ds_styled, ex.ds_rgb, ex.g_input = a1, a2, a3, g_input.data_ptr()
        N.check(N.lib().sgr_synthesis_backward_ex(C.byref(desc.struct), N.ptr(lat), batch, arr, N.ptr(gimg), N.ptr(dlat),
                                                  N.ptr(ws), ws.numel(), C.byref(ex), N.stream()), 'sgr_synthesis_backward')
        return dlat, _param_grads(g, lat, feats, noise, gimg, gfeats, ds_styled, ds_rgb, g_input)


@torch.no_grad()
def _param_grads(g, lat, feats, noise, gimg, gfeats, ds_styled, ds_rgb, g_input):
    """Gradients in synthesis_param_list order (SURVEY.md §9.4 extended to the weights):
       gt = ga * sqrt2 * (a > 0 ? 1 : 0.2);  gz = gt * d;  q = sum_p gt * (z d)
       dW = scale * [ wgrad(x s, gz)  -  W_bar * ((q d^2)^T @ s^2) ]        (convolution + demodulation terms)
       d(mod.weight) = ds^T @ w / sqrt(512), d(mod.bias) = sum_b ds, d(bias) = sum gt, d(noise.weight) = sum gt * noise"""
    from .ops import upfirdn2d
    styled, rgbs = g.styled_layers(), g.rgb_layers()
    batch = lat.shape[0]
    grads = []
    s0 = styled[0].conv.modulation(lat[:, 0])
    grads.append((g_input * s0[:, :, None, None]).sum(0, keepdim=True))                      # ConstantInput.input
    inv_sqrt_style = 1.0 / math.sqrt(lat.shape[2])
    for l, layer in enumerate(styled):
        conv = layer.conv
        row = 0 if l == 0 else l
        w_lat = lat[:, row]
        s = conv.modulation(w_lat)                                                           # [B,Cin]
        wbar = conv.weight[0] * conv.scale                                                   # [Cout,Cin,3,3]
        d = torch.rsqrt((s * s) @ wbar.pow(2).sum((2, 3)).t() + 1e-8)                        # [B,Cout]
        a, ga = feats[l], gfeats[l]
        pos = a > 0
        gt = ga * torch.where(pos, SQRT2, 0.2 * SQRT2)
        nz = noise[l]
        nzw = layer.noise.weight * nz                                                        # [1 or B,1,H,W]
        t = torch.where(pos, a / SQRT2, a * (5.0 / SQRT2))
        zd = t - nzw - layer.activate.bias.view(1, -1, 1, 1)
        q = (gt * zd).sum((2, 3))                                                            # [B,Cout]
        gz = gt * d[:, :, None, None]
        x_in = feats[l - 1] if l > 0 else g.input.input.expand(batch, -1, -1, -1)
        xs = x_in * s[:, :, None, None]
        if conv.upsample:
            gfull = upfirdn2d(gz, torch.flip(conv.blur.kernel, [0, 1]), pad=(2, 2))          # FIR^T -> (2H+1)^2
            cin, cout = conv.in_channel, conv.out_channel
            gw_t = torch.nn.grad.conv2d_weight(gfull, (cin, cout, 3, 3), xs.contiguous(), stride=2)
            gwbar = gw_t.transpose(0, 1)
        else:
            gwbar = torch.nn.grad.conv2d_weight(xs.contiguous(), tuple(wbar.shape), gz.contiguous(), padding=1)
        gwbar = gwbar - wbar * ((q * d * d).t() @ (s * s))[:, :, None, None]
        grads.append((gwbar * conv.scale).unsqueeze(0))                                      # conv.weight [1,Cout,Cin,3,3]
        ds = ds_styled[l]
        grads.append(ds.t() @ w_lat * inv_sqrt_style)                                        # modulation.weight [Cin,512]
        grads.append(ds.sum(0))                                                              # modulation.bias
        grads.append((gt * nz).sum().view(1))                                                # noise.weight
        grads.append(gt.sum((0, 2, 3)))                                                      # activate.bias
    # ToRGB: gradient of every level's rgb output = adjoint chain of the 2x FIR upsampling of the skip
    grgb = [None] * len(rgbs)
    grgb[-1] = gimg
    for r in range(len(rgbs) - 1, 0, -1):
        grgb[r - 1] = upfirdn2d(grgb[r], torch.flip(rgbs[r].upsample.kernel, [0, 1]), down=2, pad=(1, 1))
    for r, layer in enumerate(rgbs):
        conv = layer.conv
        row = 1 if r == 0 else 2 * r + 1
        w_lat = lat[:, row]
        s = conv.modulation(w_lat)
        a = feats[0 if r == 0 else 2 * r]
        gw = torch.einsum('bchw,bihw,bi->ci', grgb[r], a, s) * conv.scale
        grads.append(gw.view(1, 3, -1, 1, 1))                                                # conv.weight [1,3,Cin,1,1]
        ds = ds_rgb[r]
        grads.append(ds.t() @ w_lat * inv_sqrt_style)
        grads.append(ds.sum(0))
        grads.append(grgb[r].sum((0, 2, 3)).view(1, 3, 1, 1))                                # ToRGB.bias
    return grads

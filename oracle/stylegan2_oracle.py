"""CPU oracle for the StyleGAN2 synthesis + latent-direction hot path.

TEST INFRASTRUCTURE ONLY.  This file is a CPU restatement (torch-CPU / numpy,
fp32 or fp64) of the reference algorithm in
StelaBou/stylegan_directions_face_reenactment, written functionally over a
plain ``state_dict`` so that it shares no code with the product package.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import it; the product path (``stylegan_directions_face_reenactment_b200``) never does.

Pinning: the reference ships no tests / golden vectors, so the
oracle is pinned against outputs of the reference itself, generated in the
build container by ``oracle/make_golden.py`` (which imports the unmodified
reference modules from /root/reference) and committed under ``tests/golden/``.
``tests/test_oracle_golden.py`` checks every function here against them.

The dense convolution arithmetic is the same third-party code the reference
calls (ATen ``conv2d`` / ``conv_transpose2d``, torch 2.11 in this image;
reference call sites ``libs/gan/StyleGAN2/model.py:254,263,269``).

All ``file:line`` citations are relative to /root/reference.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

SQRT2 = math.sqrt(2.0)


# ----------------------------------------------------------------------------
# FIR helpers
# ----------------------------------------------------------------------------
def make_fir_kernel(taps):
    """Normalised 2-D FIR from 1-D taps (libs/gan/StyleGAN2/model.py:19-27)."""
    k = torch.as_tensor(taps, dtype=torch.float32)
    if k.ndim == 1:
        k = torch.outer(k, k)
    return k / k.sum()


def upfirdn2d(x, kernel, up=1, down=1, pad=(0, 0)):
    """Zero-insert upsample, pad/crop, true 2-D convolution with ``kernel``, decimate.

    Semantics of libs/gan/StyleGAN2/op/upfirdn2d.py:168-209 (``upfirdn2d_native``)
    and of the CUDA kernel op/upfirdn2d_kernel.cu:52-137 (taps flipped at :77).
    ``x`` is [B, C, H, W]; ``pad`` is (pad0, pad1) applied to both axes.
    Written as shift-and-accumulate so it does not lean on F.conv2d.
    """
    pad0, pad1 = pad
    b, c, h, w = x.shape
    kh, kw = kernel.shape
    # zero insertion: sample (i, j) lands on (i*up, j*up)
    up_h, up_w = h * up, w * up
    u = x.new_zeros(b, c, up_h, up_w)
    u[:, :, ::up, ::up] = x
    # pad (negative pad crops)
    u = F.pad(u, [max(pad0, 0), max(pad1, 0), max(pad0, 0), max(pad1, 0)])
    u = u[:, :, max(-pad0, 0): u.shape[2] - max(-pad1, 0), max(-pad0, 0): u.shape[3] - max(-pad1, 0)]
    full_h = u.shape[2] - kh + 1
    full_w = u.shape[3] - kw + 1
    kf = torch.flip(kernel, [0, 1]).to(x.dtype)
    acc = x.new_zeros(b, c, full_h, full_w)
    for ky in range(kh):
        for kx in range(kw):
            acc = acc + kf[ky, kx] * u[:, :, ky:ky + full_h, kx:kx + full_w]
    out = acc[:, :, ::down, ::down]
    out_h = (h * up + pad0 + pad1 - kh + down) // down      # upfirdn2d.py:104-105
    out_w = (w * up + pad0 + pad1 - kw + down) // down
    assert out.shape[2] == out_h and out.shape[3] == out_w
    return out


# ----------------------------------------------------------------------------
# Elementwise / linear pieces
# ----------------------------------------------------------------------------
def fused_leaky_relu(x, bias, negative_slope=0.2, scale=SQRT2):
    """y = lrelu(x + bias[c]) * scale, bias on dim 1.

    op/fused_bias_act_kernel.cu:24-47 (act*10+grad == 30) with the bias index of
    :67-71, called from op/fused_act.py:51-58,85-86.
    """
    shape = [1, -1] + [1] * (x.ndim - 2)
    v = x + bias.reshape(shape)
    return torch.where(v > 0, v, v * negative_slope) * scale


def fused_leaky_relu_backward(grad_out, out, negative_slope=0.2, scale=SQRT2):
    """(grad_input, grad_bias) keyed on the sign of the saved OUTPUT.

    op/fused_bias_act_kernel.cu:43 (case 31) and op/fused_act.py:21-39.
    """
    gx = torch.where(out > 0, grad_out, grad_out * negative_slope) * scale
    dims = [0] + list(range(2, gx.ndim))
    return gx, gx.sum(dims)


def pixel_norm(z):
    """libs/gan/StyleGAN2/model.py:11-16."""
    return z * torch.rsqrt(torch.mean(z * z, dim=1, keepdim=True) + 1e-8)


def equal_linear(x, weight, bias, lr_mul=1.0, activation=False):
    """libs/gan/StyleGAN2/model.py:129-157 (weights are stored divided by lr_mul)."""
    scale = (1.0 / math.sqrt(weight.shape[1])) * lr_mul
    if activation:
        return fused_leaky_relu(F.linear(x, weight * scale), bias * lr_mul)
    return F.linear(x, weight * scale, bias=bias * lr_mul)


def mapping_network(sd, z, n_mlp=8, lr_mlp=0.01):
    """PixelNorm + n_mlp EqualLinear('fused_lrelu') (model.py:378-387)."""
    h = pixel_norm(z)
    for i in range(1, n_mlp + 1):
        h = equal_linear(h, sd['style.%d.weight' % i], sd['style.%d.bias' % i], lr_mul=lr_mlp, activation=True)
    return h


# ----------------------------------------------------------------------------
# Modulated convolution, exactly in the reference's materialised-weight form
# ----------------------------------------------------------------------------
def modulated_conv2d(x, w_latent, weight, mod_weight, mod_bias, demodulate=True, upsample=False,
                     blur_taps=(1, 3, 3, 1)):
    """libs/gan/StyleGAN2/model.py:232-273 (plain and upsample branches).

    x [B,Cin,H,W]; w_latent [B,512]; weight [1,Cout,Cin,k,k];
    modulation EqualLinear(512->Cin, bias_init=1) given by mod_weight/mod_bias.
    """
    b, cin, h, w = x.shape
    _, cout, _, k, _ = weight.shape
    s = equal_linear(w_latent, mod_weight, mod_bias).view(b, 1, cin, 1, 1)      # :235
    scale = 1.0 / math.sqrt(cin * k * k)                                         # :213-214
    wt = scale * weight * s                                                      # :236
    if demodulate:
        d = torch.rsqrt(wt.pow(2).sum([2, 3, 4]) + 1e-8)                         # :239
        wt = wt * d.view(b, cout, 1, 1, 1)
    if upsample:
        wt_t = wt.transpose(1, 2).reshape(b * cin, cout, k, k)                   # :248-253
        out = F.conv_transpose2d(x.reshape(1, b * cin, h, w), wt_t, padding=0, stride=2, groups=b)
        out = out.view(b, cout, out.shape[2], out.shape[3])
        factor = 2
        p = (len(blur_taps) - factor) - (k - 1)                                  # :197-202
        pad0 = (p + 1) // 2 + factor - 1
        pad1 = p // 2 + 1
        fir = make_fir_kernel(blur_taps).to(x.dtype) * (factor ** 2)             # :75-79
        out = upfirdn2d(out, fir, pad=(pad0, pad1))                              # :257
    else:
        out = F.conv2d(x.reshape(1, b * cin, h, w), wt.view(b * cout, cin, k, k), padding=k // 2, groups=b)
        out = out.view(b, cout, out.shape[2], out.shape[3])                      # :267-271
    return out


def styled_conv(sd, prefix, x, w_latent, noise, upsample):
    """conv -> (+ noise.weight * noise) -> fused bias + lrelu*sqrt2 (model.py:331-337)."""
    out = modulated_conv2d(x, w_latent, sd[prefix + '.conv.weight'], sd[prefix + '.conv.modulation.weight'],
                           sd[prefix + '.conv.modulation.bias'], demodulate=True, upsample=upsample)
    out = out + sd[prefix + '.noise.weight'] * noise                             # :282-287
    return fused_leaky_relu(out, sd[prefix + '.activate.bias'])


def to_rgb(sd, prefix, x, w_latent, skip=None, blur_taps=(1, 3, 3, 1)):
    """1x1 modulated conv without demodulation + bias + upsampled skip (model.py:350-359)."""
    out = modulated_conv2d(x, w_latent, sd[prefix + '.conv.weight'], sd[prefix + '.conv.modulation.weight'],
                           sd[prefix + '.conv.modulation.bias'], demodulate=False, upsample=False)
    out = out + sd[prefix + '.bias']
    if skip is not None:
        factor = 2
        fir = make_fir_kernel(blur_taps).to(x.dtype) * (factor ** 2)            # :35
        p = fir.shape[0] - factor
        pad = ((p + 1) // 2 + factor - 1, p // 2)                                # :38-43
        out = out + upfirdn2d(skip, fir, up=factor, down=1, pad=pad)
    return out


# ----------------------------------------------------------------------------
# Generator.forward
# ----------------------------------------------------------------------------
def synthesis_config(size, channel_multiplier):
    """Channel table and layer counts (model.py:389-447)."""
    channels = {4: 512, 8: 512, 16: 512, 32: 512, 64: 256 * channel_multiplier, 128: 128 * channel_multiplier,
                256: 64 * channel_multiplier, 512: 32 * channel_multiplier, 1024: 16 * channel_multiplier}
    log_size = int(math.log(size, 2))
    return channels, log_size, (log_size - 2) * 2 + 1, log_size * 2 - 2


def generator_forward(sd, styles, size, channel_multiplier=2, n_mlp=8, truncation=1.0, truncation_latent=None,
                      input_is_latent=False, noise=None, return_latents=False, return_features=False):
    """Generator.forward (model.py:471-539), single-style branch (the only one any caller uses).

    ``sd`` maps reference state_dict names to CPU tensors.  Returns (image, latent|None)
    and, with return_features, also the list of per-layer activations.
    """
    channels, log_size, num_layers, n_latent = synthesis_config(size, channel_multiplier)
    if not input_is_latent:
        styles = [mapping_network(sd, s, n_mlp) for s in styles]                # :484-485
    if noise is None:
        noise = [sd['noises.noise_%d' % i] for i in range(num_layers)]           # :488-492
    if truncation < 1:
        styles = [truncation_latent + truncation * (s - truncation_latent) for s in styles]   # :494-500
    assert len(styles) == 1
    latent = styles[0]
    if latent.ndim < 3:
        latent = latent.unsqueeze(1).repeat(1, n_latent, 1)                      # :502-508
    feats = []
    b = latent.shape[0]
    out = sd['input.input'].repeat(b, 1, 1, 1)                                   # :296-300
    out = styled_conv(sd, 'conv1', out, latent[:, 0], noise[0], upsample=False)  # :520
    feats.append(out)
    skip = to_rgb(sd, 'to_rgb1', out, latent[:, 1])                              # :521
    i = 1
    for blk in range(log_size - 2):                                              # :526-532
        out = styled_conv(sd, 'convs.%d' % (2 * blk), out, latent[:, i], noise[2 * blk + 1], upsample=True)
        feats.append(out)
        out = styled_conv(sd, 'convs.%d' % (2 * blk + 1), out, latent[:, i + 1], noise[2 * blk + 2], upsample=False)
        feats.append(out)
        skip = to_rgb(sd, 'to_rgbs.%d' % blk, out, latent[:, i + 2], skip)
        i += 2
    ret = (skip, latent if return_latents else None)
    if return_features:
        ret = ret + (feats,)
    return ret


# ----------------------------------------------------------------------------
# Direction matrix + glue
# ----------------------------------------------------------------------------
def direction_matrix_forward(weight, bias, dp, shift_dim=512, num_layers=8, w_plus=True):
    """libs/models/direction_matrix.py:41-48: shift = dp @ W^T + b, viewed [B, num_layers, shift_dim]."""
    dp = dp.reshape(-1, weight.shape[1])
    out = F.linear(dp, weight, bias)
    if w_plus:
        out = out.view(dp.shape[0], num_layers, shift_dim)
    return out


def shifted_latent_code(latent, shift):
    """W+ branch of libs/utilities/generic.py:116-135: clone, add shift into the first rows."""
    out = latent.clone()
    out[:, :shift.shape[1], :] += shift
    return out


def generate_image(sd, latent_code, truncation, trunc, size, channel_multiplier, shift_code=None):
    """libs/utilities/generic.py:137-152 for W+ inputs (input_is_latent=True)."""
    code = latent_code if shift_code is None else shifted_latent_code(latent_code, shift_code)
    img, lat = generator_forward(sd, [code], size, channel_multiplier, truncation=truncation,
                                 truncation_latent=trunc, input_is_latent=True, return_latents=True)
    if img.shape[2] > 256:
        img = F.adaptive_avg_pool2d(img, (256, 256))                              # :146-148
    return img, lat


# ----------------------------------------------------------------------------
# Deterministic weights shared by the golden generator, the tests and the bench
# ----------------------------------------------------------------------------
def state_dict_manifest(size, channel_multiplier=2, style_dim=512, n_mlp=8):
    """(name, shape) list in the reference's state_dict order (dumped from the reference)."""
    channels, log_size, num_layers, _ = synthesis_config(size, channel_multiplier)
    m = []
    for i in range(1, n_mlp + 1):
        m += [('style.%d.weight' % i, (style_dim, style_dim)), ('style.%d.bias' % i, (style_dim,))]
    m.append(('input.input', (1, channels[4], 4, 4)))

    def styled(prefix, cin, cout, up):
        r = [(prefix + '.conv.weight', (1, cout, cin, 3, 3))]
        if up:
            r.append((prefix + '.conv.blur.kernel', (4, 4)))
        r += [(prefix + '.conv.modulation.weight', (cin, style_dim)), (prefix + '.conv.modulation.bias', (cin,)),
              (prefix + '.noise.weight', (1,)), (prefix + '.activate.bias', (cout,))]
        return r

    def rgb(prefix, cin, up):
        r = [(prefix + '.bias', (1, 3, 1, 1))]
        if up:
            r.append((prefix + '.upsample.kernel', (4, 4)))
        r += [(prefix + '.conv.weight', (1, 3, cin, 1, 1)), (prefix + '.conv.modulation.weight', (cin, style_dim)),
              (prefix + '.conv.modulation.bias', (cin,))]
        return r

    m += styled('conv1', channels[4], channels[4], False)
    m += rgb('to_rgb1', channels[4], False)
    cin = channels[4]
    convs, rgbs = [], []
    for i in range(3, log_size + 1):
        cout = channels[2 ** i]
        convs += styled('convs.%d' % (2 * (i - 3)), cin, cout, True)
        convs += styled('convs.%d' % (2 * (i - 3) + 1), cout, cout, False)
        rgbs += rgb('to_rgbs.%d' % (i - 3), cout, True)
        cin = cout
    m += convs + rgbs
    for l in range(num_layers):
        res = (l + 5) // 2
        m.append(('noises.noise_%d' % l, (1, 1, 2 ** res, 2 ** res)))
    return m


def seeded_state_dict(size, channel_multiplier=2, seed=0, style_dim=512, n_mlp=8, lr_mlp=0.01):
    """A full generator state_dict drawn from ONE numpy PCG64 stream, independent of any module class.

    Distributions follow the reference initialisers (conv/linear weights N(0,1), mapping weights
    N(0,1)/lr_mlp, modulation bias 1) except that the parameters the reference initialises to ZERO
    (noise.weight, activate.bias, ToRGB.bias; model.py:280,348, fused_act.py:77) are drawn N(0,0.1)
    so that parity tests exercise them (SURVEY §8a note 4).
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    fir4 = make_fir_kernel([1, 3, 3, 1]) * 4.0
    sd = {}
    for name, shape in state_dict_manifest(size, channel_multiplier, style_dim, n_mlp):
        if name.endswith('blur.kernel') or name.endswith('upsample.kernel'):
            t = fir4.clone()
        else:
            t = torch.from_numpy(rng.standard_normal(shape, dtype=np.float32))
            if name.startswith('style.') and name.endswith('.weight'):
                t = t / lr_mlp
            elif name.startswith('style.') and name.endswith('.bias'):
                t = t * 0.1 / lr_mlp
            elif name.endswith('modulation.bias'):
                t = 1.0 + 0.1 * t
            elif name.endswith('noise.weight') or name.endswith('activate.bias') or (
                    name.endswith('.bias') and 'to_rgb' in name):
                t = 0.1 * t
        sd[name] = t
    return sd


def seeded_wplus(sd, batch, n_latent, seed, n_mlp=8):
    """W+ codes = mapping(randn) per row (SURVEY §8d cfg 2), from a numpy stream."""
    rng = np.random.Generator(np.random.PCG64(seed))
    z = torch.from_numpy(rng.standard_normal((batch * n_latent, 512), dtype=np.float32))
    with torch.no_grad():
        return mapping_network(sd, z, n_mlp).view(batch, n_latent, 512).contiguous()


def forward_flops_per_frame(size, channel_multiplier):
    """Algorithmic forward FLOPs per frame: 2*Cin*Cout*k^2*P, P=H_in^2 for up layers."""
    channels, log_size, _, _ = synthesis_config(size, channel_multiplier)
    fl = 2 * 512 * 512 * 9 * 16 + 2 * 512 * 3 * 16
    cin = 512
    for i in range(3, log_size + 1):
        cout = channels[2 ** i]
        hin, hout = 2 ** (i - 1), 2 ** i
        fl += 2 * cin * cout * 9 * hin * hin + 2 * cout * cout * 9 * hout * hout + 2 * cout * 3 * hout * hout
        cin = cout
    return fl


def frames_to_uint8(images, size=None):
    """Output stage as the reference does it: optional AdaptiveAvgPool2d (generic.py:146-148), tensor_to_image
    (libs/utilities/image_utils.py:97-111) and np.uint8 (libs/utilities/utils_inference.py:16).  [B,3,H,W] -> uint8 [B,h,w,3]."""
    import numpy as np
    x = images.detach().float().clone()
    if size is not None and x.shape[2] > size:
        x = F.adaptive_avg_pool2d(x, (size, size))
    x.clamp_(min=-1, max=1)
    x.add_(1).div_(1 - (-1) + 1e-5)
    x = x.mul(255.0).add(0.0)
    return np.uint8(np.transpose(x.numpy(), (0, 2, 3, 1)))

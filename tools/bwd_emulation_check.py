"""CPU fp64 check of the backward formulas implemented in csrc/backward.cu against torch autograd through the oracle.
Run here (no GPU): python tools/bwd_emulation_check.py"""
import math
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import stylegan2_oracle as orc  # noqa: E402

torch.set_default_dtype(torch.float64)
size, cm, batch = 32, 2, 2
sd32 = orc.seeded_state_dict(size, cm, seed=12)
sd = {k: v.double() for k, v in sd32.items()}
channels, log_size, num_layers, n_latent = orc.synthesis_config(size, cm)
wplus = orc.seeded_wplus(sd32, batch, n_latent, seed=21).double()
torch.manual_seed(0)
r = torch.randn(batch, 3, size, size)
wr = wplus.clone().requires_grad_(True)
img, _, feats = orc.generator_forward(sd, [wr], size, cm, input_is_latent=True, return_features=True)
(img * r).sum().backward()
gref = wr.grad

names = ['conv1'] + ['convs.%d' % i for i in range(num_layers - 1)]
rgbn = ['to_rgb1'] + ['to_rgbs.%d' % i for i in range(log_size - 2)]
L, R = len(names), len(rgbn)
lat = wplus


def style(prefix, row):
    return orc.equal_linear(lat[:, row], sd[prefix + '.conv.modulation.weight'], sd[prefix + '.conv.modulation.bias'])


s = [style(n, l) for l, n in enumerate(names)]
rows_rgb = [1 if i == 0 else 2 * i + 1 for i in range(R)]
srgb = [style(n, rows_rgb[i]) for i, n in enumerate(rgbn)]
fir = orc.make_fir_kernel([1, 3, 3, 1]).double() * 4
grgb = [None] * R
grgb[R - 1] = r
for i in range(R - 1, 0, -1):
    grgb[i - 1] = orc.upfirdn2d(grgb[i], torch.flip(fir, [0, 1]), 1, 2, (1, 1))
dlat = torch.zeros_like(lat)
gx_next = None
ds_conv = [None] * L
ds_demod = [None] * L
for l in range(L - 1, -1, -1):
    n = names[l]
    W = sd[n + '.conv.weight'][0]
    cout, cin = W.shape[:2]
    up = (l % 2 == 1)
    Wb = W / math.sqrt(cin * 9)
    Wsq = (Wb ** 2).sum([2, 3])
    d = torch.rsqrt((s[l] ** 2) @ Wsq.t() + 1e-8)
    a = feats[l].detach()
    nz = sd[n + '.noise.weight'] * sd['noises.noise_%d' % l]
    bias = sd[n + '.activate.bias'].view(1, -1, 1, 1)
    ga = torch.zeros_like(a)
    if gx_next is not None:
        ga = ga + gx_next * s[l + 1].view(batch, -1, 1, 1)
        ds_conv[l + 1] = (a * gx_next).sum([2, 3])
    if not up:
        i = (l + 1) // 2
        wr_ = sd[rgbn[i] + '.conv.weight'][0, :, :, 0, 0] / math.sqrt(cout)
        rr = torch.einsum('bchw,co->bohw', grgb[i], wr_)
        ga = ga + rr * srgb[i].view(batch, -1, 1, 1)
        dsr = (a * rr).sum([2, 3])
        dlat[:, rows_rgb[i]] += dsr @ (sd[rgbn[i] + '.conv.modulation.weight'] / math.sqrt(512))
    pos = a > 0
    gt = ga * torch.where(pos, math.sqrt(2), 0.2 * math.sqrt(2))
    t = torch.where(pos, a / math.sqrt(2), a * 5 / math.sqrt(2))
    q = (gt * (t - nz - bias)).sum([2, 3])
    gz = gt * d.view(batch, -1, 1, 1)
    if up:
        gb = orc.upfirdn2d(gz, torch.flip(fir, [0, 1]), 1, 1, (2, 2))
        gx = F.conv2d(gb, Wb.transpose(0, 1).contiguous(), stride=2)
    else:
        gx = F.conv_transpose2d(gz, Wb, padding=1)
    gx_next = gx
    ds_demod[l] = -s[l] * ((q * d * d) @ Wsq)
ds_conv[0] = (sd['input.input'] * gx_next).sum([2, 3])
for l in range(L):
    dlat[:, l] += (ds_conv[l] + ds_demod[l]) @ (sd[names[l] + '.conv.modulation.weight'] / math.sqrt(512))
for row in range(n_latent):
    print('row %2d  max|ref| %.3e  err %.3e' % (row, gref[:, row].abs().max(), (dlat[:, row] - gref[:, row]).abs().max()))
print('cancellation between the conv and demod terms of ds (a bf16x3 error on either is amplified by this ratio):')
for l in range(L):
    tot = ds_conv[l] + ds_demod[l]
    print('layer %2d |ds_conv| %.3e |ds_demod| %.3e |sum| %.3e ratio %.1f' % (
        l, ds_conv[l].abs().max(), ds_demod[l].abs().max(), tot.abs().max(), ds_conv[l].abs().max() / tot.abs().max()))

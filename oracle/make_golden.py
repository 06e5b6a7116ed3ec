"""Generate tests/golden/*.npz from the UNMODIFIED reference modules in /root/reference.

Run in the build container only (the reference tree does not travel to the GPU box):

    python oracle/make_golden.py

The reference is imported as-is with two shims, both outside the arithmetic under test:
  * ``torch.utils.cpp_extension.load`` is stubbed, because the reference JIT-compiles its two CUDA
    extensions at import (op/fused_act.py:10-17, op/upfirdn2d.py:11-17) and no GPU exists here;
    the CPU branch of upfirdn2d (op/upfirdn2d.py:159-160) never touches the extension.
  * ``fused_leaky_relu`` has no CPU branch (op/fused_act.py:53-55), so it is replaced by the
    one-line torch restatement of op/fused_bias_act_kernel.cu:24-47.  Goldens that depend on it are
    therefore "reference graph + restated activation"; everything else is pure reference code.
  * ``np.product`` (removed in NumPy 2) is aliased for libs/models/direction_matrix.py:11-12.

Weights come from oracle.stylegan2_oracle.seeded_state_dict (one numpy PCG64 stream), so the
fixtures only need to hold inputs that are not re-derivable plus the reference OUTPUTS.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F
import torch.utils.cpp_extension as _ce

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import stylegan2_oracle as orc  # noqa: E402

np.product = np.prod
_ce.load = lambda *a, **k: types.SimpleNamespace()
sys.dont_write_bytecode = True
sys.path.insert(0, '/root/reference')
import libs.gan.StyleGAN2.op.fused_act as _fa  # noqa: E402
import libs.gan.StyleGAN2.model as refm  # noqa: E402
from libs.gan.StyleGAN2.op.upfirdn2d import upfirdn2d as ref_upfirdn2d  # noqa: E402
from libs.models.direction_matrix import DirectionMatrix as RefDirectionMatrix  # noqa: E402


def _flr(x, b, negative_slope=0.2, scale=2 ** 0.5):
    return F.leaky_relu(x + b.view(1, -1, *([1] * (x.ndim - 2))), negative_slope) * scale


_fa.fused_leaky_relu = _flr
refm.fused_leaky_relu = _flr
_fa.FusedLeakyReLU.forward = lambda self, x: _flr(x, self.bias, self.negative_slope, self.scale)

OUT = os.path.join(ROOT, 'tests', 'golden')
os.makedirs(OUT, exist_ok=True)


def rnd(rng, *shape):
    return torch.from_numpy(rng.standard_normal(shape, dtype=np.float32))


def save(name, **arrs):
    np.savez_compressed(os.path.join(OUT, name), **{k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v))
                                                    for k, v in arrs.items()})
    print('wrote', name, {k: tuple(np.shape(v)) for k, v in arrs.items()})


def golden_upfirdn2d():
    rng = np.random.Generator(np.random.PCG64(100))
    fir = refm.make_kernel([1, 3, 3, 1])
    cases = {}
    # (tag, shape, kernel, up, down, pad): the modes the generator hits + odd shapes
    spec = [('blur_up', (2, 5, 9, 9), fir * 4, 1, 1, (1, 1)),
            ('blur_up_big', (1, 3, 33, 33), fir * 4, 1, 1, (1, 1)),
            ('skip_up', (2, 3, 4, 4), fir * 4, 2, 1, (2, 1)),
            ('skip_up_odd', (1, 3, 7, 5), fir * 4, 2, 1, (2, 1)),
            ('down2', (2, 3, 8, 8), fir * 4, 1, 2, (1, 1)),
            ('down2_std', (1, 4, 16, 16), fir, 1, 2, (1, 1)),
            ('grad_pad22', (1, 2, 8, 8), torch.flip(fir * 4, [0, 1]), 1, 1, (2, 2)),
            ('asym3', (1, 2, 6, 6), refm.make_kernel([1, 2, 1]), 1, 1, (1, 1)),
            ('crop', (1, 2, 10, 10), fir, 1, 1, (-1, 0))]
    for tag, shape, k, up, down, pad in spec:
        x = rnd(rng, *shape)
        y = ref_upfirdn2d(x, k, up=up, down=down, pad=pad)
        cases[tag + '_x'] = x
        cases[tag + '_k'] = k
        cases[tag + '_cfg'] = np.array([up, down, pad[0], pad[1]])
        cases[tag + '_y'] = y
    save('upfirdn2d.npz', **cases)


def golden_bias_act():
    rng = np.random.Generator(np.random.PCG64(101))
    x = rnd(rng, 3, 6, 5, 7).requires_grad_(True)
    b = rnd(rng, 6).requires_grad_(True)
    y = _flr(x, b)
    g = rnd(rng, 3, 6, 5, 7)
    y.backward(g)
    x2 = rnd(rng, 4, 16)
    b2 = rnd(rng, 16)
    save('bias_act.npz', x=x, b=b, y=y, g=g, gx=x.grad, gb=b.grad, x2=x2, b2=b2, y2=_flr(x2, b2))


def golden_modconv():
    rng = np.random.Generator(np.random.PCG64(102))
    out = {}
    for tag, cin, cout, k, demod, up, h in [('plain', 16, 32, 3, True, False, 8), ('up', 16, 32, 3, True, True, 5),
                                             ('rgb', 16, 3, 1, False, False, 8), ('plain64', 64, 64, 3, True, False, 16),
                                             ('up64', 128, 64, 3, True, True, 8)]:
        m = refm.ModulatedConv2d(cin, cout, k, 512, demodulate=demod, upsample=up)
        with torch.no_grad():
            m.weight.copy_(rnd(rng, 1, cout, cin, k, k))
            m.modulation.weight.copy_(rnd(rng, cin, 512))
            m.modulation.bias.copy_(1 + 0.1 * rnd(rng, cin))
        x = rnd(rng, 2, cin, h, h)
        w = rnd(rng, 2, 512)
        with torch.no_grad():
            y = m(x, w)
        out.update({tag + '_weight': m.weight, tag + '_mw': m.modulation.weight, tag + '_mb': m.modulation.bias,
                    tag + '_x': x, tag + '_w': w, tag + '_y': y})
    save('modconv.npz', **out)


def golden_styled_block():
    """StyledConv(up) -> StyledConv -> ToRGB(+skip) with small channels, incl. autograd grads to x, w, skip."""
    rng = np.random.Generator(np.random.PCG64(103))
    cin, cmid = 32, 64
    c0 = refm.StyledConv(cin, cmid, 3, 512, upsample=True)
    c1 = refm.StyledConv(cmid, cmid, 3, 512)
    tr = refm.ToRGB(cmid, 512)
    sd = {}
    with torch.no_grad():
        for pre, mod in [('c0', c0), ('c1', c1), ('rgb', tr)]:
            for n, p in mod.named_parameters():
                if n.endswith('modulation.bias'):
                    p.copy_(1 + 0.1 * rnd(rng, *p.shape))
                elif n in ('noise.weight', 'activate.bias', 'bias'):
                    p.copy_(0.1 * rnd(rng, *p.shape))
                else:
                    p.copy_(rnd(rng, *p.shape))
                sd[pre + '.' + n] = p.detach().clone()
    x = rnd(rng, 2, cin, 8, 8).requires_grad_(True)
    ws = [rnd(rng, 2, 512).requires_grad_(True) for _ in range(3)]
    n0, n1 = rnd(rng, 1, 1, 16, 16), rnd(rng, 1, 1, 16, 16)
    skip = rnd(rng, 2, 3, 8, 8).requires_grad_(True)
    y0 = c0(x, ws[0], noise=n0)
    y1 = c1(y0, ws[1], noise=n1)
    rgb = tr(y1, ws[2], skip)
    gr = rnd(rng, *rgb.shape)
    (rgb * gr).sum().backward()
    save('styled_block.npz', x=x, w0=ws[0], w1=ws[1], w2=ws[2], n0=n0, n1=n1, skip=skip, y0=y0, y1=y1, rgb=rgb, gr=gr,
         gx=x.grad, gw0=ws[0].grad, gw1=ws[1].grad, gw2=ws[2].grad, gskip=skip.grad,
         **{'p.' + k: v for k, v in sd.items()})


def _ref_generator(size, cm, seed):
    sd = orc.seeded_state_dict(size, cm, seed=seed)
    g = refm.Generator(size, 512, 8, channel_multiplier=cm)
    g.load_state_dict(sd, strict=True)
    return g.eval(), sd


def golden_generator():
    # size 8 == BASELINE config 1 (conv1 + one 4->8 SynthesisBlock), size 32 small net, 256/cm1 = config 2 at B=1
    for size, cm, seed, batch, keep_feats in [(8, 2, 0, 1, 1), (32, 2, 1, 2, 32), (256, 1, 2, 1, 0)]:
        g, sd = _ref_generator(size, cm, seed)
        wplus = orc.seeded_wplus(sd, batch, g.n_latent, seed=1234 + size)
        feats = []
        hooks = [m.register_forward_hook(lambda mod, i, o: feats.append(o.detach())) for m in [g.conv1] + list(g.convs)]
        with torch.no_grad():
            img, lat = g([wplus], input_is_latent=True, return_latents=True)
            rng = np.random.Generator(np.random.PCG64(7))
            trunc = g.style(rnd(rng, 64, 512)).mean(0, keepdim=True)
            n_full = len(feats)
            img_t, _ = g([wplus], input_is_latent=True, truncation=0.7, truncation_latent=trunc)
            zin = rnd(rng, batch, 512)
            img_z, lat_z = g([zin], return_latents=True, truncation=0.7, truncation_latent=trunc)
        for h in hooks:
            h.remove()
        arrs = dict(wplus=wplus, img=img, trunc=trunc, img_trunc=img_t, zin=zin, img_z=img_z, lat_z=lat_z,
                    cfg=np.array([size, cm, seed, batch]))
        feats = feats[:n_full]
        arrs['feat_absmean'] = np.array([f.abs().mean().item() for f in feats])
        if keep_feats:     # channel stride of the stored per-layer activations (0 = none)
            for i, f in enumerate(feats):
                arrs['feat%d' % i] = f[:, ::keep_feats].contiguous()
            arrs['feat_stride'] = np.array(keep_feats)
        save('generator_%d_cm%d.npz' % (size, cm), **arrs)


def golden_reenact():
    """DirectionMatrix + generate_image glue + dL/dA through the generator (size 32)."""
    from libs.utilities.generic import generate_image as ref_generate_image
    size, cm, seed, batch = 32, 2, 3, 3
    g, sd = _ref_generator(size, cm, seed)
    rng = np.random.Generator(np.random.PCG64(55))
    A = RefDirectionMatrix(512, input_dim=15, out_dim=512, w_plus=True, num_layers=4)
    with torch.no_grad():
        A.linear.weight.copy_(0.03 * rnd(rng, *A.linear.weight.shape))
        A.linear.bias.copy_(0.01 * rnd(rng, *A.linear.bias.shape))
    dp = torch.from_numpy(rng.uniform(-3, 3, (batch, 15)).astype(np.float32))
    wsrc = orc.seeded_wplus(sd, 1, g.n_latent, seed=11).repeat(batch, 1, 1)
    with torch.no_grad():
        trunc = g.style(rnd(rng, 64, 512)).mean(0, keepdim=True)
    shift = A(dp)
    img, lat = ref_generate_image(g, wsrc, 0.7, trunc, w_plus=True, num_layers_shift=4, shift_code=shift,
                                  input_is_latent=True, return_latents=True)
    r = rnd(rng, *img.shape)
    loss = (img * r).sum() / img.numel()
    g.zero_grad()
    loss.backward()
    save('reenact_32.npz', cfg=np.array([size, cm, seed, batch]), A_w=A.linear.weight, A_b=A.linear.bias, dp=dp,
         wsrc=wsrc, trunc=trunc, shift=shift, img=img, lat=lat, r=r, loss=loss, gA_w=A.linear.weight.grad,
         gA_b=A.linear.bias.grad)


def golden_output_stage():
    """tensor_to_image (libs/utilities/image_utils.py:97-111) + np.uint8 (utils_inference.py:16) on seeded frames, and the
    256-pooling of generate_image (generic.py:146-148) in front of it."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('ref_image_utils', os.path.join('/root/reference', 'libs', 'utilities', 'image_utils.py'))
    iu = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(iu)
    rng = np.random.Generator(np.random.PCG64(41))
    x = rnd(rng, 2, 3, 32, 32) * 0.8
    x[0, :, 0, :4] = torch.tensor([-1.0, 1.0, -3.0, 5.0])
    y = np.stack([np.uint8(iu.tensor_to_image(x[i:i + 1].clone())) for i in range(2)])
    pooled = torch.nn.AdaptiveAvgPool2d((8, 8))(x)
    yp = np.stack([np.uint8(iu.tensor_to_image(pooled[i:i + 1].clone())) for i in range(2)])
    save('output_stage.npz', x=x, y=y, y_pooled=yp)


def golden_formats():
    """Files written the way the reference writes them: an A-matrix checkpoint with the dict of
    libs/utilities/utils_train.py:592-603 around the reference DirectionMatrix's state_dict, and one inverted-code .npy as
    invert_images.py:118-125 saves it (latent_codes[i].detach().cpu().numpy() -> np.save)."""
    torch.manual_seed(21)
    A = RefDirectionMatrix(shift_dim=512, input_dim=15, out_dim=None, w_plus=True, bias=True, num_layers=8)
    with torch.no_grad():
        A.linear.bias.normal_(0, 0.01)
    state_dict = {'step': 10, 'A_matrix': A.state_dict(), 'learned_directions': 15, 'shift_scale': 6, 'w_plus': True,
                  'num_layers_shift': 8}
    torch.save(state_dict, os.path.join(ROOT, 'tests', 'golden', 'A_matrix_000010.pt'))
    dp = rnd(np.random.Generator(np.random.PCG64(22)), 3, 15)
    shift = A(dp).detach().numpy()
    G, _ = _ref_generator(8, 2, 3)
    latent_codes = G.style(rnd(np.random.Generator(np.random.PCG64(23)), 1, 512)).unsqueeze(1).repeat(1, G.n_latent, 1)
    latent_code = latent_codes[0].detach().cpu().numpy()
    np.save(os.path.join(ROOT, 'tests', 'golden', 'latent_000.npy'), latent_code)
    save('formats.npz', dp=dp.numpy(), shift=shift, latent=latent_code)


if __name__ == '__main__':
    torch.manual_seed(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'output_stage':
        golden_output_stage()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'formats':
        golden_formats()
        sys.exit(0)
    golden_upfirdn2d()
    golden_bias_act()
    golden_modconv()
    golden_styled_block()
    golden_generator()
    golden_reenact()
    golden_output_stage()
    golden_formats()

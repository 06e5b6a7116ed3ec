"""Pin the CPU oracle (oracle/stylegan2_oracle.py) against outputs of the reference itself.

The fixtures under tests/golden/ were produced by oracle/make_golden.py, which imports the
unmodified reference modules from /root/reference in the build container.  CPU only.
"""
import os

import numpy as np
import torch

from oracle import stylegan2_oracle as orc

T = torch.from_numpy


def close(a, b, tol):
    a = a.detach().numpy() if torch.is_tensor(a) else a
    err = np.abs(a - b).max()
    assert a.shape == b.shape and err <= tol, (a.shape, b.shape, err)


def test_upfirdn2d_matches_reference(golden):
    g = golden('upfirdn2d.npz')
    tags = sorted(k[:-4] for k in g if k.endswith('_cfg'))
    assert len(tags) == 9
    for t in tags:
        up, down, p0, p1 = [int(v) for v in g[t + '_cfg']]
        y = orc.upfirdn2d(T(g[t + '_x']), T(g[t + '_k']), up, down, (p0, p1))
        close(y, g[t + '_y'], 2e-6)


def test_bias_act_matches_reference(golden):
    g = golden('bias_act.npz')
    y = orc.fused_leaky_relu(T(g['x']), T(g['b']))
    close(y, g['y'], 1e-6)
    gx, gb = orc.fused_leaky_relu_backward(T(g['g']), y)
    close(gx, g['gx'], 1e-6)
    close(gb, g['gb'], 2e-5)
    close(orc.fused_leaky_relu(T(g['x2']), T(g['b2'])), g['y2'], 1e-6)


def test_modconv_matches_reference(golden):
    g = golden('modconv.npz')
    for tag, demod, up in [('plain', True, False), ('up', True, True), ('rgb', False, False),
                           ('plain64', True, False), ('up64', True, True)]:
        y = orc.modulated_conv2d(T(g[tag + '_x']), T(g[tag + '_w']), T(g[tag + '_weight']), T(g[tag + '_mw']),
                                 T(g[tag + '_mb']), demodulate=demod, upsample=up)
        close(y, g[tag + '_y'], 2e-5)


def test_styled_block_matches_reference_with_grads(golden):
    g = golden('styled_block.npz')
    sd = {k[2:]: T(v) for k, v in g.items() if k.startswith('p.')}
    x = T(g['x']).requires_grad_(True)
    ws = [T(g['w%d' % i]).requires_grad_(True) for i in range(3)]
    skip = T(g['skip']).requires_grad_(True)
    y0 = orc.styled_conv(sd, 'c0', x, ws[0], T(g['n0']), upsample=True)
    y1 = orc.styled_conv(sd, 'c1', y0, ws[1], T(g['n1']), upsample=False)
    rgb = orc.to_rgb(sd, 'rgb', y1, ws[2], skip)
    close(y0, g['y0'], 2e-5)
    close(y1, g['y1'], 2e-5)
    close(rgb, g['rgb'], 5e-5)
    (rgb * T(g['gr'])).sum().backward()
    close(x.grad, g['gx'], 1e-4)
    close(skip.grad, g['gskip'], 1e-5)
    for i in range(3):
        close(ws[i].grad, g['gw%d' % i], 2e-3 * np.abs(g['gw%d' % i]).max())


def _gen(golden, name):
    g = golden(name)
    size, cm, seed, batch = [int(v) for v in g['cfg']]
    return g, size, cm, orc.seeded_state_dict(size, cm, seed=seed)


def test_generator_small_matches_reference(golden):
    for name in ['generator_8_cm2.npz', 'generator_32_cm2.npz']:
        g, size, cm, sd = _gen(golden, name)
        with torch.no_grad():
            img, _, feats = orc.generator_forward(sd, [T(g['wplus'])], size, cm, input_is_latent=True,
                                                  return_features=True)
            close(img, g['img'], 2e-4)
            st = int(g['feat_stride'])
            for i, f in enumerate(feats):
                close(f[:, ::st], g['feat%d' % i], 2e-4)
            img_t, _ = orc.generator_forward(sd, [T(g['wplus'])], size, cm, input_is_latent=True, truncation=0.7,
                                             truncation_latent=T(g['trunc']))
            close(img_t, g['img_trunc'], 2e-4)
            img_z, lat_z = orc.generator_forward(sd, [T(g['zin'])], size, cm, truncation=0.7,
                                                 truncation_latent=T(g['trunc']), return_latents=True)
            close(lat_z, g['lat_z'], 1e-5)
            close(img_z, g['img_z'], 2e-4)


def test_generator_256_matches_reference(golden):
    g, size, cm, sd = _gen(golden, 'generator_256_cm1.npz')
    with torch.no_grad():
        img, _, feats = orc.generator_forward(sd, [T(g['wplus'])], size, cm, input_is_latent=True,
                                              return_features=True)
    close(img, g['img'], 5e-4)
    np.testing.assert_allclose([f.abs().mean().item() for f in feats], g['feat_absmean'], rtol=1e-4)


def test_reenact_glue_and_dA_matches_reference(golden):
    g, size, cm, sd = _gen(golden, 'reenact_32.npz')
    aw = T(g['A_w']).requires_grad_(True)
    ab = T(g['A_b']).requires_grad_(True)
    shift = orc.direction_matrix_forward(aw, ab, T(g['dp']), 512, 4)
    close(shift, g['shift'], 1e-6)
    img, lat = orc.generate_image(sd, T(g['wsrc']), 0.7, T(g['trunc']), size, cm, shift_code=shift)
    close(img, g['img'], 2e-4)
    close(lat, g['lat'], 1e-5)
    loss = (img * T(g['r'])).sum() / img.numel()
    loss.backward()
    close(aw.grad, g['gA_w'], 1e-3 * np.abs(g['gA_w']).max())
    close(ab.grad, g['gA_b'], 1e-3 * np.abs(g['gA_b']).max())


def test_flop_model_matches_survey():
    assert abs(orc.forward_flops_per_frame(256, 1) / 1e9 - 29.794) < 0.01
    assert abs(orc.forward_flops_per_frame(256, 2) / 1e9 - 90.236) < 0.01
    assert abs(orc.forward_flops_per_frame(1024, 2) / 1e9 - 148.520) < 0.01


def test_scatter_upconv_algebra_matches_reference_form():
    """The parity-plane scatter + FIR bookkeeping the sm_100a kernels implement (csrc/modconv_sm100.cu mode 2,
    csrc/up_finish_sm100.cu) equals conv_transpose2d(stride 2) + upfirdn2d(pad=(1,1)) (model.py:246-257) in fp64."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        'up_scatter_emulation_check', os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tools',
                                                   'up_scatter_emulation_check.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.main() < 1e-12


def test_output_stage_golden(golden):
    """oracle.frames_to_uint8 == the reference's tensor_to_image + np.uint8 (+ AdaptiveAvgPool2d), bit-exact on CPU."""
    g = golden('output_stage.npz')
    x = torch.from_numpy(g['x'])
    assert np.array_equal(orc.frames_to_uint8(x), g['y'])
    assert np.array_equal(orc.frames_to_uint8(x, size=8), g['y_pooled'])

"""Drop-in level: the reference's OWN glue code, unmodified, driving this repository's Generator /
DirectionMatrix through the PEP-420 namespace overlay — `libs.utilities.generic.generate_image` /
`get_shifted_latent_code` (generic.py:116-152: clone + in-place `latent[:, :k] += shift`, then `G([...])`) and
`libs.optimization.optimize_g` (optimization.py:25-72: deepcopy, train(), Adam on convs[4..11], loss.backward()).

The reference tree is not in the repository; tools/make_baseline_ref.py copies it to the git-ignored baseline/_ref, which
travels to the GPU box with the snapshot.  The tests skip (with the reason) when that copy is absent.
The only stand-in is LPIPS: its AlexNet / lin weights are downloaded by the reference (lpips/utils.py:16-24), which is
impossible offline, so `libs.optimization.LPIPS` is replaced by an L2 module (outside the generator path).
"""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import stylegan2_oracle as orc

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, 'baseline', '_ref')
OVERLAY = os.path.join(ROOT, 'overlay')


@pytest.fixture(scope='module')
def ref_env():
    if not os.path.isdir(os.path.join(REF, 'libs', 'utilities')):
        pytest.skip('baseline/_ref absent (run tools/make_baseline_ref.py where /root/reference exists)')
    assert torch.cuda.is_available()
    if not hasattr(np, 'product'):
        np.product = np.prod
    saved = list(sys.path)
    dropped = {k: sys.modules.pop(k) for k in list(sys.modules) if k == 'libs' or k.startswith('libs.')}
    os.environ.setdefault('TORCH_EXTENSIONS_DIR', os.path.join(REF, '_ext'))     # the reference's JIT ops, prebuilt by make_baseline_ref
    os.environ.setdefault('TORCH_CUDA_ARCH_LIST', '10.0')
    sys.path[:0] = [OVERLAY, REF]                 # overlay first: libs.gan.StyleGAN2.model / libs.models.direction_matrix are OURS
    sys.dont_write_bytecode = True
    import stylegan_directions_face_reenactment_b200 as pkg
    import libs.gan.StyleGAN2.model as m
    import libs.models.direction_matrix as dm
    import libs.utilities.generic as generic        # the reference's own file
    assert m.Generator is pkg.Generator and dm.DirectionMatrix is pkg.DirectionMatrix
    assert os.path.realpath(generic.__file__).startswith(os.path.realpath(REF))
    yield pkg, m, dm, generic
    sys.path[:] = saved
    for k in [k for k in sys.modules if k == 'libs' or k.startswith('libs.')]:
        del sys.modules[k]
    sys.modules.update(dropped)


def _generator(pkg, size, cm, seed=0):
    sd = orc.seeded_state_dict(size, cm, seed=seed)
    G = pkg.Generator(size, 512, 8, channel_multiplier=cm)
    G.load_state_dict(sd, strict=True)
    return sd, G.cuda().eval()


def test_reference_generate_image_drives_overlay_generator(ref_env):
    pkg, m, dm, generic = ref_env
    size, cm, batch = 64, 2, 3
    sd, G = _generator(pkg, size, cm, seed=3)
    torch.manual_seed(5)
    A = dm.DirectionMatrix(512, input_dim=15, out_dim=512, w_plus=True, num_layers=8).cuda()
    trunc = orc.seeded_wplus(sd, 1, 1, seed=7)[:, 0].cuda()
    w = orc.seeded_wplus(sd, batch, G.n_latent, seed=11).cuda()
    w_before = w.clone()
    dp = (torch.rand(batch, 15, generator=torch.Generator().manual_seed(1)) * 6 - 3).cuda()
    with torch.no_grad():
        shift = A(dp)
        a = generic.generate_image(G, w, 0.7, trunc, True, 8, shift_code=shift, input_is_latent=True)      # reference glue
        b = pkg.generate_image(G, w, 0.7, trunc, w_plus=True, num_layers_shift=8, shift_code=shift, input_is_latent=True)
        code = generic.get_shifted_latent_code(G, w, shift, input_is_latent=True, w_plus=True, num_layers=8)
        ref, _ = orc.generate_image(sd, w.cpu(), 0.7, trunc.cpu(), size, cm, shift_code=shift.cpu())
    assert torch.equal(w, w_before)                         # the reference clones before its in-place add (generic.py:122)
    assert torch.equal(code, pkg.get_shifted_latent_code(G, w, shift, input_is_latent=True, w_plus=True, num_layers=8))
    assert torch.equal(a, b)                                # same kernels, same inputs: bit-identical frames
    assert float((a.cpu() - ref).abs().max()) <= 1e-3
    # Z-space entry (what libs/trainer.py:159 does) with return_latents
    z = torch.randn(batch, 512, generator=torch.Generator().manual_seed(2)).cuda()
    with torch.no_grad():
        img, lat = generic.generate_image(G, z, 0.7, trunc, input_is_latent=False, return_latents=True)
        img2, lat2 = pkg.generate_image(G, z, 0.7, trunc, input_is_latent=False, return_latents=True)
    assert torch.equal(img, img2) and torch.equal(lat, lat2) and lat.shape == (batch, G.n_latent, 512)
    # autograd through the reference glue reaches A (libs/trainer.py:175-189)
    img = generic.generate_image(G, w, 0.7, trunc, True, 8, shift_code=A(dp), input_is_latent=True)
    img.square().mean().backward()
    assert A.linear.weight.grad is not None and torch.isfinite(A.linear.weight.grad).all() and A.linear.weight.grad.abs().max() > 0


def test_reference_optimize_g_runs_on_overlay_generator(ref_env, monkeypatch):
    """optimization.py:25-72 itself (3 steps, Adam lr of the reference scaled to random-init weights): 256^2 / cm1, B=1,
    parameters of convs[4..11] — split-K and the < 128-channel paths of the weight-gradient kernels are live here."""
    pkg, m, dm, generic = ref_env
    import libs.optimization as opt_mod

    class L2AsLPIPS(torch.nn.Module):
        def __init__(self, net_type='alex'):
            super().__init__()

        def forward(self, x, y):
            return (x - y).pow(2).mean()
    monkeypatch.setattr(opt_mod, 'LPIPS', L2AsLPIPS)
    monkeypatch.setattr(opt_mod, 'tqdm', lambda it: it)
    sd, G = _generator(pkg, 256, 1, seed=0)
    latent = orc.seeded_wplus(sd, 1, G.n_latent, seed=2).cuda()
    torch.manual_seed(0)
    with torch.no_grad():
        trunc = G.mean_latent(4096)
        start = G([latent], input_is_latent=True, truncation=0.7, truncation_latent=trunc)[0]
        target = (start * 0.9 + 0.05).detach()
    before = {n: p.detach().clone() for n, p in G.named_parameters()}
    torch.manual_seed(0)                                      # optimize_g draws its own mean_latent(4096)
    G2 = opt_mod.optimize_g(G, latent, target, opt_steps=3, lr=1e-4)
    assert G2 is G and G.training                             # the reference leaves the generator in train() mode
    changed = [n for n, p in G.named_parameters() if not torch.equal(p.detach(), before[n])]
    assert changed and all(n.startswith('convs.') and 4 <= int(n.split('.')[1]) <= 11 for n in changed), changed[:5]
    assert len(changed) == 8 * 5                              # weight, modulation.{weight,bias}, noise.weight, activate.bias
    assert all(torch.isfinite(p).all() for p in G.parameters())
    with torch.no_grad():
        torch.manual_seed(0)
        trunc2 = G.mean_latent(4096)
        after = G([latent], input_is_latent=True, truncation=0.7, truncation_latent=trunc2)[0]
    l0 = float((start - target).pow(2).mean())
    l1 = float((after - target).pow(2).mean())
    print('\n[reference optimize_g on the overlay generator] L2 to target %.5f -> %.5f after 3 Adam steps' % (l0, l1))
    assert l1 < l0


def test_native_ops_against_the_reference_cuda_ops(ref_env):
    """The reference's own CUDA operators (op/fused_bias_act_kernel.cu:18-49, op/upfirdn2d_kernel.cu:52-272, JIT-built from
    its sources) as the GPU-side oracle for sgr_fused_bias_act / sgr_upfirdn2d: forward and backward on the same tensors."""
    pkg, m, dm, generic = ref_env
    from libs.gan.StyleGAN2.op import fused_leaky_relu as ref_flr, upfirdn2d as ref_upfirdn2d      # reference (not in the overlay)
    import libs.gan.StyleGAN2.op.fused_act as ref_fa
    assert os.path.realpath(ref_fa.__file__).startswith(os.path.realpath(REF))
    g = torch.Generator(device='cuda').manual_seed(3)
    for shape in [(4, 512), (3, 64, 33, 17), (2, 8, 128, 128)]:
        x = torch.randn(*shape, device='cuda', generator=g)
        b = torch.randn(shape[1], device='cuda', generator=g)
        gy = torch.randn(*shape, device='cuda', generator=g)
        xr, br = x.clone().requires_grad_(True), b.clone().requires_grad_(True)
        yr = ref_flr(xr, br)
        yr.backward(gy)
        xo, bo = x.clone().requires_grad_(True), b.clone().requires_grad_(True)
        yo = pkg.fused_leaky_relu(xo, bo)
        yo.backward(gy)
        assert torch.allclose(yo, yr, rtol=2.4e-7, atol=0), shape           # the same three fp32 operations: within 1 ulp
        assert torch.allclose(xo.grad, xr.grad, rtol=2.4e-7, atol=0), shape
        assert float((bo.grad - br.grad).abs().max()) <= 2e-5 * float(br.grad.abs().max()), shape      # reduction order
    fir = orc.make_fir_kernel([1, 3, 3, 1]).cuda() * 4
    errs = []
    for shape, up, down, pad in [((3, 7, 65, 65), 1, 1, (1, 1)), ((2, 3, 128, 128), 2, 1, (2, 1)), ((2, 3, 256, 256), 1, 2, (1, 1)),
                                 ((1, 2, 257, 257), 1, 1, (1, 1)), ((1, 3, 33, 65), 1, 1, (2, 2))]:
        # (the reference's CUDA op has kernels for these factor combinations only: any other, e.g. up = down = 2, falls through
        #  its `switch (mode)` and returns uninitialised memory, op/upfirdn2d_kernel.cu:223-270)
        x = torch.randn(*shape, device='cuda', generator=g)
        xr = x.clone().requires_grad_(True)
        yr = ref_upfirdn2d(xr, fir, up=up, down=down, pad=pad)
        gy = torch.randn_like(yr)
        yr.backward(gy)
        xo = x.clone().requires_grad_(True)
        yo = pkg.upfirdn2d(xo, fir, up=up, down=down, pad=pad)
        yo.backward(gy)
        assert yo.shape == yr.shape
        ef = float((yo.detach() - yr.detach()).abs().max()) / float(yr.detach().abs().max())
        eb = float((xo.grad - xr.grad).abs().max()) / float(xr.grad.abs().max())
        print('[upfirdn2d vs reference CUDA op] %s up %d down %d pad %s: fwd %.2e bwd %.2e (relative to max)' % (shape, up, down, pad, ef, eb))
        errs.append((shape, up, down, ef, eb))
    assert all(ef <= 2e-6 and eb <= 2e-6 for _, _, _, ef, eb in errs), errs       # fp32 sums of 16 taps in a different order

"""CPU-only checks: the C-ABI library loads and exports what include/sgr.h declares, the ctypes mirrors match the C
struct layouts, compute entry points fail loudly without a GPU, and the host-side mirror of the reference API behaves
like the oracle."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import stylegan2_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'sgr.h')


@pytest.fixture(scope='module')
def native():
    import __graft_entry__ as ge
    from stylegan_directions_face_reenactment_b200 import _native
    if not os.path.exists(_native.LIB_PATH):
        ge.build()
    _native.lib()
    return _native


def test_library_exports_every_declared_symbol(native):
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    declared = set(re.findall(r'\b(sgr_[a-z0-9_]+)\s*\(', text))
    assert len(declared) >= 15
    lib = native.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(native.SIGNATURES), declared ^ set(native.SIGNATURES)
    assert b'sm_100a' in lib.sgr_version()


def test_ctypes_structs_match_c_layout(native, tmp_path):
    src = tmp_path / 'layout.c'
    fields = {'sgr_conv_args': native.ConvArgs, 'sgr_styled_layer': native.StyledLayer, 'sgr_rgb_layer': native.RgbLayer,
              'sgr_synthesis': native.Synthesis, 'sgr_backward_extras': native.BackwardExtras,
              'sgr_wgrad_args': native.WgradArgs, 'sgr_styled_param_grads': native.StyledParamGrads,
              'sgr_rgb_param_grads': native.RgbParamGrads, 'sgr_param_grads': native.ParamGrads}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "sgr.h"', 'int main(void){']
    for cname, ct in fields.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for f in ct._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, f[0], cname, f[0]))
    lines.append('return 0;}')
    src.write_text('\n'.join(lines))
    exe = tmp_path / 'layout'
    subprocess.run(['gcc', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, ct in fields.items():
        assert int(out[cname]) == C.sizeof(ct), cname
        for f in ct._fields_:
            assert int(out['%s.%s' % (cname, f[0])]) == getattr(ct, f[0]).offset, (cname, f[0])
    assert native.MAX_STYLED == 24 and native.MAX_RGB == 12


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU error path')
def test_compute_entry_points_fail_loudly_without_gpu(native):
    lib = native.lib()
    assert lib.sgr_upfirdn2d(None, None, None, 1, 4, 4, 1, 1, 0, 0, 4, 4, None) != 0
    assert b'no CUDA device' in lib.sgr_last_error()
    assert lib.sgr_fused_bias_act(None, None, None, None, 1, 1, 1, 0, 0.2, 1.0, None) != 0
    a = native.ConvArgs()
    assert lib.sgr_modconv_forward(C.byref(a), None) != 0
    s = native.Synthesis()
    assert lib.sgr_synthesis_forward(C.byref(s), None, 1, None, None, 0, None, None) != 0
    assert lib.sgr_synthesis_backward(C.byref(s), None, 1, None, None, None, None, 0, None) != 0
    w = native.WgradArgs()
    assert lib.sgr_modconv_wgrad(C.byref(w), None) != 0
    assert b'no CUDA device' in lib.sgr_last_error()
    assert lib.sgr_synthesis_backward_ex(C.byref(s), None, 1, None, None, None, None, 0, None, None) != 0
    assert lib.sgr_wgrad_scratch_bytes(64, 64) >= 9 * 64 * 64 * 4          # at least one slice per tap
    assert lib.sgr_packed_weight_bytes(512, 512, 3, 0, 0) == 512 * 512 * 9 * 4
    assert lib.sgr_packed_weight_bytes(64, 128, 3, 1, 0) == 4 * 64 * 128 * 9 * 4


def test_python_ops_have_no_cpu_fallback():
    import stylegan_directions_face_reenactment_b200 as pkg
    with pytest.raises(RuntimeError, match='CUDA'):
        pkg.fused_leaky_relu(torch.zeros(2, 4), torch.zeros(4))
    with pytest.raises(RuntimeError, match='CUDA'):
        pkg.upfirdn2d(torch.zeros(1, 1, 4, 4), torch.ones(4, 4))
    g = pkg.Generator(8, 512, 8)
    with pytest.raises(RuntimeError, match='CUDA'):
        g([torch.zeros(1, g.n_latent, 512)], input_is_latent=True)


@pytest.mark.parametrize('size,cm', [(8, 2), (256, 1), (256, 2), (1024, 2)])
def test_state_dict_contract(size, cm):
    """Keys, order and shapes equal the reference's (manifest dumped from the reference)."""
    import stylegan_directions_face_reenactment_b200 as pkg
    g = pkg.Generator(size, 512, 8, channel_multiplier=cm)
    got = [(k, tuple(v.shape)) for k, v in g.state_dict().items()]
    assert got == orc.state_dict_manifest(size, cm)
    assert g.n_latent == orc.synthesis_config(size, cm)[3] and g.num_layers == orc.synthesis_config(size, cm)[2]
    sd = orc.seeded_state_dict(size, cm, seed=0) if size <= 256 else None
    if sd is not None:
        g.load_state_dict(sd, strict=True)
        missing = g.load_state_dict({k: v for k, v in sd.items() if not k.startswith('noises.')}, strict=False)
        assert all(k.startswith('noises.') for k in missing.missing_keys)      # 256^2 checkpoints load with strict=False
    import copy
    g2 = copy.deepcopy(g)                                                       # optimize_g deep-copies the generator
    assert [k for k, _ in g2.named_parameters()] == [k for k, _ in g.named_parameters()]
    assert len(list(g.convs[4].parameters())) == 5 if size >= 32 else True


def test_direction_matrix_and_shift_glue_match_oracle():
    import stylegan_directions_face_reenactment_b200 as pkg
    torch.manual_seed(0)
    A = pkg.DirectionMatrix(512, input_dim=15, out_dim=512, w_plus=True, num_layers=8)
    assert A.input_dim == 15 and tuple(A.linear.weight.shape) == (4096, 15) and tuple(A.linear.bias.shape) == (4096,)
    assert sorted(A.state_dict()) == ['linear.bias', 'linear.weight']
    assert abs(A.linear.weight.std().item() - 0.03) < 2e-3
    dp = torch.rand(5, 15) * 6 - 3
    ref = orc.direction_matrix_forward(A.linear.weight, A.linear.bias, dp, 512, 8)
    assert torch.equal(A(dp), ref) and tuple(ref.shape) == (5, 8, 512)
    assert pkg.DirectionMatrix((4, 8)).input_dim == 32                          # np.product-free default dims

    class FakeG:
        n_latent = 14
    z = torch.randn(5, 14, 512)
    keep = z.clone()
    out = pkg.get_shifted_latent_code(FakeG(), z, ref, input_is_latent=True, w_plus=True, num_layers=8)
    assert torch.equal(z, keep)
    assert torch.allclose(out, orc.shifted_latent_code(z, ref))
    eye = pkg.DirectionMatrix(4, input_dim=4, out_dim=4, w_plus=True, num_layers=2, initialization='eye')
    assert torch.equal(eye.linear.weight, torch.cat([torch.eye(4), torch.eye(4)]))


def test_mapping_network_semantics_on_cpu_parameters():
    """EqualLinear / PixelNorm parameters follow the reference layout (weights stored / lr_mul)."""
    import stylegan_directions_face_reenactment_b200 as pkg
    lin = pkg.EqualLinear(512, 512, lr_mul=0.01, activation='fused_lrelu')
    assert abs(lin.weight.std().item() - 100.0) < 2.0 and abs(lin.scale - 0.01 / np.sqrt(512)) < 1e-9
    mod = pkg.EqualLinear(512, 64, bias_init=1)
    x = torch.randn(3, 512)
    assert torch.allclose(mod(x), orc.equal_linear(x, mod.weight, mod.bias), atol=1e-6)
    z = torch.randn(4, 512)
    assert torch.allclose(pkg.PixelNorm()(z), orc.pixel_norm(z), atol=1e-6)


@pytest.mark.skipif(not os.path.isdir('/root/reference/libs'), reason='reference checkout not present on this box')
def test_overlay_resolves_reference_imports():
    code = ('import sys; sys.path.insert(0, %r); import run_reference_script as r; r.install("/root/reference");'
            'from libs.gan.StyleGAN2.model import Generator, EqualLinear;'
            'from libs.models.direction_matrix import DirectionMatrix;'
            'import libs.configs.config_models as cm;'
            'print(Generator.__module__, DirectionMatrix.__module__, cm.__file__)') % os.path.join(ROOT, 'tools')
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, check=True,
                         env=dict(os.environ, PYTHONDONTWRITEBYTECODE='1')).stdout.split()
    assert out[0] == 'stylegan_directions_face_reenactment_b200.model'
    assert out[1] == 'stylegan_directions_face_reenactment_b200.direction_matrix'
    assert out[2].startswith('/root/reference/')


def test_sass_is_blackwell_native():
    """The shipped library's GEMM kernels are tcgen05 / TMEM / TMA code (UTC*MMA, LDTM, UTMALDG / UBLKCP in SASS) and contain no
    legacy mma.sync (HMMA): tools/sass_summary.py over libsgr.so (cuobjdump only, no GPU)."""
    import shutil
    import subprocess
    import sys
    if shutil.which('cuobjdump') is None:
        pytest.skip('cuobjdump not on PATH')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, 'tools', 'sass_summary.py')], capture_output=True, text=True, timeout=600).stdout
    rows = {}
    for ln in out.splitlines():
        if ln.startswith('| `'):
            c = [x.strip() for x in ln.strip().strip('|').split('|')]
            rows.setdefault(c[0].strip('`').split('<')[0], []).append(c)
    for k in ('modconv_halo_kernel', 'upconv_scatter_kernel', 'modconv_kernel', 'wgrad_kernel'):
        assert k in rows, k
        for c in rows[k]:
            assert int(c[1]) > 0 and int(c[2]) > 0, (k, 'UTC*MMA / LDTM missing', c)          # tcgen05.mma, tcgen05.ld
            assert int(c[4]) + int(c[6]) > 0, (k, 'no TMA loads', c)                           # UTMALDG or UBLKCP
            assert int(c[10]) == 0, (k, 'legacy HMMA present', c)
    total = [ln for ln in out.splitlines() if ln.startswith('| **total**')]
    assert total and int(total[0].split('|')[11].strip()) == 0                               # no HMMA anywhere

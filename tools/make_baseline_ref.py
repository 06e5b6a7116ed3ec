#!/usr/bin/env python
"""Copies the part of the reference that the reference arm / drop-in tests execute into baseline/_ref (git-ignored, but it
travels to the GPU box with the gpurun snapshot) and pre-builds its two JIT CUDA extensions for sm_100 there.

The reference has no setup.py / pyproject (nothing for pip to install), so "installing" it is this copy.  Nothing under
baseline/_ref is product code; bench.py (`gpu_reference`, `--impl reference`) and tests/test_reference_dropin.py are the
only readers.  Run here (needs /root/reference): `python tools/make_baseline_ref.py [--no-ext]`.
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get('SGR_REFERENCE', '/root/reference')
DST = os.path.join(ROOT, 'baseline', '_ref')
KEEP = ['libs/gan', 'libs/models', 'libs/utilities', 'libs/configs', 'libs/criteria', 'libs/optimization.py',
        'libs/trainer.py']


def copy_tree():
    if not os.path.isdir(REF):
        return False
    for rel in KEEP:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        if os.path.isdir(src):
            shutil.copytree(src, dst, dirs_exist_ok=True, ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
        else:
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copy2(src, dst)
    return True


def prebuild_ext():
    """torch.utils.cpp_extension.load of the reference's own op/*.cu, cross-compiled for sm_100 into baseline/_ref/_ext so
    the GPU box finds them up to date (it falls back to the same JIT build there if ninja disagrees)."""
    env = dict(os.environ, TORCH_CUDA_ARCH_LIST='10.0', TORCH_EXTENSIONS_DIR=os.path.join(DST, '_ext'),
               PYTHONDONTWRITEBYTECODE='1', PYTHONPATH=DST)
    code = 'import libs.gan.StyleGAN2.op.fused_act, libs.gan.StyleGAN2.op.upfirdn2d; print("reference ops built")'
    return subprocess.run([sys.executable, '-c', code], env=env, cwd=DST).returncode == 0


if __name__ == '__main__':
    ok = copy_tree()
    print('baseline/_ref:', 'copied from ' + REF if ok else 'reference tree not present, nothing copied')
    if ok and '--no-ext' not in sys.argv:
        print('prebuilt ext:', prebuild_ext())

#!/usr/bin/env python
"""Run an UNMODIFIED script of the reference checkout (run_inference.py, run_trainer.py, ...) with the sm_100a generator
swapped in behind its own import paths.

    python tools/run_reference_script.py /path/to/reference run_inference.py --source_path ... --target_path ...

What it does, all outside the reference tree:
  * puts <repo>/overlay before the reference on sys.path, so libs.gan.StyleGAN2.model / libs.models.direction_matrix
    resolve to this repo (PEP-420 namespace packages) and everything else to the reference;
  * absorbs the API drift between the reference's pinned stack and this image (SURVEY.md "API drift" table):
    np.product (NumPy 2) and torchvision.utils.save_image(range=...) -> value_range.
"""
import os
import runpy
import sys


def install(reference_root):
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [os.path.join(repo, 'overlay'), repo, reference_root]
    import numpy as np
    if not hasattr(np, 'product'):
        np.product = np.prod
    try:
        import torchvision.utils as vu
        _save = vu.save_image

        def save_image(tensor, fp, format=None, **kw):
            if 'range' in kw:
                kw['value_range'] = kw.pop('range')
            return _save(tensor, fp, format=format, **kw)
        vu.save_image = save_image
    except ImportError:
        pass


if __name__ == '__main__':
    if len(sys.argv) < 3:
        raise SystemExit(__doc__)
    ref, script = sys.argv[1], sys.argv[2]
    install(ref)
    sys.argv = [os.path.join(ref, script)] + sys.argv[3:]
    os.chdir(ref)
    runpy.run_path(sys.argv[0], run_name='__main__')

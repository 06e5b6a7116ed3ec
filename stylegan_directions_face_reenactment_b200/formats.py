"""On-disk formats on either side of the hot path — read and written exactly as the reference does, so
files are interchangeable with it:

  generator checkpoint   torch file, dict with key 'g_ema' = Generator.state_dict()        (libs/trainer.py:106-110,
                         run_inference.py:65-70, convert_weight.py:126-185,234); 256^2 nets load with strict=False (the
                         released voxceleb checkpoint lacks the noise buffers), 1024^2 with strict=True
  direction matrix       torch file 'A_matrix_{step:06d}.pt', dict {step, A_matrix (state_dict: linear.weight [k*512, d],
                         linear.bias), learned_directions, shift_scale, w_plus, num_layers_shift[, shift_dim]}
                         (libs/utilities/utils_train.py:592-603; read back at run_inference.py:74-85)
  inverted latent code   '<frame>.npy', float32 [n_latent, 512] (W+ of one frame; invert_images.py:118-125,
                         libs/utilities/utils_inference.py:97-100; read at libs/datasets/dataloader.py:116-119 which
                         asserts ndim == 2)

The e4e encoder that produces the codes and the DECA/loss networks stay outside this package (BASELINE north_star).
Loading a generator here also warms the packed tensor-core weight layouts (hi/lo split, K-major slabs, adjoint for the
backward pass) so that the first frame does not pay for them.
"""
import os

import numpy as np
import torch

from .direction_matrix import DirectionMatrix
from .model import Generator

STYLE_DIM = 512


def _torch_load(path, map_location='cpu'):
    try:
        return torch.load(path, map_location=map_location, weights_only=False)
    except TypeError:                                   # torch < 1.13 has no weights_only
        return torch.load(path, map_location=map_location)


# ------------------------------------------------------------------------------------------------ generator
def load_generator(path, size, channel_multiplier=2, device='cuda', strict=None, warm_batch=1, backward=False):
    """Generator(size, 512, 8, channel_multiplier) with ckpt['g_ema'] loaded, on `device`, in eval() mode — the sequence
    of libs/trainer.py:106-111 / run_inference.py:65-71.  strict defaults to the reference's choice (False at 256^2).
    warm_batch > 0 (CUDA only) runs one forward of that batch so every layer's packed operand is built once at load time;
    backward=True additionally builds the adjoint-packed weights the dL/dlatent pass reads."""
    ckpt = _torch_load(path) if isinstance(path, (str, os.PathLike)) else path
    if not isinstance(ckpt, dict) or 'g_ema' not in ckpt:
        raise RuntimeError("generator checkpoint must be a dict with key 'g_ema' (got %s)"
                           % (sorted(ckpt)[:6] if isinstance(ckpt, dict) else type(ckpt).__name__))
    g = Generator(size, STYLE_DIM, 8, channel_multiplier=channel_multiplier)
    if strict is None:
        strict = size != 256
    g.load_state_dict(ckpt['g_ema'], strict=strict)
    g = g.to(device).eval()
    if warm_batch and torch.device(device).type == 'cuda':
        warm_up(g, warm_batch, backward=backward)
    return g


def warm_up(g, batch, backward=False):
    """One synthesis pass (and optionally one backward pass to the latent) at `batch`: fills the packed-weight caches,
    the workspace and the tensor-map descriptors for that batch size."""
    dev = g.input.input.device
    w = torch.zeros(batch, g.n_latent, STYLE_DIM, device=dev, requires_grad=backward)
    with torch.enable_grad() if backward else torch.no_grad():
        img, _ = g([w], input_is_latent=True)
        if backward:
            img.sum().backward()
    torch.cuda.synchronize(dev)
    return g


def save_generator(g, path, extra=None):
    """{'g_ema': state_dict} (+ extra keys such as 'latent_avg', convert_weight.py:226-234), CPU tensors."""
    ckpt = {'g_ema': {k: v.detach().cpu() for k, v in g.state_dict().items()}}
    if extra:
        ckpt.update(extra)
    torch.save(ckpt, path)


# ------------------------------------------------------------------------------------------------ direction matrix
def load_direction_matrix(path, device='cuda'):
    """-> (A, meta).  A = DirectionMatrix(shift_dim, input_dim=learned_directions, w_plus, num_layers=num_layers_shift)
    with the stored weights, eval() mode (run_inference.py:74-88); meta = the scalar fields of the checkpoint."""
    sd = _torch_load(path) if isinstance(path, (str, os.PathLike)) else path
    for key in ('A_matrix', 'learned_directions', 'shift_scale', 'w_plus', 'num_layers_shift'):
        if key not in sd:
            raise RuntimeError('direction-matrix checkpoint lacks %r (has %s)' % (key, sorted(sd)))
    w = sd['A_matrix']['linear.weight']
    w_plus, layers, k = bool(sd['w_plus']), int(sd['num_layers_shift']), int(sd['learned_directions'])
    shift_dim = int(sd.get('shift_dim', w.shape[0] // layers if w_plus else w.shape[0]))   # save_models() omits it
    if w.shape != (shift_dim * layers if w_plus else shift_dim, k):
        raise RuntimeError('A_matrix weight %s does not match shift_dim %d x layers %d, %d directions'
                           % (tuple(w.shape), shift_dim, layers, k))
    a = DirectionMatrix(shift_dim=shift_dim, input_dim=k, out_dim=None, w_plus=w_plus, bias='linear.bias' in sd['A_matrix'],
                        num_layers=layers)
    a.load_state_dict(sd['A_matrix'])
    a = a.to(device).eval()
    a.zero_grad()
    meta = {'step': int(sd.get('step', 0)), 'learned_directions': k, 'shift_scale': sd['shift_scale'], 'w_plus': w_plus,
            'num_layers_shift': layers, 'shift_dim': shift_dim}
    return a, meta


def save_direction_matrix(a, step, models_dir, learned_directions, shift_scale, w_plus, num_layers_shift):
    """utils_train.py:592-603, plus 'shift_dim' (the released checkpoints carry it: run_inference.py:80)."""
    sd = {'step': int(step), 'A_matrix': {k: v.detach().cpu() for k, v in a.state_dict().items()},
          'learned_directions': learned_directions, 'shift_scale': shift_scale, 'w_plus': w_plus,
          'num_layers_shift': num_layers_shift, 'shift_dim': a.shift_dim}
    path = os.path.join(models_dir, 'A_matrix_{:06d}.pt'.format(int(step)))
    torch.save(sd, path)
    return path


# ------------------------------------------------------------------------------------------------ latent codes
def save_latent_code(path, code):
    """One frame's W+ code -> float32 [n_latent, 512] .npy (invert_images.py:118-125)."""
    arr = code.detach().cpu().numpy() if torch.is_tensor(code) else np.asarray(code)
    if arr.ndim == 3 and arr.shape[0] == 1:
        arr = arr[0]
    if arr.ndim != 2 or arr.shape[1] != STYLE_DIM:
        raise RuntimeError('latent code must be [n_latent, 512], got %s' % (arr.shape,))
    np.save(path, arr.astype(np.float32, copy=False))


def load_latent_codes(paths, n_latent=None, pin=True):
    """Frames' codes -> float32 tensor [N, n_latent, 512] in (pinned) host memory, ready for one async copy to the GPU.
    `paths`: a directory (its *.npy in sorted order, the convention of dataloader.py:60-75), one file, or a list."""
    if isinstance(paths, (str, os.PathLike)):
        if os.path.isdir(paths):
            paths = sorted(os.path.join(paths, f) for f in os.listdir(paths) if f.endswith('.npy'))
        else:
            paths = [paths]
    if not paths:
        raise RuntimeError('no latent codes to load')
    codes = []
    for p in paths:
        c = np.load(p)
        if c.ndim != 2 or c.shape[1] != STYLE_DIM:      # the reference asserts ndim == 2 (dataloader.py:119)
            raise RuntimeError('%s: latent code dimensions should be n_latent x 512, got %s' % (p, c.shape))
        if n_latent is not None and c.shape[0] != n_latent:
            raise RuntimeError('%s: %d latent rows, generator expects %d' % (p, c.shape[0], n_latent))
        codes.append(c.astype(np.float32, copy=False))
    if len({c.shape for c in codes}) != 1:
        raise RuntimeError('latent codes of different shapes: %s' % sorted({c.shape for c in codes}))
    out = torch.from_numpy(np.stack(codes))
    if pin and torch.cuda.is_available():
        out = out.pin_memory()
    return out

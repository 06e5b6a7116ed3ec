"""On-disk formats either side of the hot path: files written by the reference load here bit-exactly,
files written here have the reference's layout.  CPU-only (no compute call)."""
import os

import numpy as np
import pytest
import torch

from oracle import stylegan2_oracle as orc
from stylegan_directions_face_reenactment_b200 import formats

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def test_direction_matrix_checkpoint_written_by_the_reference(golden):
    g = golden('formats.npz')
    a, meta = formats.load_direction_matrix(os.path.join(GOLDEN, 'A_matrix_000010.pt'), device='cpu')
    assert meta == {'step': 10, 'learned_directions': 15, 'shift_scale': 6, 'w_plus': True, 'num_layers_shift': 8,
                    'shift_dim': 512}                       # shift_dim recovered: utils_train.save_models omits it
    assert not a.training and a.input_dim == 15 and a.num_layers == 8
    shift = a(torch.from_numpy(g['dp'])).detach()
    assert tuple(shift.shape) == (3, 8, 512)
    assert float((shift - torch.from_numpy(g['shift'])).abs().max()) <= 1e-6


def test_direction_matrix_roundtrip_and_errors(tmp_path):
    from stylegan_directions_face_reenactment_b200 import DirectionMatrix
    torch.manual_seed(3)
    a = DirectionMatrix(512, input_dim=15, w_plus=True, num_layers=8)
    path = formats.save_direction_matrix(a, 1234, str(tmp_path), learned_directions=15, shift_scale=6, w_plus=True,
                                         num_layers_shift=8)
    assert os.path.basename(path) == 'A_matrix_001234.pt'            # utils_train.py:602
    raw = torch.load(path, weights_only=False)
    assert set(raw) == {'step', 'A_matrix', 'learned_directions', 'shift_scale', 'w_plus', 'num_layers_shift', 'shift_dim'}
    assert set(raw['A_matrix']) == {'linear.weight', 'linear.bias'} and tuple(raw['A_matrix']['linear.weight'].shape) == (4096, 15)
    b, meta = formats.load_direction_matrix(path, device='cpu')
    assert meta['step'] == 1234 and torch.equal(b.linear.weight, a.linear.weight) and torch.equal(b.linear.bias, a.linear.bias)
    del raw['w_plus']
    with pytest.raises(RuntimeError, match='w_plus'):
        formats.load_direction_matrix(raw, device='cpu')
    raw['w_plus'] = True
    raw['num_layers_shift'] = 7                                      # 4096 rows do not split into 7 layers of 512
    raw['shift_dim'] = 512
    with pytest.raises(RuntimeError, match='does not match'):
        formats.load_direction_matrix(raw, device='cpu')


def test_latent_codes(tmp_path, golden):
    g = golden('formats.npz')
    ref_file = os.path.join(GOLDEN, 'latent_000.npy')                # written by the reference's inversion code path
    codes = formats.load_latent_codes(ref_file, n_latent=4, pin=False)
    assert codes.dtype == torch.float32 and tuple(codes.shape) == (1, 4, 512)
    assert np.array_equal(codes[0].numpy(), g['latent'])
    d = tmp_path / 'latent_codes'
    d.mkdir()
    rng = np.random.Generator(np.random.PCG64(1))
    frames = rng.standard_normal((3, 14, 512), dtype=np.float32)
    for i in (2, 0, 1):
        formats.save_latent_code(str(d / ('%06d.npy' % i)), torch.from_numpy(frames[i:i + 1]))    # [1,14,512] is squeezed
    (d / 'notes.txt').write_text('ignored')
    back = formats.load_latent_codes(str(d), n_latent=14, pin=False)
    assert np.array_equal(back.numpy(), frames)                      # sorted file order
    assert np.load(str(d / '000000.npy')).shape == (14, 512)         # what dataloader.py:116-119 asserts (ndim == 2)
    np.save(str(tmp_path / 'bad.npy'), frames)                       # 3-D
    with pytest.raises(RuntimeError, match='n_latent x 512'):
        formats.load_latent_codes(str(tmp_path / 'bad.npy'))
    with pytest.raises(RuntimeError, match='generator expects'):
        formats.load_latent_codes(str(d), n_latent=18)
    with pytest.raises(RuntimeError, match=r'\[n_latent, 512\]'):
        formats.save_latent_code(str(tmp_path / 'x.npy'), np.zeros((14, 256), np.float32))


def test_generator_checkpoint(tmp_path):
    sd = orc.seeded_state_dict(32, 2, seed=4)
    path = str(tmp_path / 'g.pt')
    torch.save({'g_ema': sd, 'latent_avg': torch.zeros(512)}, path)  # convert_weight.py:226-234 layout
    g = formats.load_generator(path, 32, channel_multiplier=2, device='cpu', warm_batch=0)
    assert not g.training
    got = g.state_dict()
    assert set(got) == set(sd) and all(torch.equal(got[k], sd[k]) for k in sd)
    out = str(tmp_path / 'g2.pt')
    formats.save_generator(g, out, extra={'latent_avg': torch.ones(512)})
    raw = torch.load(out, weights_only=False)
    assert set(raw) == {'g_ema', 'latent_avg'} and all(torch.equal(raw['g_ema'][k], sd[k]) for k in sd)
    # a 256^2 checkpoint without noise buffers loads (strict=False, libs/trainer.py:107-108); a 1024-style strict load fails
    sd256 = {k: v for k, v in orc.seeded_state_dict(256, 1, seed=5).items() if not k.startswith('noises.')}
    g256 = formats.load_generator({'g_ema': sd256}, 256, channel_multiplier=1, device='cpu', warm_batch=0)
    assert tuple(g256.noises.noise_12.shape) == (1, 1, 256, 256)
    with pytest.raises(RuntimeError):
        formats.load_generator({'g_ema': sd256}, 256, channel_multiplier=1, device='cpu', warm_batch=0, strict=True)
    with pytest.raises(RuntimeError, match='g_ema'):
        formats.load_generator({'g': sd}, 32, device='cpu', warm_batch=0)

"""Bring-up aid (GPU box): per-latent-row error of dL/dlatent vs fp64 autograd through the oracle, next to the error of
fp32 autograd (the reference's arithmetic) against the same fp64 result."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import stylegan2_oracle as orc  # noqa: E402
import stylegan_directions_face_reenactment_b200 as pkg  # noqa: E402


def run(size, cm, batch):
    sd = orc.seeded_state_dict(size, cm, seed=12)
    G = pkg.Generator(size, 512, 8, channel_multiplier=cm)
    G.load_state_dict(sd, strict=True)
    G = G.cuda().eval()
    wplus = orc.seeded_wplus(sd, batch, G.n_latent, seed=21)
    rng = np.random.Generator(np.random.PCG64(3))
    r = torch.from_numpy(rng.standard_normal((batch, 3, size, size), dtype=np.float32))
    grads = {}
    for name, dt in [('f64', torch.float64), ('f32', torch.float32)]:
        sdd = {k: v.to(dt) for k, v in sd.items()}
        w = wplus.detach().clone().to(dt).requires_grad_(True)
        img, _ = orc.generator_forward(sdd, [w], size, cm, input_is_latent=True)
        (img * r.to(dt)).sum().backward()
        grads[name] = w.grad.double()
    wg = wplus.detach().clone().cuda().requires_grad_(True)
    img, _ = G([wg], input_is_latent=True)
    (img * r.cuda()).sum().backward()
    ours = wg.grad.double().cpu()
    ref = grads['f64']
    print('net %d cm %d B %d   max|g| = %.3e' % (size, cm, batch, ref.abs().max()))
    for row in range(G.n_latent):
        m = ref[:, row].abs().max().item()
        print('  row %2d max|g| %.3e   ours-f64 %.3e (%.1e rel)   f32-f64 %.3e (%.1e rel)' % (
            row, m, (ours[:, row] - ref[:, row]).abs().max(), (ours[:, row] - ref[:, row]).abs().max() / m,
            (grads['f32'][:, row] - ref[:, row]).abs().max(), (grads['f32'][:, row] - ref[:, row]).abs().max() / m),
            flush=True)


if __name__ == '__main__':
    run(8, 2, 2)
    run(32, 2, 3)
    run(256, 1, 1)

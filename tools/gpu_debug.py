"""Bring-up aid (run on the GPU box): per-layer error of the fused network against the CPU oracle, plus a
single-layer sweep of the tcgen05 modconv.  Not a test; prints a table."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import stylegan2_oracle as orc  # noqa: E402
import stylegan_directions_face_reenactment_b200 as pkg  # noqa: E402


def layer_sweep():
    rng = np.random.Generator(np.random.PCG64(1))
    for cin, cout, h, up, b in [(32, 32, 8, False, 2), (64, 64, 16, False, 1), (64, 64, 16, True, 1),
                                (512, 512, 4, False, 3), (512, 512, 4, True, 3), (128, 64, 32, True, 2),
                                (256, 256, 64, False, 1), (64, 64, 256, False, 1)]:
        m = pkg.ModulatedConv2d(cin, cout, 3, 512, upsample=up)
        x = torch.from_numpy(rng.standard_normal((b, cin, h, h), dtype=np.float32))
        w = torch.from_numpy(rng.standard_normal((b, 512), dtype=np.float32))
        with torch.no_grad():
            ref = orc.modulated_conv2d(x.double(), w.double(), m.weight.double(), m.modulation.weight.double(),
                                       m.modulation.bias.double(), True, up).float()
            m = m.cuda()
            t0 = time.time()
            y = m(x.cuda(), w.cuda())
            torch.cuda.synchronize()
            e = (y.cpu() - ref).abs().max().item()
            big = ref.abs() > 0.5
            bias = (((y.cpu() - ref) / ref)[big]).mean().item()      # signed: accumulate-truncation shows as shrink
        print('modconv cin=%d cout=%d h=%d up=%d B=%d  max|ref|=%.3f  err=%.3e  mean signed rel err=%.3e (%.1f ms)' %
              (cin, cout, h, up, b, ref.abs().max().item(), e, bias, (time.time() - t0) * 1e3), flush=True)


def network(size, cm, batch):
    sd = orc.seeded_state_dict(size, cm, seed=0)
    G = pkg.Generator(size, 512, 8, channel_multiplier=cm)
    G.load_state_dict(sd, strict=True)
    G = G.cuda().eval()
    wplus = orc.seeded_wplus(sd, batch, G.n_latent, seed=1234)
    with torch.no_grad():
        ref_img, _, ref_feats = orc.generator_forward(sd, [wplus], size, cm, input_is_latent=True, return_features=True)
        img, feats = G.synthesis(wplus.cuda(), return_features=True)
        torch.cuda.synchronize()
    for i, (f, r) in enumerate(zip(feats, ref_feats)):
        print('  net%d layer %2d %-18s max|ref|=%8.3f err=%.3e' % (size, i, tuple(r.shape), r.abs().max().item(),
                                                                  (f.cpu() - r).abs().max().item()), flush=True)
    print('  net%d image max|ref|=%.3f err=%.3e' % (size, ref_img.abs().max().item(),
                                                    (img.cpu() - ref_img).abs().max().item()), flush=True)


if __name__ == '__main__':
    print(torch.cuda.get_device_name(0), flush=True)
    layer_sweep()
    network(8, 2, 2)
    network(32, 2, 3)
    network(256, 1, 2)

"""Small driver for compute-sanitizer: exercises every kernel family (plain / halo MT=1,2 / concat / scatter NT=128,256 /
split-K / finish passes / backward incl. gather adjoint / parameter gradients / output stage) at modest sizes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import stylegan2_oracle as orc  # noqa: E402
import stylegan_directions_face_reenactment_b200 as pkg  # noqa: E402

CASES = [(256, 1, 2, False), (64, 2, 3, True), (1024, 2, 1, False), (256, 2, 4, False), (1024, 2, 4, False)]
if len(sys.argv) > 1:                                   # python tools/sanitizer_cases.py 1   -> only CASES[1]
    CASES = [CASES[int(v)] for v in sys.argv[1:]]
for size, cm, batch, train in CASES:
    sd = orc.seeded_state_dict(size, cm, seed=1)
    G = pkg.Generator(size, 512, 8, channel_multiplier=cm)
    G.load_state_dict(sd, strict=True)
    G = G.cuda()
    G.train() if train else G.eval()
    w = orc.seeded_wplus(sd, batch, G.n_latent, seed=2).cuda().requires_grad_(size <= 256)
    img, _ = G([w], input_is_latent=True)
    if size <= 256:
        img.square().mean().backward()
        assert torch.isfinite(w.grad).all()
        if train:
            assert all(torch.isfinite(p.grad).all() for n, p in G.named_parameters() if not n.startswith('style.'))
    u8 = pkg.frames_to_uint8(img.detach(), size=min(size, 256))
    torch.cuda.synchronize()
    print('ok', size, cm, batch, train, float(img.abs().max()), tuple(u8.shape), flush=True)
    del G, img, w
    torch.cuda.empty_cache()

#!/usr/bin/env python
"""Drives the HBM-bound kernels north_star names, at the bench shapes, for an `ncu --set full` capture:
upfirdn2d_kernel in the generator's three modes (blur after an up-conv: pad (1,1) on the (2H+1)^2 grid; RGB-skip 2x
upsampling: up 2, pad (2,1); its gradient: down 2, pad (1,1)), then one train()-mode forward + backward at 256^2/cm1
(torgb_tail_kernel, bwd_act_kernel, up_bwd_prepare_kernel, ...).

    ncu --set full --clock-control none -k regex:"upfirdn2d_kernel|torgb_tail|bwd_act|up_bwd_prepare|frames_to_uint8" \
        -o gpurun_out/prof_hbm python tools/gpu_hbm_kernels.py
    python tools/ncu_summary.py hbm gpurun_out/prof_hbm.ncu-rep profiles/r2_hbm_kernels_ncu.md
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import stylegan2_oracle as orc  # noqa: E402
import stylegan_directions_face_reenactment_b200 as pkg  # noqa: E402
from stylegan_directions_face_reenactment_b200.ops import upfirdn2d  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device('cuda')
k = orc.make_fir_kernel([1, 3, 3, 1]).to(dev)
with torch.no_grad():
    x = torch.randn(B, 64, 257, 257, device=dev)             # Blur after convs.10 (128 -> 256 up layer, 64 channels)
    upfirdn2d(x, k * 4, pad=(1, 1))
    s = torch.randn(B, 3, 128, 128, device=dev)              # ToRGB skip upsampling 128 -> 256
    y = upfirdn2d(s, k * 4, up=2, pad=(2, 1))
    upfirdn2d(y, torch.flip(k * 4, [0, 1]), down=2, pad=(1, 1))   # its gradient (kernel mode 5 of the reference)
    big = torch.randn(32 * 64, 1, 257, 257, device=dev)      # same blur at the bench batch: 1.08 GB in + out
    upfirdn2d(big, k * 4, pad=(1, 1))
    del big
sd = orc.seeded_state_dict(256, 1, seed=0)
G = pkg.Generator(256, 512, 8, channel_multiplier=1)
G.load_state_dict(sd, strict=True)
G = G.to(dev).train()
w = orc.seeded_wplus(sd, B, G.n_latent, seed=2).to(dev).requires_grad_(True)
for _ in range(2):
    img, _ = G([w], input_is_latent=True)
    img.square().mean().backward()
u8 = pkg.frames_to_uint8(img.detach())
torch.cuda.synchronize()
print('ok', float(img.abs().max()), tuple(u8.shape))

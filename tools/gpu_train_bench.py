"""Times the A-matrix training step pieces (BASELINE config 4 shape: B=16/GPU, 256^2 cm=1): no-grad forward, autograd
forward (saves every StyledConv output), backward to the latent.   python tools/gpu_train_bench.py [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import stylegan2_oracle as orc  # noqa: E402
import stylegan_directions_face_reenactment_b200 as pkg  # noqa: E402


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main(B):
    size, cm = 256, 1
    sd = orc.seeded_state_dict(size, cm, seed=0)
    G = pkg.Generator(size, 512, 8, channel_multiplier=cm)
    G.load_state_dict(sd, strict=True)
    G = G.cuda().eval()
    w = orc.seeded_wplus(sd, B, G.n_latent, seed=1).cuda()
    r = torch.randn(B, 3, size, size, device='cuda')

    def fwd_nograd():
        with torch.no_grad():
            G([w], input_is_latent=True)

    state = {}

    def fwd_grad():
        wg = w.clone().requires_grad_(True)
        img, _ = G([wg], input_is_latent=True)
        state['img'], state['wg'] = img, wg

    def fwd_bwd():
        fwd_grad()
        (state['img'] * r).sum().backward()

    t0, t1, t2 = timed(fwd_nograd), timed(fwd_grad), timed(fwd_bwd)
    print('B=%d  forward(no grad) %.3f ms   forward(autograd) %.3f ms   forward+backward %.3f ms   backward alone %.3f ms'
          % (B, t0, t1, t2, t2 - t1))


if __name__ == '__main__':
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 16)

"""Times one step of the generator fine-tuning loop (optimize_g, libs/optimization.py:45-68) on the hot
path: forward in train() mode, L2 loss to a target frame, backward with every generator-parameter gradient (tcgen05
weight-gradient GEMMs, csrc/wgrad_sm100.cu), Adam on convs[4..].   python tools/gpu_optimize_g_bench.py [B] [size] [cm]
Also prints the same step with the weight gradients through ATen/cuDNN (SGR_WGRAD=aten) for comparison."""
import copy
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import stylegan2_oracle as orc  # noqa: E402
import stylegan_directions_face_reenactment_b200 as pkg  # noqa: E402
from stylegan_directions_face_reenactment_b200 import backward as bwd  # noqa: E402


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


def main(B, size, cm):
    sd = orc.seeded_state_dict(size, cm, seed=0)
    G = pkg.Generator(size, 512, 8, channel_multiplier=cm)
    G.load_state_dict(sd, strict=True)
    G = G.cuda()
    latent = orc.seeded_wplus(sd, B, G.n_latent, seed=2).cuda()
    with torch.no_grad():
        target = copy.deepcopy(G).eval()([latent], input_is_latent=True)[0] * 0.9
    G.train()
    params = [p for i in range(4, len(G.convs)) for p in G.convs[i].parameters()]
    opt = torch.optim.Adam(params, lr=1e-5)

    def step():
        img, _ = G([latent], input_is_latent=True)
        loss = (img - target).pow(2).mean()
        opt.zero_grad()
        loss.backward()
        opt.step()

    def fwd_bwd():
        img, _ = G([latent], input_is_latent=True)
        G.zero_grad(set_to_none=True)
        (img - target).pow(2).mean().backward()

    out = {'batch': B, 'size': size, 'cm': cm}
    out['step_ms'] = timed(step)
    out['fwd_bwd_ms'] = timed(fwd_bwd)
    bwd.FORCE_ATEN_WGRAD = True
    out['fwd_bwd_ms_aten_wgrad'] = timed(fwd_bwd)
    bwd.FORCE_ATEN_WGRAD = False
    G.eval()
    lat = latent.clone().requires_grad_(True)

    def frozen():
        img, _ = G([lat], input_is_latent=True)
        lat.grad = None
        (img - target).pow(2).mean().backward()
    out['fwd_bwd_ms_frozen_generator'] = timed(frozen)
    print(json.dumps(out))


if __name__ == '__main__':
    a = [int(v) for v in sys.argv[1:]]
    main(a[0] if a else 1, a[1] if len(a) > 1 else 256, a[2] if len(a) > 2 else 1)

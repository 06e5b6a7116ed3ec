"""Batch-1 latency of Generator.forward (how run_inference.py:170-181 drives the generator: one frame per call), eager launches
vs the captured CUDA graph (`Generator.enable_cuda_graphs()`), and the same at B = 4 / 32 for context."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import stylegan2_oracle as orc  # noqa: E402
import stylegan_directions_face_reenactment_b200 as pkg  # noqa: E402

sd = orc.seeded_state_dict(256, 1, seed=0)
G = pkg.Generator(256, 512, 8, channel_multiplier=1)
G.load_state_dict(sd, strict=True)
G = G.cuda().eval().requires_grad_(False)
trunc = orc.seeded_wplus(sd, 1, 1, seed=7)[:, 0].cuda()
BATCHES = [int(v) for v in sys.argv[1].split(',')] if len(sys.argv) > 1 else [1, 4, 32]
for B in BATCHES:
    w = orc.seeded_wplus(sd, B, G.n_latent, seed=3).cuda()
    for graphs in (False, True):
        G.enable_cuda_graphs(graphs)

        def run():
            with torch.no_grad():
                return G([w], input_is_latent=True, truncation=0.7, truncation_latent=trunc)[0]
        for _ in range(5):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 200 if B == 1 else 50
        e0.record()
        for _ in range(n):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print('B=%2d %-6s %.3f ms per call, %.3f ms per frame, %.0f frames/s' % (B, 'graph' if graphs else 'eager', ms, ms / B, B / ms * 1e3), flush=True)

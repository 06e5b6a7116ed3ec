"""Experiment: one B=32 synthesis pass vs two concurrent B=16 passes on two CUDA streams (two generator replicas, so each
stream has its own workspace).  Does the HBM-bound FIR pass of one half overlap the tensor-bound GEMMs of the other?
    python tools/gpu_two_stream_check.py [B]"""
import copy
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import stylegan2_oracle as orc  # noqa: E402
import stylegan_directions_face_reenactment_b200 as pkg  # noqa: E402


def main(B, parts):
    size, cm = 256, 1
    sd = orc.seeded_state_dict(size, cm, seed=0)
    G = pkg.Generator(size, 512, 8, channel_multiplier=cm)
    G.load_state_dict(sd, strict=True)
    G = G.cuda().eval()
    Gs = [G] + [copy.deepcopy(G) for _ in range(parts - 1)]
    w = orc.seeded_wplus(sd, B, G.n_latent, seed=2).cuda()
    streams = [torch.cuda.Stream() for _ in range(parts)]
    chunks = list(w.chunk(parts))

    def one():
        with torch.no_grad():
            return G([w], input_is_latent=True)[0]

    def split():
        outs = []
        cur = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(cur)
        with torch.no_grad():
            for g, s, c in zip(Gs, streams, chunks):
                s.wait_event(ev)
                with torch.cuda.stream(s):
                    outs.append(g([c], input_is_latent=True)[0])
        for s in streams:
            cur.wait_stream(s)
        return outs

    def timed(fn, reps=20):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    a = one()
    b = torch.cat(split())
    print('max diff one vs split: %.3e' % (a - b).abs().max().item())
    t1 = timed(one)
    t2 = timed(split)
    print('B=%d  one pass %.3f ms (%.0f frames/s)   %d concurrent passes %.3f ms (%.0f frames/s)'
          % (B, t1, B / t1 * 1e3, parts, t2, B / t2 * 1e3))


if __name__ == '__main__':
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 32, int(sys.argv[2]) if len(sys.argv) > 2 else 2)

"""Synthesis throughput sweep (BASELINE config 5 and friends): frames/s of Generator.forward on random W+ for a list of
batch sizes, in the fp32-parity mode (bf16x3) and the single-pass bf16 mode, with the image-level parity of the latter.
    python tools/gpu_sweep.py [size] [channel_multiplier] [batches...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import stylegan2_oracle as orc  # noqa: E402
import stylegan_directions_face_reenactment_b200 as pkg  # noqa: E402


def timed(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    cm = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    batches = [int(v) for v in sys.argv[3:]] or [1, 2, 4, 8, 16, 32]
    sd = orc.seeded_state_dict(size, cm, seed=0)
    G = pkg.Generator(size, 512, 8, channel_multiplier=cm)
    G.load_state_dict(sd, strict=True)
    G = G.cuda().eval()
    gflop = orc.forward_flops_per_frame(size, cm) / 1e9
    print('Generator(%d, cm=%d): %.1f GFLOP/frame (algorithmic)' % (size, cm, gflop))
    for B in batches:
        w = orc.seeded_wplus(sd, B, G.n_latent, seed=B).cuda()
        out = {}
        res = {}
        for mode in ('bf16x3', 'bf16'):
            os.environ['SGR_PRECISION'] = mode

            def run():
                with torch.no_grad():
                    out[mode] = G([w], input_is_latent=True)[0]
            ms = timed(run, 5 if B >= 8 else 10)
            res[mode] = ms
        os.environ['SGR_PRECISION'] = 'bf16x3'
        err = (out['bf16'] - out['bf16x3']).abs().max().item()
        rng = out['bf16x3'].abs().max().item()
        print('B=%2d  bf16x3 %8.3f ms %8.1f frames/s %6.0f algo TFLOP/s | bf16 %8.3f ms %8.1f frames/s | bf16 vs bf16x3 max-abs '
              '%.2e (range %.1f)' % (B, res['bf16x3'], B / res['bf16x3'] * 1e3, B * gflop / res['bf16x3'], res['bf16'],
                                     B / res['bf16'] * 1e3, err, rng), flush=True)
        del out


if __name__ == '__main__':
    main()

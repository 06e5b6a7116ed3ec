#!/usr/bin/env python
"""Times the UNMODIFIED reference generator (libs/gan/StyleGAN2, its own two JIT CUDA ops + ATen/cuDNN grouped
convolutions, reference model.py:254,263,269, op/fused_act.py:10-17, op/upfirdn2d.py:11-17) on the GPU, on the bench
workload, and compares its frames with libsgr's.  This is "the kernel to beat" (SURVEY.md §2a).

The reference tree is not in the repository: tools/make_baseline_ref.py copies it to the git-ignored baseline/_ref
(which travels with the gpurun snapshot).  Prints ONE JSON line; bench.py embeds it as `gpu_reference`.

    python tools/gpu_reference_bench.py [--batch 32] [--steps 20] [--warmup 3] [--ours 1]

Two precisions are timed: `tf32` = torch's defaults (cudnn.allow_tf32 = True, what the unmodified scripts get on this
stack) and `fp32` = cudnn.allow_tf32 = False (the arithmetic the parity bar is stated against).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def ref_root():
    """baseline/_ref, through /root/repo when that is the same directory (keeps the prebuilt ninja paths valid)."""
    cand = os.path.join(ROOT, 'baseline', '_ref')
    alias = '/root/repo/baseline/_ref'
    try:
        if os.path.isdir(alias) and os.path.samefile(alias, cand):
            return alias
    except OSError:
        pass
    return cand


def import_reference():
    import numpy as np
    ref = ref_root()
    if not os.path.isdir(os.path.join(ref, 'libs', 'gan', 'StyleGAN2')):
        raise RuntimeError('baseline/_ref is missing: run tools/make_baseline_ref.py where /root/reference exists')
    os.environ.setdefault('TORCH_EXTENSIONS_DIR', os.path.join(ref, '_ext'))
    os.environ.setdefault('TORCH_CUDA_ARCH_LIST', '10.0')
    if not hasattr(np, 'product'):
        np.product = np.prod                 # libs/models/direction_matrix.py:11-12 predates NumPy 2 (environment shim)
    sys.dont_write_bytecode = True
    sys.path.insert(0, ref)
    import libs.gan.StyleGAN2.model as refm
    from libs.models.direction_matrix import DirectionMatrix
    from libs.utilities.generic import generate_image
    return refm, DirectionMatrix, generate_image


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--ours', type=int, default=1)
    ap.add_argument('--size', type=int, default=256)
    ap.add_argument('--cm', type=int, default=1)
    args = ap.parse_args()
    import torch
    from oracle import stylegan2_oracle as orc
    out = {'kind': 'unmodified reference generator on the GPU (cuDNN grouped conv + its JIT fused_bias_act / upfirdn2d ops)'}
    try:
        refm, RefA, ref_generate_image = import_reference()
    except Exception as e:  # noqa: BLE001
        print(json.dumps({'unavailable': ('%s: %s' % (type(e).__name__, e))[:300]}))
        return
    dev = torch.device('cuda')
    size, cm, B = args.size, args.cm, args.batch
    sd = orc.seeded_state_dict(size, cm, seed=0)
    G = refm.Generator(size, 512, 8, channel_multiplier=cm)
    G.load_state_dict(sd, strict=True)
    G = G.to(dev).eval()
    torch.manual_seed(5)
    A = RefA(512, input_dim=15, out_dim=512, w_plus=True, num_layers=8).to(dev)
    a_sd = {k: v.detach().clone() for k, v in A.state_dict().items()}
    torch.manual_seed(7)
    trunc = G.mean_latent(4096).detach()
    wsrc = orc.seeded_wplus(sd, 1, G.n_latent, seed=11).to(dev).repeat(B, 1, 1)
    g = torch.Generator().manual_seed(4321)
    dps = [(torch.rand(B, 15, generator=g) * 6 - 3).to(dev) for _ in range(4)]

    def step(i):
        with torch.no_grad():
            return ref_generate_image(G, wsrc, 0.7, trunc, w_plus=True, num_layers_shift=8, shift_code=A(dps[i % 4]),
                                      input_is_latent=True)

    def timed(steps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    frames = {}
    for name, tf32 in (('tf32', True), ('fp32', False)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = False          # torch default
        for i in range(args.warmup):
            step(i)
        ms = timed(args.steps)
        frames[name] = step(0).float().cpu()
        out[name] = {'ms_per_step': ms, 'frames_s': B / (ms * 1e-3), 'cudnn_allow_tf32': tf32,
                     'cudnn_benchmark': bool(torch.backends.cudnn.benchmark)}
    out['batch'] = B
    out['workload'] = 'generate_image(G, w_src, 0.7, trunc, shift_code=A(dp)), Generator(%d, cm=%d), B=%d, weights/inputs of bench.py' % (size, cm, B)
    out['tf32_vs_fp32_max_abs'] = (frames['tf32'] - frames['fp32']).abs().max().item()
    out['range'] = frames['fp32'].abs().max().item()
    if args.ours:
        import stylegan_directions_face_reenactment_b200 as pkg
        G2 = pkg.Generator(size, 512, 8, channel_multiplier=cm)
        G2.load_state_dict(sd, strict=True)
        G2 = G2.to(dev).eval()
        A2 = pkg.DirectionMatrix(512, input_dim=15, out_dim=512, w_plus=True, num_layers=8).to(dev)
        A2.load_state_dict(a_sd)
        with torch.no_grad():
            img = pkg.generate_image(G2, wsrc, 0.7, trunc, w_plus=True, num_layers_shift=8, shift_code=A2(dps[0]),
                                     input_is_latent=True).float().cpu()
        out['max_abs_ours_vs_reference_fp32'] = (img - frames['fp32']).abs().max().item()
        out['max_abs_ours_vs_reference_tf32'] = (img - frames['tf32']).abs().max().item()
    print(json.dumps(out), flush=True)


if __name__ == '__main__':
    main()

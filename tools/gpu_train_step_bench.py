"""BASELINE config 4 on the hot path: the A-matrix training step's generator part (libs/trainer.py:175-189) at
B=16 per GPU, 256^2 cm=1 — two no-grad forwards from Z (source / target), one autograd forward from the shifted W+,
backward to A, ONE flat NCCL all-reduce of A's 65 536-float gradient, Adam.  The loss heads (ArcFace / LPIPS / DECA) are
outside the path; a surrogate L1 to the target image stands in for them.  Run alone or under torchrun:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/gpu_train_step_bench.py
Prints samples/s (all ranks), step ms (max over ranks) and the all-reduce time."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import stylegan2_oracle as orc  # noqa: E402
import stylegan_directions_face_reenactment_b200 as pkg  # noqa: E402
from stylegan_directions_face_reenactment_b200 import dist as sdist  # noqa: E402


def main():
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    rank, world = sdist.init(device=dev)
    B, steps, warm = 16, 10, 3
    sd = orc.seeded_state_dict(256, 1, seed=0)
    G = pkg.Generator(256, 512, 8, channel_multiplier=1)
    G.load_state_dict(sd, strict=True)
    G = G.to(dev).eval()
    torch.manual_seed(5)
    A = pkg.DirectionMatrix(512, input_dim=15, out_dim=512, w_plus=True, num_layers=8).to(dev)
    sdist.broadcast_params_(A)
    opt = torch.optim.Adam(A.parameters(), lr=1e-4, weight_decay=5e-4)
    torch.manual_seed(7)
    trunc = G.mean_latent(4096).detach()
    g = torch.Generator(device=dev).manual_seed(100 + rank)

    def step():
        z_src = torch.randn(B, 512, device=dev, generator=g)
        z_tgt = torch.randn(B, 512, device=dev, generator=g)
        with torch.no_grad():
            _, w_src = pkg.generate_image(G, z_src, 0.7, trunc, input_is_latent=False, return_latents=True)
            tgt = pkg.generate_image(G, z_tgt, 0.7, trunc, input_is_latent=False)
        dp = torch.rand(B, 15, device=dev, generator=g) * 6 - 3
        return sdist.train_step(G, A, opt, w_src, dp, 0.7, trunc, lambda img: (img - tgt).abs().mean())

    for _ in range(warm):
        step()
    # all-reduce alone (same bucket)
    flat = torch.zeros(65536, device=dev)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(20):
            dist.all_reduce(flat)
        a1.record()
        torch.cuda.synchronize()
        ar_ms = a0.elapsed_time(a1) / 20
    else:
        ar_ms = 0.0
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss, nbytes = step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # replicas must stay identical
    w = A.linear.weight.detach().clone()
    if world > 1:
        w0 = w.clone()
        dist.broadcast(w0, src=0)
        same = bool(torch.equal(w, w0))
    else:
        same = True
    if rank == 0:
        print(json.dumps({'config': 'A-matrix train step (generator part), 256^2 cm=1, B=16/GPU', 'n_gpus': world,
                          'ms_per_step': ms.item(), 'samples_per_s': B * world / (ms.item() * 1e-3),
                          'allreduce_bytes': int(nbytes), 'allreduce_ms': ar_ms, 'loss': float(loss),
                          'replicas_identical': same}), flush=True)
    if not same:
        raise SystemExit('replicas diverged')
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

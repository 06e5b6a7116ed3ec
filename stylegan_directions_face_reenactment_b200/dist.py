"""Multi-GPU plumbing: one process per GPU, frames (batch rows) sharded over ranks.

The reference has no distributed code at all (libs/trainer.py hard-codes 'cuda').  The path shards naturally:
  * inference: rank r renders frames r, r+world, ... of the driving sequence — no collective;
  * A-matrix training: every rank holds a replica of the frozen generator and of A, runs its own batch, and ONE flat
    all-reduce carries A's gradient (linear.weight [4096,15] + linear.bias [4096] = 65 536 floats = 256 KiB, purely
    latency-bound on NVLink 5 / NVSwitch), after which identical Adam steps keep the replicas in sync
    (Adam on A only: libs/trainer.py:144,187-189).
Backend: NCCL on GPUs; the same code runs on gloo for the CPU tests of the host logic.
"""
import os

import torch
import torch.distributed as dist


def init(backend=None, device=None):
    """Initialise torch.distributed from the torchrun environment (no-op for a single process). Returns (rank, world)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        kw = {'device_id': device} if (backend == 'nccl' and device is not None) else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world


def shard_indices(n_items, rank, world):
    """Round-robin frame ownership: rank r gets r, r+world, ... (cfg 3 of SURVEY.md §8d)."""
    return list(range(rank, n_items, world))


def allreduce_mean_grads_(params, group=None):
    """Average the gradients of `params` over all ranks with ONE collective on a flat fp32 bucket, in place."""
    params = [p for p in params if p.grad is not None]
    if not params or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(dist.get_world_size(group))
    off = 0
    for p in params:
        n = p.grad.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n
    return flat.numel() * 4


def broadcast_params_(module, src=0, group=None):
    """Make every replica start from rank `src`'s parameters (A is randomly initialised, direction_matrix.py:28-32)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        for p in module.parameters():
            dist.broadcast(p.data, src=src, group=group)


def train_step(G, A, optimizer, latent_code, shift_vector, truncation, trunc, loss_fn, num_layers_shift=8,
               input_is_latent=True):
    """One data-parallel A-matrix step on this rank's batch (the generator part of libs/trainer.py:175-189):
    shift = A(dp) -> shifted image -> loss -> backward to A -> all-reduce(mean) of A's grads -> optimizer step."""
    from .reenact import generate_image
    shift = A(shift_vector)
    img = generate_image(G, latent_code, truncation, trunc, w_plus=True, num_layers_shift=num_layers_shift,
                         shift_code=shift, input_is_latent=input_is_latent)
    loss = loss_fn(img)
    A.zero_grad(set_to_none=True)
    loss.backward()
    nbytes = allreduce_mean_grads_(list(A.parameters()))
    optimizer.step()
    return loss.detach(), nbytes

"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: frame sharding, the flat-bucket gradient all-reduce and
replica consistency of the A-matrix step.  The generator itself needs a GPU, so a differentiable stand-in with the same
call signature plays its part here; the GPU path is covered by bench.py --gpus N and the -m gpu tests."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _StandInG(torch.nn.Module):
    """Smooth function of the W+ code with Generator's call signature (n_latent, forward(styles, ...))."""
    n_latent = 6

    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(3)
        self.register_buffer('proj', torch.randn(6 * 16, 3 * 4 * 4, generator=g) * 0.1)

    def forward(self, styles, return_latents=False, truncation=1, truncation_latent=None, input_is_latent=False, **kw):
        w = styles[0]
        if truncation < 1:
            w = truncation_latent + truncation * (w - truncation_latent)
        img = torch.tanh(w[:, :, :16].reshape(w.shape[0], -1) @ self.proj).view(-1, 3, 4, 4)
        return img, (w if return_latents else None)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from stylegan_directions_face_reenactment_b200 import DirectionMatrix, dist as sdist
    r, w = sdist.init(backend='gloo')
    assert (r, w) == (rank, world)
    torch.manual_seed(100 + rank)                       # replicas start DIFFERENT, then get rank 0's A
    A = DirectionMatrix(16, input_dim=15, out_dim=16, w_plus=True, num_layers=4)
    sdist.broadcast_params_(A)
    G = _StandInG()
    opt = torch.optim.Adam(A.parameters(), lr=1e-2, weight_decay=5e-4)
    g = torch.Generator().manual_seed(7)
    n_frames = 10
    dp_all = torch.rand(n_frames, 15, generator=g) * 6 - 3
    w_all = torch.randn(n_frames, 6, 16, generator=g)
    trunc = torch.zeros(1, 16)
    mine = sdist.shard_indices(n_frames, rank, world)
    dp, wsrc = dp_all[mine], w_all[mine]
    # replica of the single-process computation on the concatenated batch
    A_ref = DirectionMatrix(16, input_dim=15, out_dim=16, w_plus=True, num_layers=4)
    A_ref.load_state_dict(A.state_dict())

    def loss_fn(img):
        return (img ** 2).sum() / n_frames * world      # sum over the local shard; mean over ranks == global mean

    for step in range(3):
        loss, nbytes = sdist.train_step(G, A, opt, wsrc, dp, 0.7, trunc, loss_fn, num_layers_shift=4)
        assert nbytes == sum(p.numel() for p in A.parameters()) * 4
    from stylegan_directions_face_reenactment_b200 import generate_image
    opt_ref = torch.optim.Adam(A_ref.parameters(), lr=1e-2, weight_decay=5e-4)
    for step in range(3):
        img = generate_image(G, w_all, 0.7, trunc, num_layers_shift=4, shift_code=A_ref(dp_all), input_is_latent=True)
        l = (img ** 2).sum() / n_frames
        A_ref.zero_grad()
        l.backward()
        opt_ref.step()
    err = max((a - b).abs().max().item() for a, b in zip(A.parameters(), A_ref.parameters()))
    gathered = [None] * world
    dist.all_gather_object(gathered, [p.detach().clone() for p in A.parameters()])
    same = all(torch.equal(gathered[0][i], gathered[k][i]) for k in range(world) for i in range(len(gathered[0])))
    q.put((rank, mine, err, same))
    dist.destroy_process_group()


def test_two_rank_sharded_training_matches_single_process():
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert res[0][1] == [0, 2, 4, 6, 8] and res[1][1] == [1, 3, 5, 7, 9]          # every frame owned exactly once
    for rank, mine, err, same in res:
        assert same, 'replicas diverged'
        assert err < 1e-5, err                                                    # == single-process result (<= 1e-5)


def test_shard_indices_cover_everything():
    from stylegan_directions_face_reenactment_b200.dist import shard_indices
    for n in (0, 1, 7, 512):
        for world in (1, 2, 4, 8):
            got = sorted(sum((shard_indices(n, r, world) for r in range(world)), []))
            assert got == list(range(n))

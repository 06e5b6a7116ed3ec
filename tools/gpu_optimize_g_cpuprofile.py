"""CPU-side profile of the batch-1 optimize_g step (libs/optimization.py:45-68): the step is bound by the host launch rate
(≈240 launches of a few microseconds each), so cProfile's cumulative times are the critical path.
python tools/gpu_optimize_g_cpuprofile.py [B]"""
import cProfile
import copy
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import stylegan2_oracle as orc  # noqa: E402
import stylegan_directions_face_reenactment_b200 as pkg  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
sd = orc.seeded_state_dict(256, 1, seed=0)
G = pkg.Generator(256, 512, 8, channel_multiplier=1)
G.load_state_dict(sd, strict=True)
G = G.cuda()
latent = orc.seeded_wplus(sd, B, G.n_latent, seed=2).cuda()
with torch.no_grad():
    target = copy.deepcopy(G).eval()([latent], input_is_latent=True)[0] * 0.9
G.train()
params = [p for i in range(4, len(G.convs)) for p in G.convs[i].parameters()]
opt = torch.optim.Adam(params, lr=1e-5)


def phases(sync):
    t = [time.perf_counter()]

    def mark():
        if sync:
            torch.cuda.synchronize()
        t.append(time.perf_counter())
    img, _ = G([latent], input_is_latent=True)
    mark()
    loss = (img - target).pow(2).mean()
    mark()
    opt.zero_grad()
    mark()
    loss.backward()
    mark()
    opt.step()
    mark()
    return [b - a for a, b in zip(t, t[1:])]


def plain_step():
    img, _ = G([latent], input_is_latent=True)
    loss = (img - target).pow(2).mean()
    opt.zero_grad()
    loss.backward()
    opt.step()


def event_timed(fn, reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    walls = []
    for _ in range(reps):
        t = time.perf_counter()
        fn()
        walls.append((time.perf_counter() - t) * 1e3)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, (time.perf_counter() - t0) / reps * 1e3, walls


for _ in range(5):
    phases(False)
for reps in (10, 40):
    ev, wall, walls = event_timed(plain_step, reps)
    print('plain step x%d: %.3f ms by CUDA events, %.3f ms wall; per-step CPU ms: %s' % (reps, ev, wall, ' '.join('%.2f' % w for w in walls)))
torch.cuda.synchronize()
for sync in (False, True):
    acc = [0.0] * 5
    n = 20
    t0 = time.perf_counter()
    for _ in range(n):
        for i, v in enumerate(phases(sync)):
            acc[i] += v
    torch.cuda.synchronize()
    tot = (time.perf_counter() - t0) / n * 1e3
    print('sync=%d  step %.3f ms   forward %.3f  loss %.3f  zero_grad %.3f  backward %.3f  adam %.3f  (ms, %s)' % (
        sync, tot, *[a / n * 1e3 for a in acc], 'phase GPU+CPU time' if sync else 'CPU enqueue time'))

pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    phases(False)
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats('cumulative').print_stats(45)

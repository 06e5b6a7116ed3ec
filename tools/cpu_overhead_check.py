import os, sys, time, torch
sys.path.insert(0, '/root/repo')
from oracle import stylegan2_oracle as orc
import stylegan_directions_face_reenactment_b200 as pkg
size, cm = 256, 1
sd = orc.seeded_state_dict(size, cm, seed=0)
G = pkg.Generator(size, 512, 8, channel_multiplier=cm); G.load_state_dict(sd, strict=True); G = G.cuda().eval()
for B in (1, 32):
    w = orc.seeded_wplus(sd, B, G.n_latent, seed=1).cuda()
    with torch.no_grad():
        for _ in range(5): G([w], input_is_latent=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(50): G([w], input_is_latent=True)
        t1 = time.perf_counter()            # enqueue only
        torch.cuda.synchronize()
        t2 = time.perf_counter()
    print('B=%d: CPU enqueue %.3f ms/call, total %.3f ms/call' % (B, (t1 - t0) / 50 * 1e3, (t2 - t0) / 50 * 1e3))
G.enable_cuda_graphs(True)
for B in (1, 4):
    w = orc.seeded_wplus(sd, B, G.n_latent, seed=1).cuda()
    with torch.no_grad():
        for _ in range(5): G([w], input_is_latent=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(50): G([w], input_is_latent=True)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
    print('B=%d with CUDA graph: total %.3f ms/call' % (B, (t2 - t0) / 50 * 1e3))
G.enable_cuda_graphs(False)
import cProfile, pstats
w = orc.seeded_wplus(sd, 1, G.n_latent, seed=1).cuda()
pr = cProfile.Profile()
with torch.no_grad():
    pr.enable()
    for _ in range(50): G([w], input_is_latent=True)
    pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(14)

"""Times sgr_upfirdn2d in the Blur mode (up = down = 1, 4x4 FIR, pad (1,1)) at the bench shape: B*C = 2048 planes of 257^2 -> 256^2
(1.08 GB in + out), reports GB/s against the measured HBM copy peak.  SGR_UPFIRDN_ROWS selects the kernel (0 per-row, 1 row-walking,
2 row-walking separable)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stylegan_directions_face_reenactment_b200.ops import upfirdn2d  # noqa: E402

k1 = torch.tensor([1., 3., 3., 1.])
k = (torch.outer(k1, k1) / 64 * 4).cuda()
x = torch.randn(2048, 1, 257, 257, device='cuda')
for _ in range(3):
    y = upfirdn2d(x, k, pad=(1, 1))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    y = upfirdn2d(x, k, pad=(1, 1))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
nbytes = (x.numel() + y.numel()) * 4
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:
    peak = 6650.0
print('upfirdn2d blur mode rows=%s: %.3f ms, %.0f GB/s algorithmic = %.2f of the measured copy peak (%.0f GB/s)' % (
    os.environ.get('SGR_UPFIRDN_ROWS', 'default'), ms, nbytes / ms / 1e6, nbytes / ms / 1e6 / peak, peak))

"""Loss-head side of the A-matrix train step (BASELINE configs[3]) — the step AFTER the generator path.

The reference evaluates three heads on every shifted / source / target image (libs/utilities/utils_train.py:376-433):
  * identity:   1 - cos(ArcFace(shifted), ArcFace(source)); ArcFace = IR-SE-50 (`Backbone(112, 50, 'ir_se')`,
                libs/criteria/model_irse.py:9-48, helpers.py:24-119) on the [35:223, 32:220] crop pooled to 112^2
                (libs/criteria/id_loss.py:20-34);
  * perceptual: LPIPS with AlexNet features, taps after layers 2/5/8/10/12, unit-normalised activations, 1x1 "lin" layers
                (libs/criteria/lpips/lpips.py:28-34, networks.py:24-98);
  * shape:      DECA's ResNet50 coefficient regressor (libs/DECA/decalib/models/encoders.py:22-40: resnet50 -> 2048 ->
                1024 -> 236) on 224^2 images in [0,1], run by the reference in a PER-SAMPLE python loop
                (libs/DECA/estimate_DECA.py:30-53).

Their trained weights (model_ir_se50.pth, LPIPS lin layers, deca_model.tar) and DECA's FLAME / face-alignment dependencies
do not exist offline, so these are the SURROGATE heads SURVEY.md §8d (cfg 4) specifies: the reference's architectures with
seeded random weights, identical on every rank.  What is re-designed here is how they run next to the sm_100a generator:
BATCHED (no per-sample loop), channels-last, bf16 autocast, one forward over the concatenated [shifted; source/target]
batch per head.  They are PyTorch/cuDNN modules (the hot path this repository rewrites is the generator; north_star keeps
"the e4e/DECA calls in PyTorch"); dL/dimage flows from them into sgr_synthesis_backward.
"""
import torch
from torch import nn
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------------- ArcFace IR-SE-50
class _SE(nn.Module):
    def __init__(self, ch, reduction=16):
        super().__init__()
        self.fc1 = nn.Conv2d(ch, ch // reduction, 1, bias=False)
        self.fc2 = nn.Conv2d(ch // reduction, ch, 1, bias=False)

    def forward(self, x):
        w = torch.sigmoid(self.fc2(F.relu(self.fc1(x.mean((2, 3), keepdim=True)))))
        return x * w


class _IRSEUnit(nn.Module):
    """BN -> 3x3 conv -> PReLU -> 3x3 conv (stride) -> BN -> SE, plus shortcut (helpers.py:96-119)."""

    def __init__(self, cin, depth, stride):
        super().__init__()
        if cin == depth:
            self.shortcut = nn.MaxPool2d(1, stride)
        else:
            self.shortcut = nn.Sequential(nn.Conv2d(cin, depth, 1, stride, bias=False), nn.BatchNorm2d(depth))
        self.res = nn.Sequential(nn.BatchNorm2d(cin), nn.Conv2d(cin, depth, 3, 1, 1, bias=False), nn.PReLU(depth),
                                 nn.Conv2d(depth, depth, 3, stride, 1, bias=False), nn.BatchNorm2d(depth), _SE(depth))

    def forward(self, x):
        return self.res(x) + self.shortcut(x)


class ArcFaceIRSE50(nn.Module):
    """`Backbone(input_size=112, num_layers=50, mode='ir_se')`: units (64x3, 128x4, 256x14, 512x3), each stage stride 2."""

    def __init__(self):
        super().__init__()
        self.input_layer = nn.Sequential(nn.Conv2d(3, 64, 3, 1, 1, bias=False), nn.BatchNorm2d(64), nn.PReLU(64))
        units, cin = [], 64
        for depth, n in ((64, 3), (128, 4), (256, 14), (512, 3)):
            for i in range(n):
                units.append(_IRSEUnit(cin, depth, 2 if i == 0 else 1))
                cin = depth
        self.body = nn.Sequential(*units)
        self.output_layer = nn.Sequential(nn.BatchNorm2d(512), nn.Flatten(), nn.Linear(512 * 7 * 7, 512), nn.BatchNorm1d(512))

    def forward(self, x):
        x = self.output_layer(self.body(self.input_layer(x)))
        return x / x.norm(dim=1, keepdim=True)

    @staticmethod
    def crop(images):
        """id_loss.py:20-26: fixed face crop of a 256^2 frame, pooled to 112^2."""
        if images.shape[2] != 256:
            images = F.adaptive_avg_pool2d(images, (256, 256))
        return F.adaptive_avg_pool2d(images[:, :, 35:223, 32:220], (112, 112))


# ----------------------------------------------------------------------------------------------------- LPIPS (AlexNet)
class LPIPSAlex(nn.Module):
    TAPS = (2, 5, 8, 10, 12)
    CHANNELS = (64, 192, 384, 256, 256)

    def __init__(self):
        super().__init__()
        from torchvision import models
        self.features = models.alexnet(weights=None).features
        self.lin = nn.ModuleList([nn.Conv2d(c, 1, 1, bias=False) for c in self.CHANNELS])
        for l in self.lin:
            nn.init.constant_(l.weight, 1.0 / l.weight.shape[1])        # "unit lin layers" (SURVEY.md §8d cfg 4)
        self.register_buffer('mean', torch.tensor([-.030, -.088, -.188]).view(1, 3, 1, 1))
        self.register_buffer('std', torch.tensor([.458, .448, .450]).view(1, 3, 1, 1))

    def taps(self, x):
        x = (x - self.mean) / self.std
        out = []
        for i, layer in enumerate(self.features, 1):
            x = layer(x)
            if i in self.TAPS:
                out.append(x / (x.float().pow(2).sum(1, keepdim=True).sqrt() + 1e-10).to(x.dtype))
            if len(out) == len(self.TAPS):
                break
        return out

    def forward(self, x, y):
        """lpips.py:28-34 with both images in ONE batched feature pass."""
        n = x.shape[0]
        feats = self.taps(torch.cat([x, y], 0))
        total = 0.
        for f, lin in zip(feats, self.lin):
            d = (f[:n] - f[n:]).float().pow(2)
            total = total + lin(d).mean((2, 3)).sum()
        return total / n


# ----------------------------------------------------------------------------------------------------- shape regressor
class ShapeRegressor(nn.Module):
    """Stand-in for DECA's E_flame: resnet50 trunk -> 2048 -> 1024 -> 236 coefficients, on 224^2 images in [0,1]."""

    def __init__(self, n_out=236):
        super().__init__()
        from torchvision import models
        trunk = models.resnet50(weights=None)
        trunk.fc = nn.Identity()
        self.trunk = trunk
        self.head = nn.Sequential(nn.Linear(2048, 1024), nn.ReLU(), nn.Linear(1024, n_out))

    def forward(self, images):
        """images in the generator's [-1,1] range; the reference rescales to [0,255] then /255 (image_utils.py:87-95)."""
        x = (images.clamp(-1, 1) + 1) / (2 + 1e-5)
        if x.shape[2] != 224:
            x = F.interpolate(x, size=(224, 224), mode='bilinear', align_corners=False)
        return self.head(self.trunk(x))


# ----------------------------------------------------------------------------------------------------- the three together
class SurrogateLossHeads(nn.Module):
    """id + perceptual + shape losses of libs/utilities/utils_train.py:376-433 on (shifted, source, target) batches.

    forward(shifted, source, target) -> (loss, dict).  `shifted` carries the graph back to the generator; source / target
    are constants.  Every head runs ONCE on the concatenated batch (channels-last, bf16 autocast unless `amp=False`)."""

    def __init__(self, lambda_identity=10.0, lambda_perceptual=10.0, lambda_shape=1.0, seed=1234, amp=True):
        super().__init__()
        with torch.random.fork_rng(devices=[]):
            torch.manual_seed(seed)
            self.arcface = ArcFaceIRSE50()
            self.lpips = LPIPSAlex()
            self.shape = ShapeRegressor()
        self.lam = (lambda_identity, lambda_perceptual, lambda_shape)
        self.amp = amp
        self.eval()
        for p in self.parameters():
            p.requires_grad_(False)

    def prepare(self, device):
        return self.to(device).to(memory_format=torch.channels_last)

    def forward(self, shifted, source, target):
        n = shifted.shape[0]
        src, tgt = source.detach(), target.detach()
        with torch.autocast('cuda', dtype=torch.bfloat16, enabled=self.amp and shifted.is_cuda):
            both = torch.cat([shifted, src], 0).contiguous(memory_format=torch.channels_last)
            e = self.arcface(ArcFaceIRSE50.crop(both)).float()
            loss_id = (1 - F.cosine_similarity(e[:n], e[n:].detach(), dim=1, eps=1e-6)).mean()
            loss_lp = self.lpips(both[:n], both[n:])
            coef = self.shape(torch.cat([shifted, tgt], 0).contiguous(memory_format=torch.channels_last)).float()
            loss_sh = (coef[:n] - coef[n:].detach()).abs().mean()
        loss = self.lam[0] * loss_id + self.lam[1] * loss_lp + self.lam[2] * loss_sh
        return loss, {'loss_identity': loss_id.detach(), 'loss_perceptual': loss_lp.detach(), 'loss_shape': loss_sh.detach()}

"""StyleGAN2 generator behind the reference's class API, executed by libsgr.so (sm_100a).

Mirrors the public surface of libs/gan/StyleGAN2/model.py of the reference — Generator(size, style_dim, n_mlp,
channel_multiplier, blur_kernel, lr_mlp), .forward(styles, ...) -> (image, latent|None), .mean_latent, .get_latent,
.make_noise, .n_latent, .num_layers, .convs / .to_rgbs / .noises / .style / .input, and a state_dict with the same
keys and shapes (model.py:361-539; key list in SURVEY.md §8b) — so reference checkpoints load unchanged and
run_inference.py / run_trainer.py can import this class through the namespace overlay in overlay/libs.

The arithmetic is NOT the reference's: the per-sample modulation is moved from the weights to the activations
(y = d[b,o] * conv(x * s[b,i], W)); the transposed conv of an upsampling layer runs in scatter form (its 9
real taps as one tcgen05 GEMM onto four parity planes, csrc/modconv_scatter_sm100.cu) and the separable 4x4 blur + noise + bias
+ leaky-relu are applied on the way into the next convolution (fused producer warps or the HBM-bound up_finish_kernel; the
polyphase folding W (*) fir of SURVEY §9.2 survives only for non-separable blur kernels); a plain conv + noise + bias +
leaky-relu (+ the next ToRGB) is one tcgen05 kernel per layer (csrc/modconv_halo_sm100.cu).  The fp32 nn.Parameters stay the
source of truth; packed bf16 hi/lo weights are a cache keyed on each parameter's version counter.
"""
import ctypes as C
import math

import torch
from torch import nn
from torch.nn import functional as F

from . import _native as N
from .ops import FusedLeakyReLU, fused_leaky_relu, upfirdn2d

SQRT2 = math.sqrt(2.0)


def make_kernel(k):
    """Normalised 2-D FIR from 1-D (or 2-D) taps (reference model.py:19-27)."""
    k = torch.as_tensor(k, dtype=torch.float32)
    if k.ndim == 1:
        k = torch.outer(k, k)
    return k / k.sum()


class PixelNorm(nn.Module):
    def forward(self, input):
        return input * torch.rsqrt(torch.mean(input * input, dim=1, keepdim=True) + 1e-8)


class Upsample(nn.Module):
    """2x FIR upsampling of the RGB skip (reference model.py:30-48)."""

    def __init__(self, kernel, factor=2):
        super().__init__()
        self.factor = factor
        k = make_kernel(kernel) * (factor ** 2)
        self.register_buffer('kernel', k)
        p = k.shape[0] - factor
        self.pad = ((p + 1) // 2 + factor - 1, p // 2)

    def forward(self, input):
        return upfirdn2d(input, self.kernel, up=self.factor, down=1, pad=self.pad)


class Blur(nn.Module):
    """FIR applied after the transposed conv of an upsampling layer (reference model.py:72-88)."""

    def __init__(self, kernel, pad, upsample_factor=1):
        super().__init__()
        k = make_kernel(kernel)
        if upsample_factor > 1:
            k = k * (upsample_factor ** 2)
        self.register_buffer('kernel', k)
        self.pad = pad

    def forward(self, input):
        return upfirdn2d(input, self.kernel, pad=self.pad)


class EqualLinear(nn.Module):
    """Equalised-lr linear layer (reference model.py:129-162); imported by the e4e encoder too."""

    def __init__(self, in_dim, out_dim, bias=True, bias_init=0, lr_mul=1, activation=None):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_dim, in_dim).div_(lr_mul))
        self.bias = nn.Parameter(torch.zeros(out_dim).fill_(bias_init)) if bias else None
        self.activation = activation
        self.scale = (1 / math.sqrt(in_dim)) * lr_mul
        self.lr_mul = lr_mul

    def forward(self, input):
        if self.activation:
            out = F.linear(input, self.weight * self.scale)
            return fused_leaky_relu(out, self.bias * self.lr_mul)
        return F.linear(input, self.weight * self.scale, bias=None if self.bias is None else self.bias * self.lr_mul)

    def __repr__(self):
        return '%s(%d, %d)' % (self.__class__.__name__, self.weight.shape[1], self.weight.shape[0])


class ConstantInput(nn.Module):
    def __init__(self, channel, size=4):
        super().__init__()
        self.input = nn.Parameter(torch.randn(1, channel, size, size))

    def forward(self, input):
        return self.input.repeat(input.shape[0], 1, 1, 1)


class NoiseInjection(nn.Module):
    def __init__(self):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(1))

    def forward(self, image, noise=None):
        if noise is None:
            b, _, h, w = image.shape
            noise = image.new_empty(b, 1, h, w).normal_()
        return image + self.weight * noise


def _version_key(*tensors):
    return tuple((t.data_ptr(), t._version, t.device.index) for t in tensors)


class ModulatedConv2d(nn.Module):
    """Parameters of one modulated convolution (reference model.py:177-273) + the packed-weight cache."""

    def __init__(self, in_channel, out_channel, kernel_size, style_dim, demodulate=True, upsample=False,
                 downsample=False, blur_kernel=(1, 3, 3, 1)):
        super().__init__()
        if downsample:
            raise NotImplementedError('downsampling ModulatedConv2d is discriminator-only (outside the generator path)')
        self.eps = 1e-8
        self.kernel_size = kernel_size
        self.in_channel = in_channel
        self.out_channel = out_channel
        self.upsample = upsample
        self.downsample = downsample
        if upsample:
            factor = 2
            p = (len(blur_kernel) - factor) - (kernel_size - 1)
            self.blur = Blur(blur_kernel, pad=((p + 1) // 2 + factor - 1, p // 2 + 1), upsample_factor=factor)
        self.scale = 1 / math.sqrt(in_channel * kernel_size ** 2)
        self.padding = kernel_size // 2
        self.weight = nn.Parameter(torch.randn(1, out_channel, in_channel, kernel_size, kernel_size))
        self.modulation = EqualLinear(style_dim, in_channel, bias_init=1)
        self.demodulate = demodulate
        self._pack_cache = {}

    def __repr__(self):
        return '%s(%d, %d, %d, upsample=%s)' % (self.__class__.__name__, self.in_channel, self.out_channel,
                                                 self.kernel_size, self.upsample)

    def __deepcopy__(self, memo):          # the cache is derived data; never copy it (optimize_g deep-copies G)
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        import copy
        for k, v in self.__dict__.items():
            new.__dict__[k] = {} if k in ('_pack_cache', '_fill_plans') else copy.deepcopy(v, memo)
        return new

    def up_mode(self):
        """Packing / execution mode of an upsampling layer's forward operator: 2 (scatter conv + separable FIR pass) when
        blur.kernel is an outer product (always true for kernels built by make_kernel from 1-D taps, reference
        model.py:19-27), else 1 (polyphase folding, which takes any 4x4 FIR).  Checked once per buffer version."""
        if not self.upsample:
            return 0
        k = self.blur.kernel
        key = _version_key(k)
        cache = self.__dict__.setdefault('_fir_cache', {})       # survives weight updates (only blur.kernel matters)
        if cache.get('fir_key') != key:
            kc = k.detach().double().cpu()
            tot = kc.sum()
            sep = torch.outer(kc.sum(1), kc.sum(0)) / tot if tot != 0 else torch.zeros_like(kc)
            rank1 = tuple(kc.shape) == (4, 4) and bool((kc - sep).abs().max() <= 1e-6 * kc.abs().max())
            cache['fir_key'] = key
            cache['fir_mode'] = 2 if rank1 else 1
        return cache['fir_mode']

    def packed(self, transpose=False, fmt=None, nt=0):
        """(w_packed, wsq) device tensors for the tcgen05 kernel; refilled when the parameter changes.
        The adjoint (transpose) is always packed as bf16 hi/lo: it multiplies gradients (csrc/backward.cu).
        `nt` is the GEMM column tile the layout is built for (0 = library default, see sgr_choose_column_tile).
        An in-place update of the weight (an optimizer step: same storage, new version) refills every cached variant in its
        existing buffer, so the C descriptors and captured graphs that point at them stay valid (optimize_g,
        libs/optimization.py:45-68: one Adam step per forward); a new storage drops the cache."""
        w = self.weight
        fir = self.blur.kernel if self.upsample else None
        fmt = N.FMT_BF16 if transpose else (N.default_format() if fmt is None else fmt)
        version = _version_key(*([w] + ([fir] if fir is not None else [])))
        cache = self._pack_cache
        old = cache.get('version')
        if old != version:
            in_place = (old is not None and len(old) == len(version) and old[1:] == version[1:] and
                        old[0][0] == version[0][0] and old[0][2] == version[0][2] and len(cache) <= 7)
            # (only the weight's version moved; a cache that has collected many layouts is dropped instead of refilled)
            if in_place:
                cache['version'] = version
                self._refill_in_place(cache)
            else:                                                # new storage / new FIR: every packed variant is stale
                cache.clear()
                cache['version'] = version
        # upsampling layers: forward operator in scatter form (up=2: 9 real taps + FIR pass), adjoint in polyphase form
        # (up=2: forward = 9-tap scatter conv + FIR pass, adjoint = FIR^T to parity planes + 9-tap gather conv; up=1: both
        # in polyphase form, for a blur kernel that is not an outer product)
        up_mode = self.up_mode()
        if up_mode == 2 and not transpose:
            nt = 0                                               # fixed by the scatter layout
        slot = (bool(transpose), fmt, nt, up_mode)
        hit = cache.get(slot)
        if hit is not None:
            return hit
        cout, cin = self._packed_dims(transpose)
        nbytes = N.lib().sgr_packed_weight_bytes(cout, cin, self.kernel_size, up_mode, int(transpose))
        packed = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
        wsq = torch.empty(cout, cin, dtype=torch.float32, device=w.device) if not transpose else None
        self._fill_packed(slot, packed, wsq)
        cache[slot] = (packed, wsq)
        return packed, wsq

    def _packed_dims(self, transpose):
        cout, cin = self.out_channel, self.in_channel
        if not transpose and cout < 32:                          # ToRGB (3 channels): the GEMM columns are padded to 32
            cout = 32
        if cin % 32:                                             # zero input channels up to the K granularity
            cin += 32 - cin % 32
        return cout, cin

    def _refill_in_place(self, cache):
        """Refill every cached packed variant after an in-place weight update (one optimizer step of optimize_g).  The batch-1
        fine-tuning step is bound by host time, so the argument tuple of each sgr_pack_modconv_weight call is recorded the first
        time (`_fill_plans`: valid while the weight storage, the FIR buffer and the packed buffers stay where they are) and the
        refill is then one ctypes call per variant, with one device guard and one stream lookup for all of them."""
        plans = self.__dict__.setdefault('_fill_plans', {})
        fir = self.blur.kernel if self.upsample else None
        here = (self.weight.data_ptr(), None if fir is None else fir.data_ptr())
        with torch.cuda.device(self.weight.device):
            st = N.stream()
            lib = N.lib()
            for slot, (packed, wsq) in [(k, v) for k, v in cache.items() if k != 'version']:
                plan = plans.get(slot)
                if plan is not None and plan[0] == here and plan[1] == (packed.data_ptr(), None if wsq is None else wsq.data_ptr()):
                    N.check(lib.sgr_pack_modconv_weight(*plan[2], st), 'sgr_pack_modconv_weight')
                else:
                    plans.pop(slot, None)
                    self._fill_packed(slot, packed, wsq)

    def _fill_packed(self, slot, packed, wsq):
        transpose, fmt, nt, up_mode = slot
        cout, cin, ks = self.out_channel, self.in_channel, self.kernel_size
        wd = self.weight.detach()
        if not transpose and cout < 32:
            wd = torch.cat([wd[0], wd.new_zeros(32 - cout, cin, ks, ks)], 0)
            cout = 32
        wd = wd.reshape(cout, cin, ks, ks)
        if cin % 32:
            wd = torch.cat([wd, wd.new_zeros(cout, 32 - cin % 32, ks, ks)], 1) * math.sqrt((cin + 32 - cin % 32) / cin)
            cin = wd.shape[1]                                    # (the factor undoes the 1/sqrt(cin k^2) of the padded cin)
        if wd.dtype != torch.float32 or not wd.is_contiguous():
            wd = wd.contiguous().float()
        fir = self.blur.kernel if self.upsample else None
        firc = None if fir is None else fir.detach().contiguous().float()
        args = (N.ptr(wd), N.ptr(firc), cout, cin, ks, up_mode, int(transpose), fmt, nt, N.ptr(packed), N.ptr(wsq))
        with torch.cuda.device(wd.device):
            N.check(N.lib().sgr_pack_modconv_weight(*args, N.stream()), 'sgr_pack_modconv_weight')
        # the operands are the parameter / buffer storages themselves (no padded or converted copy): the same call refills the
        # variant after an in-place update (_refill_in_place)
        if wd.data_ptr() == self.weight.data_ptr() and (firc is None or firc.data_ptr() == fir.data_ptr()):
            self.__dict__.setdefault('_fill_plans', {})[slot] = (
                (self.weight.data_ptr(), None if fir is None else fir.data_ptr()),
                (packed.data_ptr(), None if wsq is None else wsq.data_ptr()), args)

    def forward(self, input, style):
        """Module-level call on NCHW fp32 tensors (the fused Generator.forward never goes through here)."""
        return _modconv_module_forward(self, input, style, noise=None, noise_weight=None, bias=None, act=False)


def _modconv_module_forward(conv, x, style, noise, noise_weight, bias, act):
    if not x.is_cuda:
        raise RuntimeError('ModulatedConv2d: input must be a CUDA tensor (no CPU fallback)')
    if torch.is_grad_enabled():
        # the layer-by-layer entry points launch raw kernels: there is no autograd node behind them (Generator.forward is the
        # differentiable path).  Never return a tensor that silently drops a gradient somebody asked for.
        if x.requires_grad or style.requires_grad:
            raise RuntimeError('module-level ModulatedConv2d / StyledConv / ToRGB calls are not differentiable (input or style '
                               'requires grad): run the whole network through Generator.forward, or call under torch.no_grad()')
        if any(p.requires_grad for p in conv.parameters()) and not conv.__dict__.get('_warned_nograd'):
            import warnings
            warnings.warn('module-level ModulatedConv2d / StyledConv / ToRGB calls do not record autograd history: parameter '
                          'gradients are only available through Generator.forward', stacklevel=3)
            conv.__dict__['_warned_nograd'] = True
    with torch.cuda.device(x.device):
        return _modconv_module_forward_impl(conv, x, style, noise, noise_weight, bias, act)


def _modconv_module_forward_impl(conv, x, style, noise, noise_weight, bias, act):
    lib = N.lib()
    st = N.stream()
    x = x.contiguous().float()
    style = style.contiguous().float()
    b, cin, h, w = x.shape
    dev = x.device
    mw = conv.modulation.weight.detach().contiguous()
    mb = conv.modulation.bias.detach().contiguous()
    s = torch.empty(b, cin, device=dev)
    N.check(lib.sgr_style_affine(N.ptr(style), style.shape[1], b, N.ptr(mw), N.ptr(mb), cin, N.ptr(s), st),
            'sgr_style_affine')
    fmt = N.default_format()
    packed, wsq = conv.packed(fmt=fmt)
    cout_k = wsq.shape[0]
    if wsq.shape[1] != cin:                                  # channel padding (see packed())
        padc = wsq.shape[1] - cin
        x = torch.cat([x, x.new_zeros(b, padc, h, w)], 1)
        s = torch.cat([s, s.new_zeros(b, padc)], 1)
        cin = wsq.shape[1]
    d = None
    if conv.demodulate:
        d = torch.empty(b, cout_k, device=dev)
        N.check(lib.sgr_demod(N.ptr(s), N.ptr(wsq), b, cin, cout_k, N.ptr(d), st), 'sgr_demod')
    xc8 = torch.empty(2 * b * cin * h * w, dtype=torch.bfloat16, device=dev)
    N.check(lib.sgr_nchw_to_c8(N.ptr(x), N.ptr(s), N.ptr(xc8), b, cin, h, w, 0, fmt, st), 'sgr_nchw_to_c8')
    ho, wo = (2 * h, 2 * w) if conv.upsample else (h, w)
    out = torch.empty(b, cout_k, ho, wo, device=dev)
    a = N.ConvArgs()
    a.batch, a.cin, a.cout, a.h_in, a.w_in = b, cin, cout_k, h, w
    a.ksize, a.up, a.act = conv.kernel_size, conv.up_mode(), int(act)
    a.single_pass = N.single_pass()
    if conv.up_mode() == 2:
        if d is None:
            d = torch.ones(b, cout_k, device=dev)
        scratch = torch.empty(lib.sgr_up_scratch_bytes(b, cout_k, h, w), dtype=torch.uint8, device=dev)
        firk = conv.blur.kernel.detach().contiguous().float()
        a.t_scratch, a.fir = N.ptr(scratch), N.ptr(firk)
    a.act_gain = SQRT2 if act else 1.0
    a.operand_format = a.out_format = fmt
    a.x_c8, a.w_packed, a.demod = N.ptr(xc8), N.ptr(packed), N.ptr(d)
    keep = (scratch, firk) if conv.up_mode() == 2 else None     # noqa: F841  (alive until the launches are enqueued)
    if bias is not None:
        bias = bias.detach().contiguous().float()
        a.bias = N.ptr(bias)
    if noise is not None:
        noise = noise.contiguous().float()
        nwt = noise_weight.detach().contiguous().float()
        a.noise, a.noise_weight = N.ptr(noise), N.ptr(nwt)
        a.noise_batch_stride = ho * wo if noise.shape[0] == b and b > 1 else 0
    a.out_f32 = N.ptr(out)
    N.check(lib.sgr_modconv_forward(C.byref(a), st), 'sgr_modconv_forward')
    return out[:, :conv.out_channel] if cout_k != conv.out_channel else out


class StyledConv(nn.Module):
    """conv -> +noise -> +bias -> leaky-relu*sqrt2 (reference model.py:303-337), one kernel."""

    def __init__(self, in_channel, out_channel, kernel_size, style_dim, upsample=False, blur_kernel=(1, 3, 3, 1),
                 demodulate=True):
        super().__init__()
        self.conv = ModulatedConv2d(in_channel, out_channel, kernel_size, style_dim, upsample=upsample,
                                    blur_kernel=blur_kernel, demodulate=demodulate)
        self.noise = NoiseInjection()
        self.activate = FusedLeakyReLU(out_channel)

    def forward(self, input, style, noise=None):
        if noise is None:
            b, _, h, w = input.shape
            f = 2 if self.conv.upsample else 1
            noise = input.new_empty(b, 1, h * f, w * f).normal_()
        return _modconv_module_forward(self.conv, input, style, noise, self.noise.weight, self.activate.bias, True)


class ToRGB(nn.Module):
    """1x1 modulated conv without demodulation + bias + upsampled skip (reference model.py:340-359)."""

    def __init__(self, in_channel, style_dim, upsample=True, blur_kernel=(1, 3, 3, 1)):
        super().__init__()
        if upsample:
            self.upsample = Upsample(blur_kernel)
        self.conv = ModulatedConv2d(in_channel, 3, 1, style_dim, demodulate=False)
        self.bias = nn.Parameter(torch.zeros(1, 3, 1, 1))

    def forward(self, input, style, skip=None):
        out = self.conv(input, style) + self.bias
        if skip is not None:
            out = out + self.upsample(skip)
        return out


class Generator(nn.Module):
    def __init__(self, size, style_dim, n_mlp, channel_multiplier=2, blur_kernel=(1, 3, 3, 1), lr_mlp=0.01):
        super().__init__()
        if style_dim != N.STYLE_DIM:
            raise ValueError('libsgr is built for style_dim == %d' % N.STYLE_DIM)
        self.size = size
        self.style_dim = style_dim
        layers = [PixelNorm()]
        for _ in range(n_mlp):
            layers.append(EqualLinear(style_dim, style_dim, lr_mul=lr_mlp, activation='fused_lrelu'))
        self.style = nn.Sequential(*layers)
        self.channels = {4: 512, 8: 512, 16: 512, 32: 512, 64: 256 * channel_multiplier,
                         128: 128 * channel_multiplier, 256: 64 * channel_multiplier, 512: 32 * channel_multiplier,
                         1024: 16 * channel_multiplier}
        self.input = ConstantInput(self.channels[4])
        self.conv1 = StyledConv(self.channels[4], self.channels[4], 3, style_dim, blur_kernel=blur_kernel)
        self.to_rgb1 = ToRGB(self.channels[4], style_dim, upsample=False)
        self.log_size = int(math.log(size, 2))
        self.num_layers = (self.log_size - 2) * 2 + 1
        self.convs = nn.ModuleList()
        self.upsamples = nn.ModuleList()
        self.to_rgbs = nn.ModuleList()
        self.noises = nn.Module()
        for layer_idx in range(self.num_layers):
            res = (layer_idx + 5) // 2
            self.noises.register_buffer('noise_%d' % layer_idx, torch.randn(1, 1, 2 ** res, 2 ** res))
        in_channel = self.channels[4]
        for i in range(3, self.log_size + 1):
            out_channel = self.channels[2 ** i]
            self.convs.append(StyledConv(in_channel, out_channel, 3, style_dim, upsample=True, blur_kernel=blur_kernel))
            self.convs.append(StyledConv(out_channel, out_channel, 3, style_dim, blur_kernel=blur_kernel))
            self.to_rgbs.append(ToRGB(out_channel, style_dim))
            in_channel = out_channel
        self.n_latent = self.log_size * 2 - 2
        self._workspace = {}

    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = {} if k in ('_workspace', '_desc_cache') else copy.deepcopy(v, memo)
        return new

    # ------------------------------------------------------------------ reference API
    def make_noise(self):
        device = self.input.input.device
        noises = [torch.randn(1, 1, 4, 4, device=device)]
        for i in range(3, self.log_size + 1):
            for _ in range(2):
                noises.append(torch.randn(1, 1, 2 ** i, 2 ** i, device=device))
        return noises

    def mean_latent(self, n_latent):
        latent_in = torch.randn(n_latent, self.style_dim, device=self.input.input.device)
        return self.style(latent_in).mean(0, keepdim=True)

    def get_latent(self, input):
        return self.style(input)

    def forward(self, styles, return_latents=False, return_features=False, inject_index=None, truncation=1,
                truncation_latent=None, input_is_latent=False, noise=None, randomize_noise=False):
        if not input_is_latent:
            styles = [self.style(s) for s in styles]
        if noise is None:
            if randomize_noise:
                noise = [None] * self.num_layers
            else:
                noise = [getattr(self.noises, 'noise_%d' % i) for i in range(self.num_layers)]
        if truncation < 1:
            styles = [truncation_latent + truncation * (s - truncation_latent) for s in styles]
        if len(styles) < 2:
            latent = styles[0]
            if latent.ndim < 3:
                latent = latent.unsqueeze(1).repeat(1, self.n_latent, 1)
        else:
            if inject_index is None:
                import random
                inject_index = random.randint(1, self.n_latent - 1)
            latent = torch.cat([styles[0].unsqueeze(1).repeat(1, inject_index, 1),
                                styles[1].unsqueeze(1).repeat(1, self.n_latent - inject_index, 1)], 1)
        image = self.synthesis(latent, noise)
        return (image, latent) if return_latents else (image, None)

    def invalidate_cache(self):
        """Drop every derived object (packed tensor-core weights, cached C descriptors, captured CUDA graphs, workspaces).  The
        caches are keyed on each parameter's (data_ptr, _version), which optimizers, `copy_`, `load_state_dict` and `.to()` all
        change; a write through `param.data` does not — call this after one."""
        for layer in self.styled_layers() + self.rgb_layers():
            layer.conv._pack_cache.clear()
            layer.conv.__dict__.pop('_fill_plans', None)
        self.__dict__.pop('_desc_cache', None)
        self._workspace.clear()
        return self

    def enable_cuda_graphs(self, on=True):
        """Replay the synthesis kernels of no-grad forward calls as one captured CUDA graph per (batch, weights, noise)
        configuration (also: SGR_CUDA_GRAPHS=1).  Worth it for small batches, where the ~36 launches per frame are bound by
        the CPU launch rate; weights / noise changes are detected and re-captured.  Off by default."""
        self._cuda_graphs = bool(on)
        return self

    # ------------------------------------------------------------------ fused synthesis
    def styled_layers(self):
        return [self.conv1] + list(self.convs)

    def rgb_layers(self):
        return [self.to_rgb1] + list(self.to_rgbs)

    def synthesis_uint8(self, latent, noise=None, size=None):
        """latent [B, n_latent, 512] -> uint8 HWC frames [B,size,size,3] (size divides self.size; default self.size): the
        callers' output stage (256-pooling, clamp, scale, uint8; reference generic.py:146-148, image_utils.py:97-111) fused
        into the last ToRGB tail.  No autograd."""
        from .synthesis import synthesis_forward_u8
        if noise is None:
            noise = [getattr(self.noises, 'noise_%d' % i) for i in range(self.num_layers)]
        return synthesis_forward_u8(self, latent.detach(), noise, size)

    def synthesis(self, latent, noise=None, return_features=False):
        """latent [B, n_latent, 512] -> image [B,3,size,size]; optionally also every StyledConv output (NCHW fp32)."""
        from .synthesis import run_synthesis
        if noise is None:
            noise = [getattr(self.noises, 'noise_%d' % i) for i in range(self.num_layers)]
        return run_synthesis(self, latent, noise, return_features)

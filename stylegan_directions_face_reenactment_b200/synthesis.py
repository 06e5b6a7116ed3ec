"""Whole-network execution: fills the C `sgr_synthesis` descriptor from a Generator's parameters and calls
sgr_synthesis_forward (one C call = all kernel launches of Generator.forward's synthesis part, reference
model.py:519-534)."""
import ctypes as C

import torch

from . import _native as N


def _f32c(t):
    t = t.detach()
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.contiguous().float()
    return t


class NetDescriptor:
    """Keeps the ctypes struct and every tensor it points to alive for the duration of a call."""

    def __init__(self, g, noise, batch, backward=False):
        self.keep = []
        self.noise_used = []
        self.copied = False                            # a tensor had to be converted: the struct points at a snapshot
        self.pack_calls = []                           # (conv, kwargs, struct field, pointer) of every packed operand
        s = N.Synthesis()
        styled = g.styled_layers()
        rgbs = g.rgb_layers()
        s.size, s.n_styled, s.n_rgb, s.n_latent = g.size, len(styled), len(rgbs), g.n_latent
        s.format = N.default_format()
        s.single_pass = 0 if backward else N.single_pass()
        s.const_input = self._p(g.input.input)
        for l, layer in enumerate(styled):
            conv = layer.conv
            d = s.styled[l]
            d.cin, d.cout, d.up = conv.in_channel, conv.out_channel, conv.up_mode()
            res_out = 4 << ((l + 1) // 2)
            res_in = res_out // 2 if conv.upsample else res_out
            if d.up == 2:                              # scatter layout: fixed column tile; FIR pass needs the blur taps
                d.column_tile = 0
                d.fir = self._p(conv.blur.kernel)
            else:                                      # plain layers, or polyphase fallback for a non-separable FIR
                d.column_tile = N.lib().sgr_choose_column_tile(batch, res_in, res_in, conv.out_channel * (4 if d.up else 1))
            kw = dict(fmt=s.format, nt=d.column_tile)
            packed, wsq = conv.packed(**kw)
            d.latent_row = 0 if l == 0 else l          # conv1 <- row 0, convs[j] <- row j+1 (model.py:520-531)
            d.w_packed, d.wsq = self._p(packed), self._p(wsq)
            self.pack_calls.append((conv, kw, packed.data_ptr()))
            if backward:
                d.column_tile_t = N.lib().sgr_choose_column_tile(batch, res_in, res_in, conv.in_channel)
                kw = dict(transpose=True, nt=d.column_tile_t)
                packed_t = conv.packed(**kw)[0]
                d.w_packed_t = self._p(packed_t)
                self.pack_calls.append((conv, kw, packed_t.data_ptr()))
            d.mod_weight, d.mod_bias = self._p(conv.modulation.weight), self._p(conv.modulation.bias)
            nz = noise[l]
            res = 4 << ((l + 1) // 2)
            if nz is None:                             # randomize_noise: fresh per-sample noise (model.py:282-285)
                nz = torch.randn(batch, 1, res, res, device=g.input.input.device)
            if nz.shape[-1] != res or nz.shape[-2] != res:
                raise RuntimeError('noise %d has shape %s, expected [*,1,%d,%d]' % (l, tuple(nz.shape), res, res))
            self.noise_used.append(nz)
            d.noise = self._p(nz)
            d.noise_batch_stride = res * res if (nz.shape[0] == batch and batch > 1) else 0
            d.noise_weight = self._p(layer.noise.weight)
            d.act_bias = self._p(layer.activate.bias)
        for r, layer in enumerate(rgbs):
            d = s.rgb[r]
            d.cin = layer.conv.in_channel
            d.latent_row = 1 if r == 0 else 2 * r + 1  # to_rgb1 <- row 1, to_rgbs[b] <- row 2b+3
            d.weight = self._p(layer.conv.weight)
            d.mod_weight, d.mod_bias = self._p(layer.conv.modulation.weight), self._p(layer.conv.modulation.bias)
            d.bias = self._p(layer.bias)
            d.fir = self._p(layer.upsample.kernel) if hasattr(layer, 'upsample') else None
            if backward and hasattr(layer, 'upsample'):
                d.fir_flipped = self._p(_flipped_fir(layer.upsample))
        self.struct = s

    def _p(self, t):
        if t.dtype != torch.uint8:
            c = _f32c(t)
            if c.data_ptr() != t.data_ptr():
                self.copied = True
            t = c
        self.keep.append(t)
        return N.ptr(t)

    def refresh(self):
        """After an in-place parameter update (optimizer step: same storages, new versions): the struct points straight at
        the parameters, so only the packed tensor-core weights are stale — ModulatedConv2d.packed refills them in their
        existing buffers.  False = something moved, build a new descriptor."""
        if self.copied:
            return False
        for conv, kw, ptr in self.pack_calls:
            if conv.packed(**kw)[0].data_ptr() != ptr:
                return False
        return True


def _flipped_fir(upsample):
    """flip(kernel) of a ToRGB skip upsampler (operand of the adjoint upfirdn2d, op/upfirdn2d.py:112-117), cached per buffer
    version."""
    k = upsample.kernel
    key = (k.data_ptr(), k._version)
    hit = upsample.__dict__.get('_flipped')
    if hit is None or hit[0] != key:
        hit = upsample.__dict__['_flipped'] = (key, torch.flip(k.detach(), [0, 1]).contiguous().float())
    return hit[1]


def synthesis_param_list(g):
    """The generator parameters the synthesis path reads, in a fixed order (the extra inputs of the autograd node)."""
    ps = [g.input.input]
    for layer in g.styled_layers():
        ps += [layer.conv.weight, layer.conv.modulation.weight, layer.conv.modulation.bias, layer.noise.weight,
               layer.activate.bias]
    for layer in g.rgb_layers():
        ps += [layer.conv.weight, layer.conv.modulation.weight, layer.conv.modulation.bias, layer.bias]
    return ps


def _signature(g):
    """((data_ptr...), (version...)) of every parameter, and the same of every buffer, the descriptor points at: detects
    in-place updates (Adam in optimize_g), load_state_dict, .cuda() and re-assignment.  Only the synthesis network's own
    tensors are walked (a module-tree walk costs 0.2 ms: as much as the kernels of a batch-1 frame take to launch)."""
    ps = synthesis_param_list(g)
    bufs = [layer.conv.blur.kernel for layer in g.styled_layers() if layer.conv.upsample]
    bufs += [layer.upsample.kernel for layer in g.rgb_layers() if hasattr(layer, 'upsample')]
    return (tuple([t.data_ptr() for t in ps]), tuple([t._version for t in ps]), ps[0].device.index,
            tuple([t.data_ptr() for t in bufs]), tuple([t._version for t in bufs]))


def _descriptor(g, noise, batch, backward=False):
    """NetDescriptor for this call, reused from the previous call when nothing it points at has changed (building it is
    ~0.25 ms of Python: the dominant cost of a batch-1 frame, which is how run_inference.py drives the generator), and
    refreshed in place when only parameter versions moved (an optimizer step between two calls)."""
    if any(n is None for n in noise):                  # randomize_noise: fresh tensors every call
        return NetDescriptor(g, noise, batch, backward)
    key = (batch, backward, N.default_format(), N.single_pass(), tuple((n.data_ptr(), n._version, tuple(n.shape)) for n in noise))
    sig = _signature(g)
    cache = g.__dict__.setdefault('_desc_cache', {})
    hit = cache.get(key)
    if hit is not None:
        if hit[0] == sig:
            return hit[1]
        old = hit[0]
        if old[0] == sig[0] and old[2:] == sig[2:] and hit[1].refresh():      # same storages, same buffers
            cache[key] = (sig, hit[1])
            return hit[1]
    desc = NetDescriptor(g, noise, batch, backward)
    desc.keep.extend(noise)                            # the cached struct points at them
    if len(cache) >= 8:
        cache.clear()
    cache[key] = (sig, desc)
    return desc


def _workspace(g, desc, batch, device, backward=False):
    key = (batch, device.index, backward)
    ws = g._workspace.get(key)
    if ws is None:
        fn = N.lib().sgr_synthesis_backward_workspace_bytes if backward else N.lib().sgr_synthesis_workspace_bytes
        nbytes = fn(C.byref(desc.struct), batch)
        if nbytes == 0:
            raise RuntimeError('sgr workspace query: %s' % N.lib().sgr_last_error().decode())
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        for k in [k for k in g._workspace if k[2] == backward]:     # one live workspace per direction
            del g._workspace[k]
        g._workspace[key] = ws
    return ws


def _graphs_enabled(g):
    v = g.__dict__.get('_cuda_graphs')
    if v is None:
        import os
        v = os.environ.get('SGR_CUDA_GRAPHS', '0') not in ('0', '', 'false', 'False')
    return bool(v)


class _GraphedForward:
    """One captured CUDA graph of sgr_synthesis_forward for a fixed (descriptor, batch): static latent / image buffers,
    replayed with one launch.  A batch-1 frame is ~36 kernel launches of 5-30 us each, i.e. bound by the CPU launch rate
    (0.6 ms per frame eager); the graph removes that (run_inference.py renders one frame per generator call)."""

    def __init__(self, g, desc, ws, batch, dev):
        self.desc, self.ws = desc, ws                         # keep every captured pointer alive
        self.latent = torch.empty(batch, g.n_latent, g.style_dim, device=dev, dtype=torch.float32)
        self.image = torch.empty(batch, 3, g.size, g.size, device=dev, dtype=torch.float32)
        self.graph = torch.cuda.CUDAGraph()
        lib = N.lib()

        def call():
            N.check(lib.sgr_synthesis_forward(C.byref(desc.struct), N.ptr(self.latent), batch, N.ptr(self.image), N.ptr(ws),
                                              ws.numel(), None, N.stream()), 'sgr_synthesis_forward')
        self.latent.zero_()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                         # warm-up outside capture (one-time attribute calls, packing)
            call()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        with torch.cuda.graph(self.graph):
            call()

    def run(self, lat):
        self.latent.copy_(lat)
        self.graph.replay()
        return self.image.clone()                             # callers own their frames (the reference returns fresh tensors)


def synthesis_forward(g, latent, noise, want_feats):
    if not latent.is_cuda:
        raise RuntimeError('Generator: latent must be a CUDA tensor (libsgr has no CPU fallback)')
    if latent.ndim != 3 or latent.shape[1] != g.n_latent or latent.shape[2] != g.style_dim:
        raise RuntimeError('latent must be [B,%d,%d], got %s' % (g.n_latent, g.style_dim, tuple(latent.shape)))
    lat = _f32c(latent)
    batch = lat.shape[0]
    dev = lat.device
    with torch.cuda.device(dev):
        desc = _descriptor(g, noise, batch)
        ws = _workspace(g, desc, batch, dev)
        if (not want_feats and _graphs_enabled(g) and not torch.is_grad_enabled() and not torch.cuda.is_current_stream_capturing()
                and all(n is not None for n in noise)):
            gf = desc.__dict__.get('graphed')                 # lives and dies with the cached descriptor
            if gf is None or gf.ws is not ws:
                gf = desc.__dict__['graphed'] = _GraphedForward(g, desc, ws, batch, dev)
            return gf.run(lat), None, lat, desc.noise_used
        image = torch.empty(batch, 3, g.size, g.size, device=dev, dtype=torch.float32)
        feats, feat_ptrs = None, None
        if want_feats:
            feats = []
            arr = (C.c_void_p * len(g.styled_layers()))()
            for l, layer in enumerate(g.styled_layers()):
                res = 4 << ((l + 1) // 2)
                f = torch.empty(batch, layer.conv.out_channel, res, res, device=dev, dtype=torch.float32)
                feats.append(f)
                arr[l] = f.data_ptr()
            feat_ptrs = arr
        N.check(N.lib().sgr_synthesis_forward(C.byref(desc.struct), N.ptr(lat), batch, N.ptr(image), N.ptr(ws),
                                              ws.numel(), feat_ptrs, N.stream()), 'sgr_synthesis_forward')
    return image, feats, lat, desc.noise_used


def synthesis_forward_u8(g, latent, noise, size=None):
    """No-grad forward that delivers uint8 HWC frames [B,size,size,3] straight from the last ToRGB tail
    (sgr_synthesis_forward_ex: 256-pooling + clamp + scale + uint8 fused, the fp32 frame is never written; SURVEY.md §8f-2,
    reference libs/utilities/generic.py:146-148, image_utils.py:97-111, utils_inference.py:16)."""
    if not latent.is_cuda:
        raise RuntimeError('Generator: latent must be a CUDA tensor (libsgr has no CPU fallback)')
    if latent.ndim != 3 or latent.shape[1] != g.n_latent or latent.shape[2] != g.style_dim:
        raise RuntimeError('latent must be [B,%d,%d], got %s' % (g.n_latent, g.style_dim, tuple(latent.shape)))
    lat = _f32c(latent)
    batch, dev = lat.shape[0], lat.device
    size = g.size if size is None else int(size)
    if size <= 0 or g.size % size:
        raise RuntimeError('uint8 frame size %d must divide the network size %d' % (size, g.size))
    with torch.cuda.device(dev):
        desc = _descriptor(g, noise, batch)
        ws = _workspace(g, desc, batch, dev)
        out = torch.empty(batch, size, size, 3, dtype=torch.uint8, device=dev)
        ex = N.ForwardExtras()
        ex.frames_u8, ex.u8_h, ex.u8_w = out.data_ptr(), size, size
        N.check(N.lib().sgr_synthesis_forward_ex(C.byref(desc.struct), N.ptr(lat), batch, None, N.ptr(ws), ws.numel(), None,
                                                 C.byref(ex), N.stream()), 'sgr_synthesis_forward_ex')
    return out


class _Synthesis(torch.autograd.Function):
    """Autograd node for dL/d(latent) — the only gradient the A-matrix training consumes (libs/trainer.py:144,187-189) —
    and, in train() mode, for the generator's own parameters (optimize_g, libs/optimization.py:25-72): those enter as
    extra inputs so that autograd routes their gradients (backward.py)."""

    @staticmethod
    def forward(ctx, latent, g, noise, *params):
        image, feats, lat, noise_used = synthesis_forward(g, latent, noise, want_feats=True)
        ctx.g = g
        ctx.noise = noise_used
        ctx.n_params = len(params)
        ctx.save_for_backward(lat, *feats)
        return image

    @staticmethod
    def backward(ctx, grad_image):
        from .backward import synthesis_backward
        lat, *feats = ctx.saved_tensors
        want = ctx.n_params > 0 and any(ctx.needs_input_grad[3:])
        dlat, pgrads = synthesis_backward(ctx.g, lat, feats, ctx.noise, grad_image, want_param_grads=want)
        if not ctx.needs_input_grad[0]:
            dlat = None
        if ctx.n_params == 0:
            return dlat, None, None
        if pgrads is None:
            pgrads = [None] * ctx.n_params
        pgrads = [pg if need else None for pg, need in zip(pgrads, ctx.needs_input_grad[3:])]
        return (dlat, None, None) + tuple(pgrads)


def run_synthesis(g, latent, noise, return_features=False):
    if return_features:
        image, feats, _, _ = synthesis_forward(g, latent, noise, want_feats=True)
        return image, feats
    if torch.is_grad_enabled():
        # Generator-parameter gradients (the reference's autograd forms them whenever a parameter requires grad, in train()
        # and eval() mode alike).  Policy `g.param_grads`:
        #   'auto' (default)  train() mode (optimize_g, libs/optimization.py:29): formed.  eval() mode with a latent that
        #                     requires grad (A-matrix training, libs/trainer.py:111,144: G is never stepped): NOT formed — the
        #                     weight-gradient GEMMs would double the backward for gradients nobody reads — with a one-time
        #                     warning, so it is never a silent None; eval() mode with a constant latent: formed (the only
        #                     thing backward() can be for).
        #   'always'          exactly the reference: formed whenever any synthesis parameter requires grad.
        #   'never'           latent gradient only.
        # Freezing the generator (G.requires_grad_(False)) selects the latent-only path without any warning.
        params = synthesis_param_list(g)
        mode = getattr(g, 'param_grads', 'auto')
        if mode not in ('auto', 'always', 'never'):
            raise ValueError("Generator.param_grads must be 'auto', 'always' or 'never'")
        any_req = any(p.requires_grad for p in params)
        want = any_req and (mode == 'always' or (mode == 'auto' and (g.training or not latent.requires_grad)))
        if any_req and not want and mode == 'auto' and not g.__dict__.get('_warned_frozen'):
            import warnings
            warnings.warn('Generator is in eval() mode with parameters that require grad: only dL/dlatent is computed, parameter '
                          '.grad stays None (set G.param_grads = "always" for the reference behaviour, or G.requires_grad_(False) '
                          'to silence this).', stacklevel=3)
            g.__dict__['_warned_frozen'] = True
        if want and any(n is None for n in noise):
            raise RuntimeError('generator weight gradients need fixed noise buffers (randomize_noise=False)')
        if not want:
            params = []
        if latent.requires_grad or params:
            return _Synthesis.apply(latent, g, noise, *params)
    image, _, _, _ = synthesis_forward(g, latent, noise, want_feats=False)
    return image

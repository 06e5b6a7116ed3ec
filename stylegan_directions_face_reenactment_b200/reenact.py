"""Glue between the direction matrix and the generator (reference libs/utilities/generic.py:116-152)."""
import torch


def get_shifted_latent_code(G, z, shift, input_is_latent=False, truncation=1, truncation_latent=None, w_plus=True,
                            num_layers=None):
    """latent = z.clone(); latent[:, :num_layers] += shift  (generic.py:116-135).  The caller's z is never mutated."""
    n_latent = G.n_latent
    if not input_is_latent:
        w = G.get_latent(z)
        latent = w.unsqueeze(1).repeat(1, n_latent, 1)
    else:
        latent = z.clone()
    if w_plus:                       # shift [B, k, 512] goes into the first k rows (generic.py:132-133)
        k = shift.shape[1]
        latent = torch.cat([latent[:, :k, :] + shift, latent[:, k:, :]], 1)
    else:                            # shift [B, 512]: every row, or the first num_layers rows (generic.py:123-130)
        k = n_latent if num_layers is None else num_layers
        latent = torch.cat([latent[:, :k, :] + shift.unsqueeze(1), latent[:, k:, :]], 1)
    return latent


def generate_image(G, latent_code, truncation, trunc, w_plus=True, num_layers_shift=8, shift_code=None,
                   input_is_latent=False, return_latents=False):
    """generic.py:137-152: optional shift, then G([code], truncation, ...), 256-pooling for larger nets."""
    if shift_code is None:
        imgs, latents = G([latent_code], return_latents=return_latents, truncation=truncation,
                          truncation_latent=trunc, input_is_latent=input_is_latent)
    else:
        shifted = get_shifted_latent_code(G, latent_code, shift_code, input_is_latent=input_is_latent,
                                          truncation=truncation, truncation_latent=trunc, w_plus=w_plus,
                                          num_layers=num_layers_shift)
        imgs, latents = G([shifted], return_latents=return_latents, truncation=truncation, truncation_latent=trunc,
                          input_is_latent=True)
    if imgs.shape[2] > 256:
        imgs = torch.nn.functional.adaptive_avg_pool2d(imgs, (256, 256))
    return (imgs, latents) if return_latents else imgs


def generate_frames_uint8(G, latent_code, truncation, trunc, w_plus=True, num_layers_shift=8, shift_code=None,
                          input_is_latent=False, size=256):
    """generate_image (generic.py:137-152) followed by the reference's frame post-processing (tensor_to_image + np.uint8,
    image_utils.py:97-111, utils_inference.py:16) in ONE pass: uint8 HWC frames [B,h,w,3] with h = min(G.size, size), written by
    the last ToRGB tail (G.synthesis_uint8) — a quarter of the device->host bytes and no fp32 frame in HBM.  No autograd."""
    with torch.no_grad():
        code = latent_code
        if shift_code is not None:
            code = get_shifted_latent_code(G, latent_code, shift_code, input_is_latent=input_is_latent, truncation=truncation,
                                           truncation_latent=trunc, w_plus=w_plus, num_layers=num_layers_shift)
            input_is_latent = True
        if not input_is_latent:
            code = G.get_latent(code)
        if truncation < 1:
            code = trunc + truncation * (code - trunc)
        if code.ndim < 3:
            code = code.unsqueeze(1).repeat(1, G.n_latent, 1)
        return G.synthesis_uint8(code, size=min(G.size, size))


def frames_to_uint8(images, size=None):
    """Fused output stage: [B,3,H,W] fp32 frames in [-1,1] -> [B,h,w,3] uint8 CUDA tensor with the
    arithmetic of the reference's tensor_to_image + np.uint8 (libs/utilities/image_utils.py:97-111,
    libs/utilities/utils_inference.py:16); `size` < H also applies generate_image's 256-pooling (generic.py:146-148).
    One kernel, a quarter of the device->host bytes of the fp32 frames."""
    from . import _native as N
    if not images.is_cuda:
        raise RuntimeError('frames_to_uint8: CUDA tensor required (no CPU fallback)')
    x = images.detach().contiguous().float()
    if x.ndim == 3:
        x = x.unsqueeze(0)
    b, c, h, w = x.shape
    if c != 3:
        raise RuntimeError('frames_to_uint8 expects RGB frames [B,3,H,W], got %s' % (tuple(images.shape),))
    oh = ow = size if size is not None else None
    if oh is None:
        oh, ow = h, w
    out = torch.empty(b, oh, ow, 3, dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        N.check(N.lib().sgr_frames_to_uint8(N.ptr(x), N.ptr(out), b, h, w, oh, ow, N.stream()), 'sgr_frames_to_uint8')
    return out

"""B200-native (sm_100a) StyleGAN2 synthesis + latent-direction reenactment path.

Public surface mirrors the reference's operator API for this path:
    Generator, EqualLinear, ...        <- libs/gan/StyleGAN2/model.py
    DirectionMatrix                    <- libs/models/direction_matrix.py
    generate_image, get_shifted_latent_code  <- libs/utilities/generic.py:116-152
    upfirdn2d, fused_leaky_relu, FusedLeakyReLU  <- libs/gan/StyleGAN2/op
    formats.load_generator / load_direction_matrix / load_latent_codes  <- the reference's on-disk formats
All compute goes through libsgr.so (csrc/, C ABI in include/sgr.h); there is no CPU fallback.
"""
from .direction_matrix import DirectionMatrix
from .model import (Blur, ConstantInput, EqualLinear, Generator, ModulatedConv2d, NoiseInjection, PixelNorm, StyledConv,
                    ToRGB, Upsample, make_kernel)
from .ops import FusedLeakyReLU, fused_leaky_relu, upfirdn2d
from .reenact import frames_to_uint8, generate_frames_uint8, generate_image, get_shifted_latent_code
from . import formats  # noqa: F401  (on-disk formats: generator / A-matrix checkpoints, latent-code .npy files)

__all__ = ['Generator', 'DirectionMatrix', 'generate_image', 'get_shifted_latent_code', 'frames_to_uint8', 'generate_frames_uint8', 'upfirdn2d',
           'fused_leaky_relu', 'FusedLeakyReLU', 'EqualLinear', 'ModulatedConv2d', 'StyledConv', 'ToRGB', 'Upsample',
           'Blur', 'ConstantInput', 'NoiseInjection', 'PixelNorm', 'make_kernel']

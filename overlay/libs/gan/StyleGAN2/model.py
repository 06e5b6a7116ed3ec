"""Namespace-package overlay: `from libs.gan.StyleGAN2.model import Generator` (reference libs/trainer.py:15,
run_inference.py:13, run_facial_editing.py:15, invert_images.py:24, extract_statistics.py:25) and
`from libs.gan.StyleGAN2.model import EqualLinear` (libs/gan/encoder4editing/.../psp_encoders.py:9) resolve to the
sm_100a implementation when this `overlay/` directory precedes the reference checkout on sys.path.  `libs`, `libs.gan`
and `libs.gan.StyleGAN2` have no __init__.py in the reference (PEP 420), so every other `libs.*` module still comes from
the reference tree."""
from stylegan_directions_face_reenactment_b200.model import (Blur, ConstantInput, EqualLinear, Generator,  # noqa: F401
                                                              ModulatedConv2d, NoiseInjection, PixelNorm, StyledConv,
                                                              ToRGB, Upsample, make_kernel)
from stylegan_directions_face_reenactment_b200.ops import (FusedLeakyReLU, fused_leaky_relu,  # noqa: F401
                                                            upfirdn2d)

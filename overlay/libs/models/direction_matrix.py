"""Namespace-package overlay for `from libs.models.direction_matrix import DirectionMatrix`
(reference libs/trainer.py:14, run_inference.py:12)."""
from stylegan_directions_face_reenactment_b200.direction_matrix import DirectionMatrix  # noqa: F401

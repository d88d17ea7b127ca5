"""hallo.models.mutual_self_attention (face_animate.py:44) -> hallo_b200 (mutual_self_attention.py:39-496)."""
from hallo_b200.models.mutual_self_attention import ReferenceAttentionControl  # noqa: F401

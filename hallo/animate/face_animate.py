"""hallo.animate.face_animate (scripts/inference.py:38) -> hallo_b200 (hallo/animate/face_animate.py:90-442)."""
from hallo_b200.animate.face_animate import FaceAnimatePipeline, FaceAnimatePipelineOutput  # noqa: F401

"""hallo.models.audio_proj (scripts/inference.py:42) -> hallo_b200 (hallo/models/audio_proj.py:40-124)."""
from hallo_b200.models.audio_proj import AudioProjModel  # noqa: F401

"""Overlay of `hallo.animate`: face_animate comes from hallo_b200, face_animate_static from the reference checkout."""
import os

from .. import _reference_dirs

__path__ = [os.path.dirname(os.path.abspath(__file__))] + _reference_dirs("animate")

"""Overlay of `hallo.models`: the hot-path modules come from hallo_b200, the rest from the reference checkout."""
import os

from .. import _reference_dirs

__path__ = [os.path.dirname(os.path.abspath(__file__))] + _reference_dirs("models")

"""ORACLE-ONLY compat shim: the handful of diffusers==0.27.2 symbols the reference hot path imports.

diffusers is pinned by the reference (requirements.txt:8, setup.py:22) but is absent from this
image and cannot be installed (no network).  This package restates the published semantics of the
symbols listed in SURVEY.md section 8c / Appendix A so that the UNMODIFIED reference files under
/root/reference/hallo/models can be imported and run on CPU.  It is test infrastructure: nothing
under hallo_b200/ imports it.
"""
from .models.modeling_utils import ModelMixin  # noqa: F401
from .configuration_utils import ConfigMixin, register_to_config  # noqa: F401

__version__ = "0.27.2-compat"

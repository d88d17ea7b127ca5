"""Empty stand-in: hallo/models/motion_module.py:58-59 imports xformers but never calls it at inference."""

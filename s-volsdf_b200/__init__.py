"""svolsdf_b200 — B200-native VolSDF volume-rendering hot path behind the reference's model API.

Sub-packages mirror `volsdf.model.*` / `volsdf.utils.rend_util` of cvlab-stonybrook/s-volsdf; all
compute goes through the C-ABI library built from `csrc/` (see `_lib.py`, `include/svs.h`).
"""
__version__ = '0.1.0'

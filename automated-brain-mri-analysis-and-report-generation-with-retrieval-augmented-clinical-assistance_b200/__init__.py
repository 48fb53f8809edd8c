"""brainseg_b200 — B200-native hot path of the brain-MRI pipeline.

Sliding-window Generic_UNet inference (predict_3D) and the voxel post-processing that consumes it, as
hand-written sm_100a CUDA behind a C ABI (``libbrainseg_b200.so``, declared in ``include/brainseg_b200.h``).
Import as ``brainseg_b200`` (see the shim package of that name at the repo root).
"""
__version__ = "0.1.0"

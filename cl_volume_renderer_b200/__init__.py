"""B200-native volume path tracer behind cl-volume-renderer's renderer interface.

csrc/   hand-written sm_100a kernels + the C-ABI (include/vr.h) -> libvr.so
host/   C++ mirror of the reference's renderer / reference_volume / signed_distance_field / env_map classes
api.py  ctypes view of the C-ABI for tests and bench.py
synth.py synthetic volumes / environment maps / cameras
"""
from . import api, synth  # noqa: F401

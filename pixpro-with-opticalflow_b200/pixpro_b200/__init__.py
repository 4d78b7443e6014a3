"""pixpro_b200 — Python host side of the B200-native pixel-pretext hot path.

`_cabi`  loads libpixpro_b200.so (the C ABI of include/pixpro_b200.h) through ctypes;
`ops`    wraps each entry point for torch CUDA tensors (device memory + current stream are
         the only things torch provides) and defines the autograd Functions;
`synth`  seeded synthetic inputs of the reference loader's shapes.
The drop-in mirror of the reference's modules lives next to this package in `contrast/`.
"""
from . import _cabi  # noqa: F401

__all__ = ["_cabi", "ops", "synth"]

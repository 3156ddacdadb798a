"""delora_b200 — B200-native (sm_100a) hot path of DeLORA behind the reference's operator API.

Importing this package does not load the CUDA library; `delora_b200._lib.lib()` does, and
raises if `libdelora_b200.so` has not been built (there is no CPU fallback).
"""
__version__ = "0.1.0"

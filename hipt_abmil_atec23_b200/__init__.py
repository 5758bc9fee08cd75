"""B200-native HIPT_4K + CLAM_SB hot path (hand-written sm_100a CUDA behind a C ABI, bound with ctypes)."""
__version__ = "0.1.0"

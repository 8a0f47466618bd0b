"""kai0_b200 — B200-native (sm_100a) engine for the pi0.5 hot path of OpenDriveLab/kai0.

Only what the path needs lives here: `csrc/` (hand-written CUDA + the C-ABI of include/pi05.h), the ctypes
binding (`_lib`), and the host-side mirror of the reference's `PI0Pytorch` surface (`pi0_pytorch`).
"""
__version__ = "0.1.0"

"""tiny_audio_b200 -- B200-native (sm_100a) implementation of tiny-audio's training hot path.

Host side: Python over PyTorch tensors (device memory + streams only); compute: hand-written CUDA in
csrc/ behind the C ABI declared in include/tinyaudio_b200.h.  No CPU fallback, no Triton, no torch.compile.
"""
from . import lib  # noqa: F401

__all__ = ["lib"]

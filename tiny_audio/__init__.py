"""Import shim: `scripts/train.py:44-50` of the reference does `from tiny_audio.asr_config import ASRConfig`,
`from tiny_audio.asr_modeling import ASRModel`, `from tiny_audio.projectors import ...`.  With this repository on
PYTHONPATH those imports resolve to the B200 implementation in `tiny_audio_b200/`."""
import importlib
import sys

for _name in ("asr_config", "asr_modeling", "asr_processing", "projectors"):
    _mod = importlib.import_module(f"tiny_audio_b200.{_name}")
    sys.modules[f"{__name__}.{_name}"] = _mod
    globals()[_name] = _mod

from tiny_audio_b200.asr_config import ASRConfig  # noqa: E402,F401
from tiny_audio_b200.asr_modeling import ASRModel  # noqa: E402,F401
from tiny_audio_b200.asr_processing import ASRProcessor  # noqa: E402,F401

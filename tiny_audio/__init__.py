"""Import shim: `scripts/train.py:44-50` of the reference does `from tiny_audio.asr_config import ASRConfig`,
`from tiny_audio.asr_modeling import ASRModel`, `from tiny_audio.projectors import ...`.  With this repository on
PYTHONPATH those imports resolve to the B200 implementation in `tiny_audio_b200/`."""
import importlib
import os
import sys

for _name in ("asr_config", "asr_modeling", "asr_processing", "projectors"):
    _mod = importlib.import_module(f"tiny_audio_b200.{_name}")
    sys.modules[f"{__name__}.{_name}"] = _mod
    globals()[_name] = _mod

# Everything else the reference's scripts import from `tiny_audio` (scripts/train.py:50 `tiny_audio.augmentation`, the deploy
# handler, alignment, ...) is off the hot path and stays the reference's own code: extend this package's search path with any other
# `tiny_audio` directory found later on sys.path, so `PYTHONPATH=<this repo>:<tiny-audio checkout>` resolves those submodules there
# while the four hot-path modules above (already in sys.modules) win.
_here = os.path.abspath(os.path.dirname(__file__))
for _entry in list(sys.path):
    _cand = os.path.abspath(os.path.join(_entry or ".", "tiny_audio"))
    if _cand != _here and _cand not in __path__ and os.path.isfile(os.path.join(_cand, "asr_modeling.py")):
        __path__.append(_cand)

from tiny_audio_b200.asr_config import ASRConfig  # noqa: E402,F401
from tiny_audio_b200.asr_modeling import ASRModel  # noqa: E402,F401
from tiny_audio_b200.asr_processing import ASRProcessor  # noqa: E402,F401

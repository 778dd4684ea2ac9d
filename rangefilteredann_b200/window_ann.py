"""Drop-in module name: with this directory on sys.path, `import window_ann` (what the
reference's experiments/wrapper.py does) resolves to the B200 engine."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from rangefilteredann_b200 import load_engine as _load  # noqa: E402

_eng = _load()
globals().update({k: getattr(_eng, k) for k in dir(_eng) if not k.startswith("__")})
__engine__ = _eng.__engine__
__version__ = _eng.__version__

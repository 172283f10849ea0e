"""Host-side mirror of the reference's `src/quantization` package (src/quantization/__init__.py:1-4)."""
from .modules import *  # noqa: F401,F403
from .quantizer import *  # noqa: F401,F403

"""quantized-autoregression-image-generator_b200: B200-native SOM-codebook hot path.

The directory name is not a Python identifier, so the importable package is ``somcb`` inside
it; importing this directory through importlib (or adding it to sys.path) exposes ``somcb``.
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)

import somcb  # noqa: E402,F401
from somcb import *  # noqa: E402,F401,F403

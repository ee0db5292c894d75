"""B200-native (sm_100a) drop-in for the sequence-classifier hot path of
cafe1930/MultimodalAggressionRecognition: same nn.Module API as the reference's models.py, every
forward/backward runs hand-written CUDA kernels through the C ABI of libmar.so (include/mar.h)."""
from . import ops
from .ops import (engine, get_precision, manual_seed, precision, rng_advance, set_precision, shape_probe)
from .models import *  # noqa: F401,F403  (the reference's class names)
from . import models

__version__ = "0.1.0"

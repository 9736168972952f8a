"""sparch_b200 -- B200-native implementation of sparch's surrogate-gradient SNN hot path.

``sparch_b200.snns`` mirrors ``sparch.models.snns`` (reference sparch/models/snns.py);
``sparch_b200.functional`` holds the autograd wrappers over ``libsparch_b200.so``
(hand-written sm_100a CUDA, C ABI in include/sparch_b200.h).  Importing the package does
not load the native library; the first kernel call does, and fails loudly if it is absent.
"""
from .functional import CrossEntropyLoss, set_precision
from .snns import (SNN, LIFLayer, RadLIFLayer, ReadoutLayer, RLIFLayer, SpikeFunctionBoxcar,
                   adLIFLayer, set_state_init)

__all__ = ["SNN", "LIFLayer", "adLIFLayer", "RLIFLayer", "RadLIFLayer", "ReadoutLayer",
           "SpikeFunctionBoxcar", "set_state_init", "set_precision", "CrossEntropyLoss"]
__version__ = "0.1.0"

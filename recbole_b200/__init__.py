"""recbole_b200 -- B200-native (sm_100a) embedding hot path behind RecBole's model / trainer /
evaluator API.  See DESIGN.md, INTEGRATION.md and include/recbole_b200.h.

Importing the package loads librecbole_b200.so; there is no CPU or ATen fallback.
"""
from . import _lib, ops  # noqa: F401
from .data import DeviceTrainLoader, EvalIndex  # noqa: F401
from .evaluator import FusedTopKEvaluator  # noqa: F401
from .interaction import Interaction  # noqa: F401
from .model import FusedBPR, FusedOptimizer  # noqa: F401
from .model_fm import FusedFM, FusedMFSimple  # noqa: F401
from .sampler import DeviceSampler  # noqa: F401
from .trainer import FusedTrainer  # noqa: F401

__version__ = "0.1.0"

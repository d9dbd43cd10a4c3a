"""lsml_b200 -- B200-native (sm_100a) implementation of lsml's level-set evolution hot path.

Host-side mirror of the reference package layout for that path only:

    lsml_b200.gradient.masked_gradient   <-> lsml.gradient.masked_gradient
    lsml_b200.feature                    <-> lsml.feature
    lsml_b200.LevelSetMachineLearning    <-> lsml.LevelSetMachineLearning (segment path)

Every numerical operation runs in hand-written CUDA kernels behind the C ABI of
``liblsml_b200.so`` (``include/lsml_b200.h``); there is no CPU fallback.
"""
from . import _lib  # noqa: F401  (raises ImportError when the CUDA library is not built)

__version__ = "0.1.0"


def __getattr__(name):
    # lazy: importing the model pulls in torch
    if name in ("LevelSetMachineLearning", "SegmentationModel"):
        from .core.model import LevelSetMachineLearning
        return LevelSetMachineLearning
    raise AttributeError(name)

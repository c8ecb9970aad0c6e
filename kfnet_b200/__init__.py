"""kfnet_b200 -- B200-native KFNet hot path (image sequence -> per-frame 60x80x4 scene-coordinate +
confidence maps).  Python host code over libkfnet_b200.so (hand-written sm_100a CUDA behind a C ABI,
include/kfnet_b200.h); PyTorch is only the device-memory / stream carrier.  There is no CPU fallback:
constructing an Engine without the built library or without a CUDA device raises."""
from .spec import KFNetDataSpec  # noqa: F401

__all__ = ["KFNetDataSpec", "Engine", "KFNet", "eval"]


def __getattr__(name):
    if name == "Engine":
        from .engine import Engine
        return Engine
    if name == "KFNet":
        from .kfnet import KFNet
        return KFNet
    if name == "eval":
        from .eval import eval as _eval
        return _eval
    raise AttributeError(name)

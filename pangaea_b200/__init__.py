"""pangaea_b200 - B200-native (sm_100a) read-cloud featurization for Pangaea.

Drop-in classes ``Feature`` (src/feature.py) and ``Data`` (src/data.py) over the C-ABI
in ``include/pangaea_b200.h``; the CUDA kernels live in ``pangaea_b200/csrc``.
"""
__all__ = ["Feature", "Data"]


def __getattr__(name):  # lazy: importing the package must not need torch/pandas
    if name == "Feature":
        from .feature import Feature

        return Feature
    if name == "Data":
        from .data import Data

        return Data
    raise AttributeError(name)

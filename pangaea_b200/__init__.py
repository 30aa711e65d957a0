"""pangaea_b200 - B200-native (sm_100a) read-cloud featurization for Pangaea.

Drop-in classes ``Feature`` (src/feature.py) and ``Data`` (src/data.py) over the C-ABI
in ``include/pangaea_b200.h``; the CUDA kernels live in ``pangaea_b200/csrc``.
"""
__all__ = ["Feature", "Data", "CustomWeightedRandomSampler", "DeviceBatches"]


def __getattr__(name):  # lazy: importing the package must not need torch/pandas
    if name == "Feature":
        from .feature import Feature

        return Feature
    if name == "Data":
        from .data import Data

        return Data
    if name in ("CustomWeightedRandomSampler", "DeviceBatches"):  # src/utils.py:11-23 and the DataLoader of src/pangaea.py:87-89
        from . import sampler

        return getattr(sampler, name)
    raise AttributeError(name)

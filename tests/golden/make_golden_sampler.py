"""Golden vectors of the reference's CustomWeightedRandomSampler (/root/reference/src/utils.py:11-23), generated HERE by
importing the unmodified class: weights, numpy seed -> the indices one epoch yields, with and without replacement.
    python tests/golden/make_golden_sampler.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference/src")
from utils import CustomWeightedRandomSampler  # noqa: E402  (the reference's own class)

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sampler")
cases = {}
rng = np.random.default_rng(0)
for name, n, m_rep, m_norep in (("small", 50, 80, 35), ("weights_like_data", 5000, 5000, 3500), ("zeros_inside", 300, 300, 120)):
    w = rng.random(n) ** 2
    if name == "weights_like_data":
        w = (rng.random(n) * 0.9 + 0.05) ** 2  # Data.weights = max(normalised row) ** 2
    if name == "zeros_inside":
        w[rng.random(n) < 0.4] = 0.0
    for replacement, m in ((True, m_rep), (False, m_norep)):
        np.random.seed(1234)
        s = CustomWeightedRandomSampler(torch.from_numpy(w), num_samples=m, replacement=replacement)
        first = np.array(list(iter(s)), dtype=np.int64)
        second = np.array(list(iter(s)), dtype=np.int64)  # the next epoch continues the generator
        cases[f"{name}_{'rep' if replacement else 'norep'}"] = (w, m, replacement, first, second)
for k, (w, m, rep, a, b) in cases.items():
    np.savez(os.path.join(OUT, k + ".npz"), weights=w, num_samples=m, replacement=rep, seed=1234, epoch1=a, epoch2=b)
    print(k, len(w), m, a[:8])

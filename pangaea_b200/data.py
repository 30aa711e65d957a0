"""Drop-in for Pangaea's ``src/data.py`` (class ``Data``), normalisation on the GPU.

Same constructor and attributes as /root/reference/src/data.py:8-31: ``.abd`` f32 [G, v],
``.tnf`` f32 [G, 136], ``.weights`` f64 [G], ``.bc``, ``__len__``, ``__getitem__`` ->
``{"abd", "tnf", "bc"}``.  The three arrays are computed by csrc/normalize.cuh (fp64
divide, fp32 store - bit-identical to sklearn's normalize + astype).  Like the reference,
which normalises what pandas read back from the tools' CSV text, tallies >= 10^6 enter the
division rounded to 6 significant digits (count_kmer.cpp:211) - also when the exact
device-resident tallies of ``Feature.features`` are passed.  Host copies are
kept because the unchanged consumer forks DataLoader workers (pangaea.py:87-89); the
same buffers are also available zero-copy as CUDA tensors through DLPack
(``abd_cuda`` / ``tnf_cuda`` / ``weights_cuda``) for a consumer that stays on device.
"""
from __future__ import annotations

import logging

import numpy as np
import torch
from torch.utils.data import Dataset

from . import _lib


class Data(Dataset):
    def __init__(self, barcodes, abd, tnf, features=None, device=0):
        """``abd`` / ``tnf``: raw tallies as returned by ``Feature.extract_features`` /
        ``load_features``.  ``features``: the device-resident result of the same call
        (``Feature.features``), which skips the re-upload."""
        super().__init__()
        self.bc = barcodes
        logging.info("calculate sampling weights")
        if features is None:
            # a ctx without a k-mer table: normalisation needs none (PG_TABLE_NONE)
            ctx = _lib.Context(device=device, vector_size=int(np.shape(abd)[1]), table_mode=_lib.PG_TABLE_NONE)
            features = ctx.features_from_raw(np.asarray(abd), np.asarray(tnf))
        logging.info("normalize data")
        self.abd, self.tnf, self.weights = features.normalized()
        self._features = features
        logging.info("preprocessing completed")

    # zero-copy views of the same device buffers (DLPack)
    @property
    def abd_cuda(self) -> torch.Tensor:
        return self._features.torch(_lib.ABD)

    @property
    def tnf_cuda(self) -> torch.Tensor:
        return self._features.torch(_lib.TNF)

    @property
    def weights_cuda(self) -> torch.Tensor:
        return self._features.torch(_lib.WEIGHTS)

    def __len__(self):
        return self.abd.shape[0]

    def __getitem__(self, idx):
        if torch.is_tensor(idx):
            idx = idx.tolist()
        return {"abd": self.abd[idx, :], "tnf": self.tnf[idx, :], "bc": self.bc[idx]}

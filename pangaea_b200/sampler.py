"""Step-2 input pipeline on the device: drop-ins for ``CustomWeightedRandomSampler``
(/root/reference/src/utils.py:11-23) and for the ``DataLoader`` that ``src/pangaea.py:85-89`` builds around ``Data``.

The reference draws ``numpy.random.choice(range(N), size=num_samples, p=weights / weights.sum(), replace=replacement)`` on
the host once per epoch and lets forked DataLoader workers gather the rows one by one.  Here the uniforms still come from
numpy's generator - so a seeded run draws exactly the indices the reference draws - and everything that scales with N runs
on the device (csrc/sampler.cuh): the cumulative distribution, the binary searches, the first-occurrence filter of the draws
without replacement, and the gather of each batch's rows out of the device-resident matrices (no worker processes, which
CUDA tensors inside a Dataset would not survive anyway - SURVEY §8b "hazards").
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


class CustomWeightedRandomSampler:
    """Same constructor and iteration protocol as the reference class (a ``torch.utils.data.Sampler``):
    ``CustomWeightedRandomSampler(weights, num_samples, replacement=True)``; ``iter()`` yields ``num_samples`` ints."""

    def __init__(self, weights, num_samples, replacement=True, generator=None, ctx=None, device=0):
        if not isinstance(num_samples, int) or isinstance(num_samples, bool) or num_samples <= 0:
            raise ValueError(f"num_samples should be a positive integer value, but got num_samples={num_samples}")  # torch's check
        if not isinstance(replacement, bool):
            raise ValueError(f"replacement should be a boolean value, but got replacement={replacement}")
        self.weights = torch.as_tensor(weights, dtype=torch.double)  # WeightedRandomSampler.__init__
        if self.weights.dim() != 1:
            raise ValueError(f"weights should be a 1d sequence but given weights have shape {tuple(self.weights.shape)}")
        self.num_samples, self.replacement, self.generator = num_samples, replacement, generator
        self.ctx = ctx or _lib.Context(device=device, table_mode=_lib.PG_TABLE_NONE)

    def _new_sampler(self):
        w = np.ascontiguousarray(self.weights.cpu().numpy())
        total = float(torch.sum(self.weights.cpu()).numpy())  # the reference's normaliser, computed the reference's way
        h = _lib._vp()
        self.ctx._ck(_lib.lib().pg_sampler_create(self.ctx.h, w.ctypes.data, len(w), total, C.byref(h)))
        return h

    def indices_cuda(self) -> torch.Tensor:
        """one epoch's indices as an int64 CUDA tensor (numpy.random's global generator supplies the uniforms)"""
        n, m = len(self.weights), self.num_samples
        L = _lib.lib()
        h = self._new_sampler()
        out = torch.empty(m, dtype=torch.int64, device=f"cuda:{self.ctx.params.device}")
        try:
            if self.replacement:
                u = np.random.random_sample(m)  # numpy.random.choice: uniform_samples = self.random_sample(shape)
                self.ctx._ck(L.pg_sampler_draw(self.ctx.h, h, u.ctypes.data, m, out.data_ptr()))
            else:
                if m > n:
                    raise ValueError("Cannot take a larger sample than population when 'replace=False'")
                if np.count_nonzero(self.weights.numpy() > 0) < m:
                    raise ValueError("Fewer non-zero entries in p than size")
                found = 0
                while found < m:  # numpy.random.choice's loop, one round per trip
                    u = np.random.rand(m - found)
                    new = C.c_int64(0)
                    self.ctx._ck(L.pg_sampler_draw_unique_round(self.ctx.h, h, u.ctypes.data, m - found, found, out.data_ptr(), C.byref(new)))
                    found += int(new.value)
        finally:
            L.pg_sampler_free(self.ctx.h, h)
        return out

    def __iter__(self):
        return iter(self.indices_cuda().tolist())

    def __len__(self):
        return self.num_samples


class DeviceBatches:
    """What ``DataLoader(dataset, batch_size, sampler=...)`` yields - dicts ``{"abd", "tnf", "bc"}`` - with the tensors
    gathered on the device from ``Data``'s device-resident matrices.  ``sampler=None, shuffle=False`` walks the rows in order
    (the reference's ``dataloader_original`` uses ``shuffle=True``: pass a sampler or ``shuffle=True`` for a permutation from
    torch's generator)."""

    def __init__(self, dataset, batch_size, sampler=None, shuffle=False):
        self.ds, self.bs, self.sampler, self.shuffle = dataset, int(batch_size), sampler, shuffle
        self.feats = dataset._features
        self.ctx = self.feats.ctx
        self.feats.normalize()

    def __len__(self):
        n = len(self.sampler) if self.sampler is not None else len(self.ds)
        return (n + self.bs - 1) // self.bs

    def __iter__(self):
        dev = f"cuda:{self.ctx.params.device}"
        if self.sampler is not None:
            idx = self.sampler.indices_cuda() if hasattr(self.sampler, "indices_cuda") else torch.as_tensor(list(self.sampler), dtype=torch.int64, device=dev)
        elif self.shuffle:
            idx = torch.randperm(len(self.ds)).to(dev)
        else:
            idx = torch.arange(len(self.ds), dtype=torch.int64, device=dev)
        idx = idx.contiguous()
        host_idx = idx.cpu().numpy()
        L = _lib.lib()
        bc = np.asarray(self.ds.bc)
        for lo in range(0, len(idx), self.bs):
            part = idx[lo:lo + self.bs]
            m = len(part)
            abd = torch.empty((m, self.feats.abd_dim), dtype=torch.float32, device=dev)
            tnf = torch.empty((m, self.feats.tnf_dim), dtype=torch.float32, device=dev)
            self.ctx._ck(L.pg_features_gather(self.ctx.h, self.feats.h, part.data_ptr(), m, abd.data_ptr(), tnf.data_ptr()))
            yield {"abd": abd, "tnf": tnf, "bc": list(bc[host_idx[lo:lo + m]])}

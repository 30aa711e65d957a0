"""Drop-in for Pangaea's ``src/feature.py`` (class ``Feature``), B200-native.

Same constructor, same methods, same return layout, same cache files as the reference
(/root/reference/src/feature.py:11-147) so that ``src/pangaea.py:70,84`` can import this
class instead.  What changes is underneath: where the reference spawns ``jellyfish``,
``bin/count_kmer`` and ``bin/count_tnf`` and parses their gz-CSV output with pandas, this
class parses the FASTQ once on the host, ships the bases to the GPU and runs the
sm_100a kernels behind ``include/pangaea_b200.h``.  There is no CPU compute path.
"""
from __future__ import annotations

import logging
import os

import numpy as np

from . import _lib
from . import stream as stream_mod


def _reference_text_rounding(m: np.ndarray) -> np.ndarray:
    """The reference writes tallies with ``ostream << double`` at precision 6
    (count_kmer.cpp:211, count_tnf.cpp:204), so a tally >= 1e6 reaches pandas as e.g.
    ``1.11493e+06`` and the whole matrix becomes float64 (feature.py:115).  Reproduce
    that, exactly, for the rare matrices that contain such tallies."""
    big = m >= 1_000_000
    if not big.any():
        return m
    out = m.astype(np.float64)
    out[big] = [float("%.6g" % v) for v in m[big]]
    return out


def write_reference_csv(path, names, matrix):
    """The gz-CSV artefact the reference tools leave next to the pickles (count_kmer.cpp:202-215, count_tnf.cpp:195-208):
    one line ``label,v0,v1,...`` per cloud, values printed like ``ostream << double`` at the default precision 6
    (integers below 10^6 as integers, larger ones as ``1.11493e+06``).  Optional: nothing downstream of the pickles reads it."""
    import gzip

    m = np.asarray(matrix)
    small = (m < 1_000_000).all()
    with gzip.open(path, "wt", compresslevel=1, newline="\n") as f:
        for name, row in zip(names, m):
            vals = [str(int(v)) for v in row] if small else ["%g" % float(v) for v in row]
            f.write(str(name) + "," + ",".join(vals) + "\n")


def _is_gzip(path):
    with open(path, "rb") as f:
        return f.read(2) == b"\x1f\x8b"


class Feature:
    def __init__(self, args, script_path=None, device=0, batch_seq_bytes=None, ingest=None):
        # /root/reference/src/feature.py:12-26
        self.args = args
        self.tnf_k = str(args.tnf_kmer)
        self.ws = args.window_size
        self.vs = args.vector_size
        self.kmer = args.kmer
        self.minl = args.min_length
        self.threads = args.threads
        self.device = device
        self.feature_dir = os.path.join(args.output, "1.features")
        os.makedirs(self.feature_dir, exist_ok=True)
        self.features = None  # device-resident matrices of the last extract_features()
        self.timing = {}
        # the input is streamed through the GPU in batches of about this many sequence bytes (pangaea_b200/stream.py), so a
        # file of any size works - like the reference, which holds one cloud at a time (count_kmer.cpp:236-282)
        self.batch_seq_bytes = int(batch_seq_bytes or os.environ.get("PG_BATCH_SEQ_BYTES", 0) or stream_mod.DEFAULT_BATCH_SEQ_BYTES)
        # who parses the text: "device" (pg_ingest_text: the host only moves file bytes - plain-text interleaved input without a
        # quality filter), "host" (csrc/fastq.cpp: gzip, paired files, the quality filter), "auto" = the device where it applies
        self.ingest = ingest or os.environ.get("PG_INGEST", "auto")
        if self.ingest not in ("auto", "device", "host"):
            raise ValueError("ingest must be auto, device or host")
        self.ingest_used = None

    # -- file names of the reference's cache artefacts (feature.py:42-44,68-71,126-127)
    def _abd_pkl(self):
        return os.path.join(self.feature_dir, f"abundance.k{self.kmer}.v{self.vs}.w{self.ws}.m{self.minl}.pkl")

    def _tnf_pkl(self):
        return os.path.join(self.feature_dir, f"tnf.m{self.minl}.pkl")

    def _inputs(self):
        a = self.args
        if getattr(a, "reads1", None) and getattr(a, "reads2", None):
            # the paired branch is the one that passes --min-qual-char=? to jellyfish (feature.py:79,83)
            return a.reads1, a.reads2, ord("?")
        if getattr(a, "interleaved_reads", None):
            return a.interleaved_reads, None, 0
        raise ValueError("reads must be specified")  # feature.py:99,111,135

    def _abd_csv(self):
        return os.path.join(self.feature_dir, f"abundance.k{self.kmer}.v{self.vs}.w{self.ws}.m{self.minl}.gz")

    def _tnf_csv(self):
        return os.path.join(self.feature_dir, f"tnf.m{self.minl}.gz")

    def extract_features(self, write_cache=True, reference_text_rounding=True, write_csv=False):
        """-> (names[G] object, abundance[G, v], tnf[G, 136]) in file order (feature.py:28-39)."""
        import pandas as pd

        if write_cache and os.path.isfile(self._abd_pkl()) and os.path.isfile(self._tnf_pkl()):
            logging.info("load abundance")  # the reference skips every artefact that exists (feature.py:113-119)
            names, abundance, tnf = self.load_features()
        else:
            path1, path2, min_qual = self._inputs()
            ctx = _lib.Context(device=self.device, k=int(self.kmer), tnf_k=int(self.tnf_k), window_size=int(self.ws),
                               vector_size=int(self.vs), min_length=int(self.minl), min_qual_char=min_qual)
            size = sum(os.path.getsize(p) for p in (path1, path2) if p)
            eligible = path2 is None and not min_qual and not _is_gzip(path1)
            if self.ingest == "device" and not eligible:
                raise ValueError("the device parser takes one plain-text interleaved file and no quality filter")
            names = feats = None
            if eligible and self.ingest != "host":
                window = int(min(1 << 30, max(1 << 12, self.batch_seq_bytes / 0.45)))  # file bytes per window (~45 % are sequence)
                try:
                    names, feats = stream_mod.extract_features_device_ingest(ctx, path1, window_bytes=window)
                    self.ingest_used = "device"
                except _lib.PgError as e:
                    if self.ingest == "device" or e.code not in (-1, -5):  # (PG_ERR_INVALID / PG_ERR_STATE: the text, not the GPU)
                        raise
                    logging.info(f"device parser declined the input ({e}); host reader")
            if feats is None:
                hint = None if _is_gzip(path1) else 0.45 * size  # (sequence lines are ~40 % of a plain FASTQ's bytes)
                names, feats = stream_mod.extract_features_streaming(
                    ctx, lambda: _lib.FastqStream(path1, path2, want_qual=bool(min_qual), pinned=True, target_seq_bytes=self.batch_seq_bytes),
                    seq_bytes_hint=hint)
                self.ingest_used = "host"
            names = np.array(names, dtype=object)
            abd32, tnf32 = feats.raw()
            abundance, tnf = abd32.astype(np.int64), tnf32.astype(np.int64)
            if reference_text_rounding:
                abundance, tnf = _reference_text_rounding(abundance), _reference_text_rounding(tnf)
            self.timing = {n: ctx.timing(w)[0] for n, w in (("pack", 0), ("count", 1), ("group", 2), ("featurize", 3), ("normalize", 4))}
            self.features, self._ctx = feats, ctx
            ctx.trim()  # the VAE comes next on this GPU: hand the idle scratch (partitions, packed streams) back to the driver
            if write_cache:  # same pickles the reference leaves: DataFrame, column 0 = label
                for path, m in ((self._abd_pkl(), abundance), (self._tnf_pkl(), tnf)):
                    df = pd.DataFrame(m)
                    df.columns = range(1, m.shape[1] + 1)
                    df.insert(0, 0, names)
                    df.to_pickle(path)
            if write_csv:  # the intermediate text files of the reference (feature.py:104-109,128-135); raw tallies
                write_reference_csv(self._abd_csv(), names, abd32)
                write_reference_csv(self._tnf_csv(), names, tnf32)
        logging.info(f"abundance shape {abundance.shape}")
        logging.info(f"tnf shape {tnf.shape}")
        with open(os.path.join(self.feature_dir, "feature_finished"), "w") as f:
            f.write("feature finished")  # feature.py:37-38
        return names, abundance, tnf

    def load_features(self):
        """Resume path (feature.py:41-65): read the two pickles back."""
        import pandas as pd

        try:
            tnf = pd.read_pickle(self._tnf_pkl())
        except Exception:
            raise Exception(self._tnf_pkl(), " file not found")
        readnames = tnf[0].to_numpy()
        tnf = tnf.drop(columns=0).to_numpy()
        logging.info(f"tnf shape {tnf.shape}")
        if not os.path.isfile(self._abd_pkl()):
            raise Exception(self._abd_pkl(), " file not found")
        logging.info("load features from pickle file " + self._abd_pkl())
        df = pd.read_pickle(self._abd_pkl())
        readnames = df[0].to_numpy()
        abundance = df.drop(columns=0).to_numpy()
        logging.info(f"abundance shape {abundance.shape}")
        return readnames, abundance, tnf

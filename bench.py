#!/usr/bin/env python
"""bench.py - reads/sec featurized (k-mer count + abundance + TNF) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--pairs P]

One "step" = one pass of the whole hot path over one batch of synthetic linked reads:
clear table -> 2-bit pack -> count canonical 15-mers -> group clouds -> fused abundance
histogram + TNF -> L1 normalise.  Workload at N=1 is BASELINE.json configs[1]
("synthetic stLFR 2x100bp, 50M read pairs, ~500k barcodes, 1 B200"); with N>1 every rank
holds the same amount (weak scaling), counts its shard, the dense count tables are summed
with one NCCL all-reduce, and every rank featurizes its own clouds.

Prints ONE JSON line (see DESIGN.md "Measurement").  `value` = device-resident input,
`e2e` = the same metric through the C-ABI with pinned HOST buffers (H2D of the reads and
D2H of the normalised matrices inside the timed region).  `--impl reference` times the
reference's own CPU tools (oracle/_ref, compiled from /root/reference) on a bounded sample.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "reads/sec featurized (k-mer count+abundance+TNF)"
UNIT = "reads/s"
WORKLOAD = "synthetic stLFR 2x100bp, 50M read pairs, ~500k barcodes, 1 B200"


# --------------------------------------------------------------------------------------
# synthetic input, generated in HBM (csrc/synth.cuh)
# --------------------------------------------------------------------------------------
def make_synthetic_batch(ctx, n_pairs, read_len=100, n_barcodes=None, n_genomes=200, genome_len=3_000_000,
                         frag_len=50_000, seed=2, sub_rate=0.005, n_rate=0.0005):
    """SURVEY.md §8d model.  Returns torch device tensors + a pg_reads over them."""
    import torch

    from pangaea_b200 import _lib

    dev = f"cuda:{ctx.params.device}"
    n_barcodes = n_barcodes or max(1, n_pairs // 100)
    rng = np.random.default_rng(seed)
    counts = rng.poisson(n_pairs / n_barcodes, size=n_barcodes).astype(np.int64)
    start = np.concatenate([[0], np.cumsum(counts)])
    start = np.minimum(start, n_pairs)
    start[-1] = n_pairs
    abundance = rng.lognormal(0.0, 1.0, size=n_genomes)
    genome = rng.choice(n_genomes, size=n_barcodes, p=abundance / abundance.sum()).astype(np.int32)
    frag_len = int(min(frag_len, genome_len))
    insert = int(min(max(2 * read_len, 350), frag_len))
    n_reads = 2 * n_pairs
    n_bytes = n_reads * (read_len + 1)
    d_start = torch.from_numpy(start).to(dev)
    d_genome = torch.from_numpy(genome).to(dev)
    seq = torch.empty(n_bytes + 64, dtype=torch.uint8, device=dev)
    off = torch.empty(n_reads + 1, dtype=torch.int64, device=dev)
    flag = torch.empty(max(n_reads, 1), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    ctx._ck(_lib.lib().pg_synth_generate(ctx.h, n_pairs, read_len, n_barcodes, d_start.data_ptr(), d_genome.data_ptr(), genome_len,
                                         frag_len, insert, sub_rate, n_rate, seed, seq.data_ptr(), off.data_ptr(), flag.data_ptr()))
    nonempty = int((np.diff(start) > 0).sum())
    reads = _lib.make_reads(seq, off, flag, n_reads=n_reads, n_bytes=n_bytes)
    return {"seq": seq[:n_bytes], "_seq_full": seq, "off": off, "flag": flag[:n_reads], "reads": reads, "n_groups": nonempty + 1,
            "n_pairs": n_pairs, "read_len": read_len, "bc_start": start, "n_bytes": n_bytes, "n_reads": n_reads}


def barcode_label(b: int, length=16) -> bytes:
    """barcode index -> ACGT string whose byte order equals the index order (LANG=C sort)."""
    return bytes(b"ACGT"[(b >> (2 * (length - 1 - i))) & 3] for i in range(length))


def write_sample_fastq(path, seq_host, read_len, bc_start, n_pairs):
    """first n_pairs pairs of the batch as the interleaved, barcode-sorted FASTQ pangaea.py -i gets."""
    rl = read_len + 1
    q = b"I" * read_len
    bc_of_pair = np.searchsorted(bc_start, np.arange(n_pairs), side="right") - 1
    mv = memoryview(seq_host)
    with open(path, "wb") as f:
        for p in range(n_pairs):
            h = b"@r%d\tBX:Z:%s-1\n" % (p, barcode_label(int(bc_of_pair[p])))
            for m in (0, 1):
                o = (2 * p + m) * rl
                f.write(h + bytes(mv[o:o + read_len]) + b"\n+\n" + q + b"\n")


def write_sample_fastq_fast(path, seq_host, read_len, bc_start, n_pairs):
    """Same file as write_sample_fastq, built with numpy (fixed-width read ids) - for the larger ingest sample."""
    rl = read_len + 1
    bc_of_pair = (np.searchsorted(bc_start, np.arange(n_pairs), side="right") - 1).astype(np.int64)
    bc_of_read = np.repeat(bc_of_pair, 2)
    ids = np.repeat(np.arange(n_pairs, dtype=np.int64), 2)
    n = 2 * n_pairs
    hdr = np.zeros((n, 2 + 10 + 6 + 16 + 3), dtype=np.uint8)
    hdr[:, 0:2] = np.frombuffer(b"@r", dtype=np.uint8)
    for d in range(10):
        hdr[:, 2 + d] = ord("0") + (ids // 10 ** (9 - d)) % 10
    hdr[:, 12:18] = np.frombuffer(b"\tBX:Z:", dtype=np.uint8)
    letters = np.frombuffer(b"ACGT", dtype=np.uint8)
    for i in range(16):
        hdr[:, 18 + i] = letters[(bc_of_read >> (2 * (15 - i))) & 3]
    hdr[:, 34:37] = np.frombuffer(b"-1\n", dtype=np.uint8)
    rec = np.empty((n, hdr.shape[1] + rl + 2 + rl), dtype=np.uint8)
    rec[:, :hdr.shape[1]] = hdr
    o = hdr.shape[1]
    rec[:, o:o + read_len] = np.asarray(seq_host[: n * rl]).reshape(n, rl)[:, :read_len]
    rec[:, o + read_len] = ord("\n")
    rec[:, o + rl:o + rl + 2] = np.frombuffer(b"+\n", dtype=np.uint8)
    rec[:, o + rl + 2:o + rl + 2 + read_len] = ord("I")
    rec[:, -1] = ord("\n")
    rec.tofile(path)


def ingest_from_fastq(args, ctx, batch_data, n_pairs=2_000_000):
    """The drop-in call a Pangaea user makes: a barcode-sorted interleaved FASTQ on disk -> feature matrices on the host
    (pg_fastq_parse with all host cores + pg_extract_features + copy back), on the first n_pairs pairs of the batch."""
    from pangaea_b200 import _lib

    n = min(n_pairs, batch_data["n_pairs"])
    rl = batch_data["read_len"] + 1
    host = batch_data["seq"][: 2 * n * rl].cpu().numpy()
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "sample.fq")
        write_sample_fastq_fast(path, host, batch_data["read_len"], batch_data["bc_start"], n)
        size = os.path.getsize(path)
        best = None
        for _ in range(3):  # first pass warms the page cache and the ctx workspaces
            t0 = time.perf_counter()
            fq = _lib.Fastq(path)
            t1 = time.perf_counter()
            f = ctx.extract_features(fq.reads, fq.group_keep, fq.n_groups)
            f.normalized()
            t2 = time.perf_counter()
            rows = f.rows
            f.free(); fq.close()
            if best is None or t2 - t0 < best[0]:
                best = (t2 - t0, t1 - t0, t2 - t1)
    return {"value": round(2 * n / best[0], 1), "unit": UNIT, "parse_s": round(best[1], 3), "gpu_and_copies_s": round(best[2], 3),
            "file_GB": round(size / 1e9, 3), "parse_GBps": round(size / 1e9 / best[1], 2), "host_threads": os.cpu_count(), "rows": rows,
            "sample": f"first {n} pairs of the batch as a plain-text interleaved FASTQ on local disk (page cache), best of 3"}


# --------------------------------------------------------------------------------------
# CPU reference arm: jellyfish stand-in + oracle/_ref/count_kmer ‖ oracle/_ref/count_tnf
# --------------------------------------------------------------------------------------
def run_reference_cpu(fastq, n_pairs, workdir, threads):
    """One pass of the reference's step 1 as src/feature.py:28-39 schedules it: the abundance
    chain (jellyfish count+dump -> count_kmer) and count_tnf run concurrently.  jellyfish is
    not installed: the oracle's C counter stands in for it, labelled as such.
    Returns (wall seconds, detail dict)."""
    from oracle import oracle as O

    if not O.have_ref():
        raise RuntimeError("oracle/_ref binaries missing (built by __graft_entry__.build() where /root/reference exists)")
    detail = {}

    def chain_abundance():
        t0 = time.perf_counter()
        table = O.count_fastq(fastq, 15)
        dump = os.path.join(workdir, "k15.dump")
        table.write_dump(dump, 15)
        detail["jellyfish_standin_s"] = time.perf_counter() - t0
        t1 = time.perf_counter()
        subprocess.run([O.REF_COUNT_KMER, "-i", fastq, "-t", str(threads), "-g", dump, "-k", "15", "-l", "2000", "-w", "10", "-v", "400",
                        "-o", os.path.join(workdir, "abd.gz")], check=True, stdout=subprocess.DEVNULL)
        detail["count_kmer_s"] = time.perf_counter() - t1

    def chain_tnf():
        t0 = time.perf_counter()
        subprocess.run([O.REF_COUNT_TNF, "-i", fastq, "-k", "4", "-t", str(threads), "-l", "2000", "-o", os.path.join(workdir, "tnf.gz")],
                       check=True, stdout=subprocess.DEVNULL)
        detail["count_tnf_s"] = time.perf_counter() - t0

    t0 = time.perf_counter()
    th = [threading.Thread(target=chain_abundance), threading.Thread(target=chain_tnf)]
    [t.start() for t in th]
    [t.join() for t in th]
    wall = time.perf_counter() - t0
    return wall, {k: round(v, 3) for k, v in detail.items()}


def synth_host_sample(n_pairs, read_len, seed):
    """CPU-only sample of the same model (used by --impl reference, which must not need a GPU)."""
    from pangaea_b200 import synth

    d = synth.generate(n_barcodes=max(1, n_pairs // 100), mean_pairs=100, read_len=read_len, n_genomes=200, genome_len=3_000_000,
                       frag_len=50_000, seed=seed)
    return d


# --------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        if os.environ.get("PG_NO_CLOCKS") == "1":  # experiment switch: does the sampler perturb the run?
            return
        try:
            self.path = tempfile.mktemp(prefix="clocks_", suffix=".csv")
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for n, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.path)
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        return out


# --------------------------------------------------------------------------------------
# algorithmic bytes (DESIGN.md "Roofline")
# --------------------------------------------------------------------------------------
def algorithmic_bytes(n_bytes, entries_count, windows_count, windows_feat, rows, sliced, vs=400, td=136, shared=False, count_segments=1,
                      table_bytes=2 ** 31):
    """Bytes each kernel has to move by design (DESIGN.md "Kernels and rooflines").  stream =
    2-bit codes + 1 validity bit per base position.  With the L2-sliced table (k = 15) a pass is
    two kernels: scatter writes one entry per window, apply reads it back and touches the counter."""
    stream = 0.25 + 0.125
    out = 4.0 * rows * (vs + td)
    b = {"pack": n_bytes * (1.0 + 0.25 + 2 * 0.125), "normalize": 2 * out}
    if sliced:
        # shared partition (one scatter for both passes): two mask streams in, 4 B per window + one i32 per 32 entries out
        b["count_scatter"] = n_bytes * (stream + (0.125 + 0.125 if shared else 0.0)) + (4.125 if shared else 4.0) * entries_count
        b["count_split"] = 4.0 * entries_count + 2.0 * entries_count       # second partition level: u32 in, u16 out
        # shared-memory sub-slice tables: 2 B per entry in, the table read and written once per segment
        b["count_apply"] = 2.0 * entries_count + 2.0 * table_bytes * count_segments
        b["tnf"] = n_bytes * stream + 4.0 * rows * td
        if not shared:
            b["feat_scatter"] = n_bytes * stream + 4.125 * windows_feat
        b["feat_apply"] = 4.125 * windows_feat + 4.0 * windows_feat + 4.0 * rows * vs   # entry + u32 counter read per window, tallies out
    else:
        b["count_apply"] = n_bytes * stream + 8.0 * windows_count
        b["feat_apply"] = n_bytes * stream + 4.0 * windows_feat + out
    return b


KERNEL_OF_STAGE = {"pack": "pack_kernel", "count_scatter": "bucket_scatter_kernel<15,shared>", "count_split": "bucket_split_kernel",
                   "count_apply": "sub_apply_kernel", "group": "flag_count/tile_scan/group_starts/row_assign/word_groups kernels", "tnf": "tnf_kernel<4>",
                   "feat_scatter": "bucket_scatter_kernel<15,feat>", "feat_apply": "bucket_apply_feat_kernel", "normalize": "normalize_rows_kernel"}
STAGE_SLOTS = (("pack", 0), ("count_scatter", 6), ("count_split", 9), ("count_apply", 1), ("group", 2), ("tnf", 8), ("feat_scatter", 7), ("feat_apply", 3),
               ("normalize", 4))


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic():
    p = os.path.join(ROOT, "profiles", "traffic.json")
    return json.load(open(p)) if os.path.isfile(p) else {}


# --------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=50_000_000, help="read pairs per GPU (BASELINE configs[1]: 50M)")
    ap.add_argument("--read-len", type=int, default=100)
    ap.add_argument("--cpu-sample-pairs", type=int, default=100_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": WORKLOAD if args.pairs == 50_000_000 and args.read_len == 100 else f"synthetic stLFR 2x{args.read_len}bp, {args.pairs} read pairs per GPU",
              "pairs_per_gpu": args.pairs, "read_len": args.read_len, "barcodes_per_gpu": max(1, args.pairs // 100), "k": 15, "tnf_k": 4,
              "window": 10, "vector": 400, "min_length": 2000, "l2": "inputs (>=10 GB per step) far exceed the 126 MB L2; no flush needed"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        return reference_arm(args, config)

    import torch
    import torch.distributed as dist

    from pangaea_b200 import _lib

    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    ctx = _lib.Context(device=local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=f"cuda:{local_rank}")
    batch_data = make_synthetic_batch(ctx, args.pairs, args.read_len, seed=2 + rank)
    n_groups = batch_data["n_groups"]
    keep = np.ones(n_groups, dtype=np.uint8)
    keep[0] = 0
    table_t = ctx.table_as_torch() if world > 1 else None

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    debug = os.environ.get("PG_DEBUG_STEP") == "1"
    debug_ev = [] if os.environ.get("PG_DEBUG_STEP") == "2" else None

    def step_device():
        """inputs resident in HBM -> normalised matrices in HBM"""
        marks = [("start", time.perf_counter())]

        def mark(name):
            if debug:
                ctx.synchronize()
                marks.append((name, time.perf_counter()))
            elif debug_ev is not None:  # no syncs: CUDA events on the ctx stream, read after the run
                e = torch.cuda.Event(enable_timing=True)
                with torch.cuda.stream(stream):
                    e.record()
                debug_ev.append((name, e, time.perf_counter()))

        ctx.table_clear(); mark("clear")
        b = ctx.adopt(batch_data["reads"]); mark("adopt+pack")
        ctx.count(b); mark("count")
        if world > 1:  # the one exchange step of the path: sum the dense count tables (overlaps grouping + TNF of featurize)
            ctx.all_reduce_table(table_t)
        f = ctx.featurize(b, keep); mark("featurize")
        f.normalize(); mark("normalize")
        b.free(); mark("free")
        if debug:
            free_b, total_b = torch.cuda.mem_get_info()
            print("step:", " ".join(f"{n}={1e3 * (t - marks[i][1]):.1f}" for i, (n, t) in enumerate(marks[1:])), f"free={free_b / 2**30:.1f}GiB", file=sys.stderr, flush=True)
        return f

    for _ in range(args.warmup):
        f = step_device()
        rows = f.rows
        f.free()
    barrier()
    ctx.timing_reset()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.perf_counter()
    with torch.cuda.stream(stream):
        ev0.record()
    for _ in range(args.steps):
        f = step_device()
        f.free()
    with torch.cuda.stream(stream):
        ev1.record()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    dev_ms = ev0.elapsed_time(ev1)
    if debug_ev:
        timed = [x for x in debug_ev if x[2] >= t0]
        prev_e, prev_t = ev0, t0
        line = []
        for name, e, t in timed:
            line.append(f"{name}:dev={prev_e.elapsed_time(e):.1f}/host={1e3 * (t - prev_t):.1f}")
            prev_e, prev_t = e, t
            if name == "free":
                print("step:", " ".join(line), file=sys.stderr, flush=True)
                line = []
    clocks = sampler.stop() if rank == 0 else None
    stage_ms = {n: ctx.timing(w)[0] / args.steps for n, w in STAGE_SLOTS}
    stage_launches = {n: ctx.timing(w)[1] // args.steps for n, w in STAGE_SLOTS}
    launches = ctx.timing(_lib.T_ALL)[1]

    t = torch.tensor([wall_ms, dev_ms], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall_ms, dev_ms = float(t[0]), float(t[1])
    total_reads = 2 * args.pairs * world * args.steps
    value = total_reads / (wall_ms / 1e3)

    # ---- roofline of the dominant kernel (live CUDA-event durations of this run) ----
    f = step_device()
    ctx.synchronize()
    windows_count = int(ctx.table_as_torch().to(torch.int64).sum()) if world == 1 else None
    a_raw = f.torch(_lib.ABD_RAW)
    windows_feat = int(a_raw.to(torch.int64).sum())  # look-ups that landed in a bin (>= 99.9 % of windows at this depth)
    del a_raw
    f.free()
    if windows_count is None:
        windows_count = windows_feat
    sliced = stage_ms["count_scatter"] > 0
    shared = sliced and stage_ms["feat_scatter"] == 0
    alg = algorithmic_bytes(batch_data["n_bytes"], windows_count, windows_count, windows_feat, rows, sliced, shared=shared,
                            count_segments=max(1, stage_launches["count_apply"] // 2))
    peak, peak_src = load_peaks()
    dom = max((n for n in alg if n in stage_ms), key=lambda n: stage_ms[n])
    kname = KERNEL_OF_STAGE[dom] if sliced else {"count_apply": "count_kernel", "feat_apply": "featurize_kernel"}.get(dom, KERNEL_OF_STAGE[dom])
    # spans of the scatter / split / count-apply stages also hold one housekeeping launch per segment (reset, fill save, item scan)
    per_seg = {"count_scatter": 3 if shared else 2, "feat_scatter": 2, "count_split": 2, "count_apply": 2}.get(dom, 1)
    n_launch = max(1, stage_launches[dom] // per_seg)
    achieved = alg[dom] / (stage_ms[dom] / 1e3) / 1e9
    b_pair = 2456.0 if args.read_len == 100 else 3832.0
    roofline = {"bound": "hbm", "kernel": kname, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": load_traffic().get(kname), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": int(alg[dom] / n_launch), "launches_per_step": n_launch,
                "ms_per_launch": round(stage_ms[dom] / n_launch, 3),
                "stages_ms": {k: round(v, 3) for k, v in stage_ms.items()},
                "stages_GBps": {k: round(alg[k] / (stage_ms[k] / 1e3) / 1e9, 1) for k in alg if stage_ms.get(k, 0) > 0},
                "whole_path": {"B_per_pair": b_pair, "achieved_GBps": round(b_pair * (value / world / 2) / 1e9, 1),
                               "frac": round(b_pair * (value / world / 2) / (peak * 1e9), 4),
                               "note": "SURVEY.md §8d algorithmic bytes per pair x pairs/s per GPU / measured HBM copy bandwidth"}}

    # ---- e2e: host buffers through the C-ABI ----
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, ctx, batch_data, keep, rows, world, local_rank, barrier, table_t)

    out = {"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": round(wall_ms / args.steps, 3), "device_ms_per_step": round(dev_ms / args.steps, 3), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": config, "clocks": clocks,
           "gpu_launches": launches, "rows_per_gpu": rows, "roofline": roofline}
    if e2e:
        out["e2e"] = e2e
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            out["from_fastq"] = ingest_from_fastq(args, ctx, batch_data)
        except Exception as e:  # informational: never lose the bench line over it
            out["from_fastq"] = {"value": None, "error": str(e)[:200]}
        out["cpu_baseline"] = cpu_baseline_from_batch(args, batch_data)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_e2e(args, ctx, batch_data, keep, rows, world, local_rank, barrier, table_t):
    """Same metric through pg_extract_features-equivalent calls with pinned HOST buffers:
    every step copies the reads host->device and the normalised matrices device->host."""
    import torch
    import torch.distributed as dist

    from pangaea_b200 import _lib

    h_seq = torch.empty(batch_data["n_bytes"], dtype=torch.uint8, pin_memory=True)
    h_off = torch.empty(batch_data["n_reads"] + 1, dtype=torch.int64, pin_memory=True)
    h_flag = torch.empty(batch_data["n_reads"], dtype=torch.uint8, pin_memory=True)
    h_seq.copy_(batch_data["seq"]); h_off.copy_(batch_data["off"]); h_flag.copy_(batch_data["flag"])
    torch.cuda.synchronize()
    h_abd = torch.empty((rows, 400), dtype=torch.float32, pin_memory=True)
    h_tnf = torch.empty((rows, 136), dtype=torch.float32, pin_memory=True)
    h_w = torch.empty(rows, dtype=torch.float64, pin_memory=True)
    reads = _lib.make_reads(h_seq, h_off, h_flag, n_reads=batch_data["n_reads"], n_bytes=batch_data["n_bytes"])
    h2d = batch_data["n_bytes"] + 8 * (batch_data["n_reads"] + 1) + batch_data["n_reads"] + len(keep)
    d2h = rows * (400 + 136) * 4 + rows * 8

    def step_host():
        if world == 1:
            f = ctx.extract_features(reads, keep)  # the one C-ABI call: upload + count + featurize + normalize
        else:
            ctx.table_clear()
            b = ctx.upload(reads)
            ctx.count(b)
            ctx.all_reduce_table(table_t)
            f = ctx.featurize(b, keep)
            b.free()
        f.normalized(h_abd, h_tnf, h_w)
        f.free()

    steps = max(1, min(args.steps, 5))
    step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step_host()
    barrier()
    ms = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    return {"value": round(2 * args.pairs * world * steps / (ms / 1e3), 1), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(d2h), "ms_per_step": round(ms / steps, 3), "steps": steps,
            "api": "pg_extract_features + pg_features_copy_normalized (pinned host buffers)"}


def cpu_baseline_from_batch(args, batch_data):
    """The reference's CPU tools on the first cpu_sample_pairs pairs of the very batch the GPU ran."""
    n = min(args.cpu_sample_pairs, batch_data["n_pairs"])
    rl = batch_data["read_len"] + 1
    host = batch_data["seq"][: 2 * n * rl].cpu().numpy()
    threads = os.cpu_count() or 1
    with tempfile.TemporaryDirectory() as d:
        fq = os.path.join(d, "sample.fq")
        write_sample_fastq(fq, host, batch_data["read_len"], batch_data["bc_start"], n)
        try:
            wall, detail = run_reference_cpu(fq, n, d, threads)
        except Exception as e:  # keep the GPU line even if the CPU tools are missing
            return {"value": None, "unit": UNIT, "cores": threads, "kind": "reference", "sample": f"failed: {e}"}
    return {"value": round(2 * n / wall, 1), "unit": UNIT, "cores": threads, "kind": "reference",
            "sample": f"first {n} pairs of the benchmark batch as plain-text interleaved FASTQ; count_kmer ‖ count_tnf (-t {threads}) as "
                      f"src/feature.py:28-39 runs them; jellyfish (absent) replaced by the oracle's single-thread C counter",
            "seconds": round(wall, 3), "detail": detail}


def reference_arm(args, config):
    """--impl reference: the reference's own CPU implementation of the path on the host cores."""
    # bounded sample: sized so that the whole --steps K --warmup W run stays within a few minutes
    # (one pass over 100k pairs takes ~30 s on 8 cores, most of it count_kmer's serial dump load)
    n = int(min(args.cpu_sample_pairs, max(20_000, args.cpu_sample_pairs * 5 // max(1, args.steps + args.warmup))))
    threads = os.cpu_count() or 1
    from pangaea_b200 import synth

    data = synth_host_sample(n, args.read_len, seed=2)
    n = data["n_pairs"]
    times, detail = [], {}
    with tempfile.TemporaryDirectory() as d:
        fq = synth.write_interleaved(os.path.join(d, "sample.fq"), data)
        for i in range(args.warmup + args.steps):
            wall, detail = run_reference_cpu(fq, n, d, threads)
            if i >= args.warmup:
                times.append(wall)
    total = sum(times)
    value = 2 * n * len(times) / total
    sample = (f"{n} pairs of the same synthetic model (seed 2) per step, plain-text interleaved FASTQ; count_kmer ‖ count_tnf (-t {threads}) "
              f"as src/feature.py:28-39 runs them; jellyfish (absent) replaced by the oracle's single-thread C counter")
    out = {"impl": "reference", "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": round(1e3 * total / len(times), 3), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": config,
           "cpu_baseline": {"value": round(value, 1), "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample, "detail": detail},
           "e2e": {"value": round(value, 1), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(out), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())

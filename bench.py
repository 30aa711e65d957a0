#!/usr/bin/env python
"""bench.py - reads/sec featurized (k-mer count + abundance + TNF) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c3|c4|c5]

One "step" = one pass of the whole hot path over the workload: clear table -> 2-bit pack -> count
canonical 15-mers -> group clouds -> abundance histogram + TNF -> L1 normalise.

--config (BASELINE.json configs[1..4]; SURVEY.md §8d shapes):
  c2 (default)  synthetic stLFR 2x100bp, 50M read pairs, ~500k barcodes, 1 B200.  With N > 1 every rank holds the same
                amount (weak scaling), counts its shard, the dense count tables are summed with one NCCL all-reduce and
                every rank featurizes its own clouds.  This is the driver's line.
  c3            synthetic TELL-Seq 2x150bp, 200M read pairs in total, 2M barcodes (18-bp labels), 400 genomes.
  c4            hybrid-mode 2x150bp, 300M read pairs in total, one virtual barcode per pair, -l 0: one row per pair.
  c5            synthetic stLFR 2x100bp, 1B read pairs in total, 5M barcodes, 1000 genomes.
  c3-c5 split the TOTAL over the ranks (strong scaling) and stream each rank's share through the GPU in batches, the way
  Feature.extract_features streams a file (pangaea_b200/stream.py): count pass over all batches (the packed batches stay
  in HBM), all-reduce, featurize pass batch by batch, outputs folded into a checksum.  Batches are generated on the device
  just before they are used; only the processing is timed (CUDA events), so `ms_per_step` is the sum of the timed spans.

Prints ONE JSON line (DESIGN.md "Measurement").  `value` = device-resident input, `e2e` = the same metric through the
C-ABI with pinned HOST buffers (H2D of the reads and D2H of the normalised matrices inside the timed region).
`--impl reference` times the reference's own CPU tools (oracle/_ref, compiled from /root/reference) on a bounded sample.
"""
from __future__ import annotations

import argparse
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
PARITY_NOTE = ("tests/: abundance, TNF, grouping, normalisation bit-exact against the unmodified reference tools (golden vectors) and the oracle; "
               "step 1a (global k-mer counting) is pinned only to the oracle's restatement of jellyfish, which is absent from the reference "
               "tree and this image (DESIGN.md §2: parity unpinned)")

CONFIGS = {
    # pairs: per GPU for "weak", in total for "strong"; batch_pairs: pairs per batch of the streamed configs
    "c2": dict(workload="synthetic stLFR 2x100bp, 50M read pairs, ~500k barcodes, 1 B200", pairs=50_000_000, read_len=100,
               pairs_per_barcode=100, n_genomes=200, seed=2, min_length=2000, scaling="weak", barcode_len=16, batch_pairs=50_000_000),
    "c3": dict(workload="synthetic TELL-Seq 2x150bp with barcode index, 200M read pairs, 2/4 B200", pairs=200_000_000, read_len=150,
               pairs_per_barcode=100, n_genomes=400, seed=3, min_length=2000, scaling="strong", barcode_len=18, batch_pairs=25_000_000),
    "c4": dict(workload="hybrid-mode plain short reads 2x150bp (no barcodes), 300M read pairs, per-read abundance features",
               pairs=300_000_000, read_len=150, pairs_per_barcode=1, n_genomes=400, seed=4, min_length=0, scaling="strong", barcode_len=18,
               batch_pairs=7_000_000),  # two full segments per batch; 30 GB of matrices per batch leave room in HBM next to the 43 packed batches
    "c5": dict(workload="synthetic stLFR 2x100bp, 1B read pairs, 5M barcodes, k-mer table sharded over 8 B200", pairs=1_000_000_000,
               read_len=100, pairs_per_barcode=200, n_genomes=1000, seed=5, min_length=2000, scaling="strong", barcode_len=16,
               batch_pairs=62_500_000),
}
# SURVEY.md §8d algorithmic bytes per pair, by pass: pack (ASCII in, 2-bit out), count (packed in + u32 counter read+write per
# window), featurize (packed in + u32 counter read per window), rows (int32 tallies + f32 normalised out)
def algorithmic_bytes_per_pair(read_len, rows_per_pair, vs=400, td=136):
    L = read_len
    n_k = 2 * (L - 14)
    return {"pack": 2 * L + 2 * L / 4, "count": 2 * L / 4 + 8 * n_k, "featurize": 2 * L / 4 + 4 * n_k,
            "rows_raw": 4.0 * (vs + td) * rows_per_pair, "rows_norm": 4.0 * (vs + td) * rows_per_pair}


# measured ceilings of the access patterns the kernels are built on (tools/microbench*.cu, profiles/microbench*_r01.txt), ops/s
ATTAINABLE = {"l2_gather": 289e9, "smem_slot_handout": 1212e9, "smem_atomic": 2370e9}


def make_synthetic_batch(ctx, n_pairs, **kw):
    """SURVEY.md §8d model generated in HBM (pangaea_b200/synth.py: device_batch)."""
    from pangaea_b200 import synth

    return synth.device_batch(ctx, n_pairs, **kw)


# --------------------------------------------------------------------------------------
# CPU reference arm: jellyfish stand-in + oracle/_ref/count_kmer ‖ oracle/_ref/count_tnf
# --------------------------------------------------------------------------------------
def run_reference_cpu(fastq, workdir, threads, min_length=2000):
    """One pass of the reference's step 1 as src/feature.py:28-39 schedules it: the abundance chain (jellyfish count+dump ->
    count_kmer) and count_tnf run concurrently.  jellyfish is not installed: the oracle's C counter stands in
    for it, labelled as such.  Returns (wall seconds, detail dict)."""
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
        subprocess.run([O.REF_COUNT_KMER, "-i", fastq, "-t", str(threads), "-g", dump, "-k", "15", "-l", str(min_length), "-w", "10", "-v", "400",
                        "-o", os.path.join(workdir, "abd.gz")], check=True, stdout=subprocess.DEVNULL)
        detail["count_kmer_s"] = time.perf_counter() - t1

    def chain_tnf():
        t0 = time.perf_counter()
        subprocess.run([O.REF_COUNT_TNF, "-i", fastq, "-k", "4", "-t", str(threads), "-l", str(min_length), "-o", os.path.join(workdir, "tnf.gz")],
                       check=True, stdout=subprocess.DEVNULL)
        detail["count_tnf_s"] = time.perf_counter() - t0

    t0 = time.perf_counter()
    th = [threading.Thread(target=chain_abundance), threading.Thread(target=chain_tnf)]
    [t.start() for t in th]
    [t.join() for t in th]
    wall = time.perf_counter() - t0
    return wall, {k: round(v, 3) for k, v in detail.items()}


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
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
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
# bytes each kernel moves BY DESIGN - an implementation figure, reported beside the §8d roofline, not as it
# --------------------------------------------------------------------------------------
def design_bytes(n_bytes, entries_count, windows_feat, rows, sliced, vs=400, td=136, shared=False, count_segments=1, table_bytes=2 ** 31):
    stream = 0.25 + 0.125
    out = 4.0 * rows * (vs + td)
    b = {"pack": n_bytes * (1.0 + 0.25 + 2 * 0.125), "normalize": 2 * out}
    if sliced:
        b["count_scatter"] = n_bytes * (stream + (0.125 + 0.125 if shared else 0.0)) + (4.125 if shared else 4.0) * entries_count
        b["count_split"] = 4.0 * entries_count + 2.0 * entries_count
        b["count_apply"] = 2.0 * entries_count + 2.0 * table_bytes * count_segments
        b["tnf"] = n_bytes * stream + 4.0 * rows * td
        if not shared:
            b["feat_scatter"] = n_bytes * stream + 4.125 * windows_feat
        b["feat_apply"] = 4.0 * windows_feat + 4.0 * windows_feat + 4.0 * windows_feat  # entry in, u32 counter gathered, bin written back
        b["feat_collect"] = 4.0 * windows_feat + 4.0 * rows * vs                            # entries in stream order, tallies out
    else:
        b["count_apply"] = n_bytes * stream + 8.0 * entries_count
        b["feat_apply"] = n_bytes * stream + 4.0 * windows_feat + out
    return b


KERNEL_OF_STAGE = {"pack": "pack_kernel", "count_scatter": "bucket_scatter_kernel<15,shared>", "count_split": "bucket_split_kernel",
                   "count_apply": "sub_apply_kernel", "group": "flag_count/tile_scan/group_starts/row_assign/word_groups kernels", "tnf": "tnf_kernel<4>",
                   "feat_scatter": "bucket_scatter_kernel<15,feat>", "feat_apply": "bucket_apply_feat_kernel", "feat_collect": "bucket_collect_kernel",
                   "normalize": "normalize_rows_vec_kernel"}
STAGE_SLOTS = (("pack", 0), ("count_scatter", 6), ("count_split", 9), ("count_apply", 1), ("group", 2), ("tnf", 8), ("feat_scatter", 7), ("feat_apply", 3),
               ("feat_collect", 10), ("normalize", 4))
# which §8d pass a kernel belongs to
PASS_OF_STAGE = {"pack": "pack", "count_scatter": "count", "count_split": "count", "count_apply": "count", "feat_scatter": "featurize",
                 "feat_apply": "featurize", "feat_collect": "featurize", "tnf": "featurize", "normalize": "rows_norm"}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic():
    p = os.path.join(ROOT, "profiles", "traffic.json")
    return json.load(open(p)) if os.path.isfile(p) else {}


def build_roofline(stage_ms, stage_launches, pairs_per_step, read_len, rows, n_bytes, windows_count, windows_feat, value_per_gpu):
    """§8d roofline of the dominant kernel + the whole path, the by-design figure and the attainable bounds."""
    peak, peak_src = load_peaks()
    sliced = stage_ms["count_scatter"] > 0
    shared = sliced and stage_ms["feat_scatter"] == 0
    per_pair = algorithmic_bytes_per_pair(read_len, rows / max(1, pairs_per_step))
    pass_bytes = {k: v * pairs_per_step for k, v in per_pair.items()}
    pass_bytes["featurize"] += pass_bytes.pop("rows_raw")  # the int32 tallies are written by the featurize pass
    timed = {k: v for k, v in stage_ms.items() if k in PASS_OF_STAGE and v > 0}
    dom = max(timed, key=lambda n: timed[n])
    kernel_of = dict(KERNEL_OF_STAGE)
    if stage_ms.get("feat_collect", 0) > 0:  # tiny clouds: the featurize pass is look-up + collect instead of the single sweep
        kernel_of["feat_apply"] = "bucket_lookup_kernel"
    kname = kernel_of[dom] if sliced else {"count_apply": "count_kernel", "feat_apply": "featurize_kernel"}.get(dom, kernel_of[dom])
    dom_pass = PASS_OF_STAGE[dom]
    # the dominant kernel is charged with its whole pass's §8d bytes (the other kernels of the pass are listed beside it)
    per_seg = {"count_scatter": 3 if shared else 2, "feat_scatter": 2, "count_split": 2, "count_apply": 3}.get(dom, 1)
    n_launch = max(1, stage_launches[dom] // per_seg)
    achieved = pass_bytes[dom_pass] / (stage_ms[dom] / 1e3) / 1e9
    design = design_bytes(n_bytes, windows_count, windows_feat, rows, sliced, shared=shared,
                          count_segments=max(1, stage_launches["count_apply"] // 3))
    pass_ms = {}
    for st, ms in timed.items():
        if st == "tnf" and sliced:
            continue  # runs on the second stream next to the look-up sweep
        pass_ms[PASS_OF_STAGE[st]] = pass_ms.get(PASS_OF_STAGE[st], 0.0) + ms
    b_pair = sum(per_pair.values())
    floors = {}
    if sliced:
        floors = {"feat_apply": ("l2_gather", windows_feat), "count_scatter": ("smem_slot_handout", windows_count),
                  "count_split": ("smem_slot_handout", windows_count), "count_apply": ("smem_atomic", windows_count)}
    attainable = {}
    for st, (what, ops) in floors.items():
        if stage_ms.get(st, 0) > 0:
            floor_ms = ops / ATTAINABLE[what] * 1e3
            attainable[kernel_of[st]] = {"bound": what, "ops_per_s": ATTAINABLE[what], "floor_ms": round(floor_ms, 2),
                                                "ms": round(stage_ms[st], 2), "frac_of_attainable": round(floor_ms / stage_ms[st], 3)}
    for st in ("pack", "normalize"):
        if stage_ms.get(st, 0) > 0:
            floor_ms = design[st] / (peak * 1e9) * 1e3
            attainable[kernel_of[st]] = {"bound": "hbm_stream", "floor_ms": round(floor_ms, 2), "ms": round(stage_ms[st], 2),
                                                "frac_of_attainable": round(floor_ms / stage_ms[st], 3)}
    return {"bound": "hbm", "kernel": kname, "pass": dom_pass, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
            "frac": round(achieved / peak, 4), "traffic": load_traffic().get(kname), "peak_source": peak_src,
            "basis": "SURVEY.md §8d algorithmic bytes of the kernel's pass / the kernel's CUDA-event time in this run",
            "algorithmic_bytes_per_pair": {k: round(v, 1) for k, v in per_pair.items()},
            "algorithmic_bytes_per_launch": int(pass_bytes[dom_pass] / n_launch), "launches_per_step": n_launch,
            "ms_per_launch": round(stage_ms[dom] / n_launch, 3),
            "stages_ms": {k: round(v, 3) for k, v in stage_ms.items()},
            "passes": {p: {"ms": round(ms, 3), "algorithmic_GB": round(pass_bytes[p] / 1e9, 2), "GBps": round(pass_bytes[p] / (ms / 1e3) / 1e9, 1),
                           "frac": round(pass_bytes[p] / (ms / 1e3) / 1e9 / peak, 4)} for p, ms in pass_ms.items() if p in pass_bytes},
            "whole_path": {"B_per_pair": round(b_pair, 1), "achieved_GBps": round(b_pair * (value_per_gpu / 2) / 1e9, 1),
                           "frac": round(b_pair * (value_per_gpu / 2) / (peak * 1e9), 4),
                           "note": "SURVEY.md §8d algorithmic bytes per pair x pairs/s per GPU / measured HBM copy bandwidth"},
            "by_design": {"note": "bytes each kernel moves by design (partition entries included) - an implementation figure, not the roofline",
                          "GBps": {k: round(design[k] / (stage_ms[k] / 1e3) / 1e9, 1) for k in design if stage_ms.get(k, 0) > 0}},
            "attainable": attainable}


# --------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--pairs", type=int, default=None, help="override the configuration's read pairs (per GPU for c2, in total for c3-c5)")
    ap.add_argument("--batch-pairs", type=int, default=None)
    ap.add_argument("--read-len", type=int, default=None)
    ap.add_argument("--cpu-sample-pairs", type=int, default=100_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = dict(CONFIGS[args.config])
    exact = args.pairs is None and args.read_len is None
    if args.pairs is not None:
        cfg["pairs"] = args.pairs
    if args.read_len is not None:
        cfg["read_len"] = args.read_len
    if args.batch_pairs is not None:
        cfg["batch_pairs"] = args.batch_pairs
    cfg["name"] = args.config
    weak = cfg["scaling"] == "weak"
    pairs_rank = cfg["pairs"] if weak else cfg["pairs"] // world
    workload = cfg["workload"] if exact else f"{args.config} shape, {cfg['pairs']} read pairs {'per GPU' if weak else 'in total'}, 2x{cfg['read_len']}bp"
    config = {"workload": workload, "config": args.config, "pairs_per_gpu": pairs_rank, "read_len": cfg["read_len"],
              "barcodes_per_gpu": max(1, pairs_rank // cfg["pairs_per_barcode"]), "n_genomes": cfg["n_genomes"], "k": 15, "tnf_k": 4, "window": 10,
              "vector": 400, "min_length": cfg["min_length"], "l2": "inputs (>=10 GB per step) far exceed the 126 MB L2; no flush needed"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        return reference_arm(args, cfg, config)

    import torch
    import torch.distributed as dist

    from pangaea_b200 import _lib

    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    ctx = _lib.Context(device=local_rank, min_length=cfg["min_length"])
    env = dict(args=args, cfg=cfg, config=config, rank=rank, world=world, local_rank=local_rank, ctx=ctx, pairs_rank=pairs_rank)
    if weak and pairs_rank <= cfg["batch_pairs"]:
        out = run_resident(env)
    else:
        out = run_streamed(env)
    if world > 1 and not args.no_parity:
        parity = multi_rank_parity_check(ctx, rank, world, local_rank)
        if out is not None:
            out["parity_check"] = parity
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def multi_rank_parity_check(ctx, rank, world, local_rank):
    """Results, not just speed, at N > 1: a small barcode-sorted FASTQ goes through the real multi-rank path (every rank parses
    its byte range, counts, all-reduce, featurizes; rows gathered in rank order) and rank 0 compares with the oracle."""
    import torch.distributed as dist

    from pangaea_b200 import _lib, synth
    from pangaea_b200.distributed import extract_features_from_file, gather_rows_named

    path = os.path.join(tempfile.gettempdir(), f"pg_parity_{os.environ.get('MASTER_PORT', '0')}.fq")
    want = None
    if rank == 0:
        data = synth.generate(n_barcodes=400, mean_pairs=25, read_len=100, n_genomes=4, genome_len=100_000, frag_len=12_000, seed=99,
                              unbarcoded_pairs=30, n_rate=0.002)
        synth.write_interleaved(path, data)
    dist.barrier()
    t0 = time.perf_counter()
    names, feats = extract_features_from_file(ctx, path, batch_seq_bytes=400_000)
    merged = gather_rows_named(names, feats)
    feats.free()
    res = None
    if rank == 0:
        try:
            from oracle import oracle as O

            O.build(ref=False)
            want = O.featurize(path, None)
            ok = list(merged[0]) == list(want[0]) and np.array_equal(merged[1], want[1]) and np.array_equal(merged[2], want[2])
            res = {"ok": bool(ok), "rows": int(len(want[0])), "ranks": world, "checker": "oracle/pg_oracle.c over the whole file",
                   "path": "distributed.extract_features_from_file (per-rank byte ranges, batches of 400 kB, NCCL all-reduce of the tables)",
                   "seconds": round(time.perf_counter() - t0, 2)}
        except Exception as e:
            res = {"ok": None, "error": str(e)[:200]}
    dist.barrier()
    if rank == 0 and os.path.exists(path):
        os.unlink(path)
    return res


# --------------------------------------------------------------------------------------
# c2: one batch per rank, resident in HBM (the driver's line)
# --------------------------------------------------------------------------------------
def run_resident(env):
    import torch
    import torch.distributed as dist

    from pangaea_b200 import _lib, synth

    args, cfg, config, rank, world, local_rank, ctx = (env[k] for k in ("args", "cfg", "config", "rank", "world", "local_rank", "ctx"))
    pairs = env["pairs_rank"]
    stream = torch.cuda.ExternalStream(ctx.stream, device=f"cuda:{local_rank}")
    batch_data = synth.device_batch(ctx, pairs, cfg["read_len"], n_barcodes=max(1, pairs // cfg["pairs_per_barcode"]), n_genomes=cfg["n_genomes"],
                                    seed=cfg["seed"] + rank)
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

    def step_device():
        """inputs resident in HBM -> normalised matrices in HBM"""
        marks = [("start", time.perf_counter())]

        def mark(name):
            if debug:
                ctx.synchronize()
                marks.append((name, time.perf_counter()))

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
    clocks = sampler.stop() if rank == 0 else None
    stage_ms = {n: ctx.timing(w)[0] / args.steps for n, w in STAGE_SLOTS}
    stage_launches = {n: ctx.timing(w)[1] // args.steps for n, w in STAGE_SLOTS}
    launches = ctx.timing(_lib.T_ALL)[1]

    t = torch.tensor([wall_ms, dev_ms], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall_ms, dev_ms = float(t[0]), float(t[1])
    total_reads = 2 * pairs * world * args.steps
    value = total_reads / (wall_ms / 1e3)

    # ---- roofline (live CUDA-event durations of this run) ----
    f = step_device()
    ctx.synchronize()
    windows_count = int(ctx.table_as_torch().to(torch.int64).sum()) if world == 1 else None
    a_raw = f.torch(_lib.ABD_RAW)
    windows_feat = int(a_raw.to(torch.int64).sum())  # look-ups that landed in a bin (>= 99.9 % of windows at this depth)
    checksum = {"abd_sum": windows_feat, "tnf_sum": int(f.torch(_lib.TNF_RAW).to(torch.int64).sum())}
    del a_raw
    f.free()
    if windows_count is None:
        windows_count = windows_feat
    roofline = build_roofline(stage_ms, stage_launches, pairs, cfg["read_len"], rows, batch_data["n_bytes"], windows_count, windows_feat, value / world)

    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, ctx, batch_data, keep, rows, world, local_rank, barrier, table_t, pairs)

    out = {"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": round(wall_ms / args.steps, 3), "device_ms_per_step": round(dev_ms / args.steps, 3), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": config, "clocks": clocks,
           "gpu_launches": launches, "rows_per_gpu": rows, "checksum": checksum, "parity_status": PARITY_NOTE, "roofline": roofline}
    if e2e:
        out["e2e"] = e2e
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            out["from_fastq"] = ingest_from_fastq(args, ctx, batch_data, cfg)
        except Exception as e:  # informational: never lose the bench line over it
            out["from_fastq"] = {"value": None, "error": str(e)[:200]}
        out["cpu_baseline"] = cpu_baseline_from_batch(args, batch_data, cfg)
    return out if rank == 0 else None


def run_e2e(args, ctx, batch_data, keep, rows, world, local_rank, barrier, table_t, pairs):
    """Same metric through the C-ABI with pinned HOST buffers: every step copies the reads host->device (in chunks, overlapped
    with the count pass - pg_extract_features / pg_batch_upload_count) and the normalised matrices device->host."""
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
            b = ctx.upload_count(reads, keep_partition=True)  # the same pipelined upload + count, table left for the all-reduce
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
    return {"value": round(2 * pairs * world * steps / (ms / 1e3), 1), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(d2h), "ms_per_step": round(ms / steps, 3), "steps": steps,
            "api": "pg_extract_features (N=1) / pg_batch_upload_count + all-reduce + pg_featurize (N>1) + pg_features_copy_normalized, pinned host buffers"}


def ingest_from_fastq(args, ctx, batch_data, cfg, n_pairs=20_000_000):
    """The drop-in call a Pangaea user makes: a barcode-sorted interleaved FASTQ on disk -> feature matrices on the host
    (the streaming flow of Feature.extract_features: parse on all host cores into pinned batches, H2D overlapped with the
    count pass, featurize, copy back), on the first n_pairs pairs of the batch."""
    from pangaea_b200 import _lib, stream, synth

    n = min(n_pairs, batch_data["n_pairs"])
    rl = batch_data["read_len"] + 1
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "sample.fq")
        step = 2_000_000  # built slice by slice: bounded host memory
        for p0 in range(0, n, step):
            p1 = min(n, p0 + step)
            host = batch_data["seq"][2 * p0 * rl: 2 * p1 * rl].cpu().numpy()
            synth.write_batch_fastq(path, host, batch_data["read_len"], batch_data["bc_start"], p1 - p0, barcode_len=cfg["barcode_len"],
                                    first_pair=p0, append=p0 > 0)
        del host
        size = os.path.getsize(path)
        res = {}
        import torch

        n_rows_max = batch_data["n_barcodes"] + 2  # pinned host buffers for the matrices, as a caller that cares about speed would pass
        h_abd = torch.empty((n_rows_max, 400), dtype=torch.float32, pin_memory=True)
        h_tnf = torch.empty((n_rows_max, 136), dtype=torch.float32, pin_memory=True)
        h_w = torch.empty(n_rows_max, dtype=torch.float64, pin_memory=True)
        for name, run in (("device_ingest", lambda: stream.extract_features_device_ingest(ctx, path, window_bytes=1 << 30)),
                          ("host_parser", lambda: stream.extract_features_streaming(ctx, lambda: _lib.FastqStream(path, pinned=True, target_seq_bytes=1 << 30)))):
            best = None
            try:
                for _ in range(3):  # first pass warms the page cache, the pinned buffers and the ctx workspaces
                    t0 = time.perf_counter()
                    names, f = run()
                    f.normalized(h_abd[: f.rows], h_tnf[: f.rows], h_w[: f.rows])
                    t1 = time.perf_counter()
                    rows = f.rows
                    f.free()
                    best = t1 - t0 if best is None else min(best, t1 - t0)
                res[name] = {"reads_per_s": round(2 * n / best, 1), "seconds": round(best, 3), "file_GBps": round(size / 1e9 / best, 2), "rows": rows}
            except Exception as e:
                res[name] = {"reads_per_s": None, "error": str(e)[:200]}
    ok = [v["reads_per_s"] for v in res.values() if v.get("reads_per_s")]
    return {"value": max(ok) if ok else None, "unit": UNIT, "paths": res, "file_GB": round(size / 1e9, 3), "host_threads": os.cpu_count(),
            "sample": f"first {n} pairs of the batch as a plain-text interleaved FASTQ on local disk (page cache) -> normalised matrices on the host, "
                      f"1 GiB windows, best of 3.  device_ingest: raw text staged to pinned memory and parsed in HBM (pg_ingest_text); "
                      f"host_parser: csrc/fastq.cpp on all host cores (the path for gzip / paired / hostile input)"}


def cpu_baseline_from_batch(args, batch_data, cfg):
    """The reference's CPU tools on the first cpu_sample_pairs pairs of the very batch the GPU ran."""
    from pangaea_b200 import synth

    n = min(args.cpu_sample_pairs, batch_data["n_pairs"])
    rl = batch_data["read_len"] + 1
    host = batch_data["seq"][: 2 * n * rl].cpu().numpy()
    threads = os.cpu_count() or 1
    with tempfile.TemporaryDirectory() as d:
        fq = os.path.join(d, "sample.fq")
        synth.write_batch_fastq(fq, host, batch_data["read_len"], batch_data["bc_start"], n, barcode_len=cfg["barcode_len"])
        try:
            wall, detail = run_reference_cpu(fq, d, threads, cfg["min_length"])
        except Exception as e:  # keep the GPU line even if the CPU tools are missing
            return {"value": None, "unit": UNIT, "cores": threads, "kind": "reference", "sample": f"failed: {e}"}
    return {"value": round(2 * n / wall, 1), "unit": UNIT, "cores": threads, "kind": "reference",
            "sample": f"first {n} pairs of the benchmark batch as plain-text interleaved FASTQ; count_kmer ‖ count_tnf (-t {threads}) as "
                      f"src/feature.py:28-39 runs them; jellyfish (absent) replaced by the oracle's single-thread C counter",
            "seconds": round(wall, 3), "detail": detail}


# --------------------------------------------------------------------------------------
# c3 / c4 / c5 (and c2 shares too large for one batch): each rank's share streamed in batches
# --------------------------------------------------------------------------------------
def run_streamed(env):
    import torch
    import torch.distributed as dist

    from pangaea_b200 import _lib, synth

    args, cfg, config, rank, world, local_rank, ctx = (env[k] for k in ("args", "cfg", "config", "rank", "world", "local_rank", "ctx"))
    pairs = env["pairs_rank"]
    dev = f"cuda:{local_rank}"
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    bp = min(cfg["batch_pairs"], pairs)
    n_batches = (pairs + bp - 1) // bp
    ppb = cfg["pairs_per_barcode"]
    table_t = ctx.table_as_torch() if world > 1 else None
    seed = cfg["seed"]  # one community for all ranks and batches; bc_base / pair_base make the clouds distinct

    def gen(i):
        n = min(bp, pairs - i * bp)
        pair_base = rank * pairs + i * bp
        return synth.device_batch(ctx, n, cfg["read_len"], n_barcodes=max(1, n // ppb), n_genomes=cfg["n_genomes"], seed=seed,
                                  bc_base=pair_base // ppb, pair_base=pair_base)

    def span():
        e = torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e.record()
        return e

    # the batches keep the partition of their windows for the featurize pass when all of them fit beside the packed batches
    # (the same rule as pangaea_b200/stream.py: keep_partitions)
    from pangaea_b200 import stream as stream_mod

    keep_part = n_batches == 1 or (ppb > 1 and stream_mod.keep_partitions(ctx, 2.0 * pairs * (cfg["read_len"] + 1)))

    def one_step(collect=None):
        """-> (ms timed on this rank, rows, checksum).  Generation of a batch is outside the timed spans."""
        spans = []
        ctx.table_clear()
        held = []
        for i in range(n_batches):
            bd = gen(i)
            a = span()
            b = ctx.adopt(bd["reads"])
            ctx.count(b, keep_partition=keep_part)
            if n_batches > 1:
                b.compact()
            z = span()
            spans.append((a, z))
            ctx.synchronize()
            if os.environ.get("PG_BENCH_DEBUG") == "1":
                print(f"[bench] count batch {i}: span {a.elapsed_time(z):.1f} ms, driver free {torch.cuda.mem_get_info()[0] / 2**30:.1f} GiB, "
                      f"torch reserved {torch.cuda.memory_reserved() / 2**30:.1f} GiB", file=sys.stderr)
            keep = np.ones(bd["n_groups"], dtype=np.uint8)
            keep[0] = 0  # the cloud that is open when a file starts is labelled "" and dropped (for c4, under the reference's
            # off-by-one, row i then holds pair i + 1: boundary_mode "reference")
            held.append((b, keep, bd if n_batches == 1 else None))
            if n_batches > 1:
                del bd  # the ASCII goes back to torch's allocator; the packed stream stays in the batch
        a = span()
        if world > 1:
            ctx.all_reduce_table(table_t)
        rows, abd_sum, tnf_sum, w_sum = 0, 0, 0, 0.0
        z = span()
        spans.append((a, z))
        debug = os.environ.get("PG_BENCH_DEBUG") == "1"
        if n_batches > 1:
            torch.cuda.empty_cache()  # the generator's cached blocks (not part of the path) would crowd the matrices of the featurize pass
        for b, keep, bd in held:
            if debug:
                ctx.synchronize()
                drv_free, drv_total = torch.cuda.mem_get_info()
                t_host = time.perf_counter()
            a = span()
            f = ctx.featurize(b, keep)
            f.normalize()
            z = span()
            spans.append((a, z))
            if debug:
                ctx.synchronize()
                print(f"[bench] featurize batch: span {a.elapsed_time(z):.1f} ms, host {1e3 * (time.perf_counter() - t_host):.1f} ms, driver free before "
                      f"{drv_free / 2**30:.1f} GiB, ctx free {ctx.mem_info()[0] / 2**30:.1f} GiB, torch reserved {torch.cuda.memory_reserved() / 2**30:.1f} GiB",
                      file=sys.stderr)
            rows += f.rows
            if collect is not None:  # fold the outputs into a checksum (outside the timed spans)
                ctx.synchronize()
                abd_sum += int(f.torch(_lib.ABD_RAW).to(torch.int64).sum())
                tnf_sum += int(f.torch(_lib.TNF_RAW).to(torch.int64).sum())
                w_sum += float(f.torch(_lib.WEIGHTS).sum())
            f.free()
            b.free()
        ctx.synchronize()
        ms = sum(a.elapsed_time(z) for a, z in spans)
        return ms, rows, {"abd_sum": abd_sum, "tnf_sum": tnf_sum, "weights_sum": round(w_sum, 6)}

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(args.warmup):
        one_step()
        torch.cuda.empty_cache()
    barrier()
    ctx.timing_reset()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    total_ms, rows = 0.0, 0
    for _ in range(args.steps):
        barrier()
        ms, rows, _ = one_step()
        total_ms += ms
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    stage_ms = {n: ctx.timing(w)[0] / args.steps for n, w in STAGE_SLOTS}
    stage_launches = {n: ctx.timing(w)[1] // args.steps for n, w in STAGE_SLOTS}
    launches = ctx.timing(_lib.T_ALL)[1]
    _, _, checksum = one_step(collect=True)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    r = torch.tensor([rows, checksum["abd_sum"], checksum["tnf_sum"]], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(r, op=dist.ReduceOp.SUM)
    total_ms = float(t[0])
    value = 2 * pairs * world * args.steps / (total_ms / 1e3)
    n_bytes = 2 * pairs * (cfg["read_len"] + 1)
    windows = int(r[1]) // world
    roofline = build_roofline(stage_ms, stage_launches, pairs, cfg["read_len"], rows, n_bytes, windows, windows, value / world)
    config = dict(config, batches_per_gpu=int(n_batches), batch_pairs=int(bp), partitions_kept=bool(keep_part),
                  timing="sum of CUDA-event spans around the processing of each batch (count pass, all-reduce, featurize pass), max over ranks; "
                         "batches are generated on the device between the spans")
    out = {"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": round(total_ms / args.steps, 3), "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None,
           "dtype": "u32", "data": "synthetic", "config": config, "clocks": clocks, "gpu_launches": launches,
           "rows_total": int(r[0]), "checksum": {"abd_sum": int(r[1]), "tnf_sum": int(r[2])}, "parity_status": PARITY_NOTE, "roofline": roofline}
    return out if rank == 0 else None


def reference_arm(args, cfg, config):
    """--impl reference: the reference's own CPU implementation of the path on the host cores."""
    # bounded sample: sized so that the whole --steps K --warmup W run stays within a few minutes
    # (one pass over 100k pairs takes ~30 s on 8 cores, most of it count_kmer's serial dump load)
    n = int(min(args.cpu_sample_pairs, max(20_000, args.cpu_sample_pairs * 5 // max(1, args.steps + args.warmup))))
    threads = os.cpu_count() or 1
    from pangaea_b200 import synth

    ppb = cfg["pairs_per_barcode"]
    data = synth.generate(n_barcodes=max(1, n // ppb), mean_pairs=ppb, read_len=cfg["read_len"], n_genomes=cfg["n_genomes"], genome_len=3_000_000,
                          frag_len=50_000, seed=cfg["seed"], barcode_len=cfg["barcode_len"])
    n = data["n_pairs"]
    times, detail = [], {}
    with tempfile.TemporaryDirectory() as d:
        fq = synth.write_interleaved(os.path.join(d, "sample.fq"), data)
        for i in range(args.warmup + args.steps):
            wall, detail = run_reference_cpu(fq, d, threads, cfg["min_length"])
            if i >= args.warmup:
                times.append(wall)
    total = sum(times)
    value = 2 * n * len(times) / total
    sample = (f"{n} pairs of the same synthetic model (seed {cfg['seed']}) per step, plain-text interleaved FASTQ; count_kmer ‖ count_tnf (-t {threads}) "
              f"as src/feature.py:28-39 runs them; jellyfish (absent) replaced by the oracle's single-thread C counter")
    config = dict(config, workload=f"{config['workload']} - CPU arm: a bounded sample of {n} read pairs of this workload per step",
                  sample_pairs_per_step=int(n))
    out = {"impl": "reference", "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": round(1e3 * total / len(times), 3), "higher_is_better": True, "scaling": cfg["scaling"],
           "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": config,
           "cpu_baseline": {"value": round(value, 1), "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample, "detail": detail},
           "e2e": {"value": round(value, 1), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(out), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())

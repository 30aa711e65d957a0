"""bench.py's roofline bookkeeping (CPU): §8d bytes per pair, the dominant kernel's name and ncu traffic, attainable bounds."""
import importlib.util
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


C2_STAGES = {"pack": 3.9, "count_scatter": 18.0, "count_split": 14.5, "count_apply": 8.4, "group": 1.0, "tnf": 13.0, "feat_scatter": 0.0,
             "feat_apply": 42.3, "feat_collect": 0.0, "normalize": 0.4}
C2_LAUNCHES = {"pack": 5, "count_scatter": 15, "count_split": 10, "count_apply": 15, "group": 12, "tnf": 1, "feat_scatter": 0, "feat_apply": 5,
               "feat_collect": 0, "normalize": 2}


def test_survey_8d_bytes_per_pair():
    b = _bench()
    per_pair = b.algorithmic_bytes_per_pair(100, 0.01)  # C2: 2x100 bp, 100 pairs per cloud
    assert abs(sum(per_pair.values()) - 2456.9) < 1.0  # BASELINE.json / SURVEY.md §8d
    assert per_pair["count"] == 2 * 100 / 4 + 8 * 172 and per_pair["featurize"] == 2 * 100 / 4 + 4 * 172


def test_roofline_names_the_sweep_and_carries_its_ncu_traffic():
    b = _bench()
    pairs = 50_000_000
    r = b.build_roofline(dict(C2_STAGES), dict(C2_LAUNCHES), pairs, 100, 500_000, 2 * pairs * 101, 8_600_000_000, 8_600_000_000, 1.117e9)
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["kernel"] == "bucket_apply_feat_kernel" and r["pass"] == "featurize"
    traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    assert r["traffic"] == traffic["bucket_apply_feat_kernel"]  # dram read + write bytes per launch (ncu --set full)
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3 and 0 < r["frac"] < 1
    assert abs(r["whole_path"]["frac"] - 2456.9 * (1.117e9 / 2) / (r["peak"] * 1e9)) < 1e-3
    att = r["attainable"]
    assert att["bucket_apply_feat_kernel"]["bound"] == "l2_gather" and att["bucket_scatter_kernel<15,shared>"]["bound"] == "smem_slot_handout"
    assert att["normalize_rows_vec_kernel"]["bound"] == "hbm_stream"
    assert all(0 < v["frac_of_attainable"] <= 1.0 for v in att.values())
    assert "by_design" in r and set(r["passes"]) >= {"pack", "count", "featurize"}


def test_roofline_for_tiny_clouds_names_the_lookup_kernel():
    b = _bench()
    st = dict(C2_STAGES, feat_scatter=5.3, feat_apply=11.6, feat_collect=8.0, normalize=8.9, tnf=6.0)
    ln = dict(C2_LAUNCHES, feat_scatter=6, feat_collect=3)
    pairs = 10_000_000
    r = b.build_roofline(st, ln, pairs, 150, pairs, 2 * pairs * 151, 2_720_000_000, 2_720_000_000, 2.8e8)
    assert "bucket_lookup_kernel" in r["attainable"] and "bucket_apply_feat_kernel" not in r["attainable"]
    assert r["kernel"] == "bucket_scatter_kernel<15,shared>"  # the slowest stage of this (made-up) split

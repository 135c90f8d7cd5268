"""bench.py bookkeeping that must agree with SURVEY.md section 8d and the bench contract."""
import importlib.util
import os

from conftest import ROOT


def load_bench():
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_algorithmic_bytes_match_the_survey():
    b = load_bench()
    w = b.WORKLOAD
    assert (w["num_users"], w["num_items"], w["num_ratings"], w["k"]) == (283228, 53889, 27753444, 50)
    total = b.algorithmic_bytes_per_sweep(w)
    # B_u = 11.67 GB, B_i = 11.70 GB => 23.4 GB per sweep
    assert abs(total / 1e9 - 23.37) < 0.05
    # executed fp64 FLOPs: 28 tiles x 512 FLOP per 4 ratings per side + 112 DMMA per Cholesky
    assert abs(b.executed_flops_per_sweep(w) / 1e11 - 2.18) < 0.05


def test_metric_names_follow_baseline_json():
    import json
    b = load_bench()
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert "ratings/sec per sweep" in base["metric"] and b.UNIT == "ratings/s"
    assert "per_sweep" in b.METRIC

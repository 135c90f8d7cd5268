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


def test_algorithmic_flops_match_the_survey():
    b = load_bench()
    # F_u + F_i = 1.50e11, Cholesky 1.5e10 => 1.65e11 per sweep at C3
    assert abs(b.algorithmic_flops_per_sweep(b.WORKLOAD) / 1e11 - 1.65) < 0.01


def test_reference_sample_keeps_the_shrink_rule():
    """The reference arm's subsample re-applies the reference's shrink rule: every movie keeps
    >= k ratings and every user >= k + 1, ids are dense, the sample is a subset of the workload."""
    import numpy as np
    b = load_bench()
    w = dict(b.WORKLOAD, name="dev", num_users=6000, num_items=900, num_ratings=500000, k=20)
    p, _ = b.make_problem(w, 7)
    s = b.shrunk_user_subsample(p, w, 0.5)
    assert 0 < len(s["ratings"]) < len(p["ratings"])
    cu = np.bincount(s["user_ids"], minlength=s["num_users"])
    ci = np.bincount(s["item_ids"], minlength=s["num_items"])
    assert cu.min() >= w["k"] + 1 and ci.min() >= w["k"]
    assert s["user_ids"].max() == s["num_users"] - 1 and s["item_ids"].max() == s["num_items"] - 1
    assert len(s["user_factors0"]) == s["num_users"] * (w["k"] + 1)
    assert "shrink rule" in b.sample_description(s, w)


def test_compulsory_bytes_match_the_survey():
    """SURVEY.md 8d, "each array once": per side 333 MB of grouped ids + ratings, the opposite
    factor matrix once (21.6 / 115.6 MB), the own rows read and written."""
    b = load_bench()
    w = b.WORKLOAD
    nnz, uf, itf = w["num_ratings"], 283228 * 51 * 8, 53889 * 50 * 8
    want = 2 * nnz * 12 + (itf + 2 * uf) + (uf + 2 * itf) + (283229 + 53890) * 4
    assert b.compulsory_bytes_per_sweep(w) == want
    assert abs(want / 1e9 - 1.08) < 0.01
    # far below the no-cache figure the roofline numerator uses
    assert b.compulsory_bytes_per_sweep(w) < b.algorithmic_bytes_per_sweep(w) / 20

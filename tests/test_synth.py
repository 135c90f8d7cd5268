import numpy as np

from movie_recommender_b200 import synth


def test_rating_pairs_shape_and_degrees():
    u, i = synth.rating_pairs(500, 300, 20000, 9, 8, seed=1)
    assert len(u) == 20000 and u.dtype == np.int32
    keys = u.astype(np.int64) * 300 + i
    assert len(np.unique(keys)) == len(keys)
    assert np.bincount(u, minlength=500).min() >= 9
    assert np.bincount(i, minlength=300).min() >= 8
    assert np.all(np.diff(u) >= 0)  # grouped by user


def test_determinism_and_rating_grid():
    a = synth.als_problem(100, 80, 3000, 4, seed=7)
    b = synth.als_problem(100, 80, 3000, 4, seed=7)
    for k in ("user_ids", "item_ids", "ratings", "user_factors0", "item_factors0"):
        assert np.array_equal(a[k], b[k])
    raw = synth.planted_ratings(a["user_ids"], a["item_ids"], 100, 80, seed=7, subtract_median=False)
    assert np.all(raw * 2 == np.round(raw * 2)) and raw.min() >= 0.5 and raw.max() <= 5.0
    med = synth.movie_medians(a["item_ids"], raw, 80)
    for m in range(0, 80, 13):
        assert med[m] == np.median(raw[a["item_ids"] == m])


def test_bias_system_layout():
    u, i = synth.rating_pairs(30, 20, 200, 2, 2, seed=2)
    raw = synth.planted_ratings(u, i, 30, 20, seed=2, subtract_median=False)
    rowptr, col, vals, cols, b, x0 = synth.bias_model_system(u, i, raw, 30, 20)
    assert cols == 50 and len(x0) == 50 and np.all(vals == 1.0)
    assert np.array_equal(rowptr, np.arange(201) * 2)
    assert np.array_equal(col[0::2], u) and np.array_equal(col[1::2], 30 + i)

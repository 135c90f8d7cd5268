"""Per-user fold-in against the reference's own formulation (python/app_local/models.py:657-703:
numpy.linalg.lstsq on rows [movie_factors, 1]); tolerance 1e-8 relative (well-posed rows)."""
import numpy as np
import pytest

from movie_recommender_b200 import synth

pytestmark = pytest.mark.gpu


def test_fold_in_matches_lstsq(require_gpu):
    from movie_recommender_b200.fold_in import fold_in_users
    nu, ni, k = 400, 300, 11
    p = synth.als_problem(nu, ni, 30000, k, seed=4, min_degrees=False)
    rng = np.random.default_rng(0)
    itf = rng.uniform(-1, 1, ni * k)
    u, i, r = p["user_ids"], p["item_ids"], p["ratings"]
    uf, valid = fold_in_users(u, i, r, nu, itf, k)
    V = itf.reshape(ni, k)
    deg = np.bincount(u, minlength=nu)
    assert np.array_equal(valid, deg >= k + 1) and valid.any() and (~valid).any() or valid.all()
    for a in range(nu):
        rows = np.flatnonzero(u == a)
        if len(rows) < k + 1:                                  # models.py:669
            assert np.all(np.isnan(uf[a]))
            continue
        A = np.hstack([V[i[rows]], np.ones((len(rows), 1))])
        x = np.linalg.lstsq(A, r[rows], rcond=None)[0]         # models.py:698
        assert np.linalg.norm(uf[a] - x) <= 1e-8 * max(1.0, np.linalg.norm(x))

"""Per-user evaluation on the GPU (als_predictor.als_eval -> csrc/evaluate.cu) against golden
outputs of the REAL reference functions and, on a larger seeded case, against the oracle: the
agreement values must be EQUAL (exact pair counts; predictions in the reference's arithmetic
order, so every strict comparison agrees)."""
import json
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle import eval_oracle

pytestmark = pytest.mark.gpu
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "eval_small.json")))


@pytest.mark.parametrize("name", sorted(GOLDEN))
def test_als_eval_equals_the_reference(require_gpu, name):
    from movie_recommender_b200 import als_predictor
    case = eval_oracle.synthetic_eval_case(**GOLDEN[name]["params"])
    got = als_predictor.als_eval(*case)
    assert got == [(u, float.fromhex(h)) for u, h in GOLDEN[name]["agreements"]]


def test_als_eval_equals_the_oracle_on_a_larger_case(require_gpu):
    from movie_recommender_b200 import als_predictor
    case = eval_oracle.synthetic_eval_case(num_users=300, num_movies=800, k=50, seed=9)
    assert als_predictor.als_eval(*case) == eval_oracle.als_eval(*case)


def test_edge_cases(require_gpu):
    from movie_recommender_b200 import als_predictor
    k = 3
    uf = np.arange(8, dtype=np.float64) / 7.0
    itf = np.arange(9, dtype=np.float64) / 5.0
    medians = {10: 3.0, 11: 3.5, 12: 2.5}
    movies = {10: 0, 11: 1, 12: 2}
    users = {1: 0, 2: 1}
    assert als_predictor.als_eval([], medians, uf, users, itf, movies, k) == []
    tests = [(1, []), (2, [(10, 4.0)]), (1, [(10, 4.0), (99, 1.0)]),      # nothing / one / one predictable
             (2, [(10, 2.0), (11, 2.0), (12, 2.0)]),                       # constant actual ratings
             (1, [(10, 5.0), (11, 1.0), (12, 3.0)])]
    got = als_predictor.als_eval(tests, medians, uf, users, itf, movies, k)
    assert got == eval_oracle.als_eval([tests[4]], medians, uf, users, itf, movies, k) and len(got) == 1

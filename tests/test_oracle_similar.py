"""The similarity oracle (oracle/similar_oracle.py) against golden results produced by the REAL
reference class (tests/golden/similar_*.json, made by tests/golden/make_golden_similar.py):
ids equal, scores bit-equal, including the reliability cut and the tuned parameters."""
import json
import os

import pytest

from conftest import GOLDEN
from oracle.similar_oracle import SimilarOracle, synthetic_catalogue

CASES = ["small", "dense", "cut"]


def load(name):
    with open(os.path.join(GOLDEN, "similar_%s.json" % name)) as fh:
        return json.load(fh)


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_the_real_class(name):
    g = load(name)
    genres, ratings = synthetic_catalogue(**g["params"])
    o = SimilarOracle(genres, ratings)
    for key, nres in (("results", 20), ("results_top3", 3)):
        for i_str, (ids, hexscores) in g[key].items():
            oi, os_ = o.find_similar_movie(int(i_str), num_results=nres)
            assert list(oi) == ids
            assert [float(s).hex() for s in os_] == hexscores

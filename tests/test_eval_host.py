"""Host logic of the evaluation mirror (movie_recommender_b200/als_predictor.py: als_eval) on CPU:
the flattening of the reference's per-user lists, which entries get no prediction (no training
median / no ALS factors, als_predictor.py:41-43), the "more than one prediction" and "two
different ratings" rules (worker_process.py:252, my_util.py:112-120).  The one GPU entry point
(mrb_als_rank_agreement, csrc/evaluate.cu) is replaced here by a plain-Python stand-in that works
on the SAME flattened arrays through the same C signature; the expected outputs are the golden
results of the REAL reference functions (tests/golden/eval_small.json).  The real kernel is
exercised by tests/test_gpu_eval.py."""
import ctypes
import json
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle import eval_oracle

GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "eval_small.json")))


class _StandIn:
    """include/cpp_ls_b200.h: mrb_als_rank_agreement, computed on the host from the flat arrays."""

    def __init__(self):
        self.calls = []

    def mrb_als_rank_agreement(self, user_ptr, num_users, entry_user_row, entry_movie_row, actual,
                               median, user_factors, num_user_rows, item_factors, num_items, k,
                               agree, disagree, n_pred, kernel_ms):
        as_array = np.ctypeslib.as_array
        ptr = as_array(user_ptr, shape=(num_users + 1,))
        n = int(ptr[num_users])
        urow = as_array(entry_user_row, shape=(max(n, 1),))
        mrow = as_array(entry_movie_row, shape=(max(n, 1),))
        act = as_array(actual, shape=(max(n, 1),))
        med = as_array(median, shape=(max(n, 1),))
        uf = as_array(user_factors, shape=(num_user_rows * (k + 1),))
        itf = as_array(item_factors, shape=(num_items * k,))
        self.calls.append(dict(num_users=num_users, entries=n, num_user_rows=num_user_rows,
                               num_items=num_items, k=k, urow=urow[:n].copy(), mrow=mrow[:n].copy()))
        for u in range(num_users):
            kept = []
            for e in range(int(ptr[u]), int(ptr[u + 1])):
                if mrow[e] < 0 or urow[e] < 0:
                    continue
                assert 0 <= urow[e] < num_user_rows and 0 <= mrow[e] < num_items
                f = uf[(k + 1) * urow[e]:(k + 1) * (urow[e] + 1)]
                mf = itf[k * mrow[e]:k * (mrow[e] + 1)]
                rating = 0
                for i in range(k):                       # als_predictor.py:54-58, same order
                    rating += f[i] * mf[i]
                rating += f[k]
                rating += med[e]
                kept.append((act[e], rating))
            a = d = 0
            for x in range(len(kept)):
                for y in range(x + 1, len(kept)):
                    if kept[x][0] == kept[y][0]:
                        continue
                    hi, lo = (kept[x], kept[y]) if kept[x][0] > kept[y][0] else (kept[y], kept[x])
                    if hi[1] > lo[1]:
                        a += 1
                    else:
                        d += 1
            agree[u], disagree[u], n_pred[u] = a, d, len(kept)
        ctypes.cast(kernel_ms, ctypes.POINTER(ctypes.c_float))[0] = 0.0
        return 0


@pytest.fixture()
def evaluator(monkeypatch):
    from movie_recommender_b200 import als_predictor
    stand_in = _StandIn()
    monkeypatch.setattr(als_predictor, "_dll", stand_in)
    return als_predictor, stand_in


@pytest.mark.parametrize("name", sorted(GOLDEN))
def test_als_eval_host_logic_reproduces_the_reference(evaluator, name):
    als_predictor, stand_in = evaluator
    case = eval_oracle.synthetic_eval_case(**GOLDEN[name]["params"])
    got = als_predictor.als_eval(*case)
    assert got == [(u, float.fromhex(h)) for u, h in GOLDEN[name]["agreements"]]
    call = stand_in.calls[-1]
    tests, medians, uf, user_ids, itf, movie_ids, k = case
    assert call["num_users"] == len(tests) and call["entries"] == sum(len(m) for _, m in tests)
    assert call["num_user_rows"] == len(user_ids) and call["num_items"] == len(movie_ids) and call["k"] == k
    # an entry is dropped (-1) exactly when the reference makes no prediction for it
    flat = [m for _, ms in tests for m, _ in ms]
    want_dropped = np.array([m not in medians or m not in movie_ids for m in flat])
    assert np.array_equal(call["mrow"] < 0, want_dropped) and want_dropped.any()


def test_edge_cases_host_logic(evaluator):
    als_predictor, _ = evaluator
    k = 3
    uf, itf = np.arange(8, dtype=np.float64) / 7.0, np.arange(9, dtype=np.float64) / 5.0
    medians, movies, users = {10: 3.0, 11: 3.5, 12: 2.5}, {10: 0, 11: 1, 12: 2}, {1: 0, 2: 1}
    assert als_predictor.als_eval([], medians, uf, users, itf, movies, k) == []
    tests = [(1, []), (2, [(10, 4.0)]), (1, [(10, 4.0), (99, 1.0)]),
             (2, [(10, 2.0), (11, 2.0), (12, 2.0)]),
             (1, [(10, 5.0), (11, 1.0), (12, 3.0)]),
             (7, [(10, 5.0), (11, 1.0)])]                  # a user without ALS factors: skipped
    got = als_predictor.als_eval(tests, medians, uf, users, itf, movies, k)
    assert got == eval_oracle.als_eval([tests[4]], medians, uf, users, itf, movies, k) and len(got) == 1

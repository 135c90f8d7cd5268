"""oracle/eval_oracle.py -- TEST INFRASTRUCTURE ONLY.  Pure-Python restatement of the reference's
per-user evaluation: ALS_Model.predict (python/full_data/als_predictor.py:35-60),
_test_model / _als_eval (python/full_data/worker_process.py:231-306) and
compute_ranking_agreement (python/full_data/my_util.py:101-145).  Pinned against golden outputs
of the REAL functions (tests/golden/eval_small.json, made by tests/golden/make_golden_eval.py)."""
import numpy as np


def predict(user_factors, movie_id, movie_medians, als_movie_factors, als_movie_ids):
    if movie_id not in movie_medians or movie_id not in als_movie_ids:      # als_predictor.py:41-43
        return None
    k = len(user_factors) - 1
    m = als_movie_ids[movie_id]
    mf = als_movie_factors[k * m:k * (m + 1)]
    rating = 0
    for i in range(k):                                                      # :54-55
        rating += user_factors[i] * mf[i]
    rating += user_factors[k]                                               # :57
    rating += movie_medians[movie_id]                                       # :58
    return rating


def ranking_agreement(actual, predicted):
    """my_util.py:101-145 with the grouping by rating unrolled into a pair loop."""
    if len(actual) == 1:
        return None
    if all(r == actual[0][1] for _, r in actual):
        return None
    pred = dict(predicted)
    agree = disagree = 0
    for a, (m1, r1) in enumerate(actual):
        for m2, r2 in actual[a + 1:]:
            if r1 == r2:
                continue
            hi, lo = (m1, m2) if r1 > r2 else (m2, m1)
            if pred[hi] > pred[lo]:
                agree += 1
            else:
                disagree += 1
    return agree / (agree + disagree)


def als_eval(user_ratings_test, movie_medians, uf, als_user_ids, itf, als_movie_ids, k):
    out = []
    for user_id, movie_ratings in user_ratings_test:
        row = als_user_ids[user_id]
        f = uf[(k + 1) * row:(k + 1) * (row + 1)]
        kept, predicted = [], []
        for movie_id, actual in movie_ratings:                              # worker_process.py:244-250
            p = predict(f, movie_id, movie_medians, itf, als_movie_ids)
            if p is not None:
                predicted.append((movie_id, p))
                kept.append((movie_id, actual))
        if len(predicted) > 1:
            ag = ranking_agreement(kept, predicted)
            if ag is not None:
                out.append((user_id, ag))
    return out


def synthetic_eval_case(num_users=40, num_movies=60, k=5, seed=0):
    """Small seeded evaluation inputs in the reference's in-memory formats: raw movie ids,
    a few movies without ALS factors / without a median, users with one rating or constant
    ratings (agreement undefined)."""
    rng = np.random.default_rng(seed)
    movie_ids = [int(m) for m in rng.choice(np.arange(1, 5 * num_movies), size=num_movies, replace=False)]
    als_movie_ids = {m: j for j, m in enumerate(movie_ids[:num_movies - 5])}     # 5 movies without factors
    medians = {m: float(rng.choice([2.5, 3.0, 3.5, 4.0])) for m in movie_ids[3:]}   # 3 without a median
    user_ids = [int(u) for u in rng.choice(np.arange(1, 10 * num_users), size=num_users, replace=False)]
    als_user_ids = {u: j for j, u in enumerate(user_ids)}
    uf = rng.uniform(-1, 1, num_users * (k + 1))
    itf = rng.uniform(-1, 1, len(als_movie_ids) * k)
    tests = []
    for j, u in enumerate(user_ids):
        n = 1 if j == 0 else int(rng.integers(2, 25))
        ms = [int(m) for m in rng.choice(movie_ids, size=n, replace=False)]
        if j == 1:
            rs = [3.0] * n                                   # constant ratings: agreement undefined
        else:
            rs = [float(r) for r in rng.choice(np.arange(1, 11) * 0.5, size=n)]
        tests.append((u, list(zip(ms, rs))))
    return tests, medians, uf, als_user_ids, itf, als_movie_ids, k

"""Evaluation of a trained ALS model on the GPU -- the mirror of the reference's
``python/full_data/als_predictor.py`` (``ALS_Model.predict``, :35-60) together with the worker
loop that drives it (``worker_process._als_eval``, :262-306, ``_test_model``, :231-256) and the
metric it reports (``my_util.compute_ranking_agreement``, :101-145).

The reference builds one ``ALS_Model`` per user and loops over Python lists; here every user's
test ratings are flattened once and ONE device call computes all predictions (same arithmetic
order, so the strict ``>`` comparisons agree bit for bit) and counts the agreeing / disagreeing
pairs per user (csrc/evaluate.cu).
"""
import ctypes

import numpy

from . import _lib

_dll = _lib.dll


def als_eval(user_ratings_test, movie_medians_train, als_user_factors, als_user_ids,
             als_movie_factors, als_movie_ids, num_item_factors):
    """``_als_eval`` of the reference: ``[(user id, agreement)]`` for every user of
    ``user_ratings_test = [(user_id, [(movie_id, rating)])]`` whose agreement is defined
    (more than one predictable test movie and at least two different actual ratings).

    A movie without a training median or without ALS factors gets no prediction and is dropped
    (als_predictor.py:41-43).  Unlike the reference (``als_user_ids[user_id]`` raises KeyError,
    worker_process.py:290) a user without ALS factors is skipped."""
    k = num_item_factors
    uf = numpy.ascontiguousarray(als_user_factors, dtype=numpy.double).reshape(-1)
    itf = numpy.ascontiguousarray(als_movie_factors, dtype=numpy.double).reshape(-1)
    ptr = [0]
    urow, mrow, actual, median = [], [], [], []
    for user_id, movie_ratings in user_ratings_test:
        row = als_user_ids.get(user_id, -1)
        for movie_id, rating in movie_ratings:
            ok = movie_id in movie_medians_train and movie_id in als_movie_ids
            urow.append(row)
            mrow.append(als_movie_ids[movie_id] if ok else -1)
            median.append(movie_medians_train[movie_id] if ok else 0.0)
            actual.append(rating)
        ptr.append(len(actual))
    nu = len(user_ratings_test)
    ptr = numpy.asarray(ptr, dtype=numpy.int32)
    urow = numpy.asarray(urow, dtype=numpy.int32)
    mrow = numpy.asarray(mrow, dtype=numpy.int32)
    actual = numpy.asarray(actual, dtype=numpy.double)
    median = numpy.asarray(median, dtype=numpy.double)
    agree = numpy.zeros(max(nu, 1), dtype=numpy.int64)
    disagree = numpy.zeros(max(nu, 1), dtype=numpy.int64)
    n_pred = numpy.zeros(max(nu, 1), dtype=numpy.int32)
    ms = ctypes.c_float(0)
    LL = ctypes.POINTER(ctypes.c_longlong)
    _lib.check(_dll.mrb_als_rank_agreement(
        _lib.ip(ptr), nu, _lib.ip(urow), _lib.ip(mrow), _lib.dp(actual), _lib.dp(median), _lib.dp(uf),
        len(uf) // (k + 1), _lib.dp(itf), len(itf) // k, k, agree.ctypes.data_as(LL),
        disagree.ctypes.data_as(LL), _lib.ip(n_pred), ctypes.byref(ms)))
    als_eval.last_kernel_ms = ms.value
    out = []
    for j, (user_id, _) in enumerate(user_ratings_test):
        pairs = int(agree[j]) + int(disagree[j])
        if n_pred[j] > 1 and pairs > 0:                 # worker_process.py:252, my_util.py:112-120
            out.append((user_id, int(agree[j]) / pairs))   # my_util.py:145
    return out

"""ALS data preparation on the GPU: NumPy-in / NumPy-out wrappers
of ``mrb_movie_medians`` and ``mrb_als_shrink`` (include/cpp_ls_b200.h, section 8).

They replace the multi-process Python of the reference's ``movie_lens_data.py:453-464`` (movie
medians of the training set) and ``:547-680`` (``als_data_set_shrink_mp``) on the flattened form
of its in-memory lists; ``movie_lens_data.py`` in this package keeps the reference's call
signatures and file formats on top of these.  No CPU fallback: without the CUDA library or a
device these raise.
"""
import ctypes

import numpy as np

from . import _lib


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def movie_medians(movie_ids, ratings, num_movie_slots=None):
    """``(medians f64[slots], counts i32[slots], kernel_ms)``: ``numpy.median`` of every movie's
    ratings (NaN for a slot without ratings), bit-exact (movie_lens_data_proc.py:455-471)."""
    movie_ids, ratings = _i32(movie_ids), _f64(ratings)
    if len(movie_ids) != len(ratings):
        raise ValueError("movie_ids and ratings differ in length")
    if num_movie_slots is None:
        num_movie_slots = int(movie_ids.max()) + 1 if len(movie_ids) else 0
    med = np.empty(num_movie_slots, dtype=np.float64)
    cnt = np.empty(num_movie_slots, dtype=np.int32)
    ms = ctypes.c_float(0.0)
    _lib.check(_lib.dll.mrb_movie_medians(_lib.ip(movie_ids), _lib.dp(ratings), len(ratings),
                                          num_movie_slots, _lib.dp(med), _lib.ip(cnt),
                                          ctypes.byref(ms)))
    return med, cnt, float(ms.value)


class ShrinkResult:
    """Output of :func:`als_shrink`."""
    __slots__ = ("user_ids", "movie_ids", "ratings", "keep_pos", "user_new_id", "movie_new_id",
                 "num_users", "num_movies", "rounds", "kernel_ms")


def als_shrink(user_slot_ids, movie_ids, ratings, num_user_slots, num_movie_slots, medians,
               min_user_ratings, min_movie_ratings):
    """One factor of ``als_data_set_shrink_mp`` (movie_lens_data.py:569-645): the degree filter
    run to its fixpoint, then the surviving ratings in their original order with zero-based ids
    (ascending slot order) and ``rating - medians[movie]``.  Returns a :class:`ShrinkResult`."""
    user_slot_ids, movie_ids, ratings = _i32(user_slot_ids), _i32(movie_ids), _f64(ratings)
    medians = _f64(medians)
    n = len(ratings)
    if len(user_slot_ids) != n or len(movie_ids) != n:
        raise ValueError("user_slot_ids, movie_ids and ratings differ in length")
    if len(medians) != num_movie_slots:
        raise ValueError("medians must have num_movie_slots entries")
    out_u = np.empty(n, dtype=np.int32)
    out_m = np.empty(n, dtype=np.int32)
    out_r = np.empty(n, dtype=np.float64)
    keep = np.empty(n, dtype=np.int32)
    user_new = np.empty(num_user_slots, dtype=np.int32)
    movie_new = np.empty(num_movie_slots, dtype=np.int32)
    info = _lib.ShrinkInfo()
    _lib.check(_lib.dll.mrb_als_shrink(
        _lib.ip(user_slot_ids), _lib.ip(movie_ids), _lib.dp(ratings), n, num_user_slots,
        num_movie_slots, _lib.dp(medians), int(min_user_ratings), int(min_movie_ratings),
        _lib.ip(out_u), _lib.ip(out_m), _lib.dp(out_r), _lib.ip(keep), _lib.ip(user_new),
        _lib.ip(movie_new), ctypes.byref(info)))
    m = info.num_ratings_out
    res = ShrinkResult()
    res.user_ids, res.movie_ids, res.ratings, res.keep_pos = out_u[:m], out_m[:m], out_r[:m], keep[:m]
    res.user_new_id, res.movie_new_id = user_new, movie_new
    res.num_users, res.num_movies = info.num_users_out, info.num_movies_out
    res.rounds, res.kernel_ms = info.rounds, float(info.kernel_ms)
    return res

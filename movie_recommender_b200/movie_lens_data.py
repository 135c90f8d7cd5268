"""The ALS trainer entry of the reference (``python/full_data/movie_lens_data.py:684-713``,
``als_train``) on the B200 library: same function name, same arguments, same files in and out.

The trainer glue and the data preparation right before it live here (SURVEY.md section 8 rows a9
and f2): ``compute_movie_medians`` / ``als_data_set_shrink_mp`` run the median step and the ALS
data-set "shrink" on the GPU and write the files ``als_train`` reads.  The rest of the
reference's ``movie_lens_data.py`` (CSV ingest, train/test split) is out of scope (SURVEY.md
section 2).  The files between the two steps:

  ``<als_dir>/als{k}_movie_ids.bin``, ``als{k}_user_ids.bin``   pickled id maps (only ``len`` is used)
  ``<als_dir>/als{k}_user_ratings_train.bin``                   pickled ``[user_ids int32,
                                                                item_ids int32, ratings float64]``
                                                                (median-subtracted, zero-based ids)
and it writes, like the reference, the flat float64 factor arrays

  ``<als_dir>/als{k}_user_factors.bin``, ``als{k}_item_factors.bin``

which ``python/app_local/recommend.py:152-172`` loads unchanged.  ``als_dir`` follows the
reference's ``config.py`` (``./data/als/``) and can be overridden per call.
"""
import datetime
import itertools
import os
import pickle
import time

import numpy as np

from . import cpp_ls, prep

shared_directory = "." + os.sep + "data" + os.sep     # python/full_data/config.py:3
als_dir = shared_directory + "als" + os.sep           # config.py:7
start_time = time.time()


def _current_time():
    return datetime.timedelta(seconds=int(time.time() - start_time))


def get_als_obj(object_name, directory=None):
    """Use pickle to load "object_name" from "als_dir" (movie_lens_data.py:29-36)."""
    with open((directory or als_dir) + object_name + ".bin", mode="rb") as file:
        return pickle.load(file)


def als_train(factors_list, thread_count=None, algorithm=1, *, directory=None, verbose=True):
    """Trains ALS models for each factor (movie_lens_data.py:684-713).  Each factor is a model
    with its own "user_factors" and "item_factors".  Returns {factor: iterations}."""
    directory = directory or als_dir
    if thread_count is not None:
        cpp_ls.set_thread_count(thread_count)
    iterations_by_factor = {}
    for factor in factors_list:
        num_items = len(get_als_obj("als" + str(factor) + "_movie_ids", directory))
        num_users = len(get_als_obj("als" + str(factor) + "_user_ids", directory))
        user_ids_train, item_ids_train, ratings_train = get_als_obj(
            "als" + str(factor) + "_user_ratings_train", directory)
        if verbose:
            print(_current_time(), "Building ALS factor", factor, "model")
        user_factors, item_factors, iterations = cpp_ls.als(
            user_ids_train, item_ids_train, ratings_train, factor,
            num_users, num_items, algorithm=algorithm)
        if verbose:
            print(_current_time(), "ALS took", iterations, "iterations.",
                  'Saving "user_factors" and "item_factors" to disk')
        with open(directory + "als" + str(factor) + "_user_factors.bin", mode="wb") as file:
            pickle.dump(user_factors, file)
        with open(directory + "als" + str(factor) + "_item_factors.bin", mode="wb") as file:
            pickle.dump(item_factors, file)
        iterations_by_factor[factor] = iterations
    return iterations_by_factor


# ----------------------------------------------------------------------------------------------
# Data preparation.  The reference keeps ``user_ratings_train`` and
# ``user_ratings_test`` "in process memory" of its worker pool between ``refresh_training_sets_mp``
# and ``als_data_set_shrink_mp`` (movie_lens_data.py:419-421, 549-551); ``_memory`` plays that
# role here, holding the flattened (COO) form the GPU works on.
# ----------------------------------------------------------------------------------------------
_memory = {}
_MAX_MOVIE_SLOTS = 1 << 28


def load_training_sets(user_ratings_train, user_ratings_test=None):
    """Takes the reference's in-memory lists ``[(user id, [(movie id, rating)])]`` (the test list,
    if any, is parallel to the training list: movie_lens_data_proc.py:373-374, 527-530)."""
    lens = np.fromiter((len(e) for _, e in user_ratings_train), dtype=np.int64,
                       count=len(user_ratings_train))
    n = int(lens.sum())
    flat = itertools.chain.from_iterable(e for _, e in user_ratings_train)
    pairs = np.fromiter(flat, dtype=np.dtype([("m", np.int64), ("r", np.float64)]), count=n)
    load_training_coo([u for u, _ in user_ratings_train],
                      np.repeat(np.arange(len(lens), dtype=np.int32), lens),
                      pairs["m"], pairs["r"], user_ratings_test)


def load_training_coo(user_raw_ids, user_slot_ids, movie_ids, ratings, user_ratings_test=None):
    """The same data already flattened: ``user_raw_ids[s]`` is the id of list entry ``s``,
    ``user_slot_ids[i]`` the entry of rating ``i`` (ratings in list order)."""
    movie_ids = np.asarray(movie_ids)
    if len(movie_ids) and (movie_ids.min() < 0 or movie_ids.max() >= _MAX_MOVIE_SLOTS):
        raise ValueError("movie ids must lie in [0, 2^28)")
    if user_ratings_test is not None and len(user_ratings_test) != len(user_raw_ids):
        raise ValueError("user_ratings_test must be parallel to user_ratings_train")
    _memory.clear()
    _memory.update(user_raw=list(user_raw_ids),
                   user_slot=np.ascontiguousarray(user_slot_ids, dtype=np.int32),
                   movie=np.ascontiguousarray(movie_ids, dtype=np.int32),
                   rating=np.ascontiguousarray(ratings, dtype=np.float64),
                   test=None if user_ratings_test is None else list(user_ratings_test))


def training_sets_in_memory():
    """``(user_raw_ids, user_slot_ids, movie_ids, ratings, user_ratings_test)`` as they stand."""
    m = _memory
    return m["user_raw"], m["user_slot"], m["movie"], m["rating"], m["test"]


def compute_movie_medians(directory=None, save=False):
    """The median step of ``refresh_training_sets_mp`` (movie_lens_data.py:453-471): returns
    ``{movie id: median rating}`` of the training set in memory; ``save`` also writes
    ``movie_medians_train.bin`` into ``directory``."""
    med, cnt, _ = prep.movie_medians(_memory["movie"], _memory["rating"])
    ids = np.nonzero(cnt)[0]
    movie_medians = dict(zip(ids.tolist(), med[ids].tolist()))
    if save:
        with open((directory or shared_directory) + "movie_medians_train.bin", mode="wb") as file:
            pickle.dump(movie_medians, file)
    return movie_medians


def _set_order(ids_in_first_appearance_order):
    """{id: zero based id} in the iteration order of the Python set a single-process reference
    run builds in ``_collect_ids`` (movie_lens_data_proc.py:589-608; movie_lens_data.py:596-609)."""
    s = set()
    for v in ids_in_first_appearance_order:
        s.add(v)
    return {v: i for i, v in enumerate(s)}


def als_data_set_shrink_mp(movie_medians_train, factors_list, no_test_set=False, *,
                           directory=None, id_order="sorted", verbose=True):
    """The "factors" are number of item factors (movie_lens_data.py:547-680).  Works on the
    training (and test) set in memory (``load_training_sets``): drops users with fewer than
    factor + 1 ratings and movies with fewer than factor ratings until nothing changes, and writes

      ``als{k}_user_ids.bin``, ``als{k}_movie_ids.bin``   {standard id: zero based id}
      ``als{k}_user_ratings_train.bin``                    [user ids i32, movie ids i32, ratings f64]
      ``als{k}_user_ratings_test.bin`` / ``_test_length.bin``   (unless ``no_test_set``)

    As in the reference every factor shrinks what the previous factor left in memory.
    ``id_order``: "sorted" numbers the ids in ascending order of the standard ids (users: list
    order); "reference" re-applies the set-iteration order of a single-process reference run (the
    reference's own labels depend on its worker count).  Returns {factor: (users, movies, ratings)}.
    """
    if id_order not in ("sorted", "reference"):
        raise ValueError("id_order must be 'sorted' or 'reference'")
    directory = directory or als_dir
    if verbose:
        print(_current_time(), "Shrinking training data to satisfy ALS requirements.")
    coverage = {}
    for factor in factors_list:
        m = _memory
        slots = int(m["movie"].max()) + 1 if len(m["movie"]) else 0
        medians = np.full(slots, np.nan)
        if slots:
            ids = np.fromiter(movie_medians_train.keys(), dtype=np.int64, count=len(movie_medians_train))
            vals = np.fromiter(movie_medians_train.values(), dtype=np.float64, count=len(movie_medians_train))
            inside = (ids >= 0) & (ids < slots)
            medians[ids[inside]] = vals[inside]
            if np.isnan(medians[np.unique(m["movie"])]).any():
                raise KeyError("movie_medians_train lacks a movie of the training set")
        res = prep.als_shrink(m["user_slot"], m["movie"], m["rating"], len(m["user_raw"]), slots,
                              medians, factor + 1, factor)
        kept_users = np.nonzero(res.user_new_id >= 0)[0]
        kept_movies = np.nonzero(res.movie_new_id >= 0)[0]
        user_raw = [m["user_raw"][s] for s in kept_users.tolist()]
        movie_kept = m["movie"][res.keep_pos]

        user_ids, movie_ids = res.user_ids, res.movie_ids
        if id_order == "sorted":
            als_user_ids = {u: i for i, u in enumerate(user_raw)}
            als_movie_ids = {mv: i for i, mv in enumerate(kept_movies.tolist())}
        else:
            als_user_ids = _set_order(user_raw)
            _, first = np.unique(movie_kept, return_index=True)
            als_movie_ids = _set_order(movie_kept[np.sort(first)].tolist())
            relabel_u = np.array([als_user_ids[u] for u in user_raw], dtype=np.int32)
            relabel_m = np.array([als_movie_ids[mv] for mv in kept_movies.tolist()], dtype=np.int32)
            user_ids = relabel_u[user_ids] if len(user_ids) else user_ids
            movie_ids = relabel_m[movie_ids] if len(movie_ids) else movie_ids

        if verbose:
            print(_current_time(), 'Saving "als_user_ids" and "als_movie_ids"'
                  + " for ALS factor " + str(factor))
        with open(directory + "als" + str(factor) + "_user_ids.bin", mode="wb") as file:
            pickle.dump(als_user_ids, file)
        with open(directory + "als" + str(factor) + "_movie_ids.bin", mode="wb") as file:
            pickle.dump(als_movie_ids, file)
        with open(directory + "als" + str(factor) + "_user_ratings_train.bin", mode="wb") as file:
            pickle.dump([np.ascontiguousarray(user_ids), np.ascontiguousarray(movie_ids),
                         np.ascontiguousarray(res.ratings)], file)

        # what stays in memory for the next factor: the shrunk, NOT median-subtracted data
        test = m["test"]
        if test is not None:
            test = [test[s] for s in kept_users.tolist()]
        _memory.update(user_raw=user_raw, user_slot=np.ascontiguousarray(res.user_ids),
                       movie=np.ascontiguousarray(movie_kept),
                       rating=np.ascontiguousarray(m["rating"][res.keep_pos]), test=test)

        test_file = directory + "als" + str(factor) + "_user_ratings_test.bin"
        length_file = directory + "als" + str(factor) + "_user_ratings_test_length.bin"
        if no_test_set:
            for name in (test_file, length_file):
                if os.path.exists(name):
                    os.remove(name)
        else:
            with open(test_file, mode="wb") as file:
                pickle.dump(test, file)
            with open(length_file, mode="wb") as file:
                pickle.dump(len(test) if test is not None else 0, file)
        coverage[factor] = (res.num_users, res.num_movies, len(res.ratings))

    if verbose:
        print("".ljust(10), "Users".center(15), "Movies".center(15))
        for factor in factors_list:
            print(("ALS " + str(factor)).ljust(10), str(coverage[factor][0]).center(15),
                  str(coverage[factor][1]).center(15))
    return coverage

"""The ALS trainer entry of the reference (``python/full_data/movie_lens_data.py:684-713``,
``als_train``) on the B200 library: same function name, same arguments, same files in and out.

Only the trainer glue lives here -- the rest of the reference's ``movie_lens_data.py`` (CSV
ingest, train/test split, the ALS data-set "shrink") is data preparation and out of scope
(SURVEY.md section 2); its OUTPUT files are this function's input contract:

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
import os
import pickle
import time

from . import cpp_ls

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

"""Generates tests/golden/prep_*.npz by running the REAL reference data preparation
(/root/reference/python/full_data/movie_lens_data.py: the median step of
``refresh_training_sets_mp`` (:453-464) and ``als_data_set_shrink_mp`` (:547-680)) in its own
multi-process mode (1 parent + 2 workers) on small seeded inputs, inside a scratch working
directory (the module creates ./data/* and loads ./cpp_ls_lib.so at import time).

Everything is stored in RAW id space: the zero-based labels the reference assigns follow the
iteration order of a merged Python set and depend on the number of workers; the id tables it
writes next to the arrays are used here to undo them (the labels are stored too).
Run in the build container:   python tests/golden/make_golden_prep.py
"""
import os
import pickle
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (generator kwargs, factors_list, with_test_set)
    "small": (dict(num_users=300, num_movies=200, mean_degree=14, seed=11), [3, 8], False),
    "test_set": (dict(num_users=250, num_movies=150, mean_degree=18, seed=12), [5], True),
    "descending": (dict(num_users=300, num_movies=120, mean_degree=20, seed=13), [10, 4], False),
}


def main():
    from oracle import oracle
    from oracle.prep_oracle import flatten, synthetic_user_ratings
    cwd = os.getcwd()
    work = tempfile.mkdtemp(prefix="refprep_")
    os.makedirs(os.path.join(work, "data"))
    shutil.copy(oracle.ref_path(), os.path.join(work, "cpp_ls_lib.so"))
    os.chdir(work)
    sys.path.insert(0, "/root/reference/python/full_data")
    import movie_lens_data as mld
    import movie_lens_data_proc as _proc

    _proc.start_processes(3)
    try:
        for name, (kw, factors, with_test) in CASES.items():
            train = synthetic_user_ratings(**kw)
            test = None
            if with_test:   # a parallel list, as _compute_training_set produces (one entry per user)
                test = [(u, [(m, r) for m, r in e[:2]]) for u, e in train]
            user_pos, user_raw, movie_raw, ratings = flatten(train)
            out = dict(user_pos=user_pos, user_raw=user_raw, movie_raw=movie_raw, ratings=ratings,
                       factors=np.array(factors), with_test=np.array(with_test))

            _proc.clear_all_data()
            _proc.split_list_and_send([(u, list(e)) for u, e in train], "user_ratings_train")
            if with_test:
                _proc.split_list_and_send(test, "user_ratings_test")
            else:
                _proc.send_same_data({"user_ratings_test": None})
            # the median step of refresh_training_sets_mp (movie_lens_data.py:458-464)
            _proc.run_function("_extract_movie_ratings", {})
            lists = _proc.append_var_into_list("movie_ratings")
            merged = mld._merge_movie_ratings_lists(lists)
            _proc.split_list_and_send(merged, "movie_ratings")
            _proc.run_function("_compute_medians", {})
            medians = _proc.update_var_into_dict("movie_medians")
            _proc.delete_variable("movie_ratings")
            _proc.delete_variable("movie_medians")
            ids = np.array(sorted(medians), dtype=np.int64)
            out["median_ids"] = ids
            out["median_values"] = np.array([float(medians[int(m)]) for m in ids])

            mld.als_data_set_shrink_mp(medians, factors, no_test_set=not with_test)
            for k in factors:
                als_users = mld.get_als_obj("als%d_user_ids" % k)
                als_movies = mld.get_als_obj("als%d_movie_ids" % k)
                u, m, r = mld.get_als_obj("als%d_user_ratings_train" % k)
                inv_u = np.zeros(len(als_users), dtype=np.int64)
                for raw, z in als_users.items():
                    inv_u[z] = raw
                inv_m = np.zeros(len(als_movies), dtype=np.int64)
                for raw, z in als_movies.items():
                    inv_m[z] = raw
                assert u.dtype == np.int32 and m.dtype == np.int32 and r.dtype == np.float64
                out["k%d_user_labels" % k] = u
                out["k%d_movie_labels" % k] = m
                out["k%d_users_raw" % k] = inv_u[u] if len(u) else np.zeros(0, dtype=np.int64)
                out["k%d_movies_raw" % k] = inv_m[m] if len(m) else np.zeros(0, dtype=np.int64)
                out["k%d_ratings" % k] = r
                out["k%d_num_users" % k] = np.array(len(als_users))
                out["k%d_num_movies" % k] = np.array(len(als_movies))
                if with_test:
                    t = mld.get_als_obj("als%d_user_ratings_test" % k)
                    out["k%d_test_users" % k] = np.array([e[0] for e in t], dtype=np.int64)
                print(name, "k=%d:" % k, len(train), "users ->", len(als_users), ";",
                      len(set(movie_raw.tolist())), "movies ->", len(als_movies), ";",
                      len(ratings), "ratings ->", len(r))
            np.savez_compressed(os.path.join(HERE, "prep_%s.npz" % name), **out)
    finally:
        _proc.end_processes()
        os.chdir(cwd)
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()

"""CPU restatement of the reference's ALS data preparation -- TEST INFRASTRUCTURE ONLY.

Only ``tests/`` may import this module; nothing under ``movie_recommender_b200/`` does.

It restates, for the checker, what SURVEY.md section 8 row f2 names:

* per-movie medians of the training ratings
  (``python/full_data/movie_lens_data_proc.py:393-431`` ``_extract_movie_ratings`` and ``:455-471``
  ``_compute_medians``, driven by ``movie_lens_data.py:453-464``),
* the ALS data-set "shrink" (``movie_lens_data.py:547-680`` ``als_data_set_shrink_mp`` with its
  workers ``movie_lens_data_proc.py:494-654``): users with fewer than ``factor + 1`` ratings and
  movies with fewer than ``factor`` ratings are dropped until nothing changes, the surviving ids
  are renumbered from zero, and every rating has its movie's median subtracted.

Two forms: ``*_lists`` follows the reference's list-of-tuples loops line by line (pure Python,
small cases); ``*_coo`` is the same computation on flat NumPy arrays (moderate sizes).  Both are
pinned against outputs of the REAL reference functions run in multi-process mode
(``tests/golden/prep_*.npz``, made by ``tests/golden/make_golden_prep.py``).

Id numbering.  The reference numbers the surviving ids in the iteration order of a Python ``set``
that is merged from per-process sets (``movie_lens_data.py:590-609``,
``movie_lens_data_proc.py:246-261``): the order depends on the number of worker processes, so two
runs of the reference itself differ by a relabelling.  The checker therefore compares in RAW id
space (the relabelling is undone through the id tables the reference writes next to the arrays);
``reference_set_order`` reproduces the single-process labelling for completeness.
"""
import numpy as np


# --------------------------------------------------------------------------- list form
def medians_lists(user_ratings_train):
    """{movie id: numpy.median(ratings of the movie)} (_proc.py:393-431, 455-471)."""
    movie_ratings = {}
    for _, entries in user_ratings_train:
        for movie_id, rating in entries:
            movie_ratings.setdefault(movie_id, []).append(rating)
    return {m: np.median(r) for m, r in movie_ratings.items()}


def shrink_lists(user_ratings_train, factor, user_ratings_test=None):
    """The reference's fixpoint loop (movie_lens_data.py:569-588).  Returns the shrunk copies
    ``(train, test, rounds)``; the inputs are not modified."""
    train = [(u, list(e)) for u, e in user_ratings_train]
    test = None if user_ratings_test is None else list(user_ratings_test)
    rounds = 0
    has_changed = True
    while has_changed:
        rounds += 1
        # _drop_users (_proc.py:494-535), min_ratings = factor + 1
        has_changed = any(len(e) < factor + 1 for _, e in train)
        if has_changed:
            keep = [i for i in range(len(train)) if len(train[i][1]) >= factor + 1]
            if test is not None:
                test = [test[i] for i in keep]
            train = [train[i] for i in keep]
        # _count_movies (_proc.py:538-556)
        movie_counts = {}
        for _, e in train:
            for movie_id, _r in e:
                movie_counts[movie_id] = movie_counts.get(movie_id, 0) + 1
        uncommon = {m for m, c in movie_counts.items() if c < factor}
        # _drop_movies (_proc.py:559-586)
        if uncommon:
            has_changed = True
            train = [(u, [(m, r) for m, r in e if m not in uncommon]) for u, e in train]
    return train, test, rounds


def convert_lists(train, movie_medians, als_user_ids, als_movie_ids):
    """_convert_training_data_to_numpy (_proc.py:611-654)."""
    n = sum(len(e) for _, e in train)
    users = np.zeros(n, dtype=np.int32)
    movies = np.zeros(n, dtype=np.int32)
    ratings = np.zeros(n, dtype=np.double)
    i = 0
    for user_id, e in train:
        for movie_id, rating in e:
            users[i] = als_user_ids[user_id]
            movies[i] = als_movie_ids[movie_id]
            ratings[i] = rating - movie_medians[movie_id]
            i += 1
    return users, movies, ratings


def reference_set_order(train):
    """Zero-based ids the way a SINGLE-process reference run assigns them: iteration order of the
    sets ``_collect_ids`` builds (_proc.py:589-608; movie_lens_data.py:596-609)."""
    movie_ids, user_ids = set(), set()
    for user_id, e in train:
        for movie_id, _r in e:
            movie_ids.add(movie_id)
            user_ids.add(user_id)
    return ({u: i for i, u in enumerate(user_ids)}, {m: i for i, m in enumerate(movie_ids)})


def sorted_order(train):
    """Zero-based ids in ascending raw-id order (the labelling the GPU path emits)."""
    users = sorted({u for u, e in train if e})
    movies = sorted({m for _, e in train for m, _r in e})
    return ({u: i for i, u in enumerate(users)}, {m: i for i, m in enumerate(movies)})


# --------------------------------------------------------------------------- COO form
def flatten(user_ratings_train):
    """list form -> (user_pos int32[nnz], user_raw int64[n_entries], movie_raw int32[nnz],
    ratings f64[nnz]); user_pos is the index of the rating's entry in the list."""
    lens = np.array([len(e) for _, e in user_ratings_train], dtype=np.int64)
    user_raw = np.array([u for u, _ in user_ratings_train], dtype=np.int64)
    user_pos = np.repeat(np.arange(len(lens), dtype=np.int32), lens)
    movie_raw = np.array([m for _, e in user_ratings_train for m, _r in e], dtype=np.int32)
    ratings = np.array([r for _, e in user_ratings_train for _m, r in e], dtype=np.float64)
    return user_pos, user_raw, movie_raw, ratings


def medians_coo(movie_ids, ratings, num_movie_slots):
    """medians[m] = numpy.median of movie m's ratings (NaN where the movie has none), counts."""
    med = np.full(num_movie_slots, np.nan)
    order = np.lexsort((ratings, movie_ids))
    cnt = np.bincount(movie_ids, minlength=num_movie_slots).astype(np.int64)
    ptr = np.concatenate(([0], np.cumsum(cnt)))
    rs = ratings[order]
    has = cnt > 0
    lo = rs[(ptr[:-1] + (cnt - 1) // 2)[has]]
    hi = rs[(ptr[:-1] + cnt // 2)[has]]
    # numpy.median: the middle element, or mean() of the two middle elements = (a + b) / 2
    med[has] = np.where(cnt[has] % 2 == 1, hi, (lo + hi) / 2.0)
    return med, cnt.astype(np.int32)


def shrink_coo(user_pos, movie_ids, ratings, num_user_slots, num_movie_slots, medians,
               min_user_ratings, min_movie_ratings):
    """Fixpoint degree filter + stable compaction + ascending renumbering + median subtraction.
    Returns dict(user_ids, movie_ids, ratings, keep_pos, user_new_id, movie_new_id, rounds)."""
    user_ok = np.ones(num_user_slots, dtype=bool)
    movie_ok = np.ones(num_movie_slots, dtype=bool)
    rounds = 0
    while True:
        rounds += 1
        cu = np.bincount(user_pos[movie_ok[movie_ids]], minlength=num_user_slots)
        new_user_ok = cu >= min_user_ratings
        cm = np.bincount(movie_ids[new_user_ok[user_pos] & movie_ok[movie_ids]],
                         minlength=num_movie_slots)
        new_movie_ok = cm >= min_movie_ratings
        # the reference's has_changed: a listed user fell short, or a counted movie fell short
        changed = bool(np.any(user_ok & ~new_user_ok)) or bool(np.any((cm > 0) & ~new_movie_ok))
        user_ok, movie_ok = user_ok & new_user_ok, movie_ok & new_movie_ok
        if not changed:
            break
    alive = user_ok[user_pos] & movie_ok[movie_ids]
    # ids that still carry at least one rating (_collect_ids only sees those)
    user_live = np.bincount(user_pos[alive], minlength=num_user_slots) > 0
    movie_live = np.bincount(movie_ids[alive], minlength=num_movie_slots) > 0
    user_new = np.where(user_live, np.cumsum(user_live) - 1, -1).astype(np.int32)
    movie_new = np.where(movie_live, np.cumsum(movie_live) - 1, -1).astype(np.int32)
    keep = np.nonzero(alive)[0].astype(np.int32)
    return dict(user_ids=user_new[user_pos[keep]], movie_ids=movie_new[movie_ids[keep]],
                ratings=ratings[keep] - medians[movie_ids[keep]], keep_pos=keep,
                user_new_id=user_new, movie_new_id=movie_new, rounds=rounds)


def synthetic_user_ratings(num_users, num_movies, mean_degree, seed, id_gap=3):
    """A small MovieLens-like ``[(user id, [(movie id, rating)])]`` with a long tail of rare
    movies and light users (so that the shrink needs several rounds), raw ids with gaps."""
    rng = np.random.default_rng(seed)
    pop = 1.0 / np.arange(1, num_movies + 1) ** 0.9
    pop /= pop.sum()
    out = []
    for u in range(num_users):
        deg = int(min(num_movies, max(1, rng.geometric(1.0 / mean_degree))))
        movies = rng.choice(num_movies, size=deg, replace=False, p=pop)
        vals = np.clip(np.round(rng.normal(3.5, 1.1, size=deg) * 2) / 2, 0.5, 5.0)
        out.append((1 + u * id_gap, [(int(1 + m * id_gap + (m % 2)), float(v))
                                     for m, v in zip(movies, vals)]))
    return out

"""Seeded MovieLens-shaped synthetic inputs for the ALS / least-squares / similarity path.

There is no network and the reference ships no data set, so every test and bench line uses
synthetic ratings of the reference's input contract:

  * COO triples ``user_ids int32[nnz]``, ``item_ids int32[nnz]``, ``ratings float64[nnz]``,
    ids zero-based and contiguous (python/full_data/cpp_ls.py:120-123), ordered grouped by user
    as the reference emits them (python/full_data/movie_lens_data_proc.py:641-650);
  * ratings are multiples of 0.5 in [0.5, 5] minus the per-movie median
    (movie_lens_data_proc.py:455-471, :648);
  * every user has >= k+1 ratings and every movie >= k (the reference's "shrink" rule,
    python/full_data/movie_lens_data.py:568-591) -- here by construction.

Everything is a pure function of the seed (``numpy.random.default_rng``); nothing reads the
global NumPy RNG.
"""
import numpy as np

DEFAULT_SEED = 20181001

# (num_users, num_items, num_ratings, k) of the configurations BASELINE.json names.
CONFIGS = {
    "C1": dict(num_users=610, num_items=9724, num_ratings=100836, k=10),
    "C3": dict(num_users=283228, num_items=53889, num_ratings=27753444, k=50),
}


def _power_weights(n, exponent, head, rng):
    # (rank + head)^-exponent: a power-law tail with a flattened head, so that the most popular
    # movie holds ~0.3 % of all ratings as in MovieLens instead of saturating at every user
    w = (np.arange(1, n + 1, dtype=np.float64) + head) ** (-exponent)
    rng.shuffle(w)
    return w / w.sum()


def _degrees(keys, num_users, num_items):
    u = keys // num_items
    i = keys - u * num_items
    return np.bincount(u, minlength=num_users), np.bincount(i, minlength=num_items)


def _cumcount(sorted_group):
    """Rank of each element inside its run of equal values (input must be grouped)."""
    n = len(sorted_group)
    if n == 0:
        return np.zeros(0, dtype=np.int64)
    start = np.flatnonzero(np.r_[True, sorted_group[1:] != sorted_group[:-1]])
    run_len = np.diff(np.r_[start, n])
    return np.arange(n) - np.repeat(start, run_len)


def _sorted_unique(a):
    """Sorted distinct values (numpy 2.3's hash-based np.unique is ~5x slower at 3e7 keys)."""
    a = np.sort(a, kind="stable")
    if len(a) < 2:
        return a
    return a[np.r_[True, a[1:] != a[:-1]]]


def _not_in_sorted(candidates, sorted_keys):
    pos = np.searchsorted(sorted_keys, candidates)
    pos[pos == len(sorted_keys)] = 0
    return candidates[sorted_keys[pos] != candidates] if len(sorted_keys) else candidates


def rating_pairs(num_users, num_items, num_ratings, min_user_deg=0, min_item_deg=0,
                 seed=DEFAULT_SEED, user_exponent=0.6, item_exponent=0.9, user_head=20.0,
                 item_head=30.0):
    """Unique (user, item) pairs with power-law user activity and Zipf-like item popularity,
    minimum degrees enforced, exactly ``num_ratings`` pairs when that is attainable (more if
    the minimum degrees alone need more).

    Returns ``(user_ids int32, item_ids int32)`` sorted by (user, item).
    """
    rng = np.random.default_rng(seed)
    nu, ni = int(num_users), int(num_items)
    target = int(num_ratings)
    if target > nu * ni:
        raise ValueError("more ratings than user x item pairs")
    wu = np.cumsum(_power_weights(nu, user_exponent, user_head, rng))
    wi = np.cumsum(_power_weights(ni, item_exponent, item_head, rng))

    def draw(m):
        u = np.minimum(np.searchsorted(wu, rng.random(m)), nu - 1).astype(np.int64)
        i = np.minimum(np.searchsorted(wi, rng.random(m)), ni - 1).astype(np.int64)
        return u * ni + i

    # 1. popularity-driven base, deliberately below the target
    floor_need = nu * min_user_deg + ni * min_item_deg
    base = int(max(0, min(0.85 * target, target - 0.6 * floor_need)))
    keys = _sorted_unique(draw(base)) if base > 0 else np.zeros(0, dtype=np.int64)
    # 2. minimum degrees: consecutive partners from a random start (distinct by construction)
    for _ in range(100):
        du, di = _degrees(keys, nu, ni)
        need_u = np.maximum(min_user_deg - du, 0)
        need_i = np.maximum(min_item_deg - di, 0)
        extra = []
        if need_u.any():
            users = np.flatnonzero(need_u)
            cnt = need_u[users]
            start = rng.integers(0, ni, size=len(users))
            off = _cumcount(np.repeat(np.arange(len(users)), cnt))
            items = (np.repeat(start, cnt) + off) % ni
            extra.append(np.repeat(users, cnt).astype(np.int64) * ni + items)
        if need_i.any():
            items = np.flatnonzero(need_i)
            cnt = need_i[items]
            start = rng.integers(0, nu, size=len(items))
            off = _cumcount(np.repeat(np.arange(len(items)), cnt))
            users = (np.repeat(start, cnt) + off) % nu
            extra.append(users.astype(np.int64) * ni + np.repeat(items, cnt))
        if not extra:
            break
        keys = _sorted_unique(np.concatenate([keys] + extra))
    # 3. top up with popularity-driven pairs to exactly the target (adding never breaks a minimum)
    for _ in range(100):
        missing = target - len(keys)
        if missing <= 0:
            break
        cand = _sorted_unique(draw(int(missing * 1.25) + 64))
        cand = _not_in_sorted(cand, keys)
        if len(cand) > missing:
            cand = np.sort(cand[rng.permutation(len(cand))[:missing]])
        keys = np.sort(np.concatenate([keys, cand]), kind="stable")
    u = (keys // ni).astype(np.int32)
    i = (keys - (keys // ni) * ni).astype(np.int32)
    return u, i


def movie_medians(item_ids, raw_ratings, num_items):
    """Per-movie median of the raw ratings (numpy.median semantics: mean of the two middle
    values for even counts), movie_lens_data_proc.py:455-471.  Movies without ratings get 0."""
    order = np.lexsort((raw_ratings, item_ids))
    cnt = np.bincount(item_ids, minlength=num_items)
    ptr = np.concatenate([[0], np.cumsum(cnt)])
    sr = raw_ratings[order]
    med = np.zeros(num_items, dtype=np.float64)
    has = cnt > 0
    lo = ptr[:-1] + (cnt - 1) // 2
    hi = ptr[:-1] + cnt // 2
    med[has] = 0.5 * (sr[lo[has]] + sr[hi[has]])
    return med


def _planted_raw(user_ids, item_ids, num_users, num_items, seed, rank, noise, noise_rng=None):
    """Raw (0.5-step) ratings of the planted model for the given pairs.  The model parameters
    are the first four draws of rng(seed + 1); the noise follows on the same generator unless
    ``noise_rng`` supplies another one (held-out pairs)."""
    rng = np.random.default_rng(seed + 1)
    pu = rng.standard_normal((num_users, rank)) * (0.9 / np.sqrt(rank))
    pv = rng.standard_normal((num_items, rank))
    bias = rng.standard_normal(num_users) * 0.4
    item_off = rng.standard_normal(num_items) * 0.5
    n = len(user_ids)
    raw = np.empty(n, dtype=np.float64)
    step = 1 << 22
    for s in range(0, n, step):
        u = user_ids[s:s + step]
        i = item_ids[s:s + step]
        raw[s:s + step] = 3.4 + bias[u] + item_off[i] + np.einsum("ij,ij->i", pu[u], pv[i])
    raw += (rng if noise_rng is None else noise_rng).standard_normal(n) * noise
    return np.clip(np.round(raw * 2.0) / 2.0, 0.5, 5.0)


def planted_ratings(user_ids, item_ids, num_users, num_items, seed=DEFAULT_SEED, rank=8,
                    noise=0.4, subtract_median=True):
    """Ratings from a planted rank-``rank`` model + user bias + N(0, noise), rounded to 0.5
    steps in [0.5, 5]; then minus the per-movie median (what ``cpp_ls.als`` is fed)."""
    raw = _planted_raw(user_ids, item_ids, num_users, num_items, seed, rank, noise)
    if not subtract_median:
        return raw
    return raw - movie_medians(item_ids, raw, num_items)[item_ids]


def heldout_ratings(train_user_ids, train_item_ids, num_users, num_items, count,
                    seed=DEFAULT_SEED, rank=8, noise=0.4, medians=None):
    """``count`` (user, movie) pairs that are NOT in the training set, drawn from the same
    activity / popularity laws and rated by the same planted model (fresh noise), minus the
    TRAINING set's movie medians -- the test split the reference evaluates on is prepared the
    same way (python/full_data/movie_lens_data.py: medians come from the training set).
    Returns ``(user_ids int32, item_ids int32, ratings f64)`` sorted by (user, movie)."""
    nu, ni = int(num_users), int(num_items)
    rng0 = np.random.default_rng(seed)           # the same two shuffles as rating_pairs
    wu = np.cumsum(_power_weights(nu, 0.6, 20.0, rng0))
    wi = np.cumsum(_power_weights(ni, 0.9, 30.0, rng0))
    rng = np.random.default_rng(seed + 6)
    train_keys = train_user_ids.astype(np.int64) * ni + train_item_ids
    if len(train_keys) > 1 and np.any(train_keys[1:] < train_keys[:-1]):
        train_keys = np.sort(train_keys)
    keys = np.zeros(0, dtype=np.int64)
    for _ in range(100):
        missing = int(count) - len(keys)
        if missing <= 0:
            break
        m = int(missing * 1.25) + 64
        u = np.minimum(np.searchsorted(wu, rng.random(m)), nu - 1).astype(np.int64)
        i = np.minimum(np.searchsorted(wi, rng.random(m)), ni - 1).astype(np.int64)
        cand = _not_in_sorted(_not_in_sorted(_sorted_unique(u * ni + i), train_keys), keys)
        if len(cand) > missing:
            cand = np.sort(cand[rng.permutation(len(cand))[:missing]])
        keys = np.sort(np.concatenate([keys, cand]), kind="stable")
    hu = (keys // ni).astype(np.int32)
    hi = (keys - (keys // ni) * ni).astype(np.int32)
    med = medians
    if med is None:
        raw_train = _planted_raw(train_user_ids, train_item_ids, nu, ni, seed, rank, noise)
        med = movie_medians(train_item_ids, raw_train, ni)
    raw = _planted_raw(hu, hi, nu, ni, seed, rank, noise, noise_rng=rng)
    return hu, hi, raw - med[hi]


def initial_factors(num_users, num_items, k, seed=DEFAULT_SEED):
    """U(-1, 1) initial factors, the distribution python/full_data/cpp_ls.py:150-151 draws."""
    rng = np.random.default_rng(seed + 2)
    uf = rng.uniform(-1, 1, num_users * (k + 1))
    itf = rng.uniform(-1, 1, num_items * k)
    return uf, itf


def als_problem(num_users, num_items, num_ratings, k, seed=DEFAULT_SEED, min_degrees=True,
                shuffle=False, heldout=0):
    """One ALS training problem of the given shape: dict with the COO triples, k and U(-1,1)
    initial factors.  ``shuffle=True`` randomises the COO order (the order cpp_ls_test.py
    feeds, cpp/python/cpp_ls_test.py:110-116) instead of the trainer's grouped-by-user order.
    ``heldout=N`` adds N held-out ratings of the same model (``heldout_user_ids`` ...)."""
    u, i = rating_pairs(num_users, num_items, num_ratings,
                        min_user_deg=(k + 1 if min_degrees else 0),
                        min_item_deg=(k if min_degrees else 0), seed=seed)
    raw = planted_ratings(u, i, num_users, num_items, seed=seed, subtract_median=False)
    med = movie_medians(i, raw, num_items)
    r = raw - med[i]
    held = None
    if heldout:
        held = heldout_ratings(u, i, num_users, num_items, heldout, seed=seed, medians=med)
    if shuffle:
        perm = np.random.default_rng(seed + 3).permutation(len(u))
        u, i, r = u[perm], i[perm], r[perm]
    uf, itf = initial_factors(num_users, num_items, k, seed=seed)
    out = dict(user_ids=np.ascontiguousarray(u), item_ids=np.ascontiguousarray(i),
               ratings=np.ascontiguousarray(r), k=k, num_users=num_users, num_items=num_items,
               user_factors0=uf, item_factors0=itf)
    if held is not None:
        out.update(heldout_user_ids=held[0], heldout_item_ids=held[1], heldout_ratings=held[2])
    return out


def bias_model_system(user_ids, item_ids, raw_ratings, num_users, num_items, seed=DEFAULT_SEED):
    """The user+movie-bias least-squares model expressed through the reference's generic CSR
    solver API: one row per rating, two non-zeros (value 1.0) at
    columns ``user`` and ``num_users + item``; right-hand side = raw ratings.
    Returns (row_indices, col_indices, values, num_columns, b, x0)."""
    n = len(user_ids)
    rowptr = (np.arange(n + 1, dtype=np.int64) * 2).astype(np.int32)
    col = np.empty(2 * n, dtype=np.int32)
    col[0::2] = user_ids
    col[1::2] = num_users + item_ids
    vals = np.ones(2 * n, dtype=np.float64)
    cols = num_users + num_items
    x0 = np.random.default_rng(seed + 4).uniform(-1, 1, cols)
    return rowptr, col, vals, cols, np.ascontiguousarray(raw_ratings, dtype=np.float64), x0


def random_sparse_system(rows, cols, nnz_per_row, seed=DEFAULT_SEED, noise=0.1):
    """A planted sparse least-squares problem like the reference's self-test
    (cpp/ls/main.cpp:356-464): random CSR A, x_real ~ U(-1,1), b = A x_real + N(0, noise)."""
    rng = np.random.default_rng(seed + 5)
    per = np.minimum(nnz_per_row, cols)
    col = np.empty((rows, per), dtype=np.int32)
    for s in range(0, rows, 65536):
        e = min(rows, s + 65536)
        col[s:e] = np.argsort(rng.random((e - s, cols)), axis=1)[:, :per] if cols <= 4096 else \
            np.sort(rng.integers(0, cols, size=(e - s, per)), axis=1)
    col.sort(axis=1)
    rowptr = (np.arange(rows + 1, dtype=np.int64) * per).astype(np.int32)
    vals = rng.uniform(-1, 1, rows * per)
    x_real = rng.uniform(-1, 1, cols)
    b = np.zeros(rows)
    np.add.at(b, np.repeat(np.arange(rows), per), vals * x_real[col.reshape(-1)])
    b += rng.standard_normal(rows) * noise
    x0 = rng.uniform(-1, 1, cols)
    return rowptr, col.reshape(-1).copy(), vals, cols, b, x0, x_real

"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY (the checker, never the product).

ctypes bindings for
  * ``liboracle.so``        -- our plain-C restatement (oracle/ls_oracle.c), and
  * ``_ref/cpp_ls_lib.so``  -- the UNMODIFIED reference library compiled from /root/reference
                               (see oracle/Makefile); optional, present when it was built in
                               the build container (it travels to the GPU box as a binary).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may
import this module.  Nothing under movie_recommender_b200/ imports it.

The reference's own Python wrapper (python/full_data/cpp_ls.py:114-172) draws the initial
factors from the global NumPy RNG; to be deterministic the bindings here take the initial
vectors explicitly and call the C symbols directly.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_I = ctypes.POINTER(ctypes.c_int)
_D = ctypes.POINTER(ctypes.c_double)


def build(quiet=True):
    """Compile liboracle.so (and _ref/cpp_ls_lib.so when the reference sources are present)."""
    out = subprocess.run(["make", "-C", _HERE, "all"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)


def _ip(a):
    return a.ctypes.data_as(_I)


def _dp(a):
    return a.ctypes.data_as(_D)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _lib = ctypes.CDLL(path)
        _lib.oracle_dot.restype = ctypes.c_double
        _lib.oracle_dot.argtypes = [_D, _D, ctypes.c_int, ctypes.c_int]
        _lib.oracle_cg_least_squares.restype = ctypes.c_int
        _lib.oracle_cg_least_squares.argtypes = [
            ctypes.c_int, ctypes.c_int, _I, _I, _D, _D, _D, ctypes.c_double, ctypes.c_int, _D,
            ctypes.c_int, ctypes.c_int]
        _lib.oracle_als.restype = ctypes.c_int
        _lib.oracle_als.argtypes = [
            _I, _I, ctypes.c_int, _D, ctypes.c_int, ctypes.c_int, _D, ctypes.c_int, _D,
            ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    return _lib


# ----------------------------------------------------------------------------- restatement
def chunk_table(T, length):
    out = np.zeros(T + 1, dtype=np.int32)
    lib().oracle_chunk_table(ctypes.c_int(T), ctypes.c_int(length), _ip(out))
    return out


def dot(a, b, T):
    a, b = _f64(a), _f64(b)
    return lib().oracle_dot(_dp(a), _dp(b), len(a), T)


def group_by(key, groups):
    """Stable grouping: (ptr[groups+1], idx[n]) == CSR of positions by key, input order kept."""
    key = _i32(key)
    ptr = np.zeros(groups + 1, dtype=np.int32)
    idx = np.zeros(max(len(key), 1), dtype=np.int32)
    lib().oracle_group_by(_ip(key), ctypes.c_int(len(key)), ctypes.c_int(groups), _ip(ptr), _ip(idx))
    return ptr, idx[:len(key)]


def transpose(rows, cols, rowptr, colidx, vals):
    rowptr, colidx, vals = _i32(rowptr), _i32(colidx), _f64(vals)
    nnz = int(rowptr[rows])
    t_ptr = np.zeros(cols + 1, dtype=np.int32)
    t_row = np.zeros(max(nnz, 1), dtype=np.int32)
    t_val = np.zeros(max(nnz, 1), dtype=np.float64)
    lib().oracle_transpose(ctypes.c_int(rows), ctypes.c_int(cols), _ip(rowptr), _ip(colidx),
                           _dp(vals), _ip(t_ptr), _ip(t_row), _dp(t_val))
    return t_ptr, t_row[:nnz], t_val[:nnz]


def cg_least_squares(rowptr, colidx, vals, cols, b, x0, min_r_decrease=0.01, max_iterations=200,
                     algorithm=1, thread_count=1):
    """Restated cg_least_squares / cg_least_squares2.  Returns (x, iterations, final_rr)."""
    rowptr, colidx, vals, b = _i32(rowptr), _i32(colidx), _f64(vals), _f64(b)
    x = _f64(x0).copy().reshape(-1)
    rr = ctypes.c_double(0)
    it = lib().oracle_cg_least_squares(
        len(rowptr) - 1, cols, _ip(rowptr), _ip(colidx), _dp(vals), _dp(b), _dp(x),
        min_r_decrease, max_iterations, ctypes.cast(ctypes.byref(rr), _D), thread_count,
        1 if algorithm == 1 else 2)
    return x, it, rr.value


def als(user_ids, item_ids, ratings, k, user_factors0, item_factors0, min_r_decrease=0.01,
        max_iterations=200, algorithm=1, thread_count=1):
    """Restated als().  Returns (user_factors, item_factors, iterations)."""
    user_ids, item_ids, ratings = _i32(user_ids), _i32(item_ids), _f64(ratings)
    uf = _f64(user_factors0).copy().reshape(-1)
    itf = _f64(item_factors0).copy().reshape(-1)
    it = lib().oracle_als(_ip(user_ids), _ip(item_ids), len(ratings), _dp(ratings), k,
                          len(uf), _dp(uf), len(itf), _dp(itf), min_r_decrease, max_iterations,
                          algorithm, thread_count)
    return uf, itf, it


def als_predict(user_ids, item_ids, k, user_factors, item_factors):
    user_ids, item_ids = _i32(user_ids), _i32(item_ids)
    uf, itf = _f64(user_factors).reshape(-1), _f64(item_factors).reshape(-1)
    out = np.zeros(len(user_ids), dtype=np.float64)
    lib().oracle_als_predict(_ip(user_ids), _ip(item_ids), ctypes.c_int(len(user_ids)),
                             ctypes.c_int(k), _dp(uf), _dp(itf), _dp(out))
    return out


def rmse(user_ids, item_ids, ratings, k, user_factors, item_factors):
    """Training RMSE of the model (the reference has no RMSE)."""
    pred = als_predict(user_ids, item_ids, k, user_factors, item_factors)
    d = pred - _f64(ratings)
    return float(np.sqrt(np.mean(d * d))) if len(d) else 0.0


def cosine_topk(M, topk, q_lo=0, q_hi=None):
    """Factor-cosine top-k (PARITY UNPINNED: no reference counterpart, see ls_oracle.c).
    Returns (ids[q, topk] int32 (-1 padded), scores[q, topk] float64) for queries q_lo..q_hi-1."""
    M = _f64(M)
    n, k = M.shape
    q_hi = n if q_hi is None else q_hi
    ids = np.zeros((q_hi - q_lo, topk), dtype=np.int32)
    scores = np.zeros((q_hi - q_lo, topk), dtype=np.float64)
    lib().oracle_cosine_topk(_dp(M), ctypes.c_int(n), ctypes.c_int(k), ctypes.c_int(topk),
                             ctypes.c_int(q_lo), ctypes.c_int(q_hi), _ip(ids), _dp(scores))
    return ids, scores


def normalize_rows(M):
    M = _f64(M)
    out = np.zeros_like(M)
    lib().oracle_normalize_rows(_dp(M), ctypes.c_int(M.shape[0]), ctypes.c_int(M.shape[1]), _dp(out))
    return out


# ----------------------------------------------------------------------- the real reference
_ref = None


def ref_path():
    return os.path.join(_HERE, "_ref", "cpp_ls_lib.so")


def has_ref():
    return os.path.exists(ref_path())


def ref():
    """The unmodified reference library (raises if it was not built)."""
    global _ref
    if _ref is None:
        if not has_ref():
            raise RuntimeError("oracle/_ref/cpp_ls_lib.so not built (reference sources absent)")
        _ref = ctypes.CDLL(ref_path())
    return _ref


def ref_cg_least_squares(rowptr, colidx, vals, cols, b, x0, min_r_decrease=0.01,
                         max_iterations=200, algorithm=1, thread_count=1):
    """cg_least_squares_from_python / cg_least_squares2_from_python of the real reference
    (cpp/ls_lib/ls_linux_dll.cpp:28-77) with an explicit x0."""
    r = ref()
    rowptr, colidx, vals, b = _i32(rowptr), _i32(colidx), _f64(vals), _f64(b)
    x = _f64(x0).copy().reshape(-1)
    rr = ctypes.c_double(0)
    r.set_thread_count(thread_count)
    fn = r.cg_least_squares_from_python if algorithm == 1 else r.cg_least_squares2_from_python
    it = fn(len(rowptr) - 1, cols, _ip(rowptr), _ip(colidx), _dp(vals), len(b), _dp(b), len(x),
            _dp(x), ctypes.c_double(min_r_decrease), max_iterations, ctypes.byref(rr))
    return x, it, rr.value


def ref_als(user_ids, item_ids, ratings, k, user_factors0, item_factors0, min_r_decrease=0.01,
            max_iterations=200, algorithm=1, thread_count=1):
    """als_from_python of the real reference (ls_linux_dll.cpp:81-103) with explicit factors."""
    r = ref()
    user_ids, item_ids, ratings = _i32(user_ids), _i32(item_ids), _f64(ratings)
    uf = _f64(user_factors0).copy().reshape(-1)
    itf = _f64(item_factors0).copy().reshape(-1)
    r.set_thread_count(thread_count)
    it = r.als_from_python(_ip(user_ids), _ip(item_ids), len(ratings), _dp(ratings), k, len(uf),
                           _dp(uf), len(itf), _dp(itf), ctypes.c_double(min_r_decrease),
                           max_iterations, algorithm)
    return uf, itf, it

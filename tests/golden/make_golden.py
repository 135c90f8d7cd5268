"""Generates tests/golden/*.npz by running the REAL reference library (oracle/_ref/cpp_ls_lib.so,
compiled unchanged from /root/reference/cpp/ls_lib by oracle/Makefile) on small seeded inputs.

Run in the build container (where /root/reference exists):  python tests/golden/make_golden.py
The .npz files are committed; the GPU box has no /root/reference and only reads them.

Cases mirror the reference's own self-tests:
  ls_200x50   -- cpp/python/cpp_ls_test.py:5-39  (200 x 50 dense-as-CSR, noise 0.1)
  ls_sparse   -- cpp/ls/main.cpp:356-464 shape class (1000 x 50, 5 nnz/row)
  als_planted -- cpp/python/cpp_ls_test.py:73-147 (k=5, all pairs of 38 users x 45 items, 80 %
                 shuffled training split, noise 0.1)
each at thread_count 1 and 4 and for both algorithm variants.
"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
THREADS = (1, 4)


def dense_to_csr(A):
    rows, cols = A.shape
    rowptr = (np.arange(rows + 1) * cols).astype(np.int32)
    colidx = np.tile(np.arange(cols, dtype=np.int32), rows)
    return rowptr, colidx, A.reshape(-1).copy()


def ls_cases():
    rng = np.random.default_rng(4242)
    A = rng.uniform(-1, 1, (200, 50))
    x_real = rng.uniform(-1, 1, 50)
    b = A @ x_real + rng.normal(0, 0.1, 200)
    x0 = rng.uniform(-1, 1, 50)
    rowptr, colidx, vals = dense_to_csr(A)
    yield "ls_200x50", dict(rowptr=rowptr, colidx=colidx, vals=vals, cols=50, b=b, x0=x0)

    rows, cols, per = 1000, 50, 5
    colidx = np.sort(np.argsort(rng.random((rows, cols)), axis=1)[:, :per], axis=1).astype(np.int32)
    vals = rng.uniform(-1, 1, rows * per)
    x_real = rng.uniform(-1, 1, cols)
    b = (vals.reshape(rows, per) * x_real[colidx]).sum(axis=1) + rng.normal(0, 0.1, rows)
    x0 = rng.uniform(-1, 1, cols)
    rowptr = (np.arange(rows + 1) * per).astype(np.int32)
    yield "ls_sparse", dict(rowptr=rowptr, colidx=colidx.reshape(-1).copy(), vals=vals, cols=cols,
                            b=b, x0=x0)


def als_case():
    rng = np.random.default_rng(777)
    k, ratio, mult = 5, 0.8, 6.0
    n = k + 1
    num_items = math.ceil(n * mult / ratio)
    num_users = math.ceil(k * mult / ratio)
    uf_real = rng.uniform(-1, 1, (num_users, n))
    if_real = rng.uniform(-1, 1, (num_items, k))
    u, i = np.meshgrid(np.arange(num_users), np.arange(num_items), indexing="ij")
    u, i = u.reshape(-1).astype(np.int32), i.reshape(-1).astype(np.int32)
    r = (uf_real[u, :k] * if_real[i]).sum(axis=1) + uf_real[u, k] + rng.normal(0, 0.1, len(u))
    perm = rng.permutation(len(u))
    m = math.ceil(len(u) * ratio)
    tr = perm[:m]
    uf0 = rng.uniform(-1, 1, num_users * n)
    if0 = rng.uniform(-1, 1, num_items * k)
    return "als_planted", dict(user_ids=u[tr].copy(), item_ids=i[tr].copy(), ratings=r[tr].copy(),
                               k=k, num_users=num_users, num_items=num_items, uf0=uf0, if0=if0)


def main():
    assert oracle.has_ref(), "build oracle/_ref first (make -C oracle)"
    for name, c in ls_cases():
        out = dict(c)
        for T in THREADS:
            for alg in (1, 2):
                x, it, rr = oracle.ref_cg_least_squares(c["rowptr"], c["colidx"], c["vals"],
                                                        c["cols"], c["b"], c["x0"],
                                                        algorithm=alg, thread_count=T)
                out["x_T%d_a%d" % (T, alg)] = x
                out["it_T%d_a%d" % (T, alg)] = np.int32(it)
                out["rr_T%d_a%d" % (T, alg)] = np.float64(rr)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, {k: v for k, v in out.items() if k.startswith("it_")})
    name, c = als_case()
    out = dict(c)
    for T in THREADS:
        for alg in (1, 2):
            uf, itf, it = oracle.ref_als(c["user_ids"], c["item_ids"], c["ratings"], c["k"],
                                         c["uf0"], c["if0"], algorithm=alg, thread_count=T)
            out["uf_T%d_a%d" % (T, alg)] = uf
            out["if_T%d_a%d" % (T, alg)] = itf
            out["it_T%d_a%d" % (T, alg)] = np.int32(it)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k: v for k, v in out.items() if k.startswith("it_")})


if __name__ == "__main__":
    main()

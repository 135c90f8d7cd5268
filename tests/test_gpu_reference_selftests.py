"""The reference's own self-tests (cpp/python/cpp_ls_test.py:5-39 and :73-147), re-run unchanged in
spirit through the drop-in Python boundary: same shapes, same random recipe (global NumPy RNG, as
the reference uses), same acceptance thresholds -- for every algorithm value."""
import math

import numpy
import pytest

pytestmark = pytest.mark.gpu


def convert_dense_matrix_to_sparse_format(A):
    rows, cols = A.shape
    return (numpy.arange(rows + 1, dtype=numpy.int32) * cols,
            numpy.tile(numpy.arange(cols, dtype=numpy.int32), rows),
            A.reshape(-1).astype(numpy.double))


@pytest.mark.parametrize("algorithm", [1, 2, 3])
def test_cg_least_squares_selftest(require_gpu, cpp_ls, algorithm):
    """cpp_ls_test.py:5-39: 200 x 50, noise 0.1, mean coefficient error < 0.1."""
    numpy.random.seed(100 + algorithm)
    A_size = (200, 50)
    A = numpy.random.uniform(-1, 1, A_size)
    x_real = numpy.random.uniform(-1, 1, (A_size[1], 1))
    b = A.dot(x_real) + numpy.random.normal(0, 0.1, (A_size[0], 1))
    A_row_indices, A_col_indices, A_values = convert_dense_matrix_to_sparse_format(A)
    x, iterations, final_rr = cpp_ls.cg_least_squares(
        A_row_indices, A_col_indices, A_values, A_size[1], b.reshape(-1), algorithm=algorithm)
    assert x.shape == (50, 1)
    average_error = numpy.sum(numpy.abs(x_real - x)) / len(x_real)
    assert average_error < 0.1 and 0 < iterations <= 200


@pytest.mark.parametrize("algorithm", [1, 2, 3, 4])
def test_als_selftest(require_gpu, cpp_ls, algorithm):
    """cpp_ls_test.py:73-147: planted k = 5 model, all user x item pairs, noise 0.1, shuffled,
    80 % training split; held-out mean absolute error < 0.15."""
    numpy.random.seed(200 + algorithm)
    num_item_factors, training_set_ratio, k = 5, 0.8, 6.0
    num_user_factors = num_item_factors + 1
    num_items = math.ceil(num_user_factors * k / training_set_ratio)
    num_users = math.ceil(num_item_factors * k / training_set_ratio)
    user_factors_real = numpy.random.uniform(-1, 1, num_users * num_user_factors)
    item_factors_real = numpy.random.uniform(-1, 1, num_items * num_item_factors)

    def als_predict(u, i, uf, itf):
        uf, itf = uf.reshape(num_users, num_user_factors), itf.reshape(num_items, num_item_factors)
        return (uf[u, :num_item_factors] * itf[i]).sum(axis=-1) + uf[u, num_item_factors]

    user_ids, item_ids = numpy.meshgrid(numpy.arange(num_users), numpy.arange(num_items), indexing="ij")
    user_ids = user_ids.reshape(-1).astype(numpy.int32)
    item_ids = item_ids.reshape(-1).astype(numpy.int32)
    ratings = als_predict(user_ids, item_ids, user_factors_real, item_factors_real) \
        + numpy.random.normal(0, 0.1, len(user_ids))
    perm = numpy.random.permutation(len(user_ids))
    user_ids, item_ids, ratings = user_ids[perm], item_ids[perm], ratings[perm]
    n_train = math.ceil(len(user_ids) * training_set_ratio)
    user_factors, item_factors, iterations = cpp_ls.als(
        user_ids[:n_train], item_ids[:n_train], ratings[:n_train], num_item_factors, num_users,
        num_items, algorithm=algorithm)
    pred = als_predict(user_ids[n_train:], item_ids[n_train:], user_factors, item_factors)
    average_error = numpy.sum(numpy.abs(pred - ratings[n_train:])) / len(pred)
    assert average_error < 0.15, (algorithm, average_error, iterations)


def test_has_dll_loaded(require_gpu, cpp_ls):
    assert cpp_ls.has_dll_loaded()

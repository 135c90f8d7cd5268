"""Tiny run through every kernel family, for compute-sanitizer memcheck (development aid)."""
import sys
sys.path.insert(0, ".")
import numpy as np
from movie_recommender_b200 import cpp_ls, synth, similarity
from movie_recommender_b200.build_similar_movies_db import SimilarMovieFinder
from movie_recommender_b200.fold_in import fold_in_users
from oracle.similar_oracle import synthetic_catalogue

cpp_ls.set_thread_count(3)
p = synth.als_problem(70, 90, 3000, 7, seed=1, min_degrees=False)
a = (p["user_ids"], p["item_ids"], p["ratings"], 7, 70, 90)
for alg in (1, 2, 3, 4):
    cpp_ls.als(*a, -1e300, 2, alg, user_factors=p["user_factors0"], item_factors=p["item_factors0"])
p = synth.als_problem(80, 70, 5400, 64, seed=2)            # wide rank
for alg in (3, 4):
    cpp_ls.als(p["user_ids"], p["item_ids"], p["ratings"], 64, 80, 70, -1e300, 1, alg,
               user_factors=p["user_factors0"], item_factors=p["item_factors0"])
p = synth.als_problem(40, 30, 1100, 50, seed=3, min_degrees=False)   # headline rank, ragged rows
cpp_ls.als(p["user_ids"], p["item_ids"], p["ratings"], 50, 40, 30, -1e300, 1, 4,
           user_factors=p["user_factors0"], item_factors=p["item_factors0"])
p = synth.als_problem(3000, 12, 30000, 10, seed=4, min_degrees=False)  # segmented heavy rows
cpp_ls.als(p["user_ids"], p["item_ids"], p["ratings"], 10, 3000, 12, -1e300, 1, 4,
           user_factors=p["user_factors0"], item_factors=p["item_factors0"])
rowptr, col, vals, cols, b, x0, _ = synth.random_sparse_system(700, 60, 5, seed=5)
for alg in (1, 2, 3):
    cpp_ls.cg_least_squares(rowptr, col, vals, cols, b, algorithm=alg, x0=x0)
M = np.random.default_rng(0).standard_normal((300, 50))
similarity.factor_cosine_topk(M, topk=50)
similarity.factor_cosine_topk(M[:70, :5], topk=20)
genres, ratings = synthetic_catalogue(num_movies=80, num_users=120, density=0.5, seed=1)
f = SimilarMovieFinder(genres, ratings)
f.build()
f.build(num_results=3)
f.close()
p = synth.als_problem(60, 50, 1500, 5, seed=6, min_degrees=False)
fold_in_users(p["user_ids"], p["item_ids"], p["ratings"], 60, p["item_factors0"], 5)
print("sanitize_small: done")

"""Generates tests/golden/similar_*.json by running the REAL reference class
(/root/reference/python/full_data/build_similar_movies_db.py:SimilarMovieFinder) on small seeded
catalogues.  The reference module imports its siblings at import time (movie_lens_data creates
./data/*, cpp_ls loads ./cpp_ls_lib.so), so it is imported from a scratch working directory that
holds the compiled reference library.  Run in the build container:
    python tests/golden/make_golden_similar.py
"""
import json
import os
import shutil
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference_class():
    from oracle import oracle
    work = tempfile.mkdtemp(prefix="refsim_")
    os.makedirs(os.path.join(work, "data"))
    shutil.copy(oracle.ref_path(), os.path.join(work, "cpp_ls_lib.so"))
    os.chdir(work)
    sys.path.insert(0, "/root/reference/python/full_data")
    import build_similar_movies_db as ref
    return ref.SimilarMovieFinder


def main():
    from oracle.similar_oracle import synthetic_catalogue
    cwd = os.getcwd()
    Finder = load_reference_class()
    cases = {"small": dict(num_movies=120, num_users=300, density=0.25, seed=1),
             "dense": dict(num_movies=60, num_users=150, density=0.9, seed=2),
             "cut": dict(num_movies=500, num_users=120, density=0.6, num_genres=2, seed=3)}
    for name, kw in cases.items():
        genres, ratings = synthetic_catalogue(**kw)
        f = Finder(genres, ratings)
        out = {"params": kw, "results": {}, "results_top3": {}}
        for i in range(len(ratings)):
            ids, scores = f.find_similar_movie(i)
            out["results"][str(i)] = [list(ids), [float(s).hex() for s in scores]]
            ids3, scores3 = f.find_similar_movie(i, num_results=3)      # exercises the 60-candidate cut
            out["results_top3"][str(i)] = [list(ids3), [float(s).hex() for s in scores3]]
        # tune on the two most co-rated movies of the first query's neighbourhood
        a = ratings[0][0]
        ids, _ = f.find_similar_movie(0, num_results=40)
        b = None
        for cand in ids[3:]:        # a partner outside the current top-2 with enough common raters
            if f._scaled_dot_product(0, f.find_movie_index(cand))[1] > 5:
                b = cand
                break
        if b is not None:
            f.tune(a, b, 2, 20)
            out["tune"] = {"movie_id1": a, "movie_id2": b, "buff_point": f.buff_point,
                           "buff_limit": float(f.buff_limit).hex()}
            ids2, scores2 = f.find_similar_movie(f.find_movie_index(a))
            out["tune"]["after"] = [list(ids2), [float(s).hex() for s in scores2]]
        os.chdir(cwd)
        with open(os.path.join(HERE, "similar_%s.json" % name), "w") as fh:
            json.dump(out, fh)
        print(name, "queries", len(ratings), "with results",
              sum(1 for v in out["results"].values() if v[0]), "tune" in out and out["tune"]["buff_point"])


if __name__ == "__main__":
    main()

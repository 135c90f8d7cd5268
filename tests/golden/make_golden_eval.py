"""Generates tests/golden/eval_small.json by running the REAL reference functions
(/root/reference/python/full_data/worker_process.py:_test_model with als_predictor.ALS_Model and
my_util.compute_ranking_agreement) on the seeded case of oracle/eval_oracle.py.  The reference
modules import their siblings at import time (movie_lens_data creates ./data/*, cpp_ls loads
./cpp_ls_lib.so), so they are imported from a scratch working directory that holds the compiled
reference library.  Run in the build container:  python tests/golden/make_golden_eval.py"""
import json
import os
import shutil
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    from oracle import oracle
    from oracle.eval_oracle import synthetic_eval_case
    cwd = os.getcwd()
    work = tempfile.mkdtemp(prefix="refeval_")
    os.makedirs(os.path.join(work, "data"))
    shutil.copy(oracle.ref_path(), os.path.join(work, "cpp_ls_lib.so"))
    os.chdir(work)
    sys.path.insert(0, "/root/reference/python/full_data")
    import als_predictor
    import worker_process
    out = {}
    for name, kw in {"small": dict(seed=0), "wide": dict(num_users=25, num_movies=200, k=11, seed=3)}.items():
        tests, medians, uf, als_user_ids, itf, als_movie_ids, k = synthetic_eval_case(**kw)
        res = []
        for user_id, movie_ratings in tests:
            row = als_user_ids[user_id]
            model = als_predictor.ALS_Model(uf[(k + 1) * row:(k + 1) * (row + 1)], medians, itf, als_movie_ids)
            ag = worker_process._test_model(model, movie_ratings)
            if ag is not None:
                res.append([user_id, float(ag).hex()])
        out[name] = {"params": kw, "agreements": res}
    os.chdir(cwd)
    with open(os.path.join(HERE, "eval_small.json"), "w") as f:
        json.dump(out, f, indent=1)
    print({k: len(v["agreements"]) for k, v in out.items()})


if __name__ == "__main__":
    main()

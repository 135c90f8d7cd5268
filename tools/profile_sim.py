"""Small driver for ncu: one factor-cosine top-k call (config 4 shape, one wave of query tiles).
usage: python tools/profile_sim.py [N] [Q_HI]"""
import sys
import numpy as np
sys.path.insert(0, ".")
from movie_recommender_b200 import similarity
n = int(sys.argv[1]) if len(sys.argv) > 1 else 53889
q_hi = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 128
M = np.random.default_rng(20181001).standard_normal((n, 50))
ids, scores, info = similarity.factor_cosine_topk(M, topk=50, q_lo=0, q_hi=min(q_hi, n))
print("candidates %.3f ms, total %.3f ms, fallback rows %d" % (info.candidates_ms, info.total_ms, info.fallback_rows))

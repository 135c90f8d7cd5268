"""movie_recommender_b200 -- B200-native (sm_100a CUDA) implementation of the ALS / sparse
least-squares / movie-similarity hot path of louisyang2015/movie_recommender, behind the
reference's own boundary: the C ABI of ``cpp_ls_lib.so`` (include/cpp_ls_b200.h) and the Python
signatures of ``python/full_data/cpp_ls.py``.

There is no CPU fallback: every compute call goes to the CUDA library and fails loudly when the
library or a GPU is missing.
"""
__all__ = ["cpp_ls", "synth"]

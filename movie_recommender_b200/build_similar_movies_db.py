"""Drop-in for the reference's ``SimilarMovieFinder``
(``python/full_data/build_similar_movies_db.py:21-221``): same constructor, same methods, same
return values; the all-pairs work runs on the GPU (csrc/cosim.cu) and is bit-exact.

    finder = SimilarMovieFinder(movie_genres, movie_ratings)       # same arguments as the reference
    finder.tune(1196, 1210, 2, 20)                                  # :183-221
    ids, scores = finder.find_similar_movie(index)                  # :151-180
    similar_movies = finder.build()                                 # what build_locally() pickles, :255-288

Differences, all deliberate:
* ratings must lie on the 0.5 grid with 0 <= rating <= 10 (MovieLens ratings are 0.5 ... 5.0); the
  kernel accumulates the co-rating sums as exact integers.
* a movie whose genre set is EMPTY is treated as having no genre entry (the reference divides by
  zero there, :68); at most 64 distinct genre ids.
* ``build`` computes every movie in one launch instead of the reference's multiprocessing /
  cluster fan-out (``build_locally`` / ``build_with_cluster``, out of scope as control plane).
"""
import ctypes
import math

import numpy

from . import _lib

_dll = _lib.dll


class SimilarMovieFinder:
    """ Class for finding similar movies.
    ::
        movie_genres - {movie id: set of genre ids}
        movie_ratings - [(movie_id, {user_id: rating})]
        buff_limit - maximum boost to score
        buff_point - number of common reviewers at max boost
    """

    def __init__(self, movie_genres, movie_ratings, buff_limit=0.05, buff_point=100):
        self.movie_genres = movie_genres
        self.movie_ratings = movie_ratings
        self.buff_limit = buff_limit
        self.buff_point = buff_point
        self._h = ctypes.c_void_p()
        if movie_ratings is not None:
            self._upload()

    @classmethod
    def from_arrays(cls, movie_genres, movie_ids, movie_index_of_rating, user_ids, ratings,
                    buff_limit=0.05, buff_point=100):
        """Array form of the same input (extension): ``movie_ids[i]`` is the id of list entry i;
        rating j belongs to list entry ``movie_index_of_rating[j]`` (non-decreasing), user
        ``user_ids[j]``.  Avoids building 27 M-entry Python dicts for catalogue-scale runs;
        ``tune`` needs the dict form."""
        self = cls(movie_genres, None, buff_limit, buff_point)
        self._movie_ids = numpy.asarray(movie_ids, dtype=numpy.int64)
        movie_of = numpy.ascontiguousarray(movie_index_of_rating, dtype=numpy.int32)
        if len(movie_of) and numpy.any(numpy.diff(movie_of) < 0):
            raise ValueError("ratings must be grouped by movie list index")
        self._from_arrays(len(self._movie_ids), movie_of, numpy.asarray(user_ids, dtype=numpy.int64),
                          numpy.asarray(ratings, dtype=numpy.float64))
        return self

    # ------------------------------------------------------------------ host-side marshalling
    def _upload(self):
        mr = self.movie_ratings
        n_movies = len(mr)
        self._movie_ids = numpy.array([m[0] for m in mr], dtype=numpy.int64)
        deg = numpy.array([len(m[1]) for m in mr], dtype=numpy.int64)
        nnz = int(deg.sum())
        users = numpy.empty(nnz, dtype=numpy.int64)
        rates = numpy.empty(nnz, dtype=numpy.float64)
        pos = 0
        for _, d in mr:
            k = len(d)
            users[pos:pos + k] = numpy.fromiter(d.keys(), dtype=numpy.int64, count=k)
            rates[pos:pos + k] = numpy.fromiter(d.values(), dtype=numpy.float64, count=k)
            pos += k
        movie_of = numpy.repeat(numpy.arange(n_movies, dtype=numpy.int32), deg)
        self._from_arrays(n_movies, movie_of, users, rates)

    def _from_arrays(self, n_movies, movie_of, users, rates):
        rq = numpy.round(rates * 2.0)
        if len(rates) and (numpy.any(rq != rates * 2.0) or rq.min() < 0 or rq.max() > 20):
            raise ValueError("ratings must lie on the 0.5 grid in [0, 10]")
        uniq, dense_user = numpy.unique(users, return_inverse=True)
        n_users = len(uniq)
        dense_user = dense_user.astype(numpy.int32)
        rq = rq.astype(numpy.uint8)
        self._m_ptr = numpy.concatenate([[0], numpy.cumsum(numpy.bincount(
            movie_of, minlength=n_movies))]).astype(numpy.int32)
        m_user, m_rq = dense_user, rq                                  # already grouped by movie
        order = numpy.argsort(dense_user, kind="stable")
        u_ptr = numpy.concatenate([[0], numpy.cumsum(numpy.bincount(
            dense_user, minlength=n_users))]).astype(numpy.int32)
        u_movie = numpy.ascontiguousarray(movie_of[order], dtype=numpy.int32)
        u_rq = numpy.ascontiguousarray(rq[order])
        # genres -> bit masks
        genre_bit = {}
        mask = numpy.zeros(max(n_movies, 1), dtype=numpy.uint64)
        cnt = numpy.zeros(max(n_movies, 1), dtype=numpy.int32)
        for i, mid in enumerate(self._movie_ids.tolist()):
            g = self.movie_genres.get(mid)
            if not g:
                continue
            if len(set(g)) != len(g):
                raise ValueError("genres of movie %r hold duplicates (the reference takes sets)" % (mid,))
            m = 0
            for gid in g:
                if gid not in genre_bit:
                    if len(genre_bit) >= 64:
                        raise ValueError("more than 64 distinct genre ids")
                    genre_bit[gid] = len(genre_bit)
                m |= 1 << genre_bit[gid]
            mask[i] = m
            cnt[i] = len(g)
        self._max_deg = int(numpy.diff(self._m_ptr).max()) if n_movies else 0
        self._n_movies = n_movies
        m_user = numpy.ascontiguousarray(m_user)
        m_rq = numpy.ascontiguousarray(m_rq)
        ub = ctypes.POINTER(ctypes.c_ubyte)
        _lib.check(_dll.mrb_cosim_create(
            n_movies, n_users, _lib.ip(self._m_ptr), _lib.ip(m_user), m_rq.ctypes.data_as(ub),
            _lib.ip(u_ptr), _lib.ip(u_movie), u_rq.ctypes.data_as(ub),
            mask.ctypes.data_as(ctypes.POINTER(ctypes.c_ulonglong)), _lib.ip(cnt),
            ctypes.byref(self._h)))

    def close(self):
        if self._h:
            _dll.mrb_cosim_destroy(self._h)
            self._h = ctypes.c_void_p()

    __del__ = close

    def _buff_table(self):
        """buff(n) for n = 0 .. max common raters, with the reference's own expressions and libm
        calls (build_similar_movies_db.py:109-119) so that the table is bit-identical."""
        buff_limit, buff_point = self.buff_limit, self.buff_point
        table = numpy.zeros(self._max_deg + 2, dtype=numpy.float64)
        x_limit = 3 * math.exp(buff_limit)
        for n in range(3, len(table)):
            x = 3 + (x_limit - 3) * (n - 3) / (buff_point - 3)
            buff = math.log(x) - math.log(3)
            if buff > buff_limit: buff = buff_limit
            if buff < 0: buff = 0
            table[n] = buff
        return table

    def _query(self, q_lo, q_hi, num_results):
        nq = q_hi - q_lo
        idx = numpy.full((nq, num_results), -1, dtype=numpy.int32)
        score = numpy.zeros((nq, num_results), dtype=numpy.float64)
        count = numpy.zeros(max(nq, 1), dtype=numpy.int32)
        table = self._buff_table()
        ms = ctypes.c_float(0)
        _lib.check(_dll.mrb_cosim_query(self._h, q_lo, q_hi, _lib.dp(table), len(table),
                                        num_results, _lib.ip(idx), _lib.dp(score), _lib.ip(count),
                                        ctypes.byref(ms)))
        self.last_kernel_ms = ms.value
        return idx, score, count[:nq]

    # ------------------------------------------------------------------ the reference's methods
    def _pair(self, a, b):
        """Device call: (common raters, cosine over them) of movies ``a`` and ``b``."""
        n = ctypes.c_int(0)
        sim = ctypes.c_double(0.0)
        _lib.check(_dll.mrb_cosim_pair(self._h, a, b, ctypes.byref(n), ctypes.byref(sim)))
        return n.value, sim.value

    def _scaled_dot_product(self, movie_id1_index, movie_id2_index):
        """(final_score, common_reviewers, pre_boost_score) of one pair (:72-119), used by
        ``tune`` exactly as in the reference.  The common raters, their dot product and norms --
        the part that touches the ratings -- come from the device (``mrb_cosim_pair``, exact
        integer sums, the cosine in the reference's operation order); the scalar boost below is
        the reference's own expression, textually, so that libm rounds it the same way."""
        n, similarity = self._pair(movie_id1_index, movie_id2_index)
        if n < 3: return 0.0, n, 0.0
        x_limit = 3 * math.exp(self.buff_limit)
        x = 3 + (x_limit - 3) * (n - 3) / (self.buff_point - 3)
        buff = math.log(x) - math.log(3)
        if buff > self.buff_limit: buff = self.buff_limit
        if buff < 0: buff = 0
        return similarity * (1.0 + buff), n, similarity

    @property
    def num_movies(self):
        """Length of the movie list (the range ``build`` and the multi-GPU split index)."""
        return self._n_movies

    def find_movie_index(self, movie_id: int):
        """Return the "movie_ratings" list index for movie_id, -1 if absent (:136-147)."""
        hit = numpy.flatnonzero(self._movie_ids == movie_id)
        return int(hit[0]) if len(hit) else -1

    def find_similar_movie(self, movie_id_index: int, num_results=20):
        """Returns movie_ids, similarity_scores (:151-180)."""
        idx, score, count = self._query(movie_id_index, movie_id_index + 1, num_results)
        c = int(count[0])
        if c == 0:
            return [], []
        return (tuple(int(self._movie_ids[i]) for i in idx[0, :c]),
                tuple(float(s) for s in score[0, :c]))

    def tune(self, movie_id1, movie_id2, top_n, expected_search_size):
        """Tune "buff_limit" and "buff_point" such that "movie_id2" shows up in the "top_n" of
        the movies similar to "movie_id1" (:183-221)."""
        index1 = self.find_movie_index(movie_id1)
        index2 = self.find_movie_index(movie_id2)
        final_score, common_reviewers, pre_boost_score = self._scaled_dot_product(index1, index2)
        self.buff_point = common_reviewers
        self.buff_limit = 0
        while self.buff_limit < 2:
            movie_ids, scores = self.find_similar_movie(index1, num_results=expected_search_size * 2)
            movie_ids = movie_ids[:top_n]
            for movie_id in movie_ids:
                if movie_id == movie_id2: return
            top_score = scores[0]
            score, common_reviewers, pre_boost_score = self._scaled_dot_product(index1, index2)
            self.buff_limit = self.buff_limit * top_score / score + 0.01

    def build(self, num_results=20, start=0, length=None):
        """{movie_id: (similar movie ids)} for the movies start .. start+length-1 of the list --
        the dictionary the reference's workers produce (movie_lens_data_proc.py:657-700) and
        ``build_locally`` pickles as similar_movies.bin.  The multi-GPU split is by this range."""
        length = self._n_movies - start if length is None else length
        idx, score, count = self._query(start, start + length, num_results)
        out = {}
        for q in range(length):
            c = int(count[q])
            if c > 0:
                out[int(self._movie_ids[start + q])] = tuple(int(self._movie_ids[i]) for i in idx[q, :c])
        return out

#!/bin/bash
# timing of several builds of the library (tmp_ab/lib<name>.so, see tools/build_variant.sh) on
# the same box: tools/ab.sh "<names>" <reps> <profile_als args...>; the in-tree library is restored
names=$1; reps=$2; shift 2
cp movie_recommender_b200/cpp_ls_lib.so /tmp/lib_keep.so
for r in $(seq $reps); do
  for v in $names; do
    cp tmp_ab/lib$v.so movie_recommender_b200/cpp_ls_lib.so
    echo -n "$v: "; python tools/profile_als.py "$@" | tail -1
  done
done
cp /tmp/lib_keep.so movie_recommender_b200/cpp_ls_lib.so

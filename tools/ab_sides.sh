#!/bin/bash
# per-side timing of several library builds (tmp_ab/lib<name>.so) on the same box:
# tools/ab_sides.sh "<names>" <reps> [side_times args]; the in-tree library is restored
names=$1; reps=$2; shift 2
cp movie_recommender_b200/cpp_ls_lib.so /tmp/lib_keep.so
for r in $(seq $reps); do
  for v in $names; do
    cp tmp_ab/lib$v.so movie_recommender_b200/cpp_ls_lib.so
    echo -n "$v: "; python tools/side_times.py "$@" | tail -1
  done
done
cp /tmp/lib_keep.so movie_recommender_b200/cpp_ls_lib.so

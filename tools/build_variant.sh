#!/bin/bash
# build a variant of the library with extra -D flags for als_gram.cu into tmp_ab/lib<name>.so
# usage: tools/build_variant.sh <name> [-DFOO=1 ...]
set -e
name=$1; shift
cd "$(dirname "$0")/../movie_recommender_b200/csrc"
NV=/usr/local/cuda/bin/nvcc
$NV -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr -Xptxas -v "$@" -c als_gram.cu -o build/als_gram_$name.o 2> build/als_gram_$name.log
grep -A1 "k_gramILi7ELb[01]ELi0" build/als_gram_$name.log | grep -E "registers|spill" || true
objs=$(ls build/*.o | grep -v als_gram)
$NV -gencode arch=compute_100a,code=sm_100a -shared -o ../../tmp_ab/lib$name.so $objs build/als_gram_$name.o -lcudart_static -lpthread -ldl -lrt
rm build/als_gram_$name.o

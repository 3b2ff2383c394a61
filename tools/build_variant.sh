#!/bin/bash
# tools/build_variant.sh <tag> <nvcc -D flags...>: libcoffeedb_b200_<tag>.so with locate.cu compiled under the given macros
# (A/B runs on the GPU box: CDB_LIB=$PWD/coffeedb_b200/libcoffeedb_b200_<tag>.so python bench.py ...)
set -e
cd "$(dirname "$0")/../coffeedb_b200/csrc"
tag=$1; shift
/usr/local/cuda/bin/nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wall,-Wno-unused-function \
  --expt-relaxed-constexpr "$@" -c locate.cu -o _obj/locate_$tag.o
objs=$(ls _obj/*.o | grep -v '_obj/locate' | tr '\n' ' ')
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libcoffeedb_b200_$tag.so $objs _obj/locate_$tag.o -lcudart
echo built ../libcoffeedb_b200_$tag.so

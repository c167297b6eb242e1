#!/bin/bash
# A/B builds of one kernel file: tools/build_variant.sh <tag> <file.cu> <nvcc flags...> -> segmminterest_b200/build/variants/libmmi_<tag>.so
# (the other objects come from the regular build; run the variant with MMI_LIB_PATH=<that .so>)
set -e
tag=$1; src=$2; shift 2
here=$(cd "$(dirname "$0")/.." && pwd)/segmminterest_b200
mkdir -p $here/build/variants
obj=$here/build/variants/$(basename $src .cu)_$tag.o
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c $here/csrc/$src -o $obj
others=$(ls $here/build/*.o | grep -v "/$(basename $src .cu).o")
nvcc -shared -o $here/build/variants/libmmi_$tag.so $obj $others
echo $here/build/variants/libmmi_$tag.so

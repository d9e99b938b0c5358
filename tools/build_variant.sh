#!/bin/bash
# Experiment builds: tools/build_variant.sh <name> <source.cu> [-DFLAG ...]
# compiles ONE translation unit with extra flags and links it with the other objects of the
# library into praline_b200/libpraline_b200_<name>.so (git-ignored; select it with PGPU_LIB=...).
set -e
name=$1; src=$2; shift 2
cd "$(dirname "$0")/../praline_b200/csrc"
make -s -j8 >/dev/null
obj=/tmp/variant_${name}.o
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas -v "$@" -c $src -o $obj 2> /tmp/variant_${name}.ptxas.log
others=$(ls *.o | grep -v "^${src%.cu}.o$")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libpraline_b200_${name}.so $obj $others -lcudart
echo "built praline_b200/libpraline_b200_${name}.so"

#!/bin/bash
# build a variant of the engine: tools/build_variant.sh <name> <file.cu> [nvcc -D flags ...] -> variants/libzkc_<name>.so
# (the other objects are taken from build/obj as built by build.py; select at run time with ZKC_B200_LIB=variants/libzkc_<name>.so)
set -e
cd "$(dirname "$0")/.."
name=$1; file=$2; shift 2
mkdir -p variants
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -cudart static"
nvcc $F "$@" -c era_zkevm_circuits_b200/csrc/$file -o variants/$name.$file.o
objs=$(ls build/obj/*.o | grep -v "/$file.o")
nvcc $F -shared -o variants/libzkc_$name.so $objs variants/$name.$file.o
rm variants/$name.$file.o
echo variants/libzkc_$name.so

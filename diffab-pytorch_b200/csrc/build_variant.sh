#!/bin/bash
# build_variant.sh <name> <source.cu> <extra nvcc flags...>: libdiffab_b200_<name>.so = the library with ONE source recompiled
# with extra -D flags (kernel experiments; select it with DAB_LIB_VARIANT=libdiffab_b200_<name>.so)
set -e
name=$1; src=$2; shift 2
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
$NVCC $ARCH -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c $src -o ${src%.cu}.$name.o
objs=$(ls *.o | grep -v '\.dbg\.o$' | grep -Ev '\.[A-Za-z0-9_]+\.o$' | grep -v "^${src%.cu}.o$")
$NVCC $ARCH -shared -o libdiffab_b200_$name.so $objs ${src%.cu}.$name.o -lcuda
echo built libdiffab_b200_$name.so

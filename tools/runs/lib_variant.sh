#!/bin/bash
# lib_variant.sh NAME SRC.cu "-DFLAGS"  -- build/variants/libjb_NAME.so = the product library with csrc/SRC.cu recompiled under extra flags
set -e
HERE=$(cd "$(dirname "$0")/../.." && pwd)
PKG=$HERE/juicy-audio-plugins_b200
name=$1; src=$2; flags=$3
mkdir -p $PKG/build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -ftz=true -prec-div=true -prec-sqrt=true \
     -Xcompiler -fPIC $flags -c $PKG/csrc/$src -o $PKG/build/variants/${src}_$name.o
objs=$(ls $PKG/build/*.o | grep -v "/$src.o")
g++ -shared -o $PKG/build/variants/libjb_$name.so $objs $PKG/build/variants/${src}_$name.o -L/usr/local/cuda/lib64 -lcudart_static -ldl -lrt -lpthread
echo built libjb_$name.so

#!/bin/bash
# coop_variants.sh NAME "-DFLAG=.. ..."  -- build juicy-audio-plugins_b200/build/variants/libjb_NAME.so: the product library with
# csrc/jb_coop.cu compiled under extra -D flags (kernel-variant A/B runs: JUICY_BATCH_LIB=... python tools/chain_bench.py ...).
set -e
HERE=$(cd "$(dirname "$0")/../.." && pwd)
PKG=$HERE/juicy-audio-plugins_b200
name=$1; flags=$2
mkdir -p $PKG/build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -ftz=true -prec-div=true -prec-sqrt=true \
     -Xcompiler -fPIC $flags -c $PKG/csrc/jb_coop.cu -o $PKG/build/variants/jb_coop_$name.o
objs=$(ls $PKG/build/*.o | grep -v jb_coop.cu.o)
g++ -shared -o $PKG/build/variants/libjb_$name.so $objs $PKG/build/variants/jb_coop_$name.o -L/usr/local/cuda/lib64 -lcudart_static -ldl -lrt -lpthread
echo built $PKG/build/variants/libjb_$name.so

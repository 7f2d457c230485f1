#!/bin/bash
# A/B kernel experiments: rebuild ONE source with extra flags and link it with the stock objects into
# gpurun_out/variants/libsonic_<tag>.so (select it with SONIC_LIB=...).  Usage: tools/build_variant.sh <tag> <source.cu> <flags...>
set -e
TAG=$1; SRC=$2; shift 2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
CS=$ROOT/sonicdiffusionbayeslab_b200/csrc
OUT=$ROOT/variants
mkdir -p $OUT/obj
make -C $CS -j8 > /dev/null
ARCH="-gencode arch=compute_100a,code=sm_100a"
nvcc -O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas -v "$@" -c $CS/$SRC -o $OUT/obj/${TAG}_${SRC%.cu}.o 2> $OUT/obj/${TAG}_${SRC%.cu}.ptxas.log
OBJS=""
for o in $CS/build/*.o; do
  if [ "$(basename $o)" == "${SRC%.cu}.o" ]; then OBJS="$OBJS $OUT/obj/${TAG}_${SRC%.cu}.o"; else OBJS="$OBJS $o"; fi
done
nvcc $ARCH -shared -o $OUT/libsonic_$TAG.so $OBJS -cudart shared
echo built $OUT/libsonic_$TAG.so
